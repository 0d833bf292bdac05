// kh/sharded_host.hpp -- several ranks (GPUs) driven from ONE host process, over the C ABI.
//
// The reference runs P UPC++ ranks (`srun -n P ./kmer_hash ...`): rank r parses its block of the input
// (read_kmers.hpp:55-58), the table is partitioned by owner rank (hash_map.hpp:28-30), inserts travel
// as one batch per destination (hash_map.hpp:64-77), and rank r writes `<prefix>_<r>.dat` with the
// contigs that start in its block (kmer_hash.cpp:27-31, 60-67).  This header is that flow for P tables
// in one process: rank r lives on GPU r % (visible GPUs); the batches move with device-to-device
// copies (NVLink peer copies between GPUs); lookups never leave the owner GPU (kh_shard_walk /
// kh_shard_resolve, see kh_capi.h).  The multi-process variant (one process per GPU, NCCL) is
// cs267_hw3_b200/sharded.py; both drive the same kernels.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kh_capi.h"

namespace kh_sharded {

inline void must(int status, kh_table* t, const char* what) {
    if (status == KH_OK) return;
    const char* detail = t ? kh_last_error(t) : "";
    throw std::runtime_error(std::string(detail && *detail ? detail : kh_status_string(status)) + " [" + what + "]");
}

// k-mers one shard must be able to hold: its share plus slack (whole supermers move together)
inline uint64_t shard_capacity(uint64_t n_total, int world) {
    const double share = (double)n_total / world;
    return (uint64_t)(share * 1.005 + 40.0 * std::sqrt(share + 1.0)) + 1024;
}

struct RankOutput {
    std::vector<char> text;       // contigs of this rank, '\n'-terminated, in start-line order
    uint64_t n_contigs = 0, n_nodes = 0;
};

class Cluster {
    int k_, world_;
    std::vector<kh_table*> t_;
    std::vector<int> dev_;
    std::vector<void*> recv_;                 // per-rank receive buffer (device memory on that rank's GPU)
    std::vector<uint64_t> recv_cap_;

    void barrier() { for (kh_table* t : t_) must(kh_sync(t), t, "sync"); }

    // sends[s] = (device pointer of groups in destination order, counts per destination); elements of `elem` bytes
    std::vector<std::pair<const void*, uint64_t>> exchange(const std::vector<const void*>& ptr,
                                                           const std::vector<std::vector<uint64_t>>& counts, uint64_t elem) {
        std::vector<std::pair<const void*, uint64_t>> out(world_);
        for (int d = 0; d < world_; ++d) {
            uint64_t n_recv = 0;
            for (int s = 0; s < world_; ++s) n_recv += counts[s][d];
            const uint64_t bytes = std::max<uint64_t>(n_recv * elem, 256);
            if (recv_cap_[d] < bytes) {
                if (recv_[d]) kh_device_free(recv_[d]);
                recv_[d] = nullptr;
                must(kh_device_alloc_on(dev_[d], &recv_[d], bytes + bytes / 4), t_[d], "receive buffer");
                recv_cap_[d] = bytes + bytes / 4;
            }
            uint64_t off = 0;
            for (int s = 0; s < world_; ++s) {
                uint64_t before = 0;
                for (int x = 0; x < d; ++x) before += counts[s][x];
                const char* src = static_cast<const char*>(ptr[s]) + before * elem;
                must(kh_copy_device(t_[d], static_cast<char*>(recv_[d]) + off, src, counts[s][d] * elem), t_[d], "exchange");
                off += counts[s][d] * elem;
            }
            out[d] = {recv_[d], n_recv};
        }
        barrier();
        return out;
    }

  public:
    // n_local_max: most records any rank parses; n_total: all records
    Cluster(int k, int world, uint64_t n_local_max, uint64_t n_total, double load_factor)
        : k_(k), world_(world), t_(world, nullptr), dev_(world, 0), recv_(world, nullptr), recv_cap_(world, 0) {
        const int ndev = kh_device_count();
        if (ndev <= 0) throw std::runtime_error("no CUDA device (libkh_b200 has no CPU fallback)");
        if (world < 1 || world > 8) throw std::runtime_error("KH_RANKS must be between 1 and 8");
        for (int r = 0; r < world; ++r) {
            dev_[r] = r % ndev;
            must(kh_create(k, shard_capacity(n_total, world), load_factor, dev_[r], &t_[r]), nullptr, "kh_create");
            must(kh_shard_init(t_[r], r, world, n_local_max, n_total, n_local_max), t_[r], "kh_shard_init");
        }
        for (int r = 0; r < world; ++r) must(kh_shard_connect_local(t_[r], t_.data(), world), t_[r], "kh_shard_connect_local");
    }
    ~Cluster() {
        for (void* p : recv_) if (p) kh_device_free(p);
        for (kh_table* t : t_) kh_destroy(t);
    }
    Cluster(const Cluster&) = delete;
    Cluster& operator=(const Cluster&) = delete;

    int world() const { return world_; }
    kh_table* table(int rank) { return t_[rank]; }
    int device(int rank) const { return dev_[rank]; }

    // initialize_kmers across ranks (kmer_hash.cpp:21-33): records_dev[r] = rank r's block, resident on its GPU
    void insert(const std::vector<const void*>& records_dev, const std::vector<uint64_t>& n) {
        std::vector<const void*> ptr(world_);
        std::vector<std::vector<uint64_t>> counts(world_, std::vector<uint64_t>(8, 0));
        for (int r = 0; r < world_; ++r)
            must(kh_shard_owner_partition(t_[r], records_dev[r], n[r], &ptr[r], counts[r].data()), t_[r], "owner_partition");
        auto recv = exchange(ptr, counts, kh_slot_bytes(k_));
        for (int r = 0; r < world_; ++r) must(kh_insert_slots_device(t_[r], recv[r].first, recv[r].second), t_[r], "insert");
        barrier();                                   // hash_map.hpp:79: every insert visible before any find
    }

    // assemble_contigs across ranks (kmer_hash.cpp:38-55); returns what each rank writes to <prefix>_<rank>.dat
    std::vector<RankOutput> assemble() {
        std::vector<const void*> ptr(world_);
        std::vector<std::vector<uint64_t>> counts(world_, std::vector<uint64_t>(8, 0));
        uint64_t link_bytes = 0;
        for (int r = 0; r < world_; ++r) must(kh_shard_walk(t_[r], &ptr[r], counts[r].data(), &link_bytes), t_[r], "walk");
        auto recv = exchange(ptr, counts, link_bytes);
        for (int r = 0; r < world_; ++r) must(kh_shard_resolve(t_[r], recv[r].first, recv[r].second), t_[r], "resolve");
        barrier();
        for (int round = 0; round < 12; ++round) {           // batches of pointer-jumping rounds
            int any = 0;
            for (int r = 0; r < world_; ++r) {
                int moved = 0;
                must(kh_shard_phase(t_[r], 1, &moved), t_[r], "rank");
                any |= moved;
            }
            barrier();
            if (!any) break;
        }
        for (int phase = 2; phase <= 4; ++phase)             // lengths, tail claims, offsets: no barrier needed in between
            for (int r = 0; r < world_; ++r) must(kh_shard_phase(t_[r], phase, nullptr), t_[r], "finish");
        barrier();
        for (int r = 0; r < world_; ++r) must(kh_shard_phase(t_[r], 5, nullptr), t_[r], "emit");
        barrier();
        int bits = 0;
        for (int r = 0; r < world_; ++r) {
            int b = 0;
            must(kh_shard_phase(t_[r], 6, &b), t_[r], "collect");
            bits |= b;
        }
        if (bits & 8) throw std::runtime_error("malformed input (bad base, or a backward extension that does not name the predecessor)");
        if (bits & 2) throw std::runtime_error("hash table full");
        if (bits & 1) throw std::runtime_error("Error: k-mer not found in Distributed HashMap.");      // kmer_hash.cpp:48
        if (bits & 4) throw std::runtime_error("a start-rooted chain never reaches forward extension 'F' (cycle)");
        if (bits & 16) throw std::runtime_error("two start nodes reach the same end node (chains are not linear)");
        if (bits) throw std::runtime_error("internal error: segment bookkeeping overflow");
        std::vector<RankOutput> out(world_);
        for (int r = 0; r < world_; ++r) {
            const char* dev_text = nullptr;
            uint64_t bytes = 0;
            must(kh_shard_result(t_[r], &dev_text, nullptr, &out[r].n_contigs, &bytes, &out[r].n_nodes), t_[r], "result");
            out[r].text.resize(bytes);
            must(kh_copy_to_host(t_[r], out[r].text.data(), dev_text, bytes), t_[r], "copy result");
        }
        return out;
    }
};

}  // namespace kh_sharded
