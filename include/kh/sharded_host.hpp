// kh/sharded_host.hpp -- several ranks (GPUs) driven from ONE host process, over the C ABI.
//
// The reference runs P UPC++ ranks (`srun -n P ./kmer_hash ...`): rank r parses its block of the input
// (read_kmers.hpp:55-58), the table is partitioned by owner rank (hash_map.hpp:28-30), inserts travel
// as one batch per destination (hash_map.hpp:64-77), finds are RPC round trips (hash_map.hpp:93-100), and rank r
// writes `<prefix>_<r>.dat` with the contigs that start in its block (kmer_hash.cpp:27-31, 60-67).
// This header is that flow for P shards in one process: rank r lives on GPU r % (visible GPUs) and is driven by
// its own host thread on its own stream.  A whole step (insert + traverse) is stream-ordered: the records reach
// their owner GPU through NVLink peer stores inside the grouping kernel, pending chain links and their answers
// travel the same way, and the phases are separated by an in-stream flag barrier -- the host threads only enqueue
// and wait once at the end (kh_shard_finish).  The multi-process variant (one process per GPU, torchrun) is
// cs267_hw3_b200/sharded.py; both drive the same entry points.
//
// When several ranks share one device (more ranks than GPUs: tests, or a laptop-sized box) their streams may
// share a hardware queue, where a waiting barrier kernel would block a peer's work queued behind it.  Then one
// thread enqueues the step part by part for all ranks in turn (every part ends with the barrier), which is
// deadlock-free on any queue mapping.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <exception>
#include <stdexcept>
#include <string>
#include <thread>
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
    return (uint64_t)(share * 1.02 + 64.0 * std::sqrt(share + 1.0)) + 4096;
}

struct RankOutput {
    std::vector<char> text;       // contigs of this rank, '\n'-terminated, in start-line order
    std::vector<uint64_t> offsets;   // n_contigs + 1 byte offsets into text
    std::vector<uint64_t> sorted;    // with want_sorted: this rank's contig indices in bytewise order of their lines
    uint64_t n_contigs = 0, n_nodes = 0;
};

class Cluster {
    int k_, world_;
    bool lockstep_ = false;
    std::vector<kh_table*> t_;
    std::vector<int> dev_;

    // run fn(rank) for every rank: one host thread per rank, or in turn when ranks share a device
    template <class F> void each_rank(F fn) {
        if (lockstep_ || world_ == 1) {
            for (int r = 0; r < world_; ++r) fn(r);
            return;
        }
        std::vector<std::exception_ptr> err(world_);
        std::vector<std::thread> th;
        for (int r = 0; r < world_; ++r)
            th.emplace_back([&, r] { try { fn(r); } catch (...) { err[r] = std::current_exception(); } });
        for (auto& x : th) x.join();
        for (auto& e : err) if (e) std::rethrow_exception(e);
    }

  public:
    // n_local_max: most records any rank parses; n_total: all records
    Cluster(int k, int world, uint64_t n_local_max, uint64_t n_total, double load_factor)
        : k_(k), world_(world), t_(world, nullptr), dev_(world, 0) {
        const int ndev = kh_device_count();
        if (ndev <= 0) throw std::runtime_error("no CUDA device (libkh_b200 has no CPU fallback)");
        if (world < 1 || world > 8) throw std::runtime_error("KH_RANKS must be between 1 and 8");
        lockstep_ = world > ndev;
        for (int r = 0; r < world; ++r) {
            dev_[r] = r % ndev;
            must(kh_create(k, shard_capacity(n_total, world), load_factor, dev_[r], &t_[r]), nullptr, "kh_create");
            must(kh_shard_init(t_[r], r, world, n_local_max, n_total, n_local_max), t_[r], "kh_shard_init");
        }
        for (int r = 0; r < world; ++r) must(kh_shard_connect_local(t_[r], t_.data(), world), t_[r], "kh_shard_connect_local");
    }
    ~Cluster() {
        for (kh_table* t : t_) if (t) kh_sync(t);
        for (kh_table* t : t_) kh_destroy(t);
    }
    Cluster(const Cluster&) = delete;
    Cluster& operator=(const Cluster&) = delete;

    int world() const { return world_; }
    kh_table* table(int rank) { return t_[rank]; }
    int device(int rank) const { return dev_[rank]; }

    // a fresh DistributedHashMap on every rank
    void begin() {
        each_rank([&](int r) { must(kh_shard_begin(t_[r]), t_[r], "begin"); });
    }

    // initialize_kmers across ranks (kmer_hash.cpp:21-33): records_dev[r] = rank r's block, resident on its GPU.
    // Only enqueues; the records travel to their owners inside the grouping kernel.
    void insert(const std::vector<const void*>& records_dev, const std::vector<uint64_t>& n) {
        each_rank([&](int r) { must(kh_shard_insert(t_[r], records_dev[r], n[r]), t_[r], "insert"); });
    }

    // assemble_contigs across ranks (kmer_hash.cpp:38-55); returns what each rank writes to <prefix>_<rank>.dat
    // want_sorted: also sort every rank's contigs on its GPU (kh_sorted_order), for a merged, already-sorted solution
    std::vector<RankOutput> assemble(bool want_sorted = false) {
        if (lockstep_) {
            const int parts = kh_shard_assemble_parts();
            for (int p = 0; p < parts; ++p)
                for (int r = 0; r < world_; ++r) must(kh_shard_assemble_part(t_[r], p), t_[r], "assemble");
        } else {
            each_rank([&](int r) { must(kh_shard_assemble(t_[r]), t_[r], "assemble"); });
        }
        std::vector<int> bits_of(world_, 0);
        each_rank([&](int r) { must(kh_shard_finish(t_[r], &bits_of[r]), t_[r], "finish"); });
        int bits = 0;
        for (int b : bits_of) bits |= b;
        if (bits & 8) throw std::runtime_error("malformed input (a base outside ACGT or an extension outside ACGTF)");
        if (bits & 2) throw std::runtime_error("hash table full");
        if (bits & 1) throw std::runtime_error("Error: k-mer not found in Distributed HashMap.");      // kmer_hash.cpp:48
        if (bits & 4) throw std::runtime_error("a start-rooted chain never reaches forward extension 'F' (cycle)");
        if (bits & 16) throw std::runtime_error("two start nodes reach the same end node (chains are not linear)");
        if (bits) throw std::runtime_error("internal error: a capacity was exceeded or a peer never reached a barrier");
        std::vector<RankOutput> out(world_);
        each_rank([&](int r) {
            const char* dev_text = nullptr;
            const uint64_t* dev_off = nullptr;
            uint64_t bytes = 0;
            must(kh_shard_result(t_[r], &dev_text, &dev_off, &out[r].n_contigs, &bytes, &out[r].n_nodes), t_[r], "result");
            out[r].text.resize(bytes);
            must(kh_copy_to_host(t_[r], out[r].text.data(), dev_text, bytes), t_[r], "copy result");
            if (want_sorted) {
                out[r].offsets.resize(out[r].n_contigs + 1);
                must(kh_copy_to_host(t_[r], out[r].offsets.data(), dev_off, (out[r].n_contigs + 1) * sizeof(uint64_t)), t_[r], "copy offsets");
                out[r].sorted.resize(out[r].n_contigs);
                must(kh_sorted_order(t_[r], out[r].sorted.data(), nullptr), t_[r], "sorted order");
            }
        });
        return out;
    }
};

}  // namespace kh_sharded
