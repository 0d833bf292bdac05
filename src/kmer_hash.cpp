// kmer_hash -- the reference's command line (kmer_hash.cpp:84-152) on the B200 library.
//
//   kmer_hash_<K> kmer_file [verbose|test [prefix]]
//
// Same positional arguments, same stdout lines, same `<prefix>_<rank>.dat` output (one contig
// per line, in start-node order) that scripts/check_it.sh:47-55 sorts and diffs, same
// exceptions on a K mismatch or a missing k-mer.  One process drives one GPU, so it runs as
// "rank 0 of 1"; the GPU and the load factor come from the environment (KH_DEVICE,
// KH_LOAD_FACTOR) because the positional interface has no room for them.
//
// KH_RANKS=P (1..8) runs the reference's P-rank flow in this one process: rank r parses lines
// [ceil(n/P)*r, ...) (read_kmers.hpp:55-58), lives on GPU r % (visible GPUs), and writes <prefix>_<r>.dat
// (kh/sharded_host.hpp); the printed line is rank 0's, as BUtil::print does.
//
// Stage mapping
//   read_kmers (untimed, kmer_hash.cpp:122)      -> file read + K1 pack on the GPU (kh_pack_lines)
//   initialize_kmers (timed, :131)               -> kh_insert_pairs   (records start in HOST memory,
//                                                   exactly like the reference's std::vector<kmer_pair>)
//   assemble_contigs (timed, :135)               -> kh_assemble       (contigs end in HOST memory)
//   output_results (untimed, :147)               -> one fwrite of the contig text
//
// KH_STREAM=1 (single rank): the file is not read up front; DistributedHashMap::insert_file streams it through pinned
// chunk buffers into the GPU pack + insert while a background thread reads ahead (kh/stream_reader.hpp).  The
// "insert" time then INCLUDES reading and parsing the file, which the reference leaves outside its timer.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "butil.hpp"
#include "hash_map.hpp"
#include "kh/sharded_host.hpp"
#include "kmer_t.hpp"
#include "read_kmers.hpp"

namespace {

void must(int status, kh_table* t, const char* what) {
    if (status == KH_OK) return;
    const char* detail = t ? kh_last_error(t) : "";
    throw std::runtime_error(std::string(detail && *detail ? detail : kh_status_string(status)) + " [" + what + "]");
}

// Whole file -> n kmer_pair records in pinned host memory, packed by the GPU (K1).
kmer_pair* read_and_pack(kh_table* t, const std::string& fname, size_t n_kmers) {
    const size_t line_len = KMER_LEN + 4;
    FILE* f = fopen(fname.c_str(), "r");
    if (f == nullptr) throw std::runtime_error("read_kmers: could not open " + fname);
    void* text = nullptr;
    void* pairs = nullptr;
    must(kh_host_alloc(&text, line_len * n_kmers), t, "pinned text buffer");
    must(kh_host_alloc(&pairs, sizeof(kmer_pair) * n_kmers), t, "pinned record buffer");
    const size_t got = fread(text, 1, line_len * n_kmers, f);
    fclose(f);
    if (got != line_len * n_kmers)
        throw std::runtime_error("read_kmers: " + fname + " is not " + std::to_string(n_kmers) + " lines of " +
                                 std::to_string(line_len) + " bytes");
    must(kh_pack_lines(t, static_cast<const char*>(text), n_kmers, pairs), t, "pack_lines");
    kh_host_free(text);
    return static_cast<kmer_pair*>(pairs);
}

// `cat <prefix>_*.dat | sort` (scripts/check_it.sh:47-48) without the sort: every rank's contigs were sorted on its GPU,
// the host only merges the P sorted lists
void write_merged_solution(const char* path, const std::vector<kh_sharded::RankOutput>& outs) {
    FILE* sol = fopen(path, "w");
    if (sol == nullptr) throw std::runtime_error(std::string("could not open ") + path);
    const int P = (int)outs.size();
    std::vector<size_t> pos(P, 0);
    auto line = [&](int r, size_t i, const char*& p, size_t& len) {
        const uint64_t c = outs[r].sorted[i];
        p = outs[r].text.data() + outs[r].offsets[c];
        len = (size_t)(outs[r].offsets[c + 1] - outs[r].offsets[c]);
    };
    for (;;) {
        int best = -1;
        const char* bp = nullptr;
        size_t bl = 0;
        for (int r = 0; r < P; ++r) {
            if (pos[r] >= outs[r].sorted.size()) continue;
            const char* p; size_t len;
            line(r, pos[r], p, len);
            if (best < 0) { best = r; bp = p; bl = len; continue; }
            const int c = memcmp(p, bp, std::min(len, bl));
            if (c < 0 || (c == 0 && len < bl)) { best = r; bp = p; bl = len; }
        }
        if (best < 0) break;
        fwrite(bp, 1, bl, sol);
        ++pos[best];
    }
    fclose(sol);
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        BUtil::print("Usage: srun -N nodes -n ranks ./kmer_hash kmer_file [verbose|test [prefix]]\n");
        exit(1);
    }
    const std::string kmer_fname = argv[1];
    const std::string run_type = argc >= 3 ? argv[2] : "";
    std::string test_prefix = "test";
    if (run_type == "test" && argc >= 4) test_prefix = argv[3];

    const int ks = kmer_size(kmer_fname);
    if (ks != KMER_LEN) {
        throw std::runtime_error("Error: " + kmer_fname + " contains " + std::to_string(ks) +
                                 "-mers, while this binary is compiled for " + std::to_string(KMER_LEN) +
                                 "-mers. Modify packing.hpp and recompile.");
    }

    const size_t n_kmers = line_count(kmer_fname);
    const size_t hash_table_size = n_kmers * 2;      // load factor 0.5 (kmer_hash.cpp:108-109)
    const int rank_id = 0;

    if (run_type == "verbose")
        BUtil::print("Initializing hash table of size %lu for %lu kmers.\n", hash_table_size, n_kmers);

    int n_ranks = 1;
    if (const char* e = std::getenv("KH_RANKS")) n_ranks = std::atoi(e);
    using clock = std::chrono::high_resolution_clock;

    if (n_ranks > 1) {
        double lf = 0.5;
        if (const char* e = std::getenv("KH_LOAD_FACTOR")) lf = std::atof(e);
        const size_t per_rank = (n_kmers + n_ranks - 1) / n_ranks;
        kh_sharded::Cluster cluster(KMER_LEN, n_ranks, per_rank, n_kmers, lf);
        // read_kmers: every rank's block of lines, packed on its GPU (untimed, like the reference)
        const size_t line_len = KMER_LEN + 4;
        FILE* f = fopen(kmer_fname.c_str(), "r");
        if (f == nullptr) throw std::runtime_error("read_kmers: could not open " + kmer_fname);
        std::vector<std::vector<kmer_pair>> blocks(n_ranks);
        std::vector<char> text;
        for (int r = 0; r < n_ranks; ++r) {
            const size_t first = std::min(n_kmers, per_rank * (size_t)r), count = std::min(per_rank, n_kmers - first);
            text.resize(line_len * count);
            if (fread(text.data(), 1, text.size(), f) != text.size()) throw std::runtime_error("read_kmers: short read");
            blocks[r].resize(count);
            must(kh_pack_lines(cluster.table(r), text.data(), count, blocks[r].data()), cluster.table(r), "pack_lines");
        }
        fclose(f);
        if (run_type == "verbose") BUtil::print("Finished reading kmers.\n");

        const auto start_time = clock::now();
        std::vector<void*> dev(n_ranks, nullptr);
        std::vector<const void*> cdev(n_ranks);
        std::vector<uint64_t> counts(n_ranks);
        for (int r = 0; r < n_ranks; ++r) {
            counts[r] = blocks[r].size();
            must(kh_device_alloc_on(cluster.device(r), &dev[r], std::max<size_t>(1, counts[r] * sizeof(kmer_pair))), cluster.table(r), "alloc");
            must(kh_copy_device(cluster.table(r), dev[r], blocks[r].data(), counts[r] * sizeof(kmer_pair)), cluster.table(r), "H2D");
            cdev[r] = dev[r];
        }
        cluster.begin();
        cluster.insert(cdev, counts);                              // initialize_kmers (enqueued; the build completes inside assemble)
        const auto insert_time = clock::now();
        const char* solution_path = std::getenv("KH_SOLUTION");
        std::vector<kh_sharded::RankOutput> outs = cluster.assemble(solution_path != nullptr);   // assemble_contigs
        const auto end_time = clock::now();
        for (void* p : dev) kh_device_free(p);

        const double insert_duration = std::chrono::duration<double>(insert_time - start_time).count();
        const double assembly_duration = std::chrono::duration<double>(end_time - insert_time).count();
        const double total_duration = std::chrono::duration<double>(end_time - start_time).count();
        if (run_type != "test") {
            BUtil::print("Finished inserting in %lf sec\n", insert_duration);
            BUtil::print("Assembled in %lf total\n", total_duration);
        } else {
            // output_results (kmer_hash.cpp:60-67): every rank's file from its own writer thread
            std::vector<std::string> werr(n_ranks);
            std::vector<std::thread> writers;
            for (int r = 0; r < n_ranks; ++r)
                writers.emplace_back([&, r] {
                    const std::string out_name = test_prefix + "_" + std::to_string(r) + ".dat";
                    FILE* out = fopen(out_name.c_str(), "w");
                    if (out == nullptr) { werr[r] = "output_results: could not open " + out_name; return; }
                    if (!outs[r].text.empty() && fwrite(outs[r].text.data(), 1, outs[r].text.size(), out) != outs[r].text.size())
                        werr[r] = "output_results: short write to " + out_name;
                    fclose(out);
                });
            for (auto& w : writers) w.join();
            for (const auto& e : werr) if (!e.empty()) throw std::runtime_error(e);
            if (solution_path) write_merged_solution(solution_path, outs);
            BUtil::print("Rank %d reconstructed %d contigs with %d nodes from %d start nodes. "
                         "(%lf read, %lf insert, %lf total)\n",
                         0, (int)outs[0].n_contigs, (int)outs[0].n_nodes, 0, assembly_duration, insert_duration, total_duration);
        }
        return 0;
    }

    DistributedHashMap hashmap(hash_table_size, rank_id, 1);
    kh_table* t = hashmap.handle();

    bool stream = false;
    if (const char* e = std::getenv("KH_STREAM")) stream = std::atoi(e) != 0;
    kmer_pair* kmers = nullptr;
    if (!stream) {
        kmers = read_and_pack(t, kmer_fname, n_kmers);
        if (run_type == "verbose") BUtil::print("Finished reading kmers.\n");
    }
    must(kh_sync(t), t, "sync");

    const auto start_time = clock::now();
    if (stream) {
        kh_stream::Options opt;
        if (const char* e = std::getenv("KH_STREAM_CHUNK_LINES")) opt.chunk_lines = std::strtoull(e, nullptr, 10);
        hashmap.insert_file(kmer_fname, 0, n_kmers, opt);      // read_kmers + initialize_kmers, overlapped
        if (run_type == "verbose") BUtil::print("Finished reading kmers.\n");
    } else {
        hashmap.insert_all(kmers, n_kmers);                    // initialize_kmers
    }
    hashmap.process_requests();
    const auto insert_time = clock::now();

    const char* contig_text = nullptr;
    const uint64_t* offsets = nullptr;
    uint64_t n_contigs = 0, contig_bytes = 0, n_nodes = 0;
    must(kh_assemble(t, &contig_text, &offsets, &n_contigs, &contig_bytes, &n_nodes), t, "assemble_contigs");
    const auto end_time = clock::now();

    const double insert_duration = std::chrono::duration<double>(insert_time - start_time).count();
    const double assembly_duration = std::chrono::duration<double>(end_time - insert_time).count();
    const double total_duration = std::chrono::duration<double>(end_time - start_time).count();

    if (run_type != "test") {
        BUtil::print("Finished inserting in %lf sec\n", insert_duration);
        BUtil::print("Assembled in %lf total\n", total_duration);
    } else {
        const std::string out_name = test_prefix + "_" + std::to_string(rank_id) + ".dat";
        FILE* out = fopen(out_name.c_str(), "w");
        if (out == nullptr) throw std::runtime_error("output_results: could not open " + out_name);
        if (contig_bytes && fwrite(contig_text, 1, contig_bytes, out) != contig_bytes)
            throw std::runtime_error("output_results: short write to " + out_name);
        fclose(out);
        if (const char* solution_path = std::getenv("KH_SOLUTION")) {       // the sorted contig set (check_it.sh:47-48), sorted on the GPU
            std::vector<uint64_t> order(n_contigs);
            must(kh_sorted_order(t, order.data(), nullptr), t, "sorted order");
            FILE* sol = fopen(solution_path, "w");
            if (sol == nullptr) throw std::runtime_error(std::string("could not open ") + solution_path);
            for (uint64_t c : order) fwrite(contig_text + offsets[c], 1, offsets[c + 1] - offsets[c], sol);
            fclose(sol);
        }
        // same line as kmer_hash.cpp:71-78, including its quirks: "start nodes" is a literal 0
        // and the slot labelled "read" carries the assembly time
        BUtil::print("Rank %d reconstructed %d contigs with %d nodes from %d start nodes. "
                     "(%lf read, %lf insert, %lf total)\n",
                     rank_id, (int)n_contigs, (int)n_nodes, 0, assembly_duration, insert_duration, total_duration);
    }
    if (kmers) kh_host_free(kmers);
    return 0;
}
