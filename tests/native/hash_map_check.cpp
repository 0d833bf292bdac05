// Drop-in header API on a GPU (tests/test_cli_gpu.py compiles and runs this): the member functions the reference's
// callers use -- DistributedHashMap::insert_all(vector), find(std::string, kmer_pair&) (hash_map.hpp:55, 83) -- and the
// upstream-starter HashMap::insert / find(pkmer_t, ...) (README.md:95-99), against records built with the reference's
// own type constructors.  Prints "OK <n>" or a diagnostic and a non-zero exit code.
#include <cstdio>
#include <string>
#include <vector>

#include "hash_map.hpp"
#include "kmer_t.hpp"
#include "read_kmers.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string fname = argv[1];
    if (kmer_size(fname) != KMER_LEN) { printf("K mismatch\n"); return 2; }
    std::vector<kmer_pair> kmers = read_kmers(fname, 1, 0);
    const size_t n = kmers.size();
    {
        DistributedHashMap map(2 * n, 0, 1);
        map.insert_all(kmers);
        map.process_requests();
        for (size_t i = 0; i < n; i += 3) {
            kmer_pair got;
            if (!map.find(kmers[i].kmer_str(), got)) { printf("find(std::string) missed record %zu\n", i); return 1; }
            if (!(got == kmers[i]) || got.forwardExt() != kmers[i].forwardExt() || got.backwardExt() != kmers[i].backwardExt()) {
                printf("find(std::string) returned a different record at %zu\n", i); return 1;
            }
        }
        kmer_pair dummy;
        if (map.find(std::string(KMER_LEN, 'A') , dummy) && !(dummy.kmer_str() == std::string(KMER_LEN, 'A'))) { printf("bogus hit\n"); return 1; }
        if (map.find("ACGT", dummy)) { printf("a key of the wrong length was found\n"); return 1; }
        // the reference's own walk, one find per step (kmer_hash.cpp:38-55), must reproduce the batch traversal
        std::vector<std::string> batch = map.assemble();
        size_t c = 0;
        for (const kmer_pair& start : kmers) {
            if (start.backwardExt() != 'F') continue;
            std::list<kmer_pair> contig{start};
            while (contig.back().forwardExt() != 'F') {
                kmer_pair next;
                if (!map.find(contig.back().next_kmer().get(), next)) { printf("walk: k-mer not found\n"); return 1; }
                contig.push_back(next);
            }
            if (c >= batch.size() || extract_contig(contig) != batch[c]) { printf("walk: contig %zu differs from assemble()\n", c); return 1; }
            ++c;
            if (c >= 25) break;
        }
    }
    {
        HashMap hm(2 * 500);                       // upstream-starter interface, one record per call
        const size_t m = n < 500 ? n : 500;
        for (size_t i = 0; i < m; ++i) if (!hm.insert(kmers[i])) { printf("HashMap::insert failed\n"); return 1; }
        for (size_t i = 0; i < m; ++i) {
            kmer_pair got;
            if (!hm.find(kmers[i].kmer, got) || !(got == kmers[i])) { printf("HashMap::find missed record %zu\n", i); return 1; }
        }
        kmer_pair got;
        if (n > m && hm.find(kmers[m].kmer, got)) { printf("HashMap::find found a record that was never inserted\n"); return 1; }
    }
    printf("OK %zu\n", n);
    return 0;
}
