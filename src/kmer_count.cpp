// kmer_count -- reads -> the reference's k-mer file, or straight to contigs, on the B200 library.
//
//   kmer_count K reads_file kmer_file_out   [min_count [min_ext]]
//   kmer_count K reads_file --contigs PREFIX [min_count [min_ext]]
//
// The first form writes what the reference's read_kmers parses (read_kmers.hpp:64-76: K bases, a blank, backward and
// forward extension, '\n' per unique k-mer) -- the output of "the first preprocessing stage" the assignment assumes done
// (README.md:19-21) -- so `kmer_hash_<K> kmer_file_out test` (ours or the reference's) can follow.  The second form keeps
// the records on the GPU (kh_count_extract_device -> kh_insert_pairs_device -> kh_assemble) and writes PREFIX_0.dat, one
// contig per line, the file scripts/check_it.sh:47-48 sorts and diffs: no k-mer text in between.
//
// reads_file: one read per line, FASTA or FASTQ (detected from the first byte; header / quality lines are blanked on
// the host).  Bases are upper-case ACGT; anything else (N, ...) splits a read.  Defaults: min_count = min_ext = 2.
// Environment: KH_DEVICE, KH_LOAD_FACTOR (0.5), KH_COUNT_DISTINCT = distinct k-mers to size the table for, erroneous
// ones included (default: half the input bytes).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "kh/kmer_counter.hpp"

namespace {

void must(int status, kh_table* t, const char* what) {
    if (status == KH_OK) return;
    const char* detail = t ? kh_last_error(t) : "";
    throw std::runtime_error(std::string(detail && *detail ? detail : kh_status_string(status)) + " [" + what + "]");
}

int usage() {
    fprintf(stderr, "Usage: kmer_count K reads_file (kmer_file_out | --contigs PREFIX) [min_count [min_ext]]\n");
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) return usage();
    const int k = atoi(argv[1]);
    const std::string reads_file = argv[2];
    int a = 3;
    const bool contigs = std::string(argv[a]) == "--contigs";
    if (contigs && ++a >= argc) return usage();
    const std::string out_name = argv[a++];
    const uint32_t min_count = a < argc ? (uint32_t)atoi(argv[a++]) : 2u;
    const uint32_t min_ext = a < argc ? (uint32_t)atoi(argv[a++]) : 2u;
    if (k < 2 || k > 61) { fprintf(stderr, "kmer_count: K must be 2..61\n"); return 1; }
    const int device = getenv("KH_DEVICE") ? atoi(getenv("KH_DEVICE")) : 0;
    const double lf = getenv("KH_LOAD_FACTOR") ? atof(getenv("KH_LOAD_FACTOR")) : 0.5;

    FILE* f = fopen(reads_file.c_str(), "rb");
    if (f == nullptr) throw std::runtime_error("kmer_count: could not open " + reads_file);
    fseek(f, 0, SEEK_END);
    const uint64_t file_bytes = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    const uint64_t distinct = getenv("KH_COUNT_DISTINCT") ? strtoull(getenv("KH_COUNT_DISTINCT"), nullptr, 10)
                                                          : (file_bytes / 2 > 1024 ? file_bytes / 2 : 1024);
    const auto t0 = std::chrono::high_resolution_clock::now();
    kh::KmerCounter counter(k, distinct, lf, device);

    // the file goes through in pieces that end on a line boundary
    const size_t piece = 128u << 20;
    std::vector<char> buf(piece + (1u << 20));
    size_t carry = 0;
    char format = 0;
    unsigned state = 0;
    for (;;) {
        if (carry == buf.size()) buf.resize(buf.size() * 2);                  // a single line longer than the buffer
        const size_t got = fread(buf.data() + carry, 1, buf.size() - carry, f);
        const size_t have = carry + got;
        if (have == 0) break;
        if (format == 0) format = buf[0] == '>' ? 'a' : (buf[0] == '@' ? 'q' : 'p');
        size_t cut = have;
        if (got != 0) {                                                       // not at the end of the file: stop where a new read starts
            cut = kh::sequence_cut(buf.data(), have, format);
            if (cut == 0) { carry = have; continue; }
        }
        size_t len = cut;
        kh::sequence_lines(buf.data(), len, format, state);
        counter.count_reads(buf.data(), len);
        carry = have - cut;
        memmove(buf.data(), buf.data() + cut, carry);
        if (got == 0) break;
    }
    fclose(f);
    const auto t1 = std::chrono::high_resolution_clock::now();
    kh_count_stats st = counter.stats();
    printf("Counted %llu k-mer occurrences, %llu distinct %d-mers, from %llu bytes of reads in %lf s\n",
           (unsigned long long)st.n_occurrences, (unsigned long long)st.n_distinct, k, (unsigned long long)st.n_bytes,
           std::chrono::duration<double>(t1 - t0).count());

    if (!contigs) {
        const std::string lines = counter.extract_lines(min_count, min_ext);
        FILE* o = fopen(out_name.c_str(), "wb");
        if (o == nullptr) throw std::runtime_error("kmer_count: could not open " + out_name);
        if (fwrite(lines.data(), 1, lines.size(), o) != lines.size()) throw std::runtime_error("kmer_count: short write to " + out_name);
        fclose(o);
        printf("Wrote %llu k-mers (count >= %u, extensions seen >= %u times) to %s\n",
               (unsigned long long)(lines.size() / (size_t)(k + 4)), min_count, min_ext, out_name.c_str());
        return 0;
    }
    uint64_t n = 0;
    const void* recs = counter.extract_device(min_count, min_ext, n);
    kh_table* t = nullptr;
    must(kh_create(k, n ? n : 1, lf, device, &t), nullptr, "kh_create");
    const auto t2 = std::chrono::high_resolution_clock::now();
    must(kh_insert_pairs_device(t, recs, n), t, "insert");
    const char* text = nullptr;
    const uint64_t* offs = nullptr;
    uint64_t nc = 0, bytes = 0, nodes = 0;
    must(kh_assemble(t, &text, &offs, &nc, &bytes, &nodes), t, "assemble");
    const auto t3 = std::chrono::high_resolution_clock::now();
    const std::string dat = out_name + "_0.dat";
    FILE* o = fopen(dat.c_str(), "wb");
    if (o == nullptr) throw std::runtime_error("kmer_count: could not open " + dat);
    if (bytes && fwrite(text, 1, bytes, o) != bytes) throw std::runtime_error("kmer_count: short write to " + dat);
    fclose(o);
    printf("Assembled %llu k-mers into %llu contigs with %llu nodes in %lf s; wrote %s\n", (unsigned long long)n,
           (unsigned long long)nc, (unsigned long long)nodes, std::chrono::duration<double>(t3 - t2).count(), dat.c_str());
    kh_destroy(t);
    return 0;
}
