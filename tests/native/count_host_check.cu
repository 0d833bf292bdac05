// Host-only run of the k-mer analysis functions in cs267_hw3_b200/csrc/count_core.cuh (built with nvcc, runs without a
// GPU): the loops of kc_count_kernel / kc_extract_kernel / kc_lookup_kernel (count.cu) executed serially, block by block
// and thread by thread, with the host chunking of kh_count_reads (kc_chunk_range) at a small chunk size so that chunk and
// tile boundaries fall inside the reads.  tests/test_count.py compares the output with the oracle.
//
//   count_host_check K chunk n_slots min_count min_ext misalign reads_file [grow_after_chunk]
// (grow_after_chunk: after that chunk the table is doubled and rehashed with kc_move_slot, as kc_grow does on the GPU)
// prints "n_occurrences n_distinct n_reported full" and then one line per reported k-mer:
//   record(hex)  occurrences backA backC backG backT fwdA fwdC fwdG fwdT  line       (the counters through kc_find;
//   the record as a k-mer file line through kc_record_to_line, its blank shown as '_')
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../cs267_hw3_b200/csrc/count_core.cuh"

using namespace kh;

template <int W>
static void count_piece(const unsigned char* buf, u64 n, u64 p_begin, u64 p_end, u64* table, u64 n_slots, int k, KcCounters& ctr) {
    if (p_end <= p_begin) return;
    const u64 span = p_end - (p_begin & ~15ull), blocks = (span + kKcTile - 1) / kKcTile;
    std::vector<u32> s_code(kKcWords), s_inv(kKcWords);
    for (u64 block = 0; block < blocks; ++block) {
        const u64 t0 = (p_begin & ~15ull) + block * kKcTile;
        for (u32 i = 0; i < kKcWords; ++i) kc_pack_word(buf, n, (long long)t0 - 16 + 16ll * (long long)i, s_code[i], s_inv[i]);
        // pass 1: the positions that hold a k-mer; pass 2: count them (the kernel's two passes, serially)
        std::vector<unsigned short> list;
        for (int r = 0; r < kKcPer; ++r) {
            for (u32 thread = 0; thread < (u32)kKcThreads; ++thread) {
                const u32 local = thread + (u32)r * kKcThreads;
                const u64 p = t0 + local;
                if (p >= p_begin && p < p_end && kc_kmer_valid(s_inv.data(), local, k)) list.push_back((unsigned short)local);
            }
        }
        for (unsigned short local : list) {
            const KcOcc o = kc_position(s_code.data(), s_inv.data(), local, k);
            bool f;
            if (!kc_upsert<W>(table, n_slots, o, kc_home(o.key_hi, o.key_lo, n_slots), f)) { ctr.errors |= kKcErrFull; continue; }
            ++ctr.n_occurrences;
            ctr.n_distinct += f ? 1 : 0;
        }
    }
}

template <int W>
static int run(int k, u64 chunk, u64 n_slots, u32 min_count, u32 min_ext, const unsigned char* reads, u64 n_bytes, bool one_piece, long grow_after) {
    std::vector<u64> table(n_slots * KcSlot<W>::kWords, 0ull);
    KcCounters ctr = {};
    const u64 nchunks = one_piece ? 0 : (n_bytes + chunk - 1) / chunk;
    if (one_piece) count_piece<W>(reads, n_bytes, 0, n_bytes, table.data(), n_slots, k, ctr);      // kh_count_reads_device: any pointer
    for (u64 ci = 0; ci < nchunks; ++ci) {
        u64 a, b, off, len;
        kc_chunk_range(ci, chunk, n_bytes, a, b, off, len);
        // the piece sits 16-byte aligned in its own buffer, like the device staging buffer
        unsigned char* piece = static_cast<unsigned char*>(aligned_alloc(16, ((b - a) + 15) & ~15ull));
        memcpy(piece, reads + a, b - a);
        count_piece<W>(piece, b - a, off - a, off - a + len, table.data(), n_slots, k, ctr);
        free(piece);
        if (grow_after >= 0 && ci == (u64)grow_after) {       // kc_grow: every slot moves into a table twice the size
            std::vector<u64> bigger(2 * n_slots * KcSlot<W>::kWords, 0ull);
            for (u64 i = 0; i < n_slots; ++i)
                if (!kc_move_slot<W>(table.data(), i, bigger.data(), 2 * n_slots)) ctr.errors |= kKcErrFull;
            table.swap(bigger);
            n_slots *= 2;
        }
    }
    const int pl = (k + 3) / 4, pb = pl + 2;
    std::vector<unsigned char> out;
    for (u64 i = 0; i < n_slots; ++i) {
        if (!kc_slot_reported<W>(table.data(), i, min_count)) continue;
        unsigned char rec[18];
        kc_slot_record<W>(table.data(), i, k, min_ext, rec);
        out.insert(out.end(), rec, rec + pb);
        ++ctr.n_reported;
    }
    printf("%llu %llu %llu %u\n", ctr.n_occurrences, ctr.n_distinct, ctr.n_reported, ctr.errors);
    for (u64 r = 0; r < ctr.n_reported; ++r) {
        const unsigned char* rec = out.data() + r * pb;
        for (int j = 0; j < pb; ++j) printf("%02x", rec[j]);
        u64 kh_, kl_;
        kc_packed_to_key(rec, k, kh_, kl_);
        const u64 c = kc_find<W>(table.data(), n_slots, kh_, kl_);
        printf(" %u", kc_total(c));
        for (u32 b = 0; b < 4; ++b) printf(" %u", kc_back_count(c, b));
        for (u32 b = 0; b < 4; ++b) printf(" %u", kc_fwd_count(c, b));
        unsigned char line[68];
        kc_record_to_line(rec, k, line);
        line[k + 3] = 0;
        line[k] = '_';
        printf(" %s\n", reinterpret_cast<const char*>(line));
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 8 && argc != 9) { fprintf(stderr, "usage: count_host_check K chunk n_slots min_count min_ext misalign reads_file [grow_after_chunk]\n"); return 2; }
    const long grow_after = argc == 9 ? atol(argv[8]) : -1;
    const int k = atoi(argv[1]);
    const u64 chunk = strtoull(argv[2], nullptr, 10), n_slots = strtoull(argv[3], nullptr, 10);
    const u32 min_count = (u32)atoi(argv[4]), min_ext = (u32)atoi(argv[5]);
    const int misalign = atoi(argv[6]);
    FILE* f = fopen(argv[7], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> store((size_t)n + 64);
    // misalign > 0: the reads start at an address that is NOT 16-byte aligned and go through ONE piece (the device
    // entry point kh_count_reads_device takes any pointer): the byte-wise loader of kc_pack_word
    unsigned char* base = store.data();
    while ((reinterpret_cast<uintptr_t>(base) & 15u) != (unsigned)(misalign & 15)) ++base;
    if (fread(base, 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    if (chunk % kKcTile) return 2;
    return kc_slot_words(k) == 1 ? run<1>(k, chunk, n_slots, min_count, min_ext, base, (u64)n, misalign != 0, grow_after)
                                 : run<2>(k, chunk, n_slots, min_count, min_ext, base, (u64)n, misalign != 0, grow_after);
}
