/* kmer_count_oracle.c -- CPU statement of the k-mer analysis stage that PRECEDES the reference's
 * path: reads -> unique k-mers with their backward / forward extensions (the "first preprocessing
 * stage" of README.md:19-21, whose output is the reference's input file, read_kmers.hpp:54-79).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or call this file; it
 * checks cs267_hw3_b200/csrc/count.cu from tests/ (see oracle/kmer_oracle.c for the rules).
 *
 * Parity status: the reference holds NO implementation of this stage ("the
 * Meraculous/HipMer stage this homework assumes done"), so there is no function of the reference's
 * to pin it to: PARITY UNPINNED at the unit level.  It is PINNED END TO END on files the UNMODIFIED
 * reference wrote: tests/test_count.py cuts reads from the contigs in tests/golden/<case>.dat (the
 * output of oracle/_ref/kmer_hash_ref_<K>, tests/golden/make_golden.py) and this stage must give
 * back, line for line, the k-mer file <case>.txt that the reference had consumed to produce them
 * (README.md:27's k=3 example included); and reads cut from a generated contig set, run through
 * this stage and then through insert + traverse, must give the generator's solution.
 *
 * Definition (what both this file and the CUDA path compute), for k-mer length K:
 *   - a base is one of the bytes 'A' 'C' 'G' 'T'; every other byte ('\n', 'N', lower case, ...)
 *     separates reads;
 *   - every position p whose K bytes p .. p+K-1 are bases is one OCCURRENCE of that k-mer; its
 *     backward observation is byte p-1 if that is a base (else none), its forward observation is
 *     byte p+K if that is a base (else none) -- the two extensions of the reference's text format
 *     (read_kmers.hpp:64-76: `kmer, ' ', backward, forward`), with "none" where a read begins / ends;
 *   - per distinct k-mer: count = min(occurrences, 255); per side and base: min(observations, 127);
 *   - a k-mer is REPORTED when count >= min_count; its backward extension is the base b if b is the
 *     only base on that side observed >= min_ext times, and 'F' otherwise (no base qualifies: the
 *     contig begins here, README.md:37; several qualify: a fork, where a linear contig has to end too);
 *     the forward extension likewise;
 *   - no reverse complements (the reference ignores them as well: kmer_t.hpp:51-57 only ever appends).
 * The record written per k-mer is the reference's kmer_pair (kmer_t.hpp:6-8): packKmer bytes + 2 letters.
 *
 * Method here: materialise every occurrence, sort by key, scan the runs -- deliberately not a hash
 * table, so that the check shares no algorithm with the GPU path.  Output is sorted by k-mer.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KCO_OK 0
#define KCO_ERR_ARG 1
#define KCO_ERR_ALLOC 3
#define KCO_ERR_CAP 6

typedef struct {
    uint64_t hi, lo;      /* the k-mer as a right-justified base-4 number, first base most significant */
    uint8_t back, fwd;    /* 0..3 = A C G T, 4 = none */
} kco_occ;

static int kco_code(unsigned char c) {
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
    }
}

static int kco_cmp(const void* a, const void* b) {
    const kco_occ* x = (const kco_occ*)a;
    const kco_occ* y = (const kco_occ*)b;
    if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    return 0;
}

/* key -> packKmer bytes (packing.hpp:77-92): first base in bits 7..6 of byte 0, tail padded with A */
static void kco_pack(uint64_t hi, uint64_t lo, int k, uint8_t* out) {
    const int pl = (k + 3) / 4;
    for (int i = pl - 1; i >= 0; --i) {
        /* byte i holds bases 4i .. 4i+3; base j (0-based from the left) is digit (k-1-j) of the number */
        unsigned v = 0;
        for (int j = 0; j < 4; ++j) {
            const int idx = 4 * i + j;
            unsigned code = 0;
            if (idx < k) {
                const int digit = k - 1 - idx;           /* 2 bits at position 2*digit */
                code = digit >= 32 ? (unsigned)((hi >> (2 * (digit - 32))) & 3u) : (unsigned)((lo >> (2 * digit)) & 3u);
            }
            v = (v << 2) | code;
        }
        out[i] = (uint8_t)v;
    }
}

/* Number of occurrences (positions whose K bytes are all bases). */
uint64_t kco_count_occurrences(const char* reads, uint64_t n_bytes, int k) {
    uint64_t run = 0, occ = 0;
    for (uint64_t i = 0; i < n_bytes; ++i) {
        run = kco_code((unsigned char)reads[i]) < 4 ? run + 1 : 0;
        if (run >= (uint64_t)k) ++occ;
    }
    return occ;
}

/* reads -> records.  pairs_out: room for `cap` records of (k+3)/4+2 bytes (may be NULL to size);
 * counts_out (optional): 9 x uint32 per record: count, back A C G T, forward A C G T (saturated).
 * *n_out = number of reported k-mers (also when cap is too small: then KCO_ERR_CAP). */
int kco_analyse(const char* reads, uint64_t n_bytes, int k, uint32_t min_count, uint32_t min_ext,
                uint8_t* pairs_out, uint64_t cap, uint64_t* n_out, uint32_t* counts_out) {
    if (!reads || k < 2 || k > 61 || min_count < 1 || min_count > 255 || min_ext < 1 || min_ext > 127 || !n_out)
        return KCO_ERR_ARG;
    const uint64_t n_occ = kco_count_occurrences(reads, n_bytes, k);
    kco_occ* occ = (kco_occ*)malloc((size_t)(n_occ ? n_occ : 1) * sizeof(kco_occ));
    if (!occ) return KCO_ERR_ALLOC;
    const int pb = (k + 3) / 4 + 2;
    {
        /* rolling key over the current run of bases */
        uint64_t hi = 0, lo = 0, run = 0, w = 0;
        const uint64_t hi_mask = k > 32 ? ((k == 64) ? ~0ull : ((1ull << (2 * (k - 32))) - 1ull)) : 0ull;
        const uint64_t lo_mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
        for (uint64_t i = 0; i < n_bytes; ++i) {
            const int c = kco_code((unsigned char)reads[i]);
            if (c == 4) { run = 0; hi = lo = 0; continue; }
            hi = ((hi << 2) | (lo >> 62)) & hi_mask;
            lo = ((lo << 2) | (uint64_t)c) & lo_mask;
            if (++run < (uint64_t)k) continue;
            const uint64_t p = i + 1 - (uint64_t)k;                       /* the occurrence starts here */
            occ[w].hi = hi; occ[w].lo = lo;
            occ[w].back = (uint8_t)(p > 0 ? kco_code((unsigned char)reads[p - 1]) : 4);
            occ[w].fwd = (uint8_t)(i + 1 < n_bytes ? kco_code((unsigned char)reads[i + 1]) : 4);
            ++w;
        }
    }
    qsort(occ, (size_t)n_occ, sizeof(kco_occ), kco_cmp);
    uint64_t n_rep = 0;
    int rc = KCO_OK;
    for (uint64_t a = 0; a < n_occ;) {
        uint64_t b = a;
        uint32_t cnt[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
        while (b < n_occ && occ[b].hi == occ[a].hi && occ[b].lo == occ[a].lo) {
            ++cnt[0][occ[b].back];
            ++cnt[1][occ[b].fwd];
            ++b;
        }
        uint64_t total = b - a;
        if (total > 255) total = 255;
        char ext[2];
        for (int side = 0; side < 2; ++side) {
            int pick = -1, nq = 0;
            for (int base = 0; base < 4; ++base) {
                if (cnt[side][base] > 127) cnt[side][base] = 127;
                if (cnt[side][base] >= min_ext) { pick = base; ++nq; }
            }
            ext[side] = nq == 1 ? "ACGT"[pick] : 'F';
        }
        if (total >= min_count) {
            if (pairs_out && n_rep < cap) {
                uint8_t* rec = pairs_out + n_rep * (uint64_t)pb;
                kco_pack(occ[a].hi, occ[a].lo, k, rec);
                rec[pb - 2] = (uint8_t)ext[0];                          /* kmer_t.hpp:43-45: [0] backward, [1] forward */
                rec[pb - 1] = (uint8_t)ext[1];
                if (counts_out) {
                    uint32_t* c9 = counts_out + 9 * n_rep;
                    c9[0] = (uint32_t)total;
                    for (int base = 0; base < 4; ++base) { c9[1 + base] = cnt[0][base]; c9[5 + base] = cnt[1][base]; }
                }
            } else if (pairs_out) {
                rc = KCO_ERR_CAP;
            }
            ++n_rep;
        }
        a = b;
    }
    *n_out = n_rep;
    free(occ);
    return rc;
}
