/* kmer_oracle.c -- CPU restatement of the reference's insert + traverse path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or
 * call this file; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA path.  The product fails loudly without
 * its CUDA library -- there is no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (i)  the reference's only in-repo golden vector (README.md:27, k=3), and
 *   (ii) outputs of the UNMODIFIED reference (oracle/_ref/kmer_hash_ref_<K>,
 *        built by oracle/Makefile from /root/reference/kmer_hash.cpp) committed
 *        under tests/golden/ by tests/golden/make_golden.py, and live against
 *        oracle/_ref when that binary is present.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  K is a run-time argument here (the reference fixes it at
 * compile time with -DKMER_LEN, packing.hpp:5-9).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_OK 0
#define KO_ERR_ARG 1
#define KO_ERR_NOT_FOUND 2   /* kmer_hash.cpp:47-49 "k-mer not found" */
#define KO_ERR_ALLOC 3
#define KO_ERR_CYCLE 4       /* chain longer than the table: the reference would spin forever (kmer_hash.cpp:44) */
#define KO_ERR_BAD_BASE 5    /* packing.hpp:52-70 leaves `code` uninitialised here; we refuse instead */

static int ko_packed_len(int k) { return (k + 3) / 4; }          /* packing.hpp:9 */
int ko_pair_bytes(int k) { return ko_packed_len(k) + 2; }        /* kmer_t.hpp:6-8 */

static int base_code(char b) {                                   /* packing.hpp:52-67 */
    switch (b) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
    }
}

/* packing.hpp:77-92 (packKmer) with packing.hpp:50-75 (packFourMer): four bases
 * per byte, first base in bits 7..6, the last byte's missing bases are 'A' (0).
 * Deliberate difference: when k % 4 == 0 the reference writes one extra byte
 * past the array; this writes exactly (k+3)/4 bytes. */
int ko_pack_kmer(const char* kmer, int k, uint8_t* out) {
    const int pl = ko_packed_len(k);
    for (int i = 0; i < pl; ++i) {
        unsigned v = 0;
        for (int j = 0; j < 4; ++j) {
            const int idx = 4 * i + j;
            int code = 0;
            if (idx < k) {
                code = base_code(kmer[idx]);
                if (code < 0) return KO_ERR_BAD_BASE;
            }
            v = (v << 2) | (unsigned)code;
        }
        out[i] = (uint8_t)v;
    }
    return KO_OK;
}

/* packing.hpp:94-107 (unpackKmer) + :16-48 (the byte -> 4-mer table), without the
 * table and without the overrun past kmer[k-1]. */
void ko_unpack_kmer(const uint8_t* packed, int k, char* out) {
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    for (int i = 0; i < k; ++i) out[i] = letters[(packed[i >> 2] >> (6 - 2 * (i & 3))) & 3];
}

/* read_kmers.hpp:72-76 + kmer_t.hpp:67-76: a line is K bases, one byte that is
 * skipped, backward ext, forward ext, one more byte ('\n').  Output records use
 * the reference's kmer_pair layout: packed bytes, then fb_ext[0]=backward,
 * fb_ext[1]=forward (kmer_t.hpp:43-45). */
int ko_parse_lines(const char* text, uint64_t n_lines, int k, uint8_t* pairs) {
    const int pb = ko_pair_bytes(k), pl = ko_packed_len(k);
    for (uint64_t i = 0; i < n_lines; ++i) {
        const char* line = text + i * (uint64_t)(k + 4);
        uint8_t* rec = pairs + i * (uint64_t)pb;
        const int rc = ko_pack_kmer(line, k, rec);
        if (rc) return rc;
        rec[pl] = (uint8_t)line[k + 1];
        rec[pl + 1] = (uint8_t)line[k + 2];
    }
    return KO_OK;
}

/* kmer_t.hpp:51-53 next_kmer(): unpack, drop the first base, append the forward
 * extension, repack. */
int ko_next_kmer(const uint8_t* pair, int k, uint8_t* out_packed) {
    char buf[80];
    if (k > 64) return KO_ERR_ARG;
    ko_unpack_kmer(pair, k, buf);
    memmove(buf, buf + 1, (size_t)(k - 1));
    buf[k - 1] = (char)pair[ko_packed_len(k) + 1];
    return ko_pack_kmer(buf, k, out_packed);
}

/* ---- table: semantics of unordered_map<string,kmer_pair> with map[key]=value
 * (hash_map.hpp:33-35, 67-71: last writer wins) and find (hash_map.hpp:85-92).
 * Placement is unobservable, so a flat linear-probing table
 * keyed by the packed bytes stands in for std::hash<std::string>. ---- */
typedef struct {
    int k, pl, pb;
    uint64_t cap, mask, count;
    uint8_t* rec;     /* cap * pb bytes */
    uint8_t* used;    /* cap bytes */
} ko_table;

static uint64_t bytes_hash(const uint8_t* p, int n) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001B3ull; }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}

int ko_table_create(int k, uint64_t n_expected, ko_table** out) {
    if (k < 1 || k > 64 || !out) return KO_ERR_ARG;
    ko_table* t = (ko_table*)calloc(1, sizeof *t);
    if (!t) return KO_ERR_ALLOC;
    t->k = k; t->pl = ko_packed_len(k); t->pb = t->pl + 2;
    uint64_t cap = 16;
    while (cap < 2 * n_expected + 2) cap <<= 1;     /* kmer_hash.cpp:109: load factor 0.5 */
    t->cap = cap; t->mask = cap - 1;
    t->rec = (uint8_t*)malloc(cap * (uint64_t)t->pb);
    t->used = (uint8_t*)calloc(cap, 1);
    if (!t->rec || !t->used) { free(t->rec); free(t->used); free(t); return KO_ERR_ALLOC; }
    *out = t;
    return KO_OK;
}

void ko_table_destroy(ko_table* t) {
    if (!t) return;
    free(t->rec); free(t->used); free(t);
}

uint64_t ko_table_count(const ko_table* t) { return t->count; }

/* hash_map.hpp:55-72 insert_all at one rank: every item is stored under its k-mer. */
int ko_insert_pairs(ko_table* t, const uint8_t* pairs, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* r = pairs + i * (uint64_t)t->pb;
        uint64_t s = bytes_hash(r, t->pl) & t->mask;
        for (;;) {
            if (!t->used[s]) {
                if (t->count + 1 >= t->cap) return KO_ERR_ALLOC;
                t->used[s] = 1; t->count++;
                memcpy(t->rec + s * (uint64_t)t->pb, r, (size_t)t->pb);
                break;
            }
            if (memcmp(t->rec + s * (uint64_t)t->pb, r, (size_t)t->pl) == 0) {   /* pkmer_t.hpp:41-43 */
                memcpy(t->rec + s * (uint64_t)t->pb, r, (size_t)t->pb);          /* overwrite */
                break;
            }
            s = (s + 1) & t->mask;
        }
    }
    return KO_OK;
}

/* hash_map.hpp:83-92 find(): 1 if present (record copied to out_pair), else 0. */
int ko_find(const ko_table* t, const uint8_t* packed, uint8_t* out_pair) {
    uint64_t s = bytes_hash(packed, t->pl) & t->mask;
    while (t->used[s]) {
        const uint8_t* r = t->rec + s * (uint64_t)t->pb;
        if (memcmp(r, packed, (size_t)t->pl) == 0) {
            if (out_pair) memcpy(out_pair, r, (size_t)t->pb);
            return 1;
        }
        s = (s + 1) & t->mask;
    }
    return 0;
}

/* kmer_hash.cpp:27-31 (start scan, order-preserving) + :38-55 (assemble_contigs)
 * + read_kmers.hpp:81-92 (extract_contig) + kmer_hash.cpp:64-67 (one per line).
 *
 * `pairs`/`n` is THIS rank's block of records in file order (read_kmers.hpp:55-58);
 * the table must already hold every rank's records.  Pass out == NULL to size the
 * output.  Contigs are written in start-node order, each followed by '\n'. */
int ko_assemble(const ko_table* t, const uint8_t* pairs, uint64_t n, char* out, uint64_t out_cap,
                uint64_t* out_len, uint64_t* n_contigs, uint64_t* n_nodes) {
    const int k = t->k, pl = t->pl, pb = t->pb;
    uint64_t w = 0, contigs = 0, nodes = 0;
    uint8_t cur[32], nxt[32];
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* r = pairs + i * (uint64_t)pb;
        if (r[pl] != 'F') continue;                         /* backwardExt() == 'F' */
        memcpy(cur, r, (size_t)pb);
        /* extract_contig: the front k-mer's K characters ... */
        if (out) {
            if (w + (uint64_t)k > out_cap) return KO_ERR_ARG;
            ko_unpack_kmer(cur, k, out + w);
        }
        w += (uint64_t)k;
        uint64_t steps = 1;
        while (cur[pl + 1] != 'F') {                        /* forwardExt() != 'F' */
            /* ... then every forward extension that is not 'F' */
            if (out) {
                if (w + 1 > out_cap) return KO_ERR_ARG;
                out[w] = (char)cur[pl + 1];
            }
            ++w;
            const int rc = ko_next_kmer(cur, k, nxt);
            if (rc) return rc;
            if (!ko_find(t, nxt, cur)) return KO_ERR_NOT_FOUND;
            if (++steps > t->count) return KO_ERR_CYCLE;
        }
        if (out) {
            if (w + 1 > out_cap) return KO_ERR_ARG;
            out[w] = '\n';
        }
        ++w;
        ++contigs;
        nodes += steps;
    }
    if (out_len) *out_len = w;
    if (n_contigs) *n_contigs = contigs;
    if (n_nodes) *n_nodes = nodes;
    return KO_OK;
}
