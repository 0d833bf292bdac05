// Synthetic unique-k-mer dataset generator (SURVEY.md section 8d "generator contract").
//
// Produces, from a seeded random genome, a k-mer file in the reference's text
// format (read_kmers.hpp:64-76: K bases, one separator byte, backward ext,
// forward ext, '\n'), the same records in the reference's kmer_pair byte layout
// (kmer_t.hpp:6-8 / packing.hpp:50-92), and the contigs a correct assembler must
// emit -- per rank in start-line order (kmer_hash.cpp:27-31,41,64-67) and as the
// bytewise-sorted solution file scripts/check_it.sh:47-55 diffs against.
//
// Model
//   * C contigs with node counts from a broken-stick split of N (sum is exactly
//     N, each >= 1, approximately geometric with mean N/C); optionally contig 0
//     is forced to `long_nodes` nodes (the S3-long shape).
//   * bases i.i.d. uniform over ACGT from xoshiro256**; for K <= 31 a global
//     set of packed keys enforces uniqueness: a window whose key already
//     exists gets its last base re-drawn (only not-yet-emitted windows contain
//     that base).  For K > 31 the birthday probability at these sizes is
//     < 1e-12 and the check is skipped.
//   * line order = a seeded Feistel permutation of node ids (cycle-walked to
//     [0,N)), so any line can be produced independently and in parallel.
//
// This file is a tool: nothing on the product path links it.  It exports a
// small C ABI (kg_*) so tests/bench can build datasets in memory via ctypes,
// and has a CLI (gen_kmers) when compiled with -DKG_MAIN.
#include <algorithm>
#include <atomic>
#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 33)) * 0xFF51AFD7ED558CCDull;
    z = (z ^ (z >> 33)) * 0xC4CEB9FE1A85EC53ull;
    return z ^ (z >> 33);
}

struct Xoshiro {
    uint64_t s[4];
    explicit Xoshiro(uint64_t seed) {
        for (auto& w : s) w = splitmix64(seed);
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    inline uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    // unbiased enough for dataset generation: 128-bit multiply range reduction
    inline uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
};

// Open-addressing set of 64-bit keys (key+1 stored so 0 means empty).
struct KeySet {
    std::vector<uint64_t> slot;
    uint64_t mask = 0;
    void init(uint64_t n) {
        uint64_t cap = 64;
        while (cap < 2 * n + 16) cap <<= 1;
        slot.assign(cap, 0);
        mask = cap - 1;
    }
    inline uint64_t home(uint64_t key) const { return mix64(key) & mask; }
    inline void prefetch(uint64_t key) const { __builtin_prefetch(&slot[home(key)], 1, 0); }
    // true if newly inserted, false if already present
    inline bool insert(uint64_t key) {
        const uint64_t v = key + 1;
        for (uint64_t i = home(key);; i = (i + 1) & mask) {
            if (slot[i] == 0) { slot[i] = v; return true; }
            if (slot[i] == v) return false;
        }
    }
};

const char kBase[4] = {'A', 'C', 'G', 'T'};

struct Feistel {
    uint64_t n = 0, halfmask = 0, keys[4] = {0, 0, 0, 0};
    int halfbits = 1;
    void init(uint64_t n_, uint64_t seed) {
        n = n_;
        int bits = 2;
        while (bits < 64 && (1ull << bits) < n) ++bits;
        if (bits & 1) ++bits;
        halfbits = bits / 2;
        halfmask = (1ull << halfbits) - 1;
        uint64_t s = seed ^ 0xA5A5F00DCAFEBEEFull;
        for (auto& k : keys) k = splitmix64(s);
    }
    inline uint64_t once(uint64_t x) const {
        uint64_t l = x >> halfbits, r = x & halfmask;
        for (int i = 0; i < 4; ++i) {
            const uint64_t t = l ^ (mix64(r ^ keys[i]) & halfmask);
            l = r; r = t;
        }
        return (l << halfbits) | r;
    }
    inline uint64_t operator()(uint64_t x) const {   // permutation of [0,n)
        if (n < 2) return x;
        do { x = once(x); } while (x >= n);
        return x;
    }
};

struct Gen {
    int k = 0;
    uint64_t n = 0, c = 0, seed = 0;
    std::vector<uint8_t> genome;          // codes 0..3; contig i occupies [gstart[i], gstart[i]+len[i]+k-1)
    std::vector<uint64_t> node0;          // node0[i] = id of contig i's first node; node0[c] = n
    Feistel perm;                         // line -> node id
    int threads = 1;

    inline uint64_t contig_of(uint64_t node) const {
        return (uint64_t)(std::upper_bound(node0.begin(), node0.end(), node) - node0.begin()) - 1;
    }
    inline uint64_t gpos(uint64_t node, uint64_t ci) const { return node + ci * (uint64_t)(k - 1); }
    inline uint64_t contig_nodes(uint64_t ci) const { return node0[ci + 1] - node0[ci]; }
    inline uint64_t line_bytes() const { return (uint64_t)k + 4; }
    inline int packed_len() const { return (k + 3) / 4; }
};

template <class F> void parallel_for(uint64_t n, int threads, F f) {
    threads = (int)std::max<uint64_t>(1, std::min<uint64_t>(threads, n / 4096 + 1));
    if (threads == 1) { f(0, n, 0); return; }
    std::vector<std::thread> pool;
    const uint64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const uint64_t lo = std::min(n, per * t), hi = std::min(n, lo + per);
        pool.emplace_back([=] { f(lo, hi, t); });
    }
    for (auto& th : pool) th.join();
}

int build(Gen& g, uint64_t long_nodes) {
    const int k = g.k;
    const uint64_t n = g.n, c = g.c;
    if (k < 2 || k > 64 || n == 0 || c == 0 || c > n) return 1;
    if (long_nodes > 0 && (c < 2 ? long_nodes != n : long_nodes + (c - 1) > n)) return 1;
    if (k <= 31 && n > ((1ull << (2 * k)) >> 2)) return 2;   // too dense to make unique

    Xoshiro rng(g.seed);

    // --- contig node counts: broken stick over the n-1 gaps between nodes ---
    {
        std::vector<uint8_t> cut((n + 7) / 8, 0);   // bit j set: contig boundary after node j
        uint64_t first_free = 0, cuts_needed = c - 1;
        if (long_nodes > 0 && c > 1) {
            cut[(long_nodes - 1) >> 3] |= (uint8_t)(1u << ((long_nodes - 1) & 7));
            first_free = long_nodes;          // gaps [long_nodes, n-1) are eligible
            --cuts_needed;
        }
        const uint64_t gaps = (n - 1) - first_free;   // eligible gap ids: first_free .. n-2
        if (cuts_needed > gaps) return 1;
        for (uint64_t placed = 0; placed < cuts_needed;) {
            const uint64_t j = first_free + rng.below(gaps);
            uint8_t& b = cut[j >> 3];
            const uint8_t m = (uint8_t)(1u << (j & 7));
            if (!(b & m)) { b |= m; ++placed; }
        }
        g.node0.clear();
        g.node0.reserve(c + 1);
        g.node0.push_back(0);
        for (uint64_t j = 0; j + 1 < n; ++j)
            if (cut[j >> 3] & (1u << (j & 7))) g.node0.push_back(j + 1);
        g.node0.push_back(n);
        if (g.node0.size() != c + 1) return 3;
    }

    // --- genome + uniqueness ---
    g.genome.assign(n + c * (uint64_t)(k - 1), 0);
    const bool check = (k <= 31);
    KeySet set;
    if (check) set.init(n);
    const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    constexpr int kBlock = 32;

    for (uint64_t ci = 0; ci < c; ++ci) {
        const uint64_t nodes = g.contig_nodes(ci);
        uint8_t* s = &g.genome[g.gpos(g.node0[ci], ci)];
        const uint64_t len = nodes + k - 1;
        // draw all bases of this contig (64 bits -> 32 bases)
        for (uint64_t i = 0; i < len;) {
            uint64_t r = rng.next();
            for (int j = 0; j < 32 && i < len; ++j, ++i, r >>= 2) s[i] = (uint8_t)(r & 3);
        }
        if (!check) continue;
        uint64_t key = 0;
        for (int i = 0; i < k - 1; ++i) key = (key << 2) | s[i];
        for (uint64_t w0 = 0; w0 < nodes; w0 += kBlock) {
            const uint64_t w1 = std::min(nodes, w0 + kBlock);
            uint64_t kk = key;
            for (uint64_t w = w0; w < w1; ++w) {     // speculative keys -> prefetch
                kk = ((kk << 2) | s[w + k - 1]) & kmask;
                set.prefetch(kk);
            }
            for (uint64_t w = w0; w < w1; ++w) {
                const uint64_t prefix = (key << 2) & kmask;
                uint64_t cand = prefix | s[w + k - 1];
                int tries = 0;
                while (!set.insert(cand)) {
                    if (++tries > 64) return 4;
                    s[w + k - 1] = (uint8_t)(rng.next() & 3);
                    cand = prefix | s[w + k - 1];
                }
                key = cand;
            }
        }
    }
    g.perm.init(n, g.seed);
    return 0;
}

// Emit the text line / packed record for one node.
inline void node_info(const Gen& g, uint64_t node, const uint8_t*& s, char& back, char& fwd) {
    const uint64_t ci = g.contig_of(node);
    s = &g.genome[g.gpos(node, ci)];
    back = (node == g.node0[ci]) ? 'F' : kBase[s[-1]];
    fwd = (node + 1 == g.node0[ci + 1]) ? 'F' : kBase[s[g.k]];
}

inline uint64_t fnv1a(const char* p, uint64_t len) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint64_t i = 0; i < len; ++i) { h ^= (uint8_t)p[i]; h *= 0x100000001B3ull; }
    return mix64(h ^ len);
}

}  // namespace

extern "C" {

struct kg_handle { Gen g; };

int kg_create(int k, uint64_t n, uint64_t c, uint64_t long_nodes, uint64_t seed, int threads, kg_handle** out) {
    auto* h = new kg_handle();
    h->g.k = k; h->g.n = n; h->g.c = c; h->g.seed = seed;
    if (threads <= 0) threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    h->g.threads = threads;
    const int rc = build(h->g, long_nodes);
    if (rc) { delete h; *out = nullptr; return rc; }
    *out = h;
    return 0;
}
void kg_destroy(kg_handle* h) { delete h; }

uint64_t kg_n(const kg_handle* h) { return h->g.n; }
uint64_t kg_c(const kg_handle* h) { return h->g.c; }
uint64_t kg_text_bytes(const kg_handle* h) { return h->g.n * h->g.line_bytes(); }
uint64_t kg_pair_bytes(const kg_handle* h) { return h->g.n * (uint64_t)(h->g.packed_len() + 2); }
uint64_t kg_max_contig_nodes(const kg_handle* h) {
    uint64_t m = 0;
    for (uint64_t i = 0; i < h->g.c; ++i) m = std::max(m, h->g.contig_nodes(i));
    return m;
}

// Lines [line0, line0+count) in reference text format into out (count*(k+4) bytes).
void kg_write_text(const kg_handle* h, uint64_t line0, uint64_t count, char* out) {
    const Gen& g = h->g;
    parallel_for(count, g.threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint8_t* s; char b, f;
            node_info(g, g.perm(line0 + i), s, b, f);
            char* o = out + i * g.line_bytes();
            for (int j = 0; j < g.k; ++j) o[j] = kBase[s[j]];
            o[g.k] = ' '; o[g.k + 1] = b; o[g.k + 2] = f; o[g.k + 3] = '\n';
        }
    });
}

// Same lines as reference kmer_pair bytes: (k+3)/4 packed bytes, MSB-first 2-bit
// codes A=0 C=1 G=2 T=3, tail padded with A (packing.hpp:50-92), then back, fwd chars.
void kg_write_pairs(const kg_handle* h, uint64_t line0, uint64_t count, uint8_t* out) {
    const Gen& g = h->g;
    const int pl = g.packed_len(), rec = pl + 2;
    parallel_for(count, g.threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint8_t* s; char b, f;
            node_info(g, g.perm(line0 + i), s, b, f);
            uint8_t* o = out + i * rec;
            for (int q = 0; q < pl; ++q) {
                unsigned v = 0;
                for (int j = 0; j < 4; ++j) {
                    const int idx = 4 * q + j;
                    v = (v << 2) | (idx < g.k ? s[idx] : 0u);
                }
                o[q] = (uint8_t)v;
            }
            o[pl] = (uint8_t)b; o[pl + 1] = (uint8_t)f;
        }
    });
}

// Contigs whose start line falls in rank's block [ceil(n/P)*r, ...) (read_kmers.hpp:55-58),
// in line order, each followed by '\n'.  Call with out == nullptr to get the size.
uint64_t kg_write_expected(const kg_handle* h, int nranks, int rank, char* out, uint64_t* n_contigs) {
    const Gen& g = h->g;
    const uint64_t split = (g.n + nranks - 1) / nranks;
    const uint64_t l0 = std::min(g.n, split * (uint64_t)rank), l1 = std::min(g.n, l0 + split);
    const int T = g.threads;
    std::vector<std::vector<uint64_t>> found(T);      // contig ids, in line order per chunk
    parallel_for(l1 - l0, T, [&](uint64_t lo, uint64_t hi, int t) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint64_t node = g.perm(l0 + i), ci = g.contig_of(node);
            if (node == g.node0[ci]) found[t].push_back(ci);
        }
    });
    uint64_t total = 0, count = 0;
    std::vector<uint64_t> base(T + 1, 0);
    for (int t = 0; t < T; ++t) {
        base[t] = total;
        for (uint64_t ci : found[t]) total += g.contig_nodes(ci) + g.k - 1 + 1;
        count += found[t].size();
    }
    if (n_contigs) *n_contigs = count;
    if (!out) return total;
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
        pool.emplace_back([&, t] {
            char* o = out + base[t];
            for (uint64_t ci : found[t]) {
                const uint64_t len = g.contig_nodes(ci) + g.k - 1;
                const uint8_t* s = &g.genome[g.gpos(g.node0[ci], ci)];
                for (uint64_t j = 0; j < len; ++j) o[j] = kBase[s[j]];
                o[len] = '\n';
                o += len + 1;
            }
        });
    for (auto& th : pool) th.join();
    return total;
}

// Bytewise-sorted solution (LC_ALL=C sort order), each contig followed by '\n'.
uint64_t kg_write_solution(const kg_handle* h, char* out) {
    const Gen& g = h->g;
    const uint64_t total = g.genome.size() + g.c;
    if (!out) return total;
    std::vector<uint64_t> order(g.c);
    for (uint64_t i = 0; i < g.c; ++i) order[i] = i;
    auto span = [&](uint64_t ci, const uint8_t*& s, uint64_t& len) {
        s = &g.genome[g.gpos(g.node0[ci], ci)];
        len = g.contig_nodes(ci) + g.k - 1;
    };
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        const uint8_t *sa, *sb; uint64_t la, lb;
        span(a, sa, la); span(b, sb, lb);
        const int r = memcmp(sa, sb, std::min(la, lb));   // codes are ordered like the letters
        return r != 0 ? r < 0 : la < lb;
    });
    char* o = out;
    for (uint64_t ci : order) {
        const uint8_t* s; uint64_t len;
        span(ci, s, len);
        for (uint64_t j = 0; j < len; ++j) o[j] = kBase[s[j]];
        o[len] = '\n';
        o += len + 1;
    }
    return total;
}

// Order-independent digest of a buffer of '\n'-terminated lines:
// sum (mod 2^64) of a 64-bit hash per line, plus the line count and byte total.
void kg_digest_lines(const char* buf, uint64_t len, int threads, uint64_t* sum, uint64_t* lines) {
    if (threads <= 0) threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<uint64_t> acc{0}, cnt{0};
    parallel_for(len, threads, [&](uint64_t lo, uint64_t hi, int) {
        // a chunk owns the lines that START inside it
        uint64_t p = lo;
        if (lo > 0) { while (p < len && buf[p - 1] != '\n') ++p; }
        uint64_t a = 0, n = 0;
        while (p < hi) {
            const char* e = (const char*)memchr(buf + p, '\n', len - p);
            const uint64_t q = e ? (uint64_t)(e - buf) : len;
            a += fnv1a(buf + p, q - p);
            ++n;
            p = q + 1;
        }
        acc += a; cnt += n;
    });
    *sum = acc.load(); *lines = cnt.load();
}

// Digest of the generator's own contig set (equals kg_digest_lines of any correct output).
void kg_digest_expected(const kg_handle* h, uint64_t* sum, uint64_t* lines) {
    const Gen& g = h->g;
    std::atomic<uint64_t> acc{0};
    parallel_for(g.c, g.threads, [&](uint64_t lo, uint64_t hi, int) {
        std::string tmp;
        uint64_t a = 0;
        for (uint64_t ci = lo; ci < hi; ++ci) {
            const uint64_t len = g.contig_nodes(ci) + g.k - 1;
            const uint8_t* s = &g.genome[g.gpos(g.node0[ci], ci)];
            tmp.resize(len);
            for (uint64_t j = 0; j < len; ++j) tmp[j] = kBase[s[j]];
            a += fnv1a(tmp.data(), len);
        }
        acc += a;
    });
    *sum = acc.load(); *lines = g.c;
}

}  // extern "C"

#ifdef KG_MAIN
// gen_kmers K N C out.txt [--seed S] [--long L] [--solution path] [--threads T]
int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s K N C out.txt [--seed S] [--long L] [--solution path] [--threads T]\n", argv[0]);
        return 1;
    }
    const int k = atoi(argv[1]);
    const uint64_t n = strtoull(argv[2], nullptr, 10), c = strtoull(argv[3], nullptr, 10);
    const std::string out = argv[4];
    uint64_t seed = 267, longn = 0; int threads = 0;
    std::string sol;
    for (int i = 5; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        if (a == "--seed") seed = strtoull(argv[i + 1], nullptr, 10);
        else if (a == "--long") longn = strtoull(argv[i + 1], nullptr, 10);
        else if (a == "--solution") sol = argv[i + 1];
        else if (a == "--threads") threads = atoi(argv[i + 1]);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    if (sol.empty()) {
        sol = out;
        const size_t dot = sol.rfind(".txt");
        if (dot != std::string::npos && dot + 4 == sol.size()) sol.erase(dot);
        sol += "_solution.txt";   // scripts/check_it.sh:28 naming
    }
    kg_handle* h = nullptr;
    const int rc = kg_create(k, n, c, longn, seed, threads, &h);
    if (rc) { fprintf(stderr, "gen_kmers: generation failed (code %d)\n", rc); return 2; }
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { perror(out.c_str()); return 3; }
    const uint64_t chunk = 1u << 22, lb = (uint64_t)k + 4;
    std::vector<char> buf(chunk * lb);
    for (uint64_t l = 0; l < n; l += chunk) {
        const uint64_t cnt = std::min(chunk, n - l);
        kg_write_text(h, l, cnt, buf.data());
        if (fwrite(buf.data(), 1, cnt * lb, f) != cnt * lb) { perror("write"); return 3; }
    }
    fclose(f);
    std::vector<char> s(kg_write_solution(h, nullptr));
    kg_write_solution(h, s.data());
    f = fopen(sol.c_str(), "wb");
    if (!f) { perror(sol.c_str()); return 3; }
    fwrite(s.data(), 1, s.size(), f);
    fclose(f);
    uint64_t sum, lines;
    kg_digest_expected(h, &sum, &lines);
    printf("k=%d kmers=%" PRIu64 " contigs=%" PRIu64 " max_contig_nodes=%" PRIu64 " digest=%016" PRIx64 "\n",
           k, n, c, kg_max_contig_nodes(h), sum);
    kg_destroy(h);
    return 0;
}
#endif
