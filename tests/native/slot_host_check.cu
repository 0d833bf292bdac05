// Host-only check of the __host__ __device__ helpers in slot.cuh (built with nvcc, runs without a GPU).
// Reads "k world" then hex records (kmer_pair bytes) from stdin; prints, per record:
//   slot(hex)  owner  next_key(hex)  prev_key(hex)  roundtrip_ok
// and first a line with ext_code of all 256 byte values.  tests/test_slot_host.py compares with the Python mirrors.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../cs267_hw3_b200/csrc/slot.cuh"

using namespace kh;

static void print128(u128 v) { printf("%016llx%016llx", v.hi, v.lo); }
static void print128(u64 v) { printf("%032llx", v); }

template <int W>
static void run(int k, int world) {
    typedef Slot<W> S;
    const int pl = (k + 3) / 4, pb = pl + 2;
    char line[256];
    while (scanf("%255s", line) == 1) {
        unsigned char rec[32] = {0};
        for (int i = 0; i < pb; ++i) { unsigned x; sscanf(line + 2 * i, "%2x", &x); rec[i] = (unsigned char)x; }
        bool ok = true;
        const typename S::value_t v = S::from_record(rec, k, pl, ok);
        unsigned char back[32] = {0};
        S::to_record(v, k, pl, back);
        const bool rt = memcmp(rec, back, pb) == 0;
        print128(v);
        CtGeom g = {};
        g.k = k; g.m = ct_minimizer_len(k); g.win = ct_window(k); g.world = world; g.chunks_per_rank = 1;
        u32 owner = 99u, chunk = 0;
        if (ok) ct_place(ct_min_hash<W>(S::key_only(v), g.m, g.win), g, owner, chunk);      // the rank that stores the k-mer (hash_map.hpp:28-30)
        printf(" %u ", owner);
        print128(S::next_key(v, k));
        printf(" ");
        print128(S::back(v) < 4 ? S::prev_key(v, k) : S::zero());
        printf(" %d %d\n", (int)ok, (int)rt);
    }
}

int main() {
    for (int c = 0; c < 256; ++c) printf("%u%c", ext_code((unsigned char)c), c == 255 ? '\n' : ' ');
    int k, world;
    if (scanf("%d %d", &k, &world) != 2) return 1;
    if (2 * k + 6 <= 64) run<1>(k, world); else run<2>(k, world);
    return 0;
}
