// chunk_balance_check.cu -- CPU gate for the chunk table's placement functions AS SHIPPED (csrc/slot.cuh: ct_min_hash,
// ct_place): chunk load spread against the build kernel's capacity, and k-mers per contracted segment.
// Host-only:  nvcc -O2 -std=c++17 tools/probes/chunk_balance_check.cu -o /tmp/cbc && /tmp/cbc [N] [world]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cs267_hw3_b200/csrc/slot.cuh"
using namespace kh;
static u64 rng_state = 267;
static inline u64 rng() { rng_state += 0x9E3779B97F4A7C15ull; u64 z = rng_state; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

template <int W> void run(int K, u64 N, int world, u32 max_slots, double lf) {
    typedef Slot<W> S;
    std::vector<unsigned char> g(N + 64);
    for (auto& b : g) b = rng() & 3;
    CtGeom geo = {};
    geo.k = K; geo.m = ct_minimizer_len(K); geo.win = ct_window(K); geo.world = world;
    const double mu = std::min(lf / 1.35, 0.55) * max_slots;
    geo.chunks_per_rank = (u32)std::ceil((double)N / world / mu);
    std::vector<u32> load((size_t)geo.chunks_per_rank * world, 0);
    std::vector<u64> per_rank(world, 0);
    typename S::value_t v = S::zero();
    u64 segs = 0, prev = ~0ull;
    for (u64 i = 0; i < N + K - 1; ++i) {
        // shift in base g[i]
        if constexpr (W == 1) { v = ((v >> 6 << 2 | g[i]) & ((1ull << (2 * K)) - 1)) << 6; }
        else {
            u128 key = S::shr(v, 6); key = S::shl(key, 2); key.lo |= g[i];
            const int bits = 2 * K; if (bits < 128) { if (bits <= 64) { key.hi = 0; if (bits < 64) key.lo &= (1ull << bits) - 1; } else key.hi &= (1ull << (bits - 64)) - 1; }
            v = S::shl(key, 6);
        }
        if (i + 1 < (u64)K) continue;
        u32 owner, chunk;
        ct_place(ct_min_hash<W>(v, geo.m, geo.win), geo, owner, chunk);
        const u64 gc = (u64)owner * geo.chunks_per_rank + chunk;
        ++load[gc]; ++per_rank[owner];
        if (gc != prev) ++segs;
        prev = gc;
    }
    double mean = (double)N / load.size(), var = 0; u32 mx = 0;
    for (u32 l : load) { var += ((double)l - mean) * ((double)l - mean); mx = std::max(mx, l); }
    u64 rmax = 0; for (u64 r : per_rank) rmax = std::max(rmax, r);
    printf("K=%d m=%d win=%d W=%d lf=%.2f: chunks %zu x cap %u | load mean %.0f sd %.1f%% max %u (%.0f%% of cap, %.0f%% of cap*lf) | k-mers/segment %.2f | busiest rank %+.2f%%\n",
           K, geo.m, geo.win, W, lf, load.size(), max_slots, mean, 100 * std::sqrt(var / load.size()) / mean, mx, 100.0 * mx / max_slots,
           100.0 * mx / (max_slots * lf), (double)N / segs, 100.0 * (rmax * world / (double)N - 1));
}
int main(int argc, char** argv) {
    const u64 N = argc > 1 ? strtoull(argv[1], 0, 10) : 20000000ull;
    const int world = argc > 2 ? atoi(argv[2]) : 1;
    for (double lf : {0.5, 0.9}) {
        run<1>(19, N, world, 6144, lf);
        run<1>(21, N, world, 6144, lf);
        run<2>(31, N, world, 4352, lf);
        run<2>(51, N, world, 4352, lf);
    }
    return 0;
}
