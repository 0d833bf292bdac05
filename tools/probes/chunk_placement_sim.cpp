// chunk_placement_sim.cpp -- CPU gate for the chunk-level minimizer placement ("v2 traverse").
//
// Placement under test: home CHUNK (a 64 KB piece of the table that one thread block builds in shared memory)
// = hash(minimizer of the k-mer), bucket inside the chunk = hash(key).  Consecutive k-mers of a contig share
// the minimizer for a "supermer", so they land in one chunk and can be contracted into one segment while the
// chunk sits in shared memory.  What has to hold for that to pay:
//   (1) chunk loads stay balanced (a chunk holds slots for load/LF k-mers; overflow spills),
//   (2) runs of consecutive k-mers inside one chunk are long (segments per k-mer small).
// Random genome, i.i.d. bases, N k-mers; prints per (K, m): segments per k-mer, chunk-load mean / sigma / max,
// fraction of chunks above 100 % of their slots at LF 0.5 / 0.7 / 0.9.
//   g++ -O2 -std=c++17 tools/probes/chunk_placement_sim.cpp -o /tmp/sim && /tmp/sim [N]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned long long u64;
typedef unsigned __int128 u128;
static inline u64 fmix64(u64 z) { z ^= z >> 33; z *= 0xFF51AFD7ED558CCDull; z ^= z >> 33; z *= 0xC4CEB9FE1A85EC53ull; z ^= z >> 33; return z; }
static u64 rng_state = 267;
static inline u64 rng() { rng_state += 0x9E3779B97F4A7C15ull; u64 z = rng_state; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
    const u64 N = argc > 1 ? strtoull(argv[1], 0, 10) : 89710742ull;
    std::vector<unsigned char> g(N + 64);
    for (u64 i = 0; i < g.size(); i += 32) { u64 r = rng(); for (int j = 0; j < 32 && i + j < g.size(); ++j) g[i + j] = (r >> (2 * j)) & 3; }
    const int Ks[] = {19, 31, 51};
    for (int K : Ks) {
        const int W = (2 * K + 6 <= 64) ? 1 : 2;
        const u64 slots_per_chunk = W == 1 ? 8192 : 4096;
        const int ms[] = {7, 9, 11, 13, 15, 17, 21, 25, 31};
        for (int m : ms) {
            if (m >= K) continue;
            const int w = K - m + 1;
            for (double lf : {0.5}) {
                const u64 nchunks = (u64)std::ceil((double)N / lf / slots_per_chunk);
                std::vector<unsigned> load(nchunks, 0);
                const u64 mask = (m >= 32) ? ~0ull : ((1ull << (2 * m)) - 1);
                // m-mer order values for every position
                std::vector<unsigned> ord(N + K);
                std::vector<u64> mm(N + K);
                u64 x = 0;
                for (u64 i = 0; i < N + K - 1; ++i) {
                    x = ((x << 2) | g[i]) & mask;
                    if (i + 1 >= (u64)m) { mm[i + 1 - m] = x; ord[i + 1 - m] = (unsigned)((x * 0x9E3779B97F4A7C15ull) >> 32); }
                }
                u64 segs = 0, prev_chunk = ~0ull;
                for (u64 i = 0; i < N; ++i) {          // k-mer i covers m-mers i .. i+w-1
                    unsigned best = 0xFFFFFFFFu; u64 bx = 0;
                    for (int j = 0; j < w; ++j) if (ord[i + j] < best) { best = ord[i + j]; bx = mm[i + j]; }
                    const u64 c = (u64)(((u128)fmix64(bx + 0x632BE59BD9B4E019ull) * nchunks) >> 64);
                    ++load[c];
                    if (c != prev_chunk) ++segs;
                    prev_chunk = c;
                }
                double mean = (double)N / nchunks, var = 0; unsigned mx = 0;
                for (unsigned l : load) { var += ((double)l - mean) * ((double)l - mean); mx = std::max(mx, l); }
                const double sd = std::sqrt(var / nchunks);
                auto over = [&](double f) { u64 n = 0; const double cap = slots_per_chunk * 0.5 / f; for (unsigned l : load) n += (l > cap); return (double)n / nchunks; };
                printf("K=%d m=%2d w=%2d: kmers/segment %.2f | chunks %llu load mean %.0f sd %.1f (%.1f%%) max %u (%.0f%% of slots at LF0.5) | chunks over capacity if LF were 0.7: %.4f  0.8: %.4f  0.9: %.4f\n",
                       K, m, w, (double)N / segs, nchunks, mean, sd, 100 * sd / mean, mx, 100.0 * mx / slots_per_chunk,
                       over(0.7), over(0.8), over(0.9));
                fflush(stdout);
            }
        }
    }
    return 0;
}
