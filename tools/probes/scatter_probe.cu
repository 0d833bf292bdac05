// scatter_probe.cu -- what bounds the staging pass: N records, each appended to one of C per-chunk buffers chosen by
// a hash (atomicAdd on the chunk's cursor, then one 16-byte store).  Variants isolate the atomics, the stores and the
// footprint of the buffers (TLB reach).   nvcc -O3 -arch=sm_100a tools/probes/scatter_probe.cu -o scatter_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 mix(u32 h) { h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }

template <int MODE>   // 0: atomics + stores, 1: stores only (position from a hash), 2: atomics only, 3: read only,
                      // 4 / 5: atomics + stores, buffers interleaved in groups of 8 / 32 positions (position-major layout)
__global__ void __launch_bounds__(256) scatter_kernel(const uint4* __restrict__ in, u64 n, u32 C, u32 cap, u32* cursor, uint4* fine, u32* sink) {
    const u64 i0 = (u64)blockIdx.x * 1024 + threadIdx.x;
    uint4 v[4]; u32 c[4], at[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { const u64 i = i0 + r * 256; v[r] = i < n ? in[i] : make_uint4(0, 0, 0, 0); }
    u32 acc = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const u64 i = i0 + r * 256;
        c[r] = (u32)(((u64)mix((u32)i * 0x9E3779B1u + v[r].x) * C) >> 32);
        if (MODE == 0 || MODE == 2 || MODE >= 4) at[r] = i < n ? atomicAdd(&cursor[c[r]], 1u) : 0u;
        else at[r] = mix((u32)i) % (cap / 2);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const u64 i = i0 + r * 256;
        if (i >= n) continue;
        if (MODE == 0 || MODE == 1) { if (at[r] < cap) fine[(u64)c[r] * cap + at[r]] = v[r]; }
        else if (MODE >= 4) {
            constexpr u32 G = MODE == 4 ? 8u : 32u;
            if (at[r] < cap) fine[((u64)(at[r] / G) * C + c[r]) * G + (at[r] % G)] = v[r];
        }
        else acc += at[r] + v[r].y;
    }
    if (MODE >= 2 && acc == 0x12345678u) *sink = acc;
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 89710742ull;
    uint4* in; u32 *cursor, *sink;
    cudaMalloc(&in, n * 16); cudaMemset(in, 1, n * 16);
    cudaMalloc(&sink, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    struct Cfg { u32 C, cap; const char* what; } cfgs[] = {
        {52565, 4352, "52565 chunks x 4352 slots x 16 B = 3.66 GB (the chunk buffers as built)"},
        {52565, 2048, "52565 chunks x 2048 = 1.72 GB"},
        {6571, 16384, "6571 buffers x 16384 = 1.72 GB (8x fewer frontiers)"},
        {821, 131072, "821 buffers x 131072 = 1.72 GB (64x fewer frontiers)"},
        {52565, 256, "52565 chunks x 256 slots = 215 MB (inside the TLB reach; positions wrap)"},
    };
    for (auto& cf : cfgs) {
        uint4* fine; cudaMalloc(&fine, (u64)cf.C * cf.cap * 16); cudaMalloc(&cursor, (u64)cf.C * 4);
        printf("%s\n", cf.what);
        for (int mode = 0; mode < 6; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaMemset(cursor, 0, (u64)cf.C * 4);
                cudaEventRecord(a);
                const unsigned grid = (unsigned)((n + 1023) / 1024);
                const u32 cap = cf.cap;
                if (mode == 0) scatter_kernel<0><<<grid, 256>>>(in, n, cf.C, cap, cursor, fine, sink);
                if (mode == 1) scatter_kernel<1><<<grid, 256>>>(in, n, cf.C, cap, cursor, fine, sink);
                if (mode == 2) scatter_kernel<2><<<grid, 256>>>(in, n, cf.C, cap, cursor, fine, sink);
                if (mode == 3) scatter_kernel<3><<<grid, 256>>>(in, n, cf.C, cap, cursor, fine, sink);
                if (mode == 4) scatter_kernel<4><<<grid, 256>>>(in, n, cf.C, cap, cursor, fine, sink);
                if (mode == 5) scatter_kernel<5><<<grid, 256>>>(in, n, cf.C, cap, cursor, fine, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (rep && ms < best) best = ms;
            }
            const char* names[] = {"atomics + stores", "stores only (hashed position)", "atomics only", "read only", "atomics + stores, interleaved x8", "atomics + stores, interleaved x32"};
            printf("   %-32s %7.3f ms  %6.1f G records/s\n", names[mode], best, n / best / 1e6);
        }
        cudaFree(fine); cudaFree(cursor);
        if (cudaGetLastError() != cudaSuccess) { printf("cuda error\n"); return 1; }
    }
    return 0;
}
