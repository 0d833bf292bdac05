// device_view_check.cu -- include/kh/device_table.cuh from a caller's own kernels: kh::device_find / kh::device_insert
// against the batch C ABI (kh_insert_pairs, kh_find) on the same table.  Prints "OK ..." or the first mismatch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I include tests/native/device_view_check.cu -L cs267_hw3_b200 -lkh_b200
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kh/device_table.cuh"

__global__ void find_all(kh_device_view v, const unsigned char* recs, int pb, int pl, int n, unsigned char* found, unsigned char* same) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char out[18];
    const bool f = kh::device_find(v, recs + (size_t)i * pb, out);      // the key is the first pl bytes of the record
    found[i] = f;
    bool eq = f;
    for (int j = 0; f && j < pb; ++j) eq = eq && out[j] == recs[(size_t)i * pb + j];
    same[i] = eq;
    (void)pl;
}
__global__ void insert_range(kh_device_view v, const unsigned char* recs, int pb, int first, int n, int* codes) {
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    codes[i] = kh::device_insert(v, recs + (size_t)i * pb);
}

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { printf("FAIL %s -> %d (%s)\n", #x, rc_, t ? kh_last_error(t) : ""); return 1; } } while (0)

int main(int argc, char** argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 19, n = argc > 2 ? atoi(argv[2]) : 20000;
    const int pl = (k + 3) / 4, pb = pl + 2, half = n / 2;
    kh_table* t = nullptr;
    CK(kh_create(k, n, 0.5, 0, &t));
    std::vector<unsigned char> recs((size_t)n * pb);
    unsigned long long s = 12345;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < pl; ++j) { s = s * 6364136223846793005ull + 1442695040888963407ull; recs[(size_t)i * pb + j] = (unsigned char)(s >> 33); }
        const int pad = 4 * pl - k;                                  // packing.hpp:85-91: the tail of the last byte is A-padded
        recs[(size_t)i * pb + pl - 1] &= (unsigned char)(0xFF << (2 * pad));
        recs[(size_t)i * pb + pl] = "ACGT"[(s >> 20) & 3];
        recs[(size_t)i * pb + pl + 1] = "ACGTF"[(s >> 24) % 5];
    }
    CK(kh_insert_pairs(t, recs.data(), half));
    kh_device_view v;
    CK(kh_get_device_view(t, &v));
    unsigned char *d_recs, *d_found, *d_same;
    int* d_codes;
    cudaMalloc(&d_recs, recs.size()); cudaMalloc(&d_found, n); cudaMalloc(&d_same, n); cudaMalloc(&d_codes, n * sizeof(int));
    cudaMemcpy(d_recs, recs.data(), recs.size(), cudaMemcpyHostToDevice);
    std::vector<unsigned char> found(n), same(n);
    std::vector<int> codes(n);
    find_all<<<(n + 255) / 256, 256>>>(v, d_recs, pb, pl, n, d_found, d_same);
    cudaMemcpy(found.data(), d_found, n, cudaMemcpyDeviceToHost); cudaMemcpy(same.data(), d_same, n, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i)
        if ((i < half) != (found[i] != 0) || (i < half && !same[i])) { printf("FAIL device_find record %d: found %d same %d\n", i, found[i], same[i]); return 1; }
    insert_range<<<(n - half + 255) / 256, 256>>>(v, d_recs, pb, half, n, d_codes);
    insert_range<<<1, 32>>>(v, d_recs, pb, 0, 32, d_codes);                  // already there: duplicates
    cudaMemcpy(codes.data(), d_codes, n * sizeof(int), cudaMemcpyDeviceToHost);
    for (int i = half; i < n; ++i) if (codes[i] != kh::kDevInserted) { printf("FAIL device_insert record %d -> %d\n", i, codes[i]); return 1; }
    for (int i = 0; i < 32; ++i) if (codes[i] != kh::kDevDuplicate) { printf("FAIL duplicate insert %d -> %d\n", i, codes[i]); return 1; }
    find_all<<<(n + 255) / 256, 256>>>(v, d_recs, pb, pl, n, d_found, d_same);
    cudaMemcpy(same.data(), d_same, n, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) if (!same[i]) { printf("FAIL device_find after device_insert, record %d\n", i); return 1; }
    // the batch ABI sees what the kernels inserted
    std::vector<unsigned char> keys((size_t)(n - half) * pl), got((size_t)(n - half) * pb), hf(n - half);
    for (int i = half; i < n; ++i) memcpy(&keys[(size_t)(i - half) * pl], &recs[(size_t)i * pb], pl);
    CK(kh_find(t, keys.data(), n - half, got.data(), hf.data()));
    for (int i = half; i < n; ++i)
        if (!hf[i - half] || memcmp(&got[(size_t)(i - half) * pb], &recs[(size_t)i * pb], pb)) { printf("FAIL kh_find of a device-inserted record %d\n", i); return 1; }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("FAIL cuda\n"); return 1; }
    kh_destroy(t);
    printf("OK k=%d n=%d slot_bytes=%d\n", k, n, v.slot_bytes);
    return 0;
}
