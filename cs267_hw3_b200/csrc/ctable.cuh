// ctable.cuh -- the chunk table: a batch-built hash table whose build step also contracts every contig into
// one segment per supermer, for any number of GPUs.
//
// Replaces, together with the host code in capi.cu ("ct_*"):
//   DistributedHashMap::insert_all / insert_batch_remote   hash_map.hpp:55-80, 38-46   -> ct_stage / ct_scatter / ct_build
//   DistributedHashMap::find (local + remote branch)        hash_map.hpp:83-107          -> ct_find, ct_resolve (+ inbox)
//   assemble_contigs                                        kmer_hash.cpp:38-55          -> ct_build (in shared memory) + ct_resolve
//                                                                                           + ct_rank_round + ct_lengths/claim + ct_emit
//
// Why: a contig walk over a plainly hashed table costs one random DRAM access per k-mer, and HBM serves only
// ~41 G of those per second whatever their size (profiles/r01_random_access_probe_occ8.txt) -- 2.7 ms for the
// chr14 shape, 0.17 of the bandwidth roofline.  Here the home CHUNK of a k-mer (<= 64-72 KB of table) is chosen
// by a hash of its MINIMIZER and only the bucket inside the chunk by the key (slot.cuh).  Consecutive k-mers of a
// contig share the minimizer for a supermer, so while a block holds a chunk in shared memory to build it, it also
// follows every successor link that stays inside the chunk (an LDS probe instead of a DRAM access) and emits one
// contracted SEGMENT per run: its forward-extension characters, and the key of the k-mer that follows it.  What
// is left for HBM is ONE lookup per segment (every ~4 k-mers at K=19, ~16 at K=51) to chain the segments, then
// pointer jumping over the segment list and the copy of the characters.
//
// The table is static: inserts are only STAGED (grouped by table region, ct_stage_kernel); the first find /
// assemble SEALS it: group by chunk (ct_scatter_kernel), size every chunk for its actual load (ct_layout_kernel:
// each chunk gets load / load_factor slots, so the load factor holds per chunk no matter how lumpy the minimizer
// hash is) and build + contract (ct_build_kernel).  Inserting after a seal re-seals from the staged records.
//
// Multi-GPU: chunk ids are global, rank r owns chunks [r*C, (r+1)*C).  ct_stage_kernel writes a record's slot
// value straight into the staging buffer of the owning GPU through its NVLink peer mapping while it groups
// (the all-to-all of hash_map.hpp:64-77 fused into the grouping pass, no separate exchange step); pending
// segment links travel the same way (ct_resolve_kernel -> the owner's inbox) and are answered with one peer store.
// Phases are separated by an in-stream flag barrier (ct_barrier_kernel); the host never waits inside a step.
//
// Segment ids: global id = (rank << 28) | local id.  Local ids [0, hcap) are the contig HEAD STUBS of the start
// nodes parsed by this rank (zero characters, pending link to the segment that starts with the start k-mer --
// rank r therefore emits exactly the reference's <prefix>_<r>.dat); chunk segments follow from hcap upward.
//   link[seg]  = (next segment gid << 32) | characters between this segment's start and the start of `next`
//                (markers in the high word: kLinkTail / kLinkPending / kLinkMissing / kLinkConverge / kLinkClaimed)
//   meta[seg]  = (offset of the segment's characters in the rank's character pool << 24) | number of characters
#pragma once
#include "kernels.cuh"

namespace kh {

template <int W> struct CtBuild {
    static constexpr u32 kMaxBuckets = (W == 1) ? 2048u : 2304u;                 // 64 / 72 KB of table per chunk
    static constexpr u32 kMaxSlots = kMaxBuckets * (u32)Slot<W>::kPerBucket;      // 8192 / 4608
    // table | succ u16[slots] | heads u16[slots] | characters u8[slots] | two bitmaps u32[slots / 32]
    static constexpr u32 kOffSucc = kMaxBuckets * 32u;
    static constexpr u32 kOffHeads = kOffSucc + kMaxSlots * 2u;
    static constexpr u32 kOffPool = kOffHeads + kMaxSlots * 2u;
    static constexpr u32 kOffBits = kOffPool + kMaxSlots;
    static constexpr u32 kSmem = kOffBits + (kMaxSlots / 32u) * 8u;               // 108544 / 97920 B: two blocks per SM
};
constexpr int kCtBuildThreads = 512;
constexpr u32 kSuccExt = 0xFFFFu, kSuccTail = 0xFFFEu, kSuccNone = 0xFFFDu;

template <int W> struct CtReq;                                                      // a pending link on its way to the owner
template <> struct alignas(16) CtReq<1> { u64 key; u32 src; u32 chunk; };
template <> struct alignas(16) CtReq<2> { u128 key; u32 src; u32 chunk; u64 pad; };

struct CtPeers {
    int world, rank;
    void* stage_vals[kMaxRanks];          // [R][world][cap_rs] slot values, grouped by (local region, source rank)
    unsigned short* stage_tags[kMaxRanks];   // chunk inside the region, same shape
    u32* stage_cnt[kMaxRanks];            // [R][world] published fill of each (region, source) buffer
    void* extra_vals[kMaxRanks];          // records that found their staging buffer full
    u32* extra_chunk[kMaxRanks];
    u32* extra_cnt[kMaxRanks];
    u64* link[kMaxRanks];
    u64* meta[kMaxRanks];
    void* inbox[kMaxRanks];               // [world][inbox_cap] link requests, by source rank
    u32* inbox_cnt[kMaxRanks];
    u32* contig_pre[kMaxRanks];
    u64* contig_off[kMaxRanks];
    char* out[kMaxRanks];
    u64 out_cap[kMaxRanks];
    u32* flags[kMaxRanks];                // barrier: flags[r][2 * s + parity] = (epoch << 1 | payload bit) rank s signalled to rank r
};

struct CtCaps {
    u32 cap_rs, extra_cap, inbox_cap, seg_cap, hcap;
    u64 nbuckets_alloc, pool_cap;
};

__device__ __forceinline__ void ct_internal(Counters* ctr, u32 site) {
    atomicOr(&ctr->errors, kErrInternal);
    atomicOr(&ctr->err_where, site);
}

// ---- in-stream barrier across the GPUs of one step -------------------------------------------------------
// One warp; lane r signals rank r (a store into ITS flag array) and waits for rank r's signal in ours.  Everything
// the previous kernels of this stream wrote -- including stores into peer memory -- is complete when this kernel
// starts, and the release/acquire pair orders the flag against them.  A peer that never arrives (it failed) ends
// the wait after ~4 s with kErrInternal instead of hanging the GPU.
// The flag carries one payload bit next to the epoch.  The barriers between pointer-jumping rounds (round >= 0)
// use it for "this rank moved a link in this round": every rank sees the same bits, so all of them agree when
// nobody moved -- they set rank_done and the remaining round and barrier kernels of the step return at once.
__global__ void ct_barrier_kernel(const CtPeers pe, u32 epoch, Counters* ctr, int round) {
    const int r = (int)threadIdx.x;
    if (round >= 0 && ctr->rank_done) return;                    // agreed on by all ranks: nobody waits here any more
    const u32 bit = round >= 0 ? (ctr->flags[round] ? 1u : 0u) : 0u;
    u32 peer_bit = 0;
    if (r < pe.world) {
        __threadfence_system();
        u32* dst = pe.flags[r] + 2 * pe.rank + (epoch & 1u);           // two slots per peer, by epoch parity: a peer that is already
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"((epoch << 1) | bit) : "memory");
        const u32* src = pe.flags[pe.rank] + 2 * r + (epoch & 1u);      // one barrier ahead cannot overwrite the value we still have to read
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            u32 seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(src) : "memory");
            if ((int)((seen >> 1) - epoch) >= 0) { peer_bit = ((seen >> 1) == epoch) ? (seen & 1u) : 1u; break; }
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 4000000000ull) { ct_internal(ctr, kSiteBarrier); peer_bit = 1; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    if (round >= 0) {
        const u32 any = __ballot_sync(kFullMask, peer_bit != 0);
        if (r == 0 && any == 0) ctr->rank_done = 1;
    }
}

// ---- pass A: stage records, grouped by table region, in the owner's memory --------------------------------
// Block-local counting sort of 2048 records by global region (owner rank, region of 2^cpr_shift chunks), then
// one run per (block, region) appended to this source's buffer of that region ON THE OWNER GPU (peer stores
// through NVLink when the owner is another GPU; the cursors are local because every source has its own buffer).
// Also the start bitmask / per-tile start counts for the order-preserving start scan (kmer_hash.cpp:27-31).
template <int W>
__global__ void __launch_bounds__(kPartThreads, 4)
ct_stage_kernel(const unsigned char* __restrict__ recs, u64 n, const CtGeom g, const CtPeers pe, const CtCaps caps,
                u32* __restrict__ reg_cursor, u32* __restrict__ start_mask, u32* __restrict__ tile_starts, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const u32 R = g.regions_per_rank, nreg = R * (u32)g.world;
    u32* s_hist = reinterpret_cast<u32*>(s_raw);
    u32* s_off = s_hist + nreg;
    u32* s_gbase = s_off + nreg;
    unsigned short* s_pid = reinterpret_cast<unsigned short*>(s_raw + 12 * (size_t)nreg);
    unsigned short* s_tag = s_pid + kPartTile;
    unsigned char* s_union = s_raw + ((12 * (size_t)nreg + 4 * kPartTile + 15) & ~(size_t)15);
    unsigned char* s_rec = s_union;
    V* s_sorted = reinterpret_cast<V*>(s_union);
    __shared__ u64 s_warp[33];
    __shared__ u32 s_starts[kPartTile / kInsTile], s_err;
    const int k = g.k, pl = (k + 3) >> 2, pb = pl + 2;
    const u64 rec0 = (u64)blockIdx.x * kPartTile;
    const u32 cnt = (u32)min((u64)kPartTile, n - rec0);
    for (u32 i = threadIdx.x; i < nreg; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < kPartTile / kInsTile) s_starts[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_err = 0;
    stage_in(s_rec, recs + rec0 * pb, cnt * pb);
    __syncthreads();
    V v[kPartPerThread];
    u32 pid[kPartPerThread], rk[kPartPerThread], tag[kPartPerThread];
    u32 err = 0;
    const u32 tag_mask = (1u << g.cpr_shift) - 1u;
#pragma unroll
    for (int r = 0; r < kPartPerThread; ++r) {
        const u32 j = threadIdx.x + r * kPartThreads;
        bool ok = true, live = j < cnt;
        pid[r] = 0xFFFFFFFFu; tag[r] = 0;
        v[r] = S::zero();
        if (live) v[r] = S::from_record_staged(s_rec + j * pb, k, pl, ok);
        if (!ok) { err |= kErrBadInput; live = false; }
        if (live) {
            u32 owner, chunk;
            ct_place(ct_min_hash<W>(v[r], g.m, g.win), g, owner, chunk);
            pid[r] = owner * R + (chunk >> g.cpr_shift);
            tag[r] = chunk & tag_mask;
            rk[r] = atomicAdd(&s_hist[pid[r]], 1u);
        }
        const u32 bal = __ballot_sync(kFullMask, live && S::back(v[r]) == kExtF);
        const u64 first = rec0 + (u64)r * kPartThreads + (threadIdx.x & ~31u);
        if (lane_id() == 0 && first < n) {
            start_mask[first >> 5] = bal;
            if (bal) atomicAdd(&s_starts[(r * kPartThreads + threadIdx.x) / kInsTile], (u32)__popc(bal));
        }
    }
    err = __reduce_or_sync(kFullMask, err);
    if (lane_id() == 0 && err) atomicOr(&s_err, err);
    __syncthreads();
    {   // exclusive scan of the histogram (nreg <= 1024 = 4 per thread) + reservations in this source's buffers
        u32 h[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u32 i = threadIdx.x * 4 + q;
            h[q] = i < nreg ? s_hist[i] : 0u;
            sum += h[q];
        }
        u64 total;
        u64 run = block_exclusive_scan((u64)sum, s_warp, total);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u32 i = threadIdx.x * 4 + q;
            if (i < nreg) {
                s_off[i] = (u32)run;
                s_gbase[i] = h[q] ? atomicAdd(&reg_cursor[i], h[q]) : 0u;
            }
            run += h[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kPartPerThread; ++r) {
        if (pid[r] != 0xFFFFFFFFu) {
            const u32 pos = s_off[pid[r]] + rk[r];
            s_sorted[pos] = v[r];
            s_pid[pos] = (unsigned short)pid[r];
            s_tag[pos] = (unsigned short)tag[r];
        }
    }
    __syncthreads();
    const u32 good = s_off[nreg - 1] + s_hist[nreg - 1];
    for (u32 pos = threadIdx.x; pos < good; pos += blockDim.x) {
        const u32 p = s_pid[pos];
        const u32 owner = p / R, lr = p - owner * R;
        const u32 at = s_gbase[p] + (pos - s_off[p]);
        if (at < caps.cap_rs) {
            const u64 slot = ((u64)lr * (u32)g.world + (u32)g.rank) * caps.cap_rs + at;
            static_cast<V*>(pe.stage_vals[owner])[slot] = s_sorted[pos];
            pe.stage_tags[owner][slot] = s_tag[pos];
        } else {                                           // this (region, source) buffer is full: hand it to the owner one by one
            const u32 o = atomicAdd_system(pe.extra_cnt[owner], 1u);
            if (o < caps.extra_cap) {
                static_cast<V*>(pe.extra_vals[owner])[o] = s_sorted[pos];
                pe.extra_chunk[owner][o] = (lr << g.cpr_shift) | s_tag[pos];
            } else {
                atomicOr(&s_err, kErrTableFull);            // far more records than the table was sized for
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < kPartTile / kInsTile) {
        const u64 tile = (u64)blockIdx.x * (kPartTile / kInsTile) + threadIdx.x;
        if (tile * kInsTile < n) tile_starts[tile] = s_starts[threadIdx.x];
    }
    if (threadIdx.x == 0 && s_err) atomicOr(&ctr->errors, s_err);
}

// tell every owner how much this source has put into each of its regions' buffers
__global__ void __launch_bounds__(256)
ct_publish_stage_kernel(const CtGeom g, const CtPeers pe, const CtCaps caps, const u32* __restrict__ reg_cursor) {
    const u32 R = g.regions_per_rank, nreg = R * (u32)g.world;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nreg; i += gridDim.x * blockDim.x) {
        const u32 owner = i / R, lr = i - owner * R;
        pe.stage_cnt[owner][lr * (u32)g.world + (u32)g.rank] = min(reg_cursor[i], caps.cap_rs);
    }
}

// start nodes seen so far live on the device (no host round trip per insert call in the sharded path)
__global__ void ct_bump_starts_kernel(Counters* ctr, u32 hcap) {
    const u64 total = ctr->n_starts_dev + ctr->scan_total;
    if (total > hcap) { ct_internal(ctr, kSiteStarts); ctr->n_starts_dev = hcap; }
    else ctr->n_starts_dev = total;
}

// ---- pass B: the staged values of one (region, source) buffer -> per-chunk buffers --------------------------
template <int W>
__global__ void __launch_bounds__(kSubThreads, 4)
ct_scatter_kernel(const typename Slot<W>::value_t* __restrict__ stage_vals, const unsigned short* __restrict__ stage_tags,
                  const u32* __restrict__ stage_cnt, const CtGeom g, const CtCaps caps, u32 blocks_per_rs,
                  u32* __restrict__ chunk_cursor, typename Slot<W>::value_t* __restrict__ fine, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const u32 cpr = 1u << g.cpr_shift;
    u32* s_hist = reinterpret_cast<u32*>(s_raw);
    u32* s_gbase = s_hist + cpr;
    const u32 rs = blockIdx.x / blocks_per_rs, jblk = blockIdx.x % blocks_per_rs;      // rs = lr * world + src
    const u32 n = min(stage_cnt[rs], caps.cap_rs);
    const u32 base = jblk * kSubTile;
    if (base >= n) return;
    const u32 lr = rs / (u32)g.world;
    const V* __restrict__ src = stage_vals + (u64)rs * caps.cap_rs;
    const unsigned short* __restrict__ tags = stage_tags + (u64)rs * caps.cap_rs;
    for (u32 i = threadIdx.x; i < cpr; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    V v[kSubPerThread];
    u32 sid[kSubPerThread], rk[kSubPerThread];
#pragma unroll
    for (int r = 0; r < kSubPerThread; ++r) {
        const u32 i = base + r * kSubThreads + threadIdx.x;
        sid[r] = 0xFFFFFFFFu;
        if (i < n) { v[r] = src[i]; sid[r] = tags[i]; }
    }
#pragma unroll
    for (int r = 0; r < kSubPerThread; ++r)
        if (sid[r] != 0xFFFFFFFFu) rk[r] = atomicAdd(&s_hist[sid[r]], 1u);
    __syncthreads();
    const u32 first_chunk = lr << g.cpr_shift;
    for (u32 i = threadIdx.x; i < cpr; i += blockDim.x)
        s_gbase[i] = s_hist[i] ? atomicAdd(&chunk_cursor[first_chunk + i], s_hist[i]) : 0u;
    __syncthreads();
    constexpr u32 kMaxSlots = CtBuild<W>::kMaxSlots;
#pragma unroll
    for (int r = 0; r < kSubPerThread; ++r) {
        if (sid[r] == 0xFFFFFFFFu) continue;
        const u32 at = s_gbase[sid[r]] + rk[r];
        if (at < kMaxSlots) fine[(u64)(first_chunk + sid[r]) * kMaxSlots + at] = v[r];
        else atomicOr(&ctr->errors, kErrTableFull);           // more k-mers share this chunk than a chunk can hold
    }
}

template <int W>
__global__ void __launch_bounds__(256)
ct_extra_kernel(const typename Slot<W>::value_t* __restrict__ extra_vals, const u32* __restrict__ extra_chunk,
                const u32* __restrict__ extra_cnt, const CtCaps caps, u32* __restrict__ chunk_cursor,
                typename Slot<W>::value_t* __restrict__ fine, Counters* ctr) {
    constexpr u32 kMaxSlots = CtBuild<W>::kMaxSlots;
    const u32 n = min(*extra_cnt, caps.extra_cap);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 c = extra_chunk[i];
        const u32 at = atomicAdd(&chunk_cursor[c], 1u);
        if (at < kMaxSlots) fine[(u64)c * kMaxSlots + at] = extra_vals[i];
        else atomicOr(&ctr->errors, kErrTableFull);
    }
}

// ---- layout: every chunk gets load / load_factor slots ------------------------------------------------------
// chunk_base[c] = first bucket of chunk c (C + 1 entries), pool_off[c] = offset of its characters (16-byte aligned)
__global__ void __launch_bounds__(1024)
ct_layout_kernel(const u32* __restrict__ chunk_cursor, u32 nchunks, const CtGeom g, const CtCaps caps, u32 per_bucket, u32 max_slots,
                 u32* __restrict__ chunk_base, u32* __restrict__ pool_off, Counters* ctr) {
    __shared__ u64 s_warp[33];
    u64 carry_b = 0, carry_p = 0;
    for (u32 c0 = 0; c0 < nchunks; c0 += blockDim.x) {
        const u32 c = c0 + threadIdx.x;
        u32 capb = 0, chars = 0;
        if (c < nchunks) {
            const u32 load = min(chunk_cursor[c], max_slots);
            if (load) {
                const u64 want_slots = ((u64)load * g.lf_inv_q16 + 65535u) >> 16;
                capb = (u32)min((u64)g.max_buckets, max((want_slots + per_bucket - 1) / per_bucket, (u64)(load / per_bucket + 1)));
                capb = min(capb, g.max_buckets);
            }
            chars = (load + 15u) & ~15u;
        }
        u64 tot_b, tot_p;
        const u64 eb = block_exclusive_scan((u64)capb, s_warp, tot_b);
        const u64 ep = block_exclusive_scan((u64)chars, s_warp, tot_p);
        if (c < nchunks) { chunk_base[c] = (u32)(carry_b + eb); pool_off[c] = (u32)(carry_p + ep); }
        carry_b += tot_b; carry_p += tot_p;
    }
    if (threadIdx.x == 0) {
        chunk_base[nchunks] = (u32)carry_b;
        pool_off[nchunks] = (u32)carry_p;
        if (carry_b > caps.nbuckets_alloc || carry_p > caps.pool_cap) atomicOr(&ctr->errors, kErrTableFull);
    }
}

// ---- pass C: build a chunk in shared memory and contract its chains ------------------------------------------
template <int W>
__device__ __forceinline__ void ct_lds_bucket(unsigned saddr, u64 (&q)[4]) {
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(q[0]), "=l"(q[1]) : "r"(saddr));
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2+16];" : "=l"(q[2]), "=l"(q[3]) : "r"(saddr));
}

// insert into the chunk held in shared memory; linear probing that wraps inside the chunk
template <int W>
__device__ __forceinline__ int ct_smem_insert(typename Slot<W>::value_t* s_tab, u32 nb, typename Slot<W>::value_t v) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_tab);
    u32 b = ct_bucket_in_chunk(CtSlot<W>::hash32(S::key_only(v)), nb);
    for (u32 tries = 0; tries < nb;) {
        u64 q[4];
        lds_bucket(s_base + b * 32u, q);
        int j = -1;
        bool dup = false;
#pragma unroll
        for (int i = S::kPerBucket - 1; i >= 0; --i) {
            const V c = S::from_bucket(q, i);
            if (S::empty(c)) j = i;
            else if (S::same_key(c, v)) dup = true;
        }
        if (dup) return kInsDuplicate;
        if (j < 0) { b = (b + 1 == nb) ? 0u : b + 1; ++tries; continue; }
        const V old = S::cas_shared(s_tab + b * S::kPerBucket + j, S::zero(), v);
        if (S::empty(old)) return kInsInserted;
        if (S::same_key(old, v)) return kInsDuplicate;
    }
    return kInsFull;
}

// slot index of `key` (ext and index bits 0) in the finished chunk, or -1
template <int W>
__device__ __forceinline__ int ct_smem_find(unsigned s_base, u32 nb, typename Slot<W>::value_t key) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    u32 b = ct_bucket_in_chunk(CtSlot<W>::hash32(key), nb);
    for (u32 tries = 0; tries < nb; ++tries) {
        u64 q[4];
        ct_lds_bucket<W>(s_base + b * 32u, q);
#pragma unroll
        for (int i = 0; i < S::kPerBucket; ++i) {
            const V c = S::from_bucket(q, i);
            if (S::empty(c)) return -1;
            if (S::same_key(c, key)) return (int)(b * S::kPerBucket + i);
        }
        b = (b + 1 == nb) ? 0u : b + 1;
    }
    return -1;
}

template <int W>
__global__ void __launch_bounds__(kCtBuildThreads, 2)
ct_build_kernel(const typename Slot<W>::value_t* __restrict__ fine, const u32* __restrict__ chunk_cursor,
                const u32* __restrict__ chunk_base, const u32* __restrict__ pool_off,
                typename Slot<W>::value_t* __restrict__ table, u32* __restrict__ seg_base,
                u64* __restrict__ link, u64* __restrict__ meta, typename Slot<W>::value_t* __restrict__ ext_key,
                unsigned char* __restrict__ pool, const CtGeom g, const CtCaps caps, Counters* ctr) {
    typedef Slot<W> S;
    typedef CtSlot<W> CS;
    typedef typename S::value_t V;
    typedef CtBuild<W> B;
    extern __shared__ __align__(16) unsigned char s_raw[];
    V* s_tab = reinterpret_cast<V*>(s_raw);
    unsigned short* s_succ = reinterpret_cast<unsigned short*>(s_raw + B::kOffSucc);
    unsigned short* s_heads = reinterpret_cast<unsigned short*>(s_raw + B::kOffHeads);
    unsigned char* s_pool = s_raw + B::kOffPool;
    u32* s_haspred = reinterpret_cast<u32*>(s_raw + B::kOffBits);
    u32* s_multi = s_haspred + B::kMaxSlots / 32u;
    __shared__ u32 s_inserted, s_dups, s_nheads, s_seg0, s_chars, s_err;
    const u32 c = blockIdx.x;
    const u32 b0 = chunk_base[c], nb = chunk_base[c + 1] - b0;
    const u32 cnt = min(chunk_cursor[c], B::kMaxSlots);
    if (nb == 0) {                                   // nothing hashed here
        if (threadIdx.x == 0) seg_base[c] = 0;
        return;
    }
    const u32 nslots = nb * S::kPerBucket;
    const V* __restrict__ recs = fine + (u64)c * B::kMaxSlots;
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_tab);
    const int k = g.k;
    constexpr int kBatch = 4;
    V v[kBatch];
#pragma unroll
    for (int r = 0; r < kBatch; ++r) {
        const u32 i = threadIdx.x + r * kCtBuildThreads;
        v[r] = i < cnt ? recs[i] : S::zero();
    }
    if (threadIdx.x == 0) { s_inserted = 0; s_dups = 0; s_nheads = 0; s_chars = 0; s_err = 0; }
    {
        uint4* s4 = reinterpret_cast<uint4*>(s_raw);
        for (u32 i = threadIdx.x; i < nb * 2u; i += blockDim.x) s4[i] = make_uint4(0, 0, 0, 0);
        for (u32 i = threadIdx.x; i < B::kMaxSlots / 16u; i += blockDim.x) s_haspred[i] = 0;      // both bitmaps
    }
    __syncthreads();
    // ---- 1. insert ----
    u32 inserted = 0, dups = 0, err = 0;
    for (u32 base = 0; base < cnt; base += kCtBuildThreads * kBatch) {
        V nxt[kBatch];
#pragma unroll
        for (int r = 0; r < kBatch; ++r) {
            const u32 i = base + kCtBuildThreads * kBatch + threadIdx.x + r * kCtBuildThreads;
            nxt[r] = i < cnt ? recs[i] : S::zero();
        }
#pragma unroll
        for (int r = 0; r < kBatch; ++r) {
            if (S::empty(v[r])) continue;
            const int rc = ct_smem_insert<W>(s_tab, nb, v[r]);
            inserted += (rc == kInsInserted);
            dups += (rc == kInsDuplicate);
            if (rc == kInsFull) err |= kErrTableFull;
        }
#pragma unroll
        for (int r = 0; r < kBatch; ++r) v[r] = nxt[r];
    }
    inserted = __reduce_add_sync(kFullMask, inserted);
    dups = __reduce_add_sync(kFullMask, dups);
    err = __reduce_or_sync(kFullMask, err);
    if (lane_id() == 0) {
        if (inserted) atomicAdd(&s_inserted, inserted);
        if (dups) atomicAdd(&s_dups, dups);
        if (err) atomicOr(&s_err, err);
    }
    __syncthreads();
    // ---- 2. successor of every k-mer, if it lives in this chunk (kmer_hash.cpp:44-51 as an LDS probe) ----
    for (u32 i = threadIdx.x; i < nslots; i += blockDim.x) {
        const V cur = s_tab[i];
        u32 s = kSuccNone;
        if (!S::empty(cur)) {
            if (S::fwd(cur) == kExtF) {
                s = kSuccTail;
            } else {
                const int j = ct_smem_find<W>(s_base, nb, S::next_key(cur, k));
                if (j < 0) {
                    s = kSuccExt;
                } else {
                    s = (u32)j;
                    const u32 bit = 1u << (j & 31);
                    const u32 old = atomicOr(&s_haspred[j >> 5], bit);
                    if (old & bit) atomicOr(&s_multi[j >> 5], bit);       // two predecessors: j must start a segment of its own
                }
            }
        }
        s_succ[i] = (unsigned short)s;
    }
    __syncthreads();
    // ---- 3. heads: k-mers no chain of this chunk runs into (or that two run into, or with backward ext 'F') ----
    for (u32 base = 0; base < nslots; base += blockDim.x) {
        const u32 i = base + threadIdx.x;
        bool head = false;
        V cur = S::zero();
        if (i < nslots && s_succ[i] != kSuccNone) {
            cur = s_tab[i];
            const u32 bit = 1u << (i & 31);
            head = !(s_haspred[i >> 5] & bit) || (s_multi[i >> 5] & bit) || S::back(cur) == kExtF;
        }
        const u32 bal = __ballot_sync(kFullMask, head);
        u32 first = 0;
        if (lane_id() == 0 && bal) first = atomicAdd(&s_nheads, (u32)__popc(bal));
        first = __shfl_sync(kFullMask, first, 0);
        if (head) {
            const u32 idx = first + __popc(bal & ((1u << lane_id()) - 1u));
            s_heads[idx] = (unsigned short)i;
            s_tab[i] = CS::with_idx(cur, idx + 1u);
        }
    }
    __syncthreads();
    const u32 nheads = s_nheads;
    if (threadIdx.x == 0) {
        u32 seg0 = nheads ? atomicAdd(&ctr->next_seg, nheads) : 0u;
        if (seg0 + nheads > caps.seg_cap) { ct_internal(ctr, kSiteSegCap); seg0 = caps.seg_cap; }
        s_seg0 = seg0;
        seg_base[c] = seg0;
    }
    __syncthreads();
    const u32 seg0 = s_seg0;
    const u32 my_bits = (u32)g.rank << kRankShift;
    const u64 pool0 = pool_off[c];
    // ---- 4. one lane per head: walk its chain through shared memory, write the contracted segment ----
    if (seg0 < caps.seg_cap) {
        for (u32 h = threadIdx.x; h < nheads; h += blockDim.x) {
            const u32 i0 = s_heads[h];
            u32 j = i0, last = i0, n = 0, next_hi = kLinkTail;
            for (u32 guard = 0; guard <= cnt; ++guard) {
                const V cur = s_tab[j];
                if (S::fwd(cur) == kExtF) { next_hi = kLinkTail; break; }
                ++n; last = j;
                const u32 s = s_succ[j];
                if (s == kSuccExt) { next_hi = kLinkPending; break; }
                const u32 sidx = CS::idx(s_tab[s]);
                if (sidx) { next_hi = my_bits | (seg0 + sidx - 1u); break; }      // the successor starts its own segment
                j = s;
            }
            const u32 off = atomicAdd(&s_chars, n);
            j = i0;
            for (u32 t = 0; t < n; ++t) {
                s_pool[off + t] = ext_char(S::fwd(s_tab[j]));                      // extract_contig: read_kmers.hpp:86-90
                j = s_succ[j];
            }
            const u32 gseg = seg0 + h;
            link[gseg] = ((u64)next_hi << 32) | (next_hi == kLinkTail ? 0u : n);
            meta[gseg] = ((pool0 + off) << 24) | n;
            if (next_hi == kLinkPending) ext_key[gseg] = S::next_key(CS::strip(s_tab[last]), k);
        }
    }
    __syncthreads();
    // ---- 5. characters and the finished chunk go to HBM in coalesced sweeps ----
    {
        const u32 nv = (s_chars + 15u) >> 4;
        const uint4* sp = reinterpret_cast<const uint4*>(s_pool);
        uint4* gp = reinterpret_cast<uint4*>(pool + pool0);
        for (u32 i = threadIdx.x; i < nv; i += blockDim.x) gp[i] = sp[i];
        const uint4* s4 = reinterpret_cast<const uint4*>(s_raw);
        uint4* g4 = reinterpret_cast<uint4*>(table + (u64)b0 * S::kPerBucket);
        for (u32 i = threadIdx.x; i < nb * 2u; i += blockDim.x) g4[i] = s4[i];
    }
    if (threadIdx.x == 0) {
        if (s_inserted) atomicAdd(&ctr->n_inserted, (u64)s_inserted);
        if (s_dups) atomicAdd(&ctr->n_duplicates, (u64)s_dups);
        if (s_err) atomicOr(&ctr->errors, s_err);
    }
}

// ---- lookups in the sealed table (HBM) -------------------------------------------------------------------------
// key: ext and index bits 0.  Returns the stored slot value (with its index bits) or false.
template <int W>
__device__ __forceinline__ bool ct_lookup_chunk(const typename Slot<W>::value_t* __restrict__ table, const u32* __restrict__ chunk_base,
                                                u32 chunk, typename Slot<W>::value_t key, typename Slot<W>::value_t& found) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    const u32 b0 = chunk_base[chunk], nb = chunk_base[chunk + 1] - b0;
    if (nb == 0) return false;
    u32 b = ct_bucket_in_chunk(CtSlot<W>::hash32(key), nb);
    for (u32 tries = 0; tries < nb; ++tries) {
        u64 q[4];
        load256_nc(table + (u64)(b0 + b) * S::kPerBucket, q);
        bool hole = false, hit = false;
#pragma unroll
        for (int i = 0; i < S::kPerBucket; ++i) {                 // branch-free inside the bucket: every lane leaves together
            const V cur = S::from_bucket(q, i);
            const bool e = S::empty(cur);
            if (!hole && !hit && !e && CtSlot<W>::same_key(cur, key)) { found = cur; hit = true; }
            hole = hole || e;
        }
        if (hit) return true;
        if (hole) return false;
        b = (b + 1 == nb) ? 0u : b + 1;
    }
    return false;
}

// batch find (hash_map.hpp:83-92) on a sealed single-GPU chunk table
template <int W>
__global__ void __launch_bounds__(256)
ct_find_kernel(const typename Slot<W>::value_t* __restrict__ table, const u32* __restrict__ chunk_base, const CtGeom g,
               const unsigned char* __restrict__ pkmers, u64 n, unsigned char* __restrict__ pairs_out,
               unsigned char* __restrict__ found_out) {
    typedef Slot<W> S;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = g.k, pl = (k + 3) >> 2, pb = pl + 2;
    unsigned char key[16];
    for (int j = 0; j < pl; ++j) key[j] = pkmers[i * pl + j];
    const typename S::value_t kb = S::from_packed(key, k, pl);
    u32 owner, chunk;
    ct_place(ct_min_hash<W>(kb, g.m, g.win), g, owner, chunk);
    typename S::value_t hit;
    const bool ok = ct_lookup_chunk<W>(table, chunk_base, chunk, kb, hit);
    unsigned char rec[18];
    for (int j = 0; j < pb; ++j) rec[j] = 0;
    if (ok) S::to_record(CtSlot<W>::strip(hit), k, pl, rec);
    for (int j = 0; j < pb; ++j) pairs_out[i * pb + j] = rec[j];
    found_out[i] = ok ? 1 : 0;
}

// ---- contig head stubs: one zero-length segment per start node parsed by this rank ----------------------------
template <int W>
__global__ void __launch_bounds__(256)
ct_stub_kernel(const typename Slot<W>::value_t* __restrict__ starts, const Counters* __restrict__ ctr, u32 hcap,
               u64* __restrict__ link, u64* __restrict__ meta, typename Slot<W>::value_t* __restrict__ ext_key) {
    typedef Slot<W> S;
    const u32 n_starts = (u32)min(ctr->n_starts_dev, (u64)hcap);
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < hcap; c += gridDim.x * blockDim.x) {
        if (c < n_starts) {
            link[c] = (u64)kLinkPending << 32;
            meta[c] = 0;
            ext_key[c] = S::key_only(starts[c]);
        } else {
            link[c] = (u64)kLinkUnused << 32;
        }
    }
}

// resolve one pending link against THIS rank's table: the segment that starts with `key`
template <int W>
__device__ __forceinline__ u32 ct_resolve_one(const typename Slot<W>::value_t* __restrict__ table, const u32* __restrict__ chunk_base,
                                              const u32* __restrict__ seg_base, u32 chunk, typename Slot<W>::value_t key, u32 my_bits) {
    typename Slot<W>::value_t hit;
    if (!ct_lookup_chunk<W>(table, chunk_base, chunk, key, hit)) return kLinkMissing;      // kmer_hash.cpp:47-49, raised later if start-rooted
    const u32 idx = CtSlot<W>::idx(hit);
    if (idx == 0) return kLinkConverge;                // inside another segment: this k-mer has a second predecessor
    return my_bits | (seg_base[chunk] + idx - 1u);
}

// ---- chain the segments: ONE table lookup per segment -------------------------------------------------------------
// Pending links whose successor lives in this rank's chunks are resolved on the spot; the others are grouped by
// owner (tile-wise, so the peer stores are runs) straight into the owner's inbox.
constexpr int kResPerThread = 4;
constexpr int kResTile = 256 * kResPerThread;
template <int W>
__global__ void __launch_bounds__(256)
ct_resolve_kernel(const typename Slot<W>::value_t* __restrict__ table, const u32* __restrict__ chunk_base,
                  const u32* __restrict__ seg_base, u64* __restrict__ link, const typename Slot<W>::value_t* __restrict__ ext_key,
                  const CtGeom g, const CtPeers pe, const CtCaps caps, u32* __restrict__ out_cursor, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    __shared__ u32 s_cnt[kMaxRanks], s_base[kMaxRanks];
    const u32 nseg = min(ctr->next_seg, caps.seg_cap);
    const u32 my_bits = (u32)g.rank << kRankShift;
    const u32 ntiles = (nseg + kResTile - 1) / kResTile;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        CtReq<W> e[kResPerThread];
        u32 dest[kResPerThread], rk[kResPerThread];
#pragma unroll
        for (int r = 0; r < kResPerThread; ++r) {
            const u32 s = tile * kResTile + r * 256 + threadIdx.x;
            u32 d = 0xFFFFFFFFu;
            if (s < nseg && (u32)(link[s] >> 32) == kLinkPending) {
                const V key = ext_key[s];
                u32 owner, chunk;
                ct_place(ct_min_hash<W>(key, g.m, g.win), g, owner, chunk);
                if (owner == (u32)g.rank) {
                    reinterpret_cast<u32*>(link + s)[1] = ct_resolve_one<W>(table, chunk_base, seg_base, chunk, key, my_bits);
                } else {
                    d = owner;
                    e[r].key = key; e[r].src = my_bits | s; e[r].chunk = chunk;
                }
            }
            dest[r] = d;
            const unsigned same = __match_any_sync(kFullMask, d);
            const int leader = __ffs(same) - 1;
            u32 first = 0;
            if ((int)lane_id() == leader && d != 0xFFFFFFFFu) first = atomicAdd(&s_cnt[d], (u32)__popc(same));
            rk[r] = __shfl_sync(kFullMask, first, leader) + (u32)__popc(same & ((1u << lane_id()) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < kMaxRanks) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&out_cursor[threadIdx.x], s_cnt[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kResPerThread; ++r) {
            if (dest[r] == 0xFFFFFFFFu) continue;
            const u32 at = s_base[dest[r]] + rk[r];
            if (at < caps.inbox_cap) static_cast<CtReq<W>*>(pe.inbox[dest[r]])[(u64)g.rank * caps.inbox_cap + at] = e[r];
            else ct_internal(ctr, kSiteInbox);
        }
        __syncthreads();
    }
}

__global__ void ct_publish_inbox_kernel(const CtPeers pe, const CtCaps caps, const u32* __restrict__ out_cursor) {
    const int d = (int)threadIdx.x;
    if (d < pe.world) pe.inbox_cnt[d][pe.rank] = min(out_cursor[d], caps.inbox_cap);
}

// the owner answers: local lookup, then one peer store into the high word of the sender's link
template <int W>
__global__ void __launch_bounds__(256)
ct_answer_kernel(const typename Slot<W>::value_t* __restrict__ table, const u32* __restrict__ chunk_base,
                 const u32* __restrict__ seg_base, const CtReq<W>* __restrict__ inbox, const u32* __restrict__ inbox_cnt,
                 const CtGeom g, const CtPeers pe, const CtCaps caps) {
    const u32 my_bits = (u32)g.rank << kRankShift;
    for (int src = 0; src < g.world; ++src) {
        if (src == g.rank) continue;
        const u32 n = min(inbox_cnt[src], caps.inbox_cap);
        const CtReq<W>* __restrict__ box = inbox + (u64)src * caps.inbox_cap;
        for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const CtReq<W> e = box[i];
            const u32 hi = ct_resolve_one<W>(table, chunk_base, seg_base, e.chunk, e.key, my_bits);
            reinterpret_cast<u32*>(pe.link[e.src >> kRankShift] + (e.src & kLocalMask))[1] = hi;
        }
    }
}

// ---- pointer jumping over segment lists that span GPUs --------------------------------------------------------------
// A link is FINAL (top bit of the distance word) once its pointer is the chain's last segment (high word a marker);
// a link that is not final always moves when it is visited, so "nothing moved here" means this rank's links are final
// for good and it only keeps the others company at the barriers.  The rounds of all ranks run in lockstep (a barrier
// after each, ct_barrier_kernel): jumping over a peer's link doubles the distance only if the peer jumps too --
// ranks that ran their rounds on their own would walk a chain one remote segment per round.
__device__ __forceinline__ u64* ct_peer_link(const CtPeers& pe, u32 gid) { return pe.link[gid >> kRankShift] + (gid & kLocalMask); }
__device__ __forceinline__ bool ct_is_end_marker(u32 hi) { return hi == kLinkTail || hi == kLinkClaimed || hi == kLinkMissing || hi == kLinkConverge; }

__global__ void __launch_bounds__(256)
ct_rank_round_kernel(const CtPeers pe, u64* __restrict__ link, const CtCaps caps, Counters* ctr, u32* __restrict__ moved, int round) {
    if (ctr->rank_done || (round > 0 && moved[round - 1] == 0)) return;     // everybody / this rank has only final links
    const u32 nseg = min(ctr->next_seg, caps.seg_cap);
    bool any = false;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += (u64)gridDim.x * blockDim.x) {
        const u64 li = __ldcg(link + i);
        const u32 pi = (u32)(li >> 32), si = (u32)li;
        if (pi >= kLinkCtFirstMarker) continue;               // ends here / unused / still pending (reported by ct_lengths)
        if (si & kLinkFinalBit) continue;
        const u64 lp = __ldcg(ct_peer_link(pe, pi));
        const u32 pp = (u32)(lp >> 32), sp = (u32)lp;
        if (ct_is_end_marker(pp)) { __stcg(link + i, li | kLinkFinalBit); continue; }
        if (pp >= kLinkCtFirstMarker) continue;               // pending / unused target: left open, ct_lengths reports it
        __stcg(link + i, ((u64)pp << 32) | (u32)((si + (sp & kLinkDistMask)) | (sp & kLinkFinalBit)));
        any = true;
    }
    if (__any_sync(kFullMask, any) && lane_id() == 0) moved[round] = 1;
}

// contig lengths for the start nodes of this rank (kmer_hash.cpp:38-55: one contig per start node)
__global__ void __launch_bounds__(256)
ct_lengths_kernel(const CtPeers pe, const u64* __restrict__ link, const CtCaps caps, int k,
                  u32* __restrict__ contig_len, u32* __restrict__ contig_pre, Counters* ctr) {
    const u32 n_starts = (u32)min(ctr->n_starts_dev, (u64)caps.hcap);
    u64 nodes = 0;
    u32 err = 0;
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c <= caps.hcap; c += (u64)gridDim.x * blockDim.x) {
        if (c >= n_starts) { contig_len[c] = 0; continue; }           // the offsets scan runs over hcap + 1 entries
        const u64 lc = __ldcg(link + c);
        const u32 pc = (u32)(lc >> 32);
        u32 e = 0, tail_hi = 0, chars = 0, pre = 0;
        if (pc == kLinkMissing) e = kErrNotFound;
        else if (pc == kLinkConverge) e = kErrConverge;
        else if (pc >= kLinkCtFirstMarker) { e = kErrInternal; atomicOr(&ctr->err_where, kSiteStubOpen); }                     // a stub is never a tail, and nobody may leave it pending
        else if (!((u32)lc & kLinkFinalBit)) e = kErrCycle;                      // still moving after the last round: kmer_hash.cpp:44 never exits
        else {
            tail_hi = (u32)(__ldcg(ct_peer_link(pe, pc)) >> 32);
            if (tail_hi == kLinkMissing) e = kErrNotFound;                       // kmer_hash.cpp:47-49
            else if (tail_hi == kLinkConverge) e = kErrConverge;
            else {
                pre = (u32)lc & kLinkDistMask;
                chars = pre + (u32)(pe.meta[pc >> kRankShift][pc & kLocalMask] & 0xFFFFFFu);
            }
        }
        if (e) { err |= e; contig_pre[c] = 0; contig_len[c] = 0; continue; }
        contig_pre[c] = pre;
        contig_len[c] = (u32)k + chars + 1u;
        nodes += (u64)chars + 1u;
    }
    nodes = warp_sum_u64(nodes);
    err = __reduce_or_sync(kFullMask, err);
    if (lane_id() == 0) {
        if (nodes) atomicAdd(&ctr->n_nodes, nodes);
        if (err) atomicOr(&ctr->errors, err);
    }
}

// each contig claims its last segment (possibly on another GPU): link[tail] = (CLAIMED, home rank << 28 | contig)
__global__ void __launch_bounds__(256)
ct_claim_kernel(const CtPeers pe, const u64* __restrict__ link, const CtCaps caps, const u32* __restrict__ contig_len, Counters* ctr) {
    const u32 n_starts = (u32)min(ctr->n_starts_dev, (u64)caps.hcap);
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < n_starts; c += (u64)gridDim.x * blockDim.x) {
        if (contig_len[c] == 0) continue;
        const u32 tail = (u32)(__ldcg(link + c) >> 32);
        const u64 want = (u64)kLinkTail << 32;
        const u64 old = atomicCAS(ct_peer_link(pe, tail), want, ((u64)kLinkClaimed << 32) | (((u32)pe.rank << kRankShift) | (u32)c));
        if (old != want) atomicOr(&ctr->errors, kErrConverge);       // two start nodes, one end node
    }
}

// ---- emit: every GPU copies the characters of its segments into the output of the contig's home rank ---------------
// One lane resolves a segment's destination (its last segment's claim -> contig -> offset), then the warp copies.
__global__ void __launch_bounds__(256)
ct_emit_kernel(const CtPeers pe, const u64* __restrict__ link, const u64* __restrict__ meta, const unsigned char* __restrict__ pool,
               const CtCaps caps, Counters* ctr, int k) {
    const u32 nseg = min(ctr->next_seg, caps.seg_cap);
    if (ctr->errors & (kErrConverge | kErrCycle | kErrInternal | kErrNotFound)) return;       // the host reports it; write nothing
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 seg0 = (u64)caps.hcap + (u64)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); seg0 < nseg; seg0 += stride) {   // stubs carry no characters
        const u64 seg = seg0 + lane_id();
        u32 len = 0;
        char* dst = nullptr;
        const unsigned char* src = nullptr;
        if (seg < nseg) {
            const u64 li = link[seg];
            const u32 pi = (u32)(li >> 32);
            const u64 mt = meta[seg];
            u32 cg = 0, dist = 0;
            bool ok = false, is_tail = false;
            if (pi == kLinkClaimed) { cg = (u32)li; ok = true; is_tail = true; }
            else if (pi < kLinkCtFirstMarker && ((u32)li & kLinkFinalBit)) {
                const u64 lt = *ct_peer_link(pe, pi);
                if ((u32)(lt >> 32) == kLinkClaimed) { cg = (u32)lt; dist = (u32)li & kLinkDistMask; ok = true; }
            }
            if (ok) {                                          // otherwise: on no start-rooted chain, ignored like the reference
                const u32 r = cg >> kRankShift, c = cg & kLocalMask;
                const u32 pre = pe.contig_pre[r][c];
                if (!is_tail && dist > pre) { atomicOr(&ctr->errors, kErrConverge); ok = false; }
                if (ok) {
                    const u64 off = pe.contig_off[r][c] + (u64)k + (is_tail ? pre : pre - dist);
                    len = (u32)(mt & 0xFFFFFFu);
                    if (off + len > pe.out_cap[r]) { ct_internal(ctr, kSiteOutCap); len = 0; }
                    dst = pe.out[r] + off;
                    src = pool + (mt >> 24);
                }
            }
        }
        u32 todo = __ballot_sync(kFullMask, len > 0);
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            const u32 n = __shfl_sync(kFullMask, len, sl);
            char* d = reinterpret_cast<char*>(__shfl_sync(kFullMask, reinterpret_cast<u64>(dst), sl));
            const unsigned char* s = reinterpret_cast<const unsigned char*>(__shfl_sync(kFullMask, reinterpret_cast<u64>(src), sl));
            for (u32 j = lane_id(); j < n; j += 32) d[j] = (char)s[j];
        }
    }
}

// first k-mer of each contig + '\n' (read_kmers.hpp:84, kmer_hash.cpp:66); n_starts read on the device
template <int W>
__global__ void __launch_bounds__(256)
ct_emit_heads_kernel(const typename Slot<W>::value_t* __restrict__ starts, const CtCaps caps, int k,
                     const u32* __restrict__ contig_len, const u64* __restrict__ contig_off,
                     const Counters* __restrict__ ctr, u64 out_cap, char* __restrict__ out) {
    typedef Slot<W> S;
    if (ctr->errors & (kErrConverge | kErrCycle | kErrInternal | kErrNotFound)) return;
    const u32 n_starts = (u32)min(ctr->n_starts_dev, (u64)caps.hcap);
    const u64 total = (u64)n_starts * (u32)(k + 1);
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x) {
        const u64 c = t / (u32)(k + 1);
        const u32 j = (u32)(t - c * (u32)(k + 1));
        const u32 len = contig_len[c];
        if (len == 0) continue;
        const u64 off = contig_off[c];
        if (off + len > out_cap) continue;
        if (j < (u32)k) out[off + j] = (char)ext_char(S::base_at(starts[c], k, (int)j));
        else out[off + len - 1] = '\n';
    }
}

// start list with the write position read from the device (ct path: no host round trip per insert call)
template <int W>
__global__ void __launch_bounds__(256)
ct_scatter_starts_kernel(const unsigned char* __restrict__ recs, u64 n, int k, const u32* __restrict__ start_mask,
                         const u64* __restrict__ tile_offsets, u64 ntiles, typename Slot<W>::value_t* __restrict__ starts_out,
                         const Counters* __restrict__ ctr, u32 hcap) {
    typedef Slot<W> S;
    const u64 tile = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    const int pl = (k + 3) >> 2, pb = pl + 2;
    const u64 word = tile * (kInsTile / 32) + lane_id();
    const u64 nwords = (n + 31) >> 5;
    u32 m = word < nwords ? start_mask[word] : 0u;
    u32 inc = __popc(m);
    const u32 mine = inc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 y = __shfl_up_sync(kFullMask, inc, d);
        if (lane_id() >= (u32)d) inc += y;
    }
    u64 dst = ctr->n_starts_dev + tile_offsets[tile] + (inc - mine);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const u64 rec = (word << 5) + bit;
        bool ok;
        unsigned char tmp[18];
        const unsigned char* src = recs + rec * pb;
        for (int i = 0; i < pb; ++i) tmp[i] = src[i];
        if (dst < hcap) starts_out[dst] = S::from_record(tmp, k, pl, ok);
        ++dst;
    }
}

}  // namespace kh
