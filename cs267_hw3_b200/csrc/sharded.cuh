// sharded.cuh -- kernels of the multi-GPU (hash-sharded) path.
//
// Replaces the UPC++ side of the reference: owner = hash(key) % ranks
// (hash_map.hpp:28-30), one batched RPC per destination for inserts (hash_map.hpp:38-46, 64-77)
// and one blocking RPC round trip per remote lookup (hash_map.hpp:93-100).
//
//   K7 owner_count / owner_scatter   records -> slot values grouped by owning GPU (+ start bitmask);
//                                    the groups travel with one NCCL all-to-all (torch.distributed)
//   K2 insert_slots_direct           insert received slot values into the local shard
//   K8 walk_sharded                  like walk_kernel, but a successor lookup reads the OWNER's table
//                                    directly through its NVLink peer mapping: the one-sided analogue of
//                                    the reference's find() RPC, with ~300k lookups in flight per GPU
//                                    instead of one blocking round trip per rank
//      rank_round / lengths / claim  pointer jumping over segment lists that live on different GPUs
//      emit_*_sharded                each GPU copies its segments' characters into the output buffer
//                                    of the GPU that owns the contig's start node (peer stores)
//
// Global segment id = (rank << kRankShift) | local id.  Local ids: [0, n_split) splitters of the
// local shard, [n_split, n_split + n_starts) start nodes parsed by this rank, then overflow.
#pragma once
#include "kernels.cuh"

namespace kh {

constexpr int kMaxRanks = 8;
constexpr u32 kRankShift = 28;
constexpr u32 kLocalMask = (1u << kRankShift) - 1u;
// pointer jumping across GPUs: a link whose pointer already is its contig's tail is flagged in the top bit of the
// distance word, so later rounds do not pay a remote read for it
constexpr u32 kLinkFinalBit = 0x80000000u;
constexpr u32 kLinkDistMask = 0x7FFFFFFFu;

struct Peers {
    const void* table[kMaxRanks];
    u64 nbuckets[kMaxRanks];
    u64* link[kMaxRanks];
    unsigned char* seglen[kMaxRanks];
    u32* contig_pre[kMaxRanks];
    u64* contig_off[kMaxRanks];
    char* out[kMaxRanks];
    u64 out_cap[kMaxRanks];
    int world, rank;
};

// ---- K7: group records by owner ---------------------------------------------------------
// pass 1: per-owner counts + start bitmask / per-tile start counts (same tiles as insert_kernel).  The owner of
// every record is kept (one byte) so that pass 2 does not repeat the minimizer scan.
template <int W>
__global__ void __launch_bounds__(kInsThreads)
owner_count_kernel(const unsigned char* __restrict__ recs, u64 n, int k, int mo, int world,
                   u32* __restrict__ start_mask, u32* __restrict__ tile_starts, u64* __restrict__ owner_counts,
                   unsigned char* __restrict__ owner_byte, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    extern __shared__ __align__(16) unsigned char s_rec[];
    __shared__ u32 s_cnt[kMaxRanks], s_starts, s_err;
    const int pl = (k + 3) >> 2, pb = pl + 2;
    const u64 rec0 = (u64)blockIdx.x * kInsTile;
    const u32 cnt = (u32)min((u64)kInsTile, n - rec0);
    if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_starts = 0; s_err = 0; }
    stage_in(s_rec, recs + rec0 * pb, cnt * pb);
    __syncthreads();
    u32 err = 0, starts = 0;
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        const u32 j = threadIdx.x + r * kInsThreads;
        bool live = j < cnt, ok = true;
        V v = S::zero();
        if (live) v = S::from_record_staged(s_rec + j * pb, k, pl, ok);
        if (!ok) { err |= kErrBadInput; live = false; }
        const u32 o = live ? owner_of<W>(v, world, k, mo) : 0u;
        if (j < cnt) owner_byte[rec0 + j] = live ? (unsigned char)o : (unsigned char)0xFF;
        // warp-aggregate per owner (world <= 8)
        for (int w = 0; w < world; ++w) {
            const u32 bo = __ballot_sync(kFullMask, live && o == (u32)w);
            if (lane_id() == 0 && bo) atomicAdd(&s_cnt[w], (u32)__popc(bo));
        }
        const u32 bal = __ballot_sync(kFullMask, live && S::back(v) == kExtF);
        const u64 word = (rec0 + (u64)r * kInsThreads + (threadIdx.x & ~31u)) >> 5;
        if (lane_id() == 0 && (word << 5) < n) { start_mask[word] = bal; starts += __popc(bal); }
    }
    err = __reduce_or_sync(kFullMask, err);
    if (lane_id() == 0) {
        if (starts) atomicAdd(&s_starts, starts);
        if (err) atomicOr(&s_err, err);
    }
    __syncthreads();
    if (threadIdx.x < (u32)world && s_cnt[threadIdx.x]) atomicAdd(&owner_counts[threadIdx.x], (u64)s_cnt[threadIdx.x]);
    if (threadIdx.x == 0) {
        tile_starts[blockIdx.x] = s_starts;
        if (s_err) atomicOr(&ctr->errors, s_err);
    }
}

// pass 2: scatter slot values to their owner's group (owner_base = exclusive prefix of the counts)
template <int W>
__global__ void __launch_bounds__(kInsThreads)
owner_scatter_kernel(const unsigned char* __restrict__ recs, u64 n, int k, const unsigned char* __restrict__ owner_byte, int world,
                     const u64* __restrict__ owner_base, u64* __restrict__ owner_cursor,
                     typename Slot<W>::value_t* __restrict__ grouped) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    extern __shared__ __align__(16) unsigned char s_rec[];
    __shared__ u32 s_cnt[kMaxRanks];
    __shared__ u64 s_base[kMaxRanks];
    const int pl = (k + 3) >> 2, pb = pl + 2;
    const u64 rec0 = (u64)blockIdx.x * kInsTile;
    const u32 cnt = (u32)min((u64)kInsTile, n - rec0);
    if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
    stage_in(s_rec, recs + rec0 * pb, cnt * pb);
    __syncthreads();
    V v[kInsPerThread];
    u32 own[kInsPerThread], rk[kInsPerThread];
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        const u32 j = threadIdx.x + r * kInsThreads;
        bool ok = true;
        own[r] = 0xFFFFFFFFu;
        if (j < cnt) {
            const u32 o = owner_byte[rec0 + j];
            if (o < (u32)world) {
                v[r] = S::from_record_staged(s_rec + j * pb, k, pl, ok);
                own[r] = o;
                rk[r] = atomicAdd(&s_cnt[o], 1u);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < (u32)world)
        s_base[threadIdx.x] = owner_base[threadIdx.x] +
                              (s_cnt[threadIdx.x] ? atomicAdd(&owner_cursor[threadIdx.x], (u64)s_cnt[threadIdx.x]) : 0ull);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r)
        if (own[r] != 0xFFFFFFFFu) grouped[s_base[own[r]] + rk[r]] = v[r];
}

// ---- K2 on received slot values ------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kInsThreads)
insert_slots_direct_kernel(const typename Slot<W>::value_t* __restrict__ slots, u64 n, int k, int m,
                           typename Slot<W>::value_t* table, u64 nbuckets, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    __shared__ u32 s_inserted, s_dups, s_err;
    if (threadIdx.x == 0) { s_inserted = 0; s_dups = 0; s_err = 0; }
    __syncthreads();
    const u64 base = (u64)blockIdx.x * kInsTile;
    V v[kInsPerThread];
    u64 b[kInsPerThread];
    u64 q[kInsPerThread][4];
    bool live[kInsPerThread];
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        const u64 i = base + (u64)r * kInsThreads + threadIdx.x;
        live[r] = i < n;
        v[r] = live[r] ? slots[i] : S::zero();
        live[r] = live[r] && !S::empty(v[r]);
        b[r] = live[r] ? place_bucket<W>(v[r], k, m, nbuckets) : 0;
        if (live[r]) load256_cg(table + b[r] * S::kPerBucket, q[r]);
    }
    u32 inserted = 0, dups = 0, err = 0;
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        if (!live[r]) continue;
        const int rc = insert_one<W>(table, nbuckets, b[r], v[r], q[r]);
        inserted += (rc == kInsInserted);
        dups += (rc == kInsDuplicate);
        if (rc == kInsFull) err |= kErrTableFull;
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
    if (threadIdx.x == 0) {
        if (s_inserted) atomicAdd(&ctr->n_inserted, (u64)s_inserted);
        if (s_dups) atomicAdd(&ctr->n_duplicates, (u64)s_dups);
        if (s_err) atomicOr(&ctr->errors, s_err);
    }
}

// ---- K8: walk with peer lookups -----------------------------------------------------------------
struct ShardWalkParams {
    Peers peers;
    const void* starts;
    u64* link;                 // local segment arrays
    unsigned char* seglen;
    unsigned char* tmp;
    Counters* ctr;
    u32 n_starts, n_split;     // local
    u32 split_shift, seg_chars, seg_cap;
    int k, m, mo;          // m: minimizer length of the in-shard placement (0 = key hash); mo: of the owner function
};

template <int W>
__device__ __forceinline__ bool lookup_sharded(const Peers& pe, int k, int m, int mo, typename Slot<W>::value_t keybits,
                                               typename Slot<W>::value_t& found, u32& owner, u64& bucket, int& slot) {
    typedef Slot<W> S;
    // the owner is a hash of the minimizer: the successor of a k-mer usually shares it, so most steps stay on-GPU
    const u64 oh = mo ? fmix64(minimizer_value<W>(keybits, k, mo) + 0x632BE59BD9B4E019ull) : S::owner_hash(keybits);
    owner = (u32)__umul64hi(oh, (u64)pe.world);
    const typename S::value_t* __restrict__ table = static_cast<const typename S::value_t*>(pe.table[owner]);
    const u64 nb = pe.nbuckets[owner];
    u64 b = (m == 0) ? bucket_of(S::hash(keybits), nb)
                     : place_bucket_from<W>(m == mo ? oh : fmix64(minimizer_value<W>(keybits, k, m) + 0x632BE59BD9B4E019ull), keybits, nb);
    for (u64 tries = 0; tries < nb; ++tries) {
        u64 q[4];
        load256_ro(table + b * S::kPerBucket, q);          // local HBM or a peer's HBM over NVLink
#pragma unroll
        for (int i = 0; i < S::kPerBucket; ++i) {
            const typename S::value_t cur = S::from_bucket(q, i);
            if (S::empty(cur)) return false;
            if (S::same_key(cur, keybits)) { found = cur; bucket = b; slot = i; return true; }
        }
        b = (b + 1 == nb) ? 0 : b + 1;
    }
    return false;
}

template <int W>
__global__ void __launch_bounds__(kWalkThreads)
walk_sharded_kernel(const ShardWalkParams p) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    const V* __restrict__ table = static_cast<const V*>(p.peers.table[p.peers.rank]);
    const V* __restrict__ starts = static_cast<const V*>(p.starts);
    const u32 total = p.n_split + p.n_starts;
    const u32 lane = lane_id();
    const u32 lt_mask = (1u << lane) - 1u;
    const u64 split_mask = (1ull << p.split_shift) - 1ull;
    const u32 my_rank_bits = (u32)p.peers.rank << kRankShift;

    u32 w_next = 0, w_end = 0;
    bool exhausted = false;
    u32 o_next = 0, o_end = 0;
    bool active = false;
    V cur = S::zero(), chk = S::zero();
    u32 seg = 0, n = 0, steps = 0, limit = 256;
    u64 acc = 0;

    auto close = [&](u32 next) {       // next: global id or kLinkTail
        if (n & 7u) *reinterpret_cast<u64*>(p.tmp + (u64)seg * p.seg_chars + (n & ~7u)) = acc;
        p.seglen[seg] = (unsigned char)n;
        p.link[seg] = ((u64)next << 32) | (next == kLinkTail ? 0u : n);
        active = false;
    };

    for (;;) {
        __syncwarp();
        const u32 idle = __ballot_sync(kFullMask, !active);
        if (idle && !exhausted) {
            if (w_next == w_end) {
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->next_walker, kWalkBatch);
                base = __shfl_sync(kFullMask, base, 0);
                w_next = min(base, total);
                w_end = min(base + kWalkBatch, total);
                exhausted = (w_next == w_end);
            }
            const u32 avail = w_end - w_next;
            const u32 rank = __popc(idle & lt_mask);
            if (!active && rank < avail) {
                const u32 w = w_next + rank;
                seg = w; n = 0; acc = 0; steps = 0; limit = 256;
                if (w >= p.n_split) {
                    cur = starts[w - p.n_split];
                    active = true;
                } else {
                    const u64 b = (u64)w << p.split_shift;
                    cur = S::load_one_nc(table + b * S::kPerBucket);
                    if (S::empty(cur)) p.link[seg] = (u64)kLinkUnused << 32;
                    else active = true;
                }
                chk = cur;
            }
            w_next += min((u32)__popc(idle), avail);
        }
        if (__ballot_sync(kFullMask, active) == 0) {
            if (exhausted) break;
            continue;
        }
        const u32 full = __ballot_sync(kFullMask, active && n == p.seg_chars && S::fwd(cur) != kExtF);
        if (full) {
            const u32 want = __popc(full);
            if (o_end - o_next < want) {
                for (u32 id = o_next + lane; id < o_end; id += 32) p.link[id] = (u64)kLinkUnused << 32;
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->next_seg, kSegBatch);
                base = __shfl_sync(kFullMask, base, 0);
                o_next = base; o_end = base + kSegBatch;
            }
            if (full & (1u << lane)) {
                const u32 ns = o_next + __popc(full & lt_mask);
                if (ns >= p.seg_cap) {
                    atomicOr(&p.ctr->errors, kErrInternal);
                    close(kLinkTail);
                } else {
                    close(my_rank_bits | ns);
                    active = true; seg = ns; n = 0; acc = 0;
                }
            }
            o_next += want;
        }
        if (active) {
            const u32 f = S::fwd(cur);
            if (f == kExtF) {
                close(kLinkTail);
            } else {
                acc |= (u64)ext_char(f) << (8u * (n & 7u));
                ++n;
                if ((n & 7u) == 0) {
                    *reinterpret_cast<u64*>(p.tmp + (u64)seg * p.seg_chars + n - 8) = acc;
                    acc = 0;
                }
                V nxt; u64 b; int s; u32 owner;
                if (!lookup_sharded<W>(p.peers, p.k, p.m, p.mo, S::next_key(cur, p.k), nxt, owner, b, s)) {
                    atomicOr(&p.ctr->errors, kErrNotFound);
                    close(kLinkTail);
                } else if (s == 0 && (b & split_mask) == 0) {
                    close((owner << kRankShift) | (u32)(b >> p.split_shift));   // the owner's splitter walker takes over
                } else {
                    cur = nxt;
                    if (S::same_key(cur, chk)) {
                        atomicOr(&p.ctr->errors, kErrCycle);
                        close(kLinkTail);
                    } else if (++steps == limit) {
                        chk = cur; steps = 0; limit <<= 1;
                    }
                }
            }
        }
    }
    for (u32 id = o_next + lane; id < o_end; id += 32)
        if (id < p.seg_cap) p.link[id] = (u64)kLinkUnused << 32;
}

// ---- pointer jumping across GPUs -------------------------------------------------------------
__device__ __forceinline__ u64* peer_link(const Peers& pe, u32 gid) { return pe.link[gid >> kRankShift] + (gid & kLocalMask); }

// one round over the local segments; *changed = 1 if any local link moved
__global__ void __launch_bounds__(256)
rank_round_sharded_kernel(const Peers pe, u64* link, u32 seg_cap, Counters* ctr, u32* changed) {
    const u32 nseg = min(ctr->next_seg, seg_cap);
    bool moved = false;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += (u64)gridDim.x * blockDim.x) {
        const u64 li = __ldcg(link + i);
        const u32 pi = (u32)(li >> 32);
        if (pi >= kLinkClaimed) continue;
        const u64 lp = __ldcg(peer_link(pe, pi));
        const u32 pp = (u32)(lp >> 32);
        if (pp == kLinkTail) continue;
        if (pp >= kLinkClaimed) { atomicOr(&ctr->errors, kErrInternal); continue; }
        __stcg(link + i, ((u64)pp << 32) | (u32)((u32)li + (u32)lp));
        moved = true;
    }
    if (__any_sync(kFullMask, moved) && lane_id() == 0) *changed = 1;
}

// contig lengths for the local start nodes (local ids n_split + c)
__global__ void __launch_bounds__(256)
contig_lengths_sharded_kernel(const Peers pe, const u64* link, u32 n_split, u32 n_starts, int k,
                              u32* contig_len, u32* contig_pre, Counters* ctr) {
    u64 nodes = 0;
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < n_starts; c += (u64)gridDim.x * blockDim.x) {
        const u64 lc = __ldcg(link + n_split + c);
        const u32 pc = (u32)(lc >> 32);
        u32 tail_gid = ((u32)pe.rank << kRankShift) | (u32)(n_split + c), pre = 0;
        bool ok = true;
        if (pc != kLinkTail && pc != kLinkClaimed) {
            tail_gid = pc; pre = (u32)lc & kLinkDistMask;
            const u32 tp = pc < kLinkFirstMarker ? (u32)(__ldcg(peer_link(pe, pc)) >> 32) : 0u;
            ok = tp == kLinkTail || tp == kLinkClaimed;       // another rank may already have claimed its own tails
        }
        if (!ok) {
            atomicOr(&ctr->errors, kErrCycle);
            contig_pre[c] = 0; contig_len[c] = 0;
            continue;
        }
        const u32 chars = pre + pe.seglen[tail_gid >> kRankShift][tail_gid & kLocalMask];
        contig_pre[c] = pre;
        contig_len[c] = (u32)k + chars + 1u;
        nodes += (u64)chars + 1u;
    }
    nodes = warp_sum_u64(nodes);
    if (lane_id() == 0 && nodes) atomicAdd(&ctr->n_nodes, nodes);
}

// each contig claims its tail (possibly on another GPU): link[tail] = (CLAIMED, rank<<28 | contig)
__global__ void __launch_bounds__(256)
claim_tails_sharded_kernel(const Peers pe, const u64* link, u32 n_split, u32 n_starts, const u32* contig_len, Counters* ctr) {
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < n_starts; c += (u64)gridDim.x * blockDim.x) {
        if (contig_len[c] == 0) continue;
        const u32 pc = (u32)(__ldcg(link + n_split + c) >> 32);
        const u32 tail_gid = (pc == kLinkTail || pc == kLinkClaimed) ? (((u32)pe.rank << kRankShift) | (u32)(n_split + c)) : pc;
        const u64 want = (u64)kLinkTail << 32;
        const u64 old = atomicCAS(peer_link(pe, tail_gid), want,
                                  ((u64)kLinkClaimed << 32) | (((u32)pe.rank << kRankShift) | (u32)c));
        if (old != want) atomicOr(&ctr->errors, kErrConverge);
    }
}

// copy the local segments' characters into the output buffer of the GPU that owns the contig
__global__ void __launch_bounds__(256)
emit_segments_sharded_kernel(const Peers pe, const u64* __restrict__ link, const unsigned char* __restrict__ seglen,
                             const unsigned char* __restrict__ tmp, u32 seg_chars, u32 seg_cap, Counters* ctr, int k) {
    const u64 seg = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 nseg = min(ctr->next_seg, seg_cap);
    u32 len = 0;
    char* dst = nullptr;
    if (seg < nseg) {
        const u64 li = link[seg];
        const u32 pi = (u32)(li >> 32);
        if (pi != kLinkUnused && pi != kLinkTail) {
            bool ok = true, is_tail = (pi == kLinkClaimed);
            u32 cg = 0;
            if (is_tail) {
                cg = (u32)li;
            } else {
                const u64 lt = *peer_link(pe, pi);
                if ((u32)(lt >> 32) != kLinkClaimed) ok = false;
                else cg = (u32)lt;
            }
            if (ok) {
                const u32 r = cg >> kRankShift, c = cg & kLocalMask;
                const u32 pre = pe.contig_pre[r][c];
                u32 pos = pre;
                if (!is_tail) {
                    const u32 dist = (u32)li & kLinkDistMask;
                    if (dist > pre) { atomicOr(&ctr->errors, kErrConverge); ok = false; }
                    else pos = pre - dist;
                }
                if (ok) {
                    const u64 off = pe.contig_off[r][c] + (u64)k + pos;
                    len = seglen[seg];
                    if (off + len > pe.out_cap[r]) { atomicOr(&ctr->errors, kErrInternal); len = 0; }
                    dst = pe.out[r] + off;
                }
            }
        }
    }
    u32 todo = __ballot_sync(kFullMask, len > 0);
    const u64 seg0 = seg - lane_id();
    while (todo) {
        const int src_lane = __ffs(todo) - 1;
        todo &= todo - 1;
        const u32 n = __shfl_sync(kFullMask, len, src_lane);
        char* d = reinterpret_cast<char*>(__shfl_sync(kFullMask, reinterpret_cast<u64>(dst), src_lane));
        const unsigned char* src = tmp + (seg0 + src_lane) * (u64)seg_chars;
        for (u32 j = lane_id(); j < n; j += 32) d[j] = (char)src[j];
    }
}


// =========================================================================================
// Migrating walk: every GPU walks only the nodes it owns
// =========================================================================================
// With peer lookups (walk_sharded_kernel above) a walker stays on the GPU where it started, so after its
// first supermer (P-1)/P of its lookups cross NVLink no matter how owners are chosen; measured on
// B200s the walk gets SLOWER with every GPU added (3.4 ms at 2 GPUs, 7.9 ms at 4, ~18 ms at 8 for the
// chr14 k=19 shape).  Here the walk follows the data instead:
//   * a node whose predecessor lives on another GPU (or that has backward ext 'F') is a "boundary
//     start": its owner registers it as a walker at insert time (the predecessor key follows from
//     the backward extension, kmer_pair::last_kmer, kmer_t.hpp:55-57);
//   * a walker follows successors only while they are local; when the successor belongs to another
//     GPU it closes its segment with a PENDING link and drops (successor key, own segment id) into an
//     outbox;
//   * one all-to-all delivers the outboxes; the owner looks each key up locally, finds the boundary
//     walker that started there and writes its segment id into the sender's link with one peer store.
// With the minimizer-keyed owner function segments break only where the minimizer changes AND the new
// owner differs, so the number of cross-GPU links is ~N/((K-m+2)/2) * (P-1)/P, each 16-32 bytes.
// The start nodes parsed by a rank become zero-length "head stubs" that link to their first k-mer's
// walker the same way, so contigs stay with the rank that parsed their start line.
//
// Local segment ids:  [0, n_split) splitters | [n_split, n_split+bcap) boundary starts (n_boundary used)
//                     | [n_split+bcap, walk_cap) overflow | [walk_cap, walk_cap+n_starts) head stubs
struct MigLayout {
    u32 n_split, bcap, walk_cap, hcap;
    u32 outbox_cap;
};

constexpr u32 kOutBatch = 256;          // outbox entries a warp reserves per global atomic
constexpr u32 kOutVoid = 0xFFu;          // dest of a reserved-but-unused outbox entry
constexpr u32 kOutUnrouted = 0xFEu;      // dest not computed yet: the grouping pass derives it from the key
template <int W> struct OutEntry;
template <> struct alignas(16) OutEntry<1> { u64 key; u32 src; u32 dest; };
template <> struct alignas(16) OutEntry<2> { u128 key; u32 src; u32 dest; u64 pad; };

// ---- K2 on received slot values, registering boundary starts ---------------------------------------
template <int W>
__global__ void __launch_bounds__(kInsThreads)
insert_slots_shard_kernel(const typename Slot<W>::value_t* __restrict__ slots, u64 n, int k, int m, int mo,
                          int rank, int world, typename Slot<W>::value_t* table, u64 nbuckets,
                          u32* __restrict__ seg_of_slot, typename Slot<W>::value_t* __restrict__ boundary_list,
                          u32 bcap, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    __shared__ u32 s_inserted, s_dups, s_err;
    if (threadIdx.x == 0) { s_inserted = 0; s_dups = 0; s_err = 0; }
    __syncthreads();
    const u64 base = (u64)blockIdx.x * kInsTile;
    V v[kInsPerThread];
    u64 b[kInsPerThread], pos[kInsPerThread];
    u64 q[kInsPerThread][4];
    bool live[kInsPerThread], fresh[kInsPerThread];
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        const u64 i = base + (u64)r * kInsThreads + threadIdx.x;
        live[r] = i < n;
        v[r] = live[r] ? slots[i] : S::zero();
        live[r] = live[r] && !S::empty(v[r]);
        b[r] = live[r] ? place_bucket<W>(v[r], k, m, nbuckets) : 0;
        if (live[r]) load256_cg(table + b[r] * S::kPerBucket, q[r]);
    }
    u32 inserted = 0, dups = 0, err = 0;
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        fresh[r] = false; pos[r] = 0;
        if (!live[r]) continue;
        const int rc = insert_one<W>(table, nbuckets, b[r], v[r], q[r], &pos[r]);
        fresh[r] = (rc == kInsInserted);
        inserted += fresh[r];
        dups += (rc == kInsDuplicate);
        if (rc == kInsFull) err |= kErrTableFull;
    }
    // boundary starts: backward ext 'F', or the predecessor (backward ext + first K-1 bases) lives elsewhere
#pragma unroll
    for (int r = 0; r < kInsPerThread; ++r) {
        bool bnd = false;
        if (fresh[r]) {
            bnd = S::back(v[r]) == kExtF;
            if (!bnd && world > 1) bnd = owner_of<W>(S::prev_key(v[r], k), world, k, mo) != (u32)rank;
        }
        const u32 bal = __ballot_sync(kFullMask, bnd);
        if (bal) {
            u32 first = 0;
            if (lane_id() == 0) first = atomicAdd(&ctr->n_boundary, (u32)__popc(bal));
            first = __shfl_sync(kFullMask, first, 0);
            if (bnd) {
                const u32 id = first + __popc(bal & ((1u << lane_id()) - 1u));
                if (id < bcap) { boundary_list[id] = v[r]; seg_of_slot[pos[r]] = id; }
                else err |= kErrInternal;
            }
        }
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
    if (threadIdx.x == 0) {
        if (s_inserted) atomicAdd(&ctr->n_inserted, (u64)s_inserted);
        if (s_dups) atomicAdd(&ctr->n_duplicates, (u64)s_dups);
        if (s_err) atomicOr(&ctr->errors, s_err);
    }
}

// the few values the chunked owner-side build could not place (see build_chunks_kernel): ordinary global insert,
// after every chunk is final, with the same boundary registration as insert_slots_shard_kernel
template <int W>
__global__ void __launch_bounds__(256)
insert_overflow_shard_kernel(const typename Slot<W>::value_t* __restrict__ overflow, u32 overflow_cap, int k, int m, int mo,
                             int rank, int world, typename Slot<W>::value_t* table, u64 nbuckets,
                             u32* __restrict__ seg_of_slot, typename Slot<W>::value_t* __restrict__ boundary_list,
                             u32 bcap, Counters* ctr) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    const u32 n = min(ctr->n_outbox, overflow_cap);
    for (u32 base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // warp-uniform trip count
        const u32 i = base + threadIdx.x;
        bool bnd = false;
        V v = S::zero();
        u64 pos = 0;
        if (i < n) {
            v = overflow[i];
            const u64 b = place_bucket<W>(v, k, m, nbuckets);
            u64 q[4];
            load256_cg(table + b * S::kPerBucket, q);
            const int rc = insert_one<W>(table, nbuckets, b, v, q, &pos);
            if (rc == kInsInserted) {
                atomicAdd(&ctr->n_inserted, 1ull);
                bnd = S::back(v) == kExtF;
                if (!bnd && world > 1) bnd = owner_of<W>(S::prev_key(v, k), world, k, mo) != (u32)rank;
            } else if (rc == kInsDuplicate) {
                atomicAdd(&ctr->n_duplicates, 1ull);
            } else {
                atomicOr(&ctr->errors, kErrTableFull);
            }
        }
        const u32 bal = __ballot_sync(kFullMask, bnd);
        if (bal) {
            u32 first = 0;
            if (lane_id() == 0) first = atomicAdd(&ctr->n_boundary, (u32)__popc(bal));
            first = __shfl_sync(kFullMask, first, 0);
            if (bnd) {
                const u32 id = first + __popc(bal & ((1u << lane_id()) - 1u));
                if (id < bcap) { boundary_list[id] = v; seg_of_slot[pos] = id; }
                else atomicOr(&ctr->errors, kErrInternal);
            }
        }
    }
}

__global__ void init_mig_kernel(Counters* c, u32 first_overflow_seg) {
    c->next_walker = 0;
    c->next_seg = first_overflow_seg;
    c->rank_rounds = 0;
    c->n_nodes = 0;
    c->contig_bytes = 0;
    c->n_outbox = 0;
    for (int i = 0; i < 40; ++i) c->flags[i] = 0;
}

// ---- head stubs: one zero-length segment per start node parsed here --------------------------------
// Contig c of this rank is segment walk_cap + c; it emits nothing itself and links to the walker of its
// first k-mer on that k-mer's owner (every node with backward ext 'F' is a boundary start there).
template <int W>
__global__ void __launch_bounds__(256)
head_stub_kernel(const typename Slot<W>::value_t* __restrict__ starts, u32 n_starts, int k, int mo, int rank, int world,
                 MigLayout lay, u64* __restrict__ link, unsigned char* __restrict__ seglen,
                 OutEntry<W>* __restrict__ outbox, Counters* ctr) {
    typedef Slot<W> S;
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = c < n_starts;
    const u32 bal = __ballot_sync(kFullMask, on);
    if (!bal) return;
    u32 first = 0;
    if (lane_id() == 0) first = atomicAdd(&ctr->n_outbox, (u32)__popc(bal));
    first = __shfl_sync(kFullMask, first, 0);
    if (!on) return;
    const u32 lid = lay.walk_cap + c;
    const typename S::value_t key = S::key_only(starts[c]);
    seglen[lid] = 0;
    link[lid] = (u64)kLinkPending << 32;
    const u32 at = first + __popc(bal & ((1u << lane_id()) - 1u));
    if (at < lay.outbox_cap) {
        OutEntry<W> e;
        e.key = key; e.src = ((u32)rank << kRankShift) | lid; e.dest = owner_of<W>(key, world, k, mo);
        outbox[at] = e;
    } else {
        atomicOr(&ctr->errors, kErrInternal);
    }
}

// ---- the local walk --------------------------------------------------------------------------------
struct MigWalkParams {
    const void* table;
    u64 nbuckets;
    const void* boundary_list;
    u64* link;
    unsigned char* seglen;
    unsigned char* tmp;
    void* outbox;
    Counters* ctr;
    MigLayout lay;
    u32 split_shift, seg_chars;
    int k, m, mo, rank, world;
};

template <int W>
__global__ void __launch_bounds__(kWalkThreads, W == 1 ? 6 : 4)
walk_mig_kernel(const MigWalkParams p) {
    typedef Slot<W> S;
    typedef typename S::value_t V;
    const V* __restrict__ table = static_cast<const V*>(p.table);
    const V* __restrict__ bstart = static_cast<const V*>(p.boundary_list);
    OutEntry<W>* __restrict__ outbox = static_cast<OutEntry<W>*>(p.outbox);
    const u32 n_boundary = min(p.ctr->n_boundary, p.lay.bcap);
    const u32 total = p.lay.n_split + n_boundary;
    const u32 lane = lane_id();
    const u32 lt_mask = (1u << lane) - 1u;
    const u64 split_mask = (1ull << p.split_shift) - 1ull;
    const u32 my_bits = (u32)p.rank << kRankShift;

    u32 w_next = 0, w_end = 0;
    bool exhausted = false;
    u32 o_next = 0, o_end = 0;
    u32 x_next = 0, x_end = 0;          // this warp's block of outbox entries (one global atomic per kOutBatch links)
    bool active = false;
    V cur = S::zero(), chk = S::zero();
    u32 seg = 0, n = 0, steps = 0, limit = 256;
    u64 acc = 0;

    auto close = [&](u32 next) {       // next: global id, kLinkTail or kLinkPending
        if (n & 7u) *reinterpret_cast<u64*>(p.tmp + (u64)seg * p.seg_chars + (n & ~7u)) = acc;
        p.seglen[seg] = (unsigned char)n;
        p.link[seg] = ((u64)next << 32) | (next == kLinkTail ? 0u : n);
        active = false;
    };

    for (;;) {
        __syncwarp();
        const u32 idle = __ballot_sync(kFullMask, !active);
        if (idle && !exhausted) {
            if (w_next == w_end) {
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->next_walker, kWalkBatch);
                base = __shfl_sync(kFullMask, base, 0);
                w_next = min(base, total);
                w_end = min(base + kWalkBatch, total);
                exhausted = (w_next == w_end);
            }
            const u32 avail = w_end - w_next;
            const u32 rank_in = __popc(idle & lt_mask);
            if (!active && rank_in < avail) {
                const u32 w = w_next + rank_in;
                seg = w; n = 0; acc = 0; steps = 0; limit = 256;          // boundary id b has local id n_split + b = w
                if (w >= p.lay.n_split) {
                    cur = bstart[w - p.lay.n_split];
                    active = true;
                } else {
                    const u64 b = (u64)w << p.split_shift;
                    cur = S::load_one_nc(table + b * S::kPerBucket);
                    if (S::empty(cur)) p.link[seg] = (u64)kLinkUnused << 32;
                    else active = true;
                }
                chk = cur;
            }
            w_next += min((u32)__popc(idle), avail);
        }
        if (__ballot_sync(kFullMask, active) == 0) {
            if (exhausted) break;
            continue;
        }
        const u32 full = __ballot_sync(kFullMask, active && n == p.seg_chars && S::fwd(cur) != kExtF);
        if (full) {
            const u32 want = __popc(full);
            if (o_end - o_next < want) {
                for (u32 id = o_next + lane; id < o_end; id += 32) p.link[id] = (u64)kLinkUnused << 32;
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->next_seg, kSegBatch);
                base = __shfl_sync(kFullMask, base, 0);
                o_next = base; o_end = base + kSegBatch;
            }
            if (full & (1u << lane)) {
                const u32 ns = o_next + __popc(full & lt_mask);
                if (ns >= p.lay.walk_cap) {
                    atomicOr(&p.ctr->errors, kErrInternal);
                    close(kLinkTail);
                } else {
                    close(my_bits | ns);
                    active = true; seg = ns; n = 0; acc = 0;
                }
            }
            o_next += want;
        }
        // ---- one step.  Inserts are routed by owner, so a k-mer stored here IS local: look it up in the local
        // table first; only a miss leaves the GPU.  The lane parks the key for the outbox, and the dense grouping
        // pass (outbox_count_kernel) derives the destination from it -- the owner function (a minimizer scan) is
        // ~150 instructions and would run here with a handful of lanes active.
        bool send = false;
        V send_key = S::zero();
        u32 send_src = 0;
        if (active) {
            const u32 f = S::fwd(cur);
            if (f == kExtF) {
                close(kLinkTail);
            } else {
                acc |= (u64)ext_char(f) << (8u * (n & 7u));
                ++n;
                if ((n & 7u) == 0) {
                    *reinterpret_cast<u64*>(p.tmp + (u64)seg * p.seg_chars + n - 8) = acc;
                    acc = 0;
                }
                const V nk = S::next_key(cur, p.k);
                {
                    const u64 home = place_bucket<W>(nk, p.k, p.m, p.nbuckets);
                    V nxt = S::zero();
                    u64 b = home; int s = -1;
                    for (u64 tries = 0; tries < p.nbuckets && s < 0; ++tries) {
                        u64 q[4];
                        load256_ro(table + b * S::kPerBucket, q);
                        bool hole = false;
#pragma unroll
                        for (int i = 0; i < S::kPerBucket; ++i) {
                            const V c2 = S::from_bucket(q, i);
                            if (s < 0 && !hole) {
                                if (S::empty(c2)) hole = true;
                                else if (S::same_key(c2, nk)) { nxt = c2; s = i; }
                            }
                        }
                        if (hole) break;
                        if (s < 0) b = (b + 1 == p.nbuckets) ? 0 : b + 1;
                    }
                    if (s < 0) {
                        if (p.world > 1) {                                  // not here: its owner resolves it (or reports it missing)
                            send = true; send_key = nk; send_src = my_bits | seg;
                            close(kLinkPending);
                        } else {
                            atomicOr(&p.ctr->errors, kErrNotFound);         // kmer_hash.cpp:47-49
                            close(kLinkTail);
                        }
                    } else if (s == 0 && (b & split_mask) == 0) {
                        close(my_bits | (u32)(b >> p.split_shift));
                    } else {
                        cur = nxt;
                        if (S::same_key(cur, chk)) {
                            atomicOr(&p.ctr->errors, kErrCycle);
                            close(kLinkTail);
                        } else if (++steps == limit) {
                            chk = cur; steps = 0; limit <<= 1;
                        }
                    }
                }
            }
        }
        const u32 sends = __ballot_sync(kFullMask, send);
        if (sends) {
            const u32 want = __popc(sends);
            if (x_end - x_next < want) {
                // unused tail of the old block: mark the entries void (dest = kOutVoid) so grouping skips them
                for (u32 at = x_next + lane; at < x_end; at += 32)
                    if (at < p.lay.outbox_cap) outbox[at].dest = kOutVoid;
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&p.ctr->n_outbox, kOutBatch);
                base = __shfl_sync(kFullMask, base, 0);
                x_next = base; x_end = base + kOutBatch;
            }
            if (send) {
                const u32 at = x_next + __popc(sends & lt_mask);
                if (at < p.lay.outbox_cap) {
                    OutEntry<W> e;
                    e.key = send_key; e.src = send_src; e.dest = kOutUnrouted;
                    outbox[at] = e;
                } else {
                    atomicOr(&p.ctr->errors, kErrInternal);
                }
            }
            x_next += want;
        }
    }
    for (u32 id = o_next + lane; id < o_end; id += 32)
        if (id < p.lay.walk_cap) p.link[id] = (u64)kLinkUnused << 32;
    for (u32 at = x_next + lane; at < x_end; at += 32)
        if (at < p.lay.outbox_cap) outbox[at].dest = kOutVoid;
}

// ---- outbox -> groups per destination (count, then scatter) ----------------------------------------
// One tile = kOutTile entries: destinations are ranked in shared memory, then one global atomic per destination
// reserves the tile's run in the group (a per-entry atomicAdd on `world` addresses serialises in L2).
constexpr int kOutPerThread = 4;
constexpr int kOutTile = 256 * kOutPerThread;
template <int W>
__global__ void __launch_bounds__(256)
outbox_count_kernel(OutEntry<W>* __restrict__ outbox, Counters* __restrict__ ctr, u32 cap, int k, int mo, int rank, int world,
                    u64* __restrict__ counts) {
    __shared__ u32 s_cnt[kMaxRanks];
    if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 n = min(ctr->n_outbox, cap);
    u32 mine[kMaxRanks];
#pragma unroll
    for (int w = 0; w < kMaxRanks; ++w) mine[w] = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u32 d = outbox[i].dest;
        if (d == kOutUnrouted) {                   // parked by the walk after a local miss: route it now
            d = owner_of<W>(outbox[i].key, world, k, mo);
            if (d == (u32)rank) {                  // it would live here, and it does not: the k-mer is in no table
                atomicOr(&ctr->errors, kErrNotFound);          // kmer_hash.cpp:47-49
                d = kOutVoid;
            }
            outbox[i].dest = d;
        }
#pragma unroll
        for (int w = 0; w < kMaxRanks; ++w) mine[w] += (d == (u32)w);
    }
#pragma unroll
    for (int w = 0; w < kMaxRanks; ++w) {
        const u32 tot = __reduce_add_sync(kFullMask, mine[w]);
        if (lane_id() == 0 && tot) atomicAdd(&s_cnt[w], tot);
    }
    __syncthreads();
    if (threadIdx.x < kMaxRanks && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (u64)s_cnt[threadIdx.x]);
}
template <int W>
__global__ void __launch_bounds__(256)
outbox_scatter_kernel(const OutEntry<W>* __restrict__ outbox, const Counters* __restrict__ ctr, u32 cap,
                      const u64* __restrict__ base, u64* __restrict__ cursor, OutEntry<W>* __restrict__ grouped) {
    __shared__ u32 s_cnt[kMaxRanks];
    __shared__ u64 s_base[kMaxRanks];
    const u32 n = min(ctr->n_outbox, cap);
    const u32 ntiles = (n + kOutTile - 1) / kOutTile;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        OutEntry<W> e[kOutPerThread];
        u32 rk[kOutPerThread];
#pragma unroll
        for (int r = 0; r < kOutPerThread; ++r) {
            const u32 i = tile * kOutTile + r * 256 + threadIdx.x;
            u32 d = 0xFFFFFFFFu;
            if (i < n) { e[r] = outbox[i]; d = e[r].dest; }
            if (d >= (u32)kMaxRanks) d = 0xFFFFFFFFu;              // voided entry (kOutVoid) or past the end
            e[r].dest = d;
            // lanes of a warp that share a destination take consecutive ranks from one shared-memory atomic
            const unsigned same = __match_any_sync(kFullMask, d);
            const int leader = __ffs(same) - 1;
            u32 first = 0;
            if ((int)lane_id() == leader && d != 0xFFFFFFFFu) first = atomicAdd(&s_cnt[d], (u32)__popc(same));
            rk[r] = __shfl_sync(kFullMask, first, leader) + (u32)__popc(same & ((1u << lane_id()) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < kMaxRanks)
            s_base[threadIdx.x] = base[threadIdx.x] +
                                  (s_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (u64)s_cnt[threadIdx.x]) : 0ull);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kOutPerThread; ++r)
            if (e[r].dest != 0xFFFFFFFFu) grouped[s_base[e[r].dest] + rk[r]] = e[r];
        __syncthreads();
    }
}

// ---- resolve the links that arrived: local lookup, then one peer store into the sender's link --------
template <int W>
__global__ void __launch_bounds__(256)
resolve_links_kernel(const Peers pe, const OutEntry<W>* __restrict__ inbox, u64 n,
                     const typename Slot<W>::value_t* __restrict__ table, u64 nbuckets, int k, int m,
                     const u32* __restrict__ seg_of_slot, const typename Slot<W>::value_t* __restrict__ boundary_list,
                     MigLayout lay, Counters* ctr) {
    typedef Slot<W> S;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const OutEntry<W> e = inbox[i];
    typename S::value_t hit;
    u64 b; int s;
    if (!lookup<W>(table, nbuckets, k, m, e.key, hit, b, s)) {
        atomicOr(&ctr->errors, kErrNotFound);              // kmer_hash.cpp:47-49
        return;
    }
    const u32 id = seg_of_slot[b * S::kPerBucket + s];
    const u32 n_boundary = min(ctr->n_boundary, lay.bcap);
    if (id >= n_boundary || !S::same_key(boundary_list[id], e.key)) {
        // the successor was not registered as a walker start: its backward extension does not name its predecessor
        atomicOr(&ctr->errors, kErrBadInput);
        return;
    }
    const u32 next_gid = ((u32)pe.rank << kRankShift) | (lay.n_split + id);
    reinterpret_cast<u32*>(peer_link(pe, e.src))[1] = next_gid;       // high word of the sender's link
}

// valid local ids: splitters, used boundary starts, allocated overflow segments, head stubs
__device__ __forceinline__ bool mig_live_id(u32 lid, const MigLayout& lay, u32 n_boundary, u32 next_seg, u32 n_starts) {
    if (lid < lay.n_split + n_boundary) return true;
    if (lid < lay.n_split + lay.bcap) return false;
    if (lid < lay.walk_cap) return lid < next_seg;
    return lid < lay.walk_cap + n_starts;
}

__global__ void __launch_bounds__(256)
rank_round_mig_kernel(const Peers pe, u64* link, MigLayout lay, u32 n_starts, Counters* ctr, u32* changed) {
    const u32 n_boundary = min(ctr->n_boundary, lay.bcap), next_seg = min(ctr->next_seg, lay.walk_cap);
    const u32 end = lay.walk_cap + n_starts;
    bool moved = false;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (u64)gridDim.x * blockDim.x) {
        if (!mig_live_id((u32)i, lay, n_boundary, next_seg, n_starts)) continue;
        const u64 li = __ldcg(link + i);
        const u32 pi = (u32)(li >> 32), si = (u32)li;
        if (pi >= kLinkFirstMarker) {
            if (pi == kLinkPending) atomicOr(&ctr->errors, kErrInternal);     // a link nobody resolved
            continue;
        }
        if (si & kLinkFinalBit) continue;                     // already points at its tail: no remote read
        const u64 lp = __ldcg(peer_link(pe, pi));
        const u32 pp = (u32)(lp >> 32), sp = (u32)lp;
        if (pp == kLinkTail || pp == kLinkClaimed) {          // pi is the tail
            __stcg(link + i, li | kLinkFinalBit);
            continue;
        }
        if (pp >= kLinkFirstMarker) { atomicOr(&ctr->errors, kErrInternal); continue; }
        // jump; if the segment we jumped over was final, its pointer is the tail and so is ours now
        __stcg(link + i, ((u64)pp << 32) | (u32)((si + (sp & kLinkDistMask)) | (sp & kLinkFinalBit)));
        moved = true;
    }
    if (__any_sync(kFullMask, moved) && lane_id() == 0) *changed = 1;
}

__global__ void __launch_bounds__(256)
emit_segments_mig_kernel(const Peers pe, const u64* __restrict__ link, const unsigned char* __restrict__ seglen,
                         const unsigned char* __restrict__ tmp, u32 seg_chars, MigLayout lay, Counters* ctr, int k) {
    const u64 seg = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n_boundary = min(ctr->n_boundary, lay.bcap), next_seg = min(ctr->next_seg, lay.walk_cap);
    u32 len = 0;
    char* dst = nullptr;
    if (seg < lay.walk_cap && mig_live_id((u32)seg, lay, n_boundary, next_seg, 0)) {
        const u64 li = link[seg];
        const u32 pi = (u32)(li >> 32);
        if (pi != kLinkUnused && pi != kLinkTail && pi != kLinkPending) {
            bool ok = true, is_tail = (pi == kLinkClaimed);
            u32 cg = 0;
            if (is_tail) {
                cg = (u32)li;
            } else {
                const u64 lt = *peer_link(pe, pi);
                if ((u32)(lt >> 32) != kLinkClaimed) ok = false;
                else cg = (u32)lt;
            }
            if (ok) {
                const u32 r = cg >> kRankShift, c = cg & kLocalMask;
                const u32 pre = pe.contig_pre[r][c];
                u32 pos = pre;
                if (!is_tail) {
                    const u32 dist = (u32)li & kLinkDistMask;
                    if (dist > pre) { atomicOr(&ctr->errors, kErrConverge); ok = false; }
                    else pos = pre - dist;
                }
                if (ok) {
                    const u64 off = pe.contig_off[r][c] + (u64)k + pos;
                    len = seglen[seg];
                    if (off + len > pe.out_cap[r]) { atomicOr(&ctr->errors, kErrInternal); len = 0; }
                    dst = pe.out[r] + off;
                }
            }
        }
    }
    u32 todo = __ballot_sync(kFullMask, len > 0);
    const u64 seg0 = seg - lane_id();
    while (todo) {
        const int src_lane = __ffs(todo) - 1;
        todo &= todo - 1;
        const u32 n = __shfl_sync(kFullMask, len, src_lane);
        char* d = reinterpret_cast<char*>(__shfl_sync(kFullMask, reinterpret_cast<u64>(dst), src_lane));
        const unsigned char* src = tmp + (seg0 + src_lane) * (u64)seg_chars;
        for (u32 j = lane_id(); j < n; j += 32) d[j] = (char)src[j];
    }
}

}  // namespace kh
