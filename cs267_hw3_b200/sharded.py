"""Multi-GPU (hash-sharded) insert + traverse: host-side orchestration over the C ABI.

Replaces what the reference does with UPC++: the table is sharded by owner rank
(hash_map.hpp:28-30); inserts travel as one batch per destination (hash_map.hpp:64-77), finds are RPC round
trips (hash_map.hpp:93-100).  Here a rank owns a contiguous range of table CHUNKS (csrc/ctable.cuh) and the
owner of a k-mer follows from its minimizer, so consecutive k-mers of a contig live on one GPU.  A step is
SPMD and stream-ordered, exactly as in include/kh_capi.h:

    kh_shard_begin -> kh_shard_insert -> kh_shard_assemble -> kh_shard_finish

The records reach their owner GPU through NVLink peer stores inside the grouping kernel (no separate all-to-all
step), chain links that leave a GPU and their answers travel the same way, phases are separated by an in-stream
barrier kernel; the host only waits in kh_shard_finish.  Rank r emits the contigs whose start node lies in its
block of input lines, in input order (kmer_hash.cpp:27-31, read_kmers.hpp:55-58) -- the reference's
``<prefix>_<r>.dat``.

Two ways to host the ranks:
  * TorchComm  -- one process per GPU (torchrun): torch.distributed moves the CUDA IPC handles once at start-up
                  and reduces the error bits at the end; nothing else crosses the host.  This is what bench.py runs.
  * LocalComm  -- all ranks in this process (even on ONE GPU): the same calls, enqueued part by part for all
                  ranks in turn (kh_shard_assemble_part) so that barrier kernels of ranks that share a device
                  cannot block each other's queued work.  Used by the tests on single-GPU boxes.
The C++ twin (one host thread per GPU) is include/kh/sharded_host.hpp.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

MAX_RANKS = 8
_M64 = (1 << 64) - 1
_M32 = (1 << 32) - 1


# ---------------------------------------------------------------------------- host mirrors ----
def _fmix64(z: int) -> int:
    z ^= z >> 33
    z = (z * 0xFF51AFD7ED558CCD) & _M64
    z ^= z >> 33
    z = (z * 0xC4CEB9FE1A85EC53) & _M64
    z ^= z >> 33
    return z


def ext_code(ch: int) -> int:
    return {65: 0, 67: 1, 71: 2, 84: 3, 70: 4}[ch]


def slot_from_pair(pair: bytes, k: int) -> int:
    """kmer_pair bytes -> device slot value (csrc/slot.cuh): (key << 6) | (back << 3) | (fwd + 1)."""
    pl = (k + 3) // 4
    key = int.from_bytes(pair[:pl], "big") >> (8 * pl - 2 * k)
    return (key << 6) | (ext_code(pair[pl]) << 3) | (ext_code(pair[pl + 1]) + 1)


def ct_minimizer_len(k: int) -> int:
    return 15 if k >= 31 else (13 if k >= 23 else (k - 6 if k >= 17 else k))


def ct_window(k: int) -> int:
    return min(32, k - ct_minimizer_len(k) + 1)


def ct_min_hash(slot: int, k: int) -> int:
    """Mirror of ct_min_hash<W>() (csrc/slot.cuh): min over the `win` rightmost m-mers x of
    (x * odd << (32 - 2m)) + c mod 2^32."""
    key, m, win = slot >> 6, ct_minimizer_len(k), ct_window(k)
    mul = (0x9E3779B1 << (32 - 2 * m)) & _M32
    return min((((key >> (2 * j)) & _M32) * mul + 0x7F4A7C15) & _M32 for j in range(win))


def _fmix32_hi(h: int) -> int:
    h ^= h >> 15
    h = (h * 0x85EBCA6B) & _M32
    h ^= h >> 13
    return (h * 0xC2B2AE35) & _M32


def owner_of_slot(slot: int, k: int, world: int) -> int:
    """Mirror of ct_place() (csrc/slot.cuh): the rank that owns a k-mer (tests compare it with where the GPU put it)."""
    mh = ct_min_hash(slot, k)
    return (_fmix32_hi(mh ^ (mh >> 16)) * world) >> 32


def slot_bytes(k: int) -> int:
    return 8 if (17 <= k <= 54 and 2 * k + 6 + 14 <= 64) or (not 17 <= k <= 54 and 2 * k + 6 <= 64) else 16


def block_of_rank(n: int, world: int, rank: int) -> tuple[int, int]:
    """read_kmers.hpp:55-58: rank r parses lines [ceil(n/P)*r, min(n, ceil(n/P)*(r+1)))."""
    split = (n + world - 1) // world
    lo = min(n, split * rank)
    return lo, min(n, lo + split)


def shard_capacity(n_total: int, world: int) -> int:
    """k-mers a shard must be able to hold: its expected share plus slack.  Ownership follows the minimizer, so whole
    supermers (runs of ~4-16 k-mers) move together and the spread is a few times the binomial one."""
    share = n_total / world
    return int(share * 1.02 + 64.0 * math.sqrt(share + 1.0)) + 4096


# ---------------------------------------------------------------------------- shard object ----
class Shard:
    """One rank: a kh_table in sharded mode."""

    def __init__(self, k: int, rank: int, world: int, n_local_max: int, n_total: int,
                 load_factor: float = 0.5, device: int = 0, n_starts_max: int | None = None):
        import cs267_hw3_b200 as kh

        self.kh, self.k, self.rank, self.world, self.device = kh, k, rank, world, device
        self.n_local_max, self.n_total = n_local_max, n_total
        self.n_starts_max = n_local_max if n_starts_max is None else n_starts_max
        self.tab = kh.KmerHashTable(k, shard_capacity(n_total, world), load_factor, device)
        L = kh.lib()
        self._declare(L)
        self.reinit()

    def reinit(self) -> None:
        """Fix the capacities (after setting options, before connecting the peers)."""
        self.tab._check(self.kh.lib().kh_shard_init(self.tab._h, self.rank, self.world, self.n_local_max, self.n_total,
                                                    self.n_starts_max))

    @staticmethod
    def _declare(L):
        if getattr(L, "_shard_declared", False):
            return
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.kh_shard_init.argtypes = [vp, i32, i32, u64, u64, u64]
        L.kh_shard_export_count.restype = i32
        L.kh_shard_export.argtypes = [vp, vp, C.POINTER(u64)]
        L.kh_shard_connect.argtypes = [vp, vp, C.POINTER(u64)]
        L.kh_shard_connect_local.argtypes = [vp, C.POINTER(vp), i32]
        L.kh_shard_begin.argtypes = [vp]
        L.kh_shard_insert.argtypes = [vp, vp, u64]
        L.kh_shard_assemble.argtypes = [vp]
        L.kh_shard_assemble_parts.restype = i32
        L.kh_shard_assemble_part.argtypes = [vp, i32]
        L.kh_shard_finish.argtypes = [vp, C.POINTER(i32)]
        L.kh_shard_result.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.kh_slot_bytes.argtypes = [i32]
        L.kh_slot_bytes.restype = u64
        L.kh_device_alloc.argtypes = [C.POINTER(vp), u64]
        L.kh_device_alloc_on.argtypes = [i32, C.POINTER(vp), u64]
        L.kh_device_free.argtypes = [vp]
        L.kh_copy_to_host.argtypes = [vp, vp, vp, u64]
        L.kh_copy_device.argtypes = [vp, vp, vp, u64]
        L._shard_declared = True

    def begin(self) -> None:
        self.tab._check(self.kh.lib().kh_shard_begin(self.tab._h))

    def insert(self, pairs_dev: int, n: int) -> None:
        self.tab._check(self.kh.lib().kh_shard_insert(self.tab._h, pairs_dev, n))

    def assemble(self) -> None:
        self.tab._check(self.kh.lib().kh_shard_assemble(self.tab._h))

    def assemble_part(self, part: int) -> None:
        self.tab._check(self.kh.lib().kh_shard_assemble_part(self.tab._h, part))

    def finish(self) -> int:
        bits = C.c_int()
        self.tab._check(self.kh.lib().kh_shard_finish(self.tab._h, C.byref(bits)))
        return bits.value

    def result(self):
        cp, op = C.c_void_p(), C.c_void_p()
        nc, nb, nn = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.tab._check(self.kh.lib().kh_shard_result(self.tab._h, C.byref(cp), C.byref(op), C.byref(nc), C.byref(nb), C.byref(nn)))
        return cp.value, op.value, nc.value, nb.value, nn.value

    def result_host(self) -> tuple[np.ndarray, int, int]:
        cp, _, nc, nb, nn = self.result()
        buf = np.empty(nb, dtype=np.uint8)
        self.tab._check(self.kh.lib().kh_copy_to_host(self.tab._h, buf.ctypes.data, cp, nb))
        return buf, nc, nn

    def export(self) -> tuple[bytes, tuple[int, int]]:
        handles = (C.c_uint8 * (self.kh.lib().kh_shard_export_count() * 64))()
        meta = (C.c_uint64 * 2)()
        self.tab._check(self.kh.lib().kh_shard_export(self.tab._h, handles, meta))
        return bytes(handles), (int(meta[0]), int(meta[1]))

    def connect(self, all_handles: list[bytes], all_meta: list[tuple[int, int]]) -> None:
        blob = b"".join(all_handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        meta = (C.c_uint64 * (2 * len(all_meta)))(*[x for m in all_meta for x in m])
        self.tab._check(self.kh.lib().kh_shard_connect(self.tab._h, arr, meta))

    def connect_local(self, shards: list["Shard"]) -> None:
        arr = (C.c_void_p * len(shards))(*[s.tab._h for s in shards])
        self.tab._check(self.kh.lib().kh_shard_connect_local(self.tab._h, arr, len(shards)))

    def close(self) -> None:
        self.tab.close()


ERR_BITS = {1: ("KH_ERR_NOT_FOUND", "Error: k-mer not found in Distributed HashMap."), 2: ("KH_ERR_TABLE_FULL", "hash table full"),
            4: ("KH_ERR_CYCLE", "chain never terminates (cycle)"), 8: ("KH_ERR_BAD_INPUT", "malformed input"),
            16: ("KH_ERR_CONVERGE", "chains are not linear"), 32: ("KH_ERR_CUDA", "internal: a capacity was exceeded or a peer never reached a barrier")}


class ShardedError(RuntimeError):
    def __init__(self, bits: int):
        names = [v for b, v in ERR_BITS.items() if bits & b]
        super().__init__("; ".join(f"{n}: {m}" for n, m in names))
        self.bits = bits


# ---------------------------------------------------------------------------- hosting the ranks ----
class LocalComm:
    """All ranks in this process; shards[i] is rank i.  Calls that contain a barrier are enqueued for every rank
    before the next one (see the module docstring)."""

    def __init__(self, shards: list[Shard]):
        self.shards = shards

    def connect(self):
        for s in self.shards:
            s.connect_local(self.shards)

    def sync(self):
        for s in self.shards:
            s.tab.sync()

    def begin(self):
        for s in self.shards:
            s.begin()

    def assemble(self):
        parts = self.shards[0].kh.lib().kh_shard_assemble_parts()
        for p in range(parts):
            for s in self.shards:
                s.assemble_part(p)

    def finish(self) -> int:
        bits = 0
        for s in self.shards:
            bits |= s.finish()
        return bits

    def close(self):
        pass


def gather_peer_handles(handles: bytes, meta: tuple[int, int], dist, world: int):
    """Every rank's CUDA IPC handle blob and metadata, in rank order (any torch.distributed backend)."""
    gathered = [None] * world
    dist.all_gather_object(gathered, (handles, meta))
    return [g[0] for g in gathered], [tuple(g[1]) for g in gathered]


def reduce_error_bits(bits: int, dist, device="cpu") -> int:
    """OR of the per-rank error bits (kh_shard_finish) over all ranks."""
    import torch

    t = torch.tensor([bits], dtype=torch.int32, device=device)
    gathered = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, t)
    out = 0
    for g in gathered:
        out |= int(g.item())
    return out


class TorchComm:
    """One rank per process (torchrun).  torch.distributed carries the CUDA IPC handles at start-up and the error
    bits at the end of a step; the data path never touches it."""

    def __init__(self, shard: Shard):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.shards = [shard]

    def connect(self):
        s = self.shards[0]
        h, m = s.export()
        hs, ms = gather_peer_handles(h, m, self.dist, s.world)
        s.connect(hs, ms)
        self.dist.barrier()

    def sync(self):
        self.shards[0].tab.sync()

    def begin(self):
        self.shards[0].begin()

    def assemble(self):
        self.shards[0].assemble()

    def finish(self) -> int:
        bits = self.shards[0].finish()
        dev = self.torch.device("cuda", self.torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else "cpu"
        return reduce_error_bits(bits, self.dist, dev)

    def close(self):
        pass


# ---------------------------------------------------------------------------- the algorithm ----
def sharded_insert(comm, blocks: list[tuple[int, int]]) -> None:
    """blocks[i] = (device ptr of this rank's kmer_pair records, count) for comm.shards[i].
    initialize_kmers (kmer_hash.cpp:21-33) across ranks; only enqueues."""
    for s, (ptr, n) in zip(comm.shards, blocks):
        s.insert(ptr, n)


def sharded_assemble(comm, finish: bool = True) -> int:
    """assemble_contigs (kmer_hash.cpp:38-55) across ranks.  Returns the number of pointer-jumping rounds of rank 0's
    shard.  Raises ShardedError if any rank flagged an error."""
    comm.assemble()
    if not finish:
        return 0
    bits = comm.finish()
    if bits:
        raise ShardedError(bits)
    return comm.shards[0].tab.stats()["rank_rounds"]


# ---------------------------------------------------------------------------- bench (N > 1) ----
def shutdown():
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def bench(args, workload, rank, world, local_rank, steps, warmup):
    """One workload on `world` GPUs, one process per GPU (torchrun).  Strong scaling: the same synthetic file as at
    N=1, block-partitioned over the ranks (read_kmers.hpp:55-58), table sharded by chunk range.  A step is enqueued
    without any host synchronisation; the host waits once per step (kh_shard_finish) to read the error bits."""
    import time

    import torch
    import torch.distributed as dist

    import cs267_hw3_b200 as kh
    from bench import METRIC, UNIT, WORKLOADS, ClockSampler, alg_bytes_per_kmer, measured_peak_gbs
    from tools import kmergen

    k, n_total, c_total, longn = WORKLOADS[workload]
    if args.n:
        c_total = max(1, int(round(args.n * c_total / n_total)))
        n_total = args.n
    t_gen = time.time()
    weak = getattr(args, "scaling", "strong") == "weak"
    if weak:
        # per-GPU work fixed: every rank brings its own file of n_total k-mers (its block of a world-times larger
        # input).  Independent files are only collision-free for long k-mers, so this mode needs K >= 31.
        if k < 31:
            raise ValueError("--scaling weak needs K >= 31 (independent per-rank files must not share k-mers)")
        data = kmergen.Dataset(k, n_total, c_total, seed=267 + 1000 * rank, long_nodes=longn)
        n_local, lo = n_total, 0
        n_total, c_total = n_total * world, c_total * world
        exp_buf, exp_nc = data.expected_array(1, 0)
    else:
        data = kmergen.Dataset(k, n_total, c_total, seed=267, long_nodes=longn)     # every rank derives the same file
        lo, hi = block_of_rank(n_total, world, rank)
        n_local = hi - lo
        exp_buf, exp_nc = data.expected_array(world, rank)
    pb = kh.pair_bytes(k)
    host = kh.PinnedBuffer(max(n_local, 1) * pb)
    data.pairs_into(host.ptr, lo, n_local)
    t_gen = time.time() - t_gen

    stream = torch.cuda.current_stream()
    n_local_max = n_local if weak else (n_total + world - 1) // world
    shard = Shard(k, rank, world, n_local_max, n_total, args.load_factor, device=local_rank,
                  n_starts_max=int(exp_nc * 1.25) + 4096)
    shard.tab.set_stream(stream.cuda_stream)
    comm = TorchComm(shard)
    comm.connect()
    dev = torch.empty(max(n_local, 1) * pb, dtype=torch.uint8, device="cuda")
    dev.copy_(torch.from_numpy(host.array))
    torch.cuda.synchronize()

    def step():
        comm.begin()                                   # fresh table + barrier: the ranks start the step together
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        sharded_insert(comm, [(dev.data_ptr(), n_local)])
        e1.record(stream)
        comm.assemble()
        e2.record(stream)
        bits = shard.finish()                          # the step's only host wait
        if bits:
            raise ShardedError(bits)
        return e0, e1, e2

    sampler = ClockSampler(local_rank)        # warm-up + timed steps
    sampler.start()
    for _ in range(warmup):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = shard.tab.stats()["n_launches"]
    evs, stage = [], {"ms_stage": [], "ms_build": [], "ms_walk": [], "ms_rank": [], "ms_emit": []}
    wall0 = time.perf_counter()
    for _ in range(steps):
        evs.append(step())
        st = shard.tab.stats()
        for key in stage:
            stage[key].append(st[key])
    dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    st_last = shard.tab.stats()
    launches = st_last["n_launches"] - launches0
    clocks = sampler.stop()
    ms_total = float(np.mean([a.elapsed_time(c) for a, b, c in evs]))
    ms_ins = float(np.mean([a.elapsed_time(b) for a, b, c in evs]))
    tm = torch.tensor([ms_total, ms_ins] + [float(np.mean(v)) for v in stage.values()], device="cuda", dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)                      # slowest rank defines the step
    ms_total, ms_ins = tm[:2].tolist()
    stage_max = dict(zip(stage.keys(), tm[2:].tolist()))

    got, nc, nn = shard.result_host()
    ok = bool(nc == exp_nc and got.size == exp_buf.size and np.array_equal(got, exp_buf))

    # end to end: this rank's records start in pinned HOST memory, its contigs end in host memory
    e2e_ms_max, out_host = None, got
    if not args.no_e2e:
        host_t = torch.from_numpy(host.array)
        e2e_ms = []
        for i in range(3):
            comm.begin()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            dev.copy_(host_t, non_blocking=True)                      # H2D of the step's input
            sharded_insert(comm, [(dev.data_ptr(), n_local)])
            comm.assemble()
            if shard.finish():
                raise ShardedError(1 << 5)
            out_host, _, _ = shard.result_host()                        # D2H of the step's result
            b.record(stream)
            torch.cuda.synchronize()
            if i:
                e2e_ms.append(a.elapsed_time(b))
        te = torch.tensor([float(np.mean(e2e_ms))], device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms_max = float(te.item())
        ok = ok and bool(np.array_equal(out_host, exp_buf))
    tot = torch.tensor([float(nn), 1.0 if ok else 0.0, float(got.size), float(st_last["n_segments"])], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    nodes_all, ok_all, bytes_all, segs_all = tot.tolist()
    verified = bool(int(nodes_all) == n_total and int(ok_all) == world)
    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        alg = alg_bytes_per_kmer(k)
        path_gbs = n_total * alg["total"] / (ms_total * 1e-3) / 1e9
        sb = int(st_last["slot_bits"]) // 8
        build_ms = stage_max["ms_build"]
        build_bytes = (n_total / world) * (alg["insert"] + alg["lookup"] + alg["output"])
        line = {
            "metric": METRIC, "value": n_total / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_total, "higher_is_better": True,
            "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "u64" if sb == 8 else "u128", "data": "synthetic",
            "config": {"workload": workload, "k": k, "n_kmers": n_total, "n_contigs": c_total, "load_factor": args.load_factor,
                       "seed": 267, "per_gpu_kmers": n_local, "slot_bytes": sb,
                       "sharding": f"chunk table sharded over {world} GPUs (a rank owns a contiguous range of chunks; the chunk, and "
                       "so the owner, follows from the k-mer's minimizer); input lines block-partitioned (read_kmers.hpp:55-58); "
                       "records reach the owner as owner-sorted runs of NVLink peer stores inside the staging kernel; one lookup per "
                       "segment chains the segments, requests to other GPUs and their answers travel as coalesced runs; segment "
                       "links, meta words and characters are gathered on every rank, contigs ranked and emitted by per-contig walks "
                       "over local memory; in-stream flag barriers, two host waits per step",
                       "timing": "CUDA events on the launching stream per step (after the step's opening barrier .. after its "
                                 "closing barrier), max over ranks, mean of steps",
                       "l2": "per-GPU staging buffers and table larger than L2; every step starts from an empty table"},
            "stages_ms": {"ms_insert_staging_incl_peer_stores": ms_ins, "ms_seal_and_traverse": ms_total - ms_ins,
                          "max_over_ranks": stage_max},
            "n_segments": int(segs_all), "rank_rounds": int(st_last["rank_rounds"]), "assembly_time_s": ms_total * 1e-3,
            "wall_s_timed_loop": wall, "gen_s": t_gen, "verified": verified,
            "roofline": {"bound": "hbm", "kernel": "ct_build_kernel (per GPU, slowest rank)", "achieved": build_bytes / (build_ms * 1e-3) / 1e9 if build_ms else None,
                         "peak": peak, "unit": "GB/s", "frac": build_bytes / (build_ms * 1e-3) / 1e9 / peak if build_ms else None,
                         "traffic": None, "peak_source": peak_src, "kernel_ms": build_ms,
                         "units_per_launch": n_total // world, "alg_bytes_per_unit": alg["insert"] + alg["lookup"] + alg["output"],
                         "alg_bytes_per_kmer": alg,
                         "path": {"achieved": path_gbs, "peak": peak * world, "frac": path_gbs / (peak * world),
                                  "note": "N x B_alg / t(insert+traverse) against the measured HBM peak x GPUs"},
                         "nvlink_bytes_per_step": int(n_total * (world - 1) / world * (sb + 4) + (world - 1) * (16 * segs_all + n_total)),
                         "nvlink_note": "(P-1)/P of the slot values (+ 4-byte chunk ids) cross in the staging pass; per segment "
                                        "that continues on another GPU one 16/32-byte request and a 4-byte answer (not counted); "
                                        "the gather sends 16 bytes per segment + 1 byte per k-mer to each of the P-1 peers"},
            "e2e": ({"value": n_total / (e2e_ms_max * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_max,
                     "h2d_bytes_per_step": int(n_total * pb), "d2h_bytes_per_step": int(bytes_all),
                     "note": "each rank copies its block of records from pinned host memory and its contigs back"}
                    if e2e_ms_max else None),
            "gpu_launches": int(launches),
            "gpu_launches_note": "kernels launched by rank 0 in the timed steps (counted by the library)",
            "clocks": clocks,
        }
    comm.close()
    shard.close()
    host.free()
    del dev
    torch.cuda.empty_cache()
    dist.barrier()
    return line
