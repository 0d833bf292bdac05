"""GPU: the hash-sharded (multi-rank) path.  All ranks run in this one process on cuda:0 through
LocalComm -- same kernels, same phases, the all-to-all done by device copies -- so the multi-rank logic
is exercised on a single-GPU box.  Rank r must emit exactly the reference's `<prefix>_<r>.dat`: the contigs
whose start line lies in its block of the input, in input order."""
import ctypes as C

import numpy as np
import pytest

from tools import kmergen

pytestmark = pytest.mark.gpu


def _run(k, n, c, world, seed=1, longn=0, lf=0.5, options=None):
    import cs267_hw3_b200 as kh
    from cs267_hw3_b200 import sharded as sh

    d = kmergen.Dataset(k, n, c, seed=seed, long_nodes=longn)
    pairs = d.pairs()
    n_local_max = (n + world - 1) // world
    shards = [sh.Shard(k, r, world, n_local_max, n, lf, device=0) for r in range(world)]
    for s in shards:
        for name, val in (options or {}).items():
            s.tab.set_option(name, val)
    if options:       # capacities depend on the options: re-init
        for s in shards:
            s.reinit()
    comm = sh.LocalComm(shards)
    comm.connect()
    L = kh.lib()
    blocks, bufs = [], []
    for r, s in enumerate(shards):
        lo, hi = sh.block_of_rank(n, world, r)
        p = C.c_void_p()
        assert L.kh_device_alloc(C.byref(p), max(1, (hi - lo) * pairs.shape[1])) == 0
        blk = np.ascontiguousarray(pairs[lo:hi])
        s.tab._check(L.kh_copy_device(s.tab._h, p, blk.ctypes.data, blk.nbytes))
        s.tab.sync()
        blocks.append((p.value, hi - lo))
        bufs.append(p)
    outs = None
    for rep in range(2):                       # second pass: clear + redo on the same handles
        for s in shards:
            s.tab.clear()
        comm.barrier()
        sh.sharded_insert(comm, blocks)
        rounds = sh.sharded_assemble(comm)
        outs = [s.result_host() for s in shards]
        st = [s.tab.stats() for s in shards]
        assert sum(x["n_inserted"] for x in st) == n and sum(x["n_duplicates"] for x in st) == 0
        for r in range(world):
            want, want_nc = d.expected(world, r)
            got, nc, nn = outs[r]
            assert nc == want_nc
            assert got.tobytes() == want, f"rank {r} of {world} (k={k}) differs from the reference's per-rank output"
        assert sum(o[2] for o in outs) == n
    comm.close()
    for p in bufs:
        L.kh_device_free(p)
    for s in shards:
        s.close()
    return st, rounds


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("k", [19, 51])
def test_sharded_matches_per_rank_reference_output(k, world):
    st, rounds = _run(k, 60000, 400, world, seed=world)
    # the shards really are shards: nobody holds everything (world > 1) and the split is even
    if world > 1:
        share = 60000 / world
        assert all(abs(x["n_inserted"] - share) < 30 * share ** 0.5 + 10 for x in st)   # supermers move as a unit


def test_sharded_long_contig_and_dense_splitters():
    _run(51, 80000, 6, 4, seed=3, longn=60000, options={"split_buckets": 1, "seg_chars": 8})
    _run(19, 80000, 6, 3, seed=4, longn=60000, options={"split_buckets": 1 << 20})


def test_sharded_many_tiny_contigs():
    _run(19, 30000, 30000, 4, seed=5)
    _run(31, 50000, 10000, 2, seed=6)


def test_sharded_medium():
    _run(19, 2_000_000, 19_000, 4, seed=7)
    _run(51, 1_000_000, 9_500, 8, seed=8)


@pytest.mark.parametrize("mode", ["0", "2"])
@pytest.mark.parametrize("k,world", [(19, 1), (19, 2), (19, 5), (31, 3), (51, 2), (51, 8)])
def test_owner_side_build_modes(monkeypatch, k, world, mode):
    """KH_SHARD_BUILD: 0 = atomic insert of the received slot values, 2 = always the chunked shared-memory build
    (default 1 picks it for shards larger than L2).  Same per-rank bytes either way."""
    monkeypatch.setenv("KH_SHARD_BUILD", mode)
    _run(k, 90000, 700, world, seed=20 + world)


@pytest.mark.parametrize("pct", ["1", "40"])
def test_owner_side_build_overflow_paths(monkeypatch, pct):
    """Undersized grouping buffers (KH_DEBUG_CAP_PCT): values that do not fit go through the overflow fix-up,
    which registers boundary starts exactly like the main path."""
    monkeypatch.setenv("KH_SHARD_BUILD", "2")
    monkeypatch.setenv("KH_DEBUG_CAP_PCT", pct)
    _run(19, 120000, 900, 3, seed=31)
    _run(51, 60000, 300, 2, seed=32)


def test_owner_side_build_default_large_shard():
    # shard tables of 160 MB: the default picks the chunked build
    _run(19, 20_000_000, 190_000, 2, seed=33)


def test_owner_function_mirror_matches_gpu():
    """cs267_hw3_b200.sharded.owner_of_slot (host mirror used for planning and the gloo tests) == the GPU's grouping."""
    import cs267_hw3_b200 as kh
    from cs267_hw3_b200 import sharded as sh

    for k in (19, 51):
        world, n = 5, 4000
        d = kmergen.Dataset(k, n, 40, seed=k)
        pairs = d.pairs()
        s = sh.Shard(k, 0, world, n, n, 0.5, 0)
        L = kh.lib()
        p = C.c_void_p()
        assert L.kh_device_alloc(C.byref(p), pairs.nbytes) == 0
        s.tab._check(L.kh_copy_device(s.tab._h, p, pairs.ctypes.data, pairs.nbytes))
        ptr, counts = s.owner_partition(p.value, n)
        eb = sh.slot_bytes(k)
        raw = np.empty(n * eb, dtype=np.uint8)
        s.tab._check(L.kh_copy_to_host(s.tab._h, raw.ctypes.data, ptr, raw.nbytes))
        slots = [int.from_bytes(raw[i * eb:(i + 1) * eb].tobytes(), "little") for i in range(n)]
        want = sorted(sh.slot_from_pair(r.tobytes(), k) for r in pairs)
        assert sorted(slots) == want
        owners = [sh.owner_of_slot(v, k, world) for v in slots]
        assert owners == sorted(owners)                                   # grouped in owner order
        assert [owners.count(w) for w in range(world)] == counts
        L.kh_device_free(p)
        s.close()


def _run_raw(k, pairs, world, n_total):
    """insert + assemble given explicit records; returns per-rank outputs or raises ShardedError."""
    import cs267_hw3_b200 as kh
    from cs267_hw3_b200 import sharded as sh

    n = pairs.shape[0]
    shards = [sh.Shard(k, r, world, (n + world - 1) // world, n_total, 0.5, device=0) for r in range(world)]
    comm = sh.LocalComm(shards)
    comm.connect()
    L = kh.lib()
    blocks, bufs = [], []
    try:
        for r, s in enumerate(shards):
            lo, hi = sh.block_of_rank(n, world, r)
            p = C.c_void_p()
            assert L.kh_device_alloc(C.byref(p), max(1, (hi - lo) * pairs.shape[1])) == 0
            blk = np.ascontiguousarray(pairs[lo:hi])
            s.tab._check(L.kh_copy_device(s.tab._h, p, blk.ctypes.data, blk.nbytes))
            s.tab.sync()
            blocks.append((p.value, hi - lo))
            bufs.append(p)
        sh.sharded_insert(comm, blocks)
        sh.sharded_assemble(comm)
        return [s.result_host() for s in shards]
    finally:
        comm.close()
        for p in bufs:
            L.kh_device_free(p)
        for s in shards:
            s.close()


@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("world", [1, 4])
def test_sharded_missing_successor_is_reported(monkeypatch, world, mode):
    # kmer_hash.cpp:47-49 -- drop one interior k-mer: the chain through it cannot be completed
    from cs267_hw3_b200 import sharded as sh
    monkeypatch.setenv("KH_SHARD_BUILD", mode)
    k = 19
    d = kmergen.Dataset(k, 20000, 40, seed=9)
    pairs = d.pairs()
    pl = (k + 3) // 4
    interior = np.where((pairs[:, pl] != ord("F")) & (pairs[:, pl + 1] != ord("F")))[0]
    broken = np.delete(pairs, interior[7], axis=0)
    with pytest.raises(sh.ShardedError) as e:
        _run_raw(k, broken, world, 20000)
    assert e.value.bits & 1 and "k-mer not found in Distributed HashMap" in str(e.value)


def test_sharded_needs_truthful_backward_extensions():
    """The migrating walk finds a shard's walker starts from the backward extensions (kmer_t.hpp:55-57).  The
    reference never reads them except for 'F'; files that lie about them are refused, not mis-assembled."""
    from cs267_hw3_b200 import sharded as sh
    k = 19
    d = kmergen.Dataset(k, 30000, 60, seed=10)
    pairs = d.pairs().copy()
    pl = (k + 3) // 4
    rot = {ord("A"): ord("C"), ord("C"): ord("G"), ord("G"): ord("T"), ord("T"): ord("A")}
    for i in range(pairs.shape[0]):
        if pairs[i, pl] != ord("F"):
            pairs[i, pl] = rot[int(pairs[i, pl])]
    with pytest.raises(sh.ShardedError) as e:
        _run_raw(k, pairs, 4, 30000)
    assert e.value.bits & 8
    outs = _run_raw(k, pairs, 1, 30000)           # one rank: nothing crosses, output as the reference's
    assert outs[0][0].tobytes() == d.expected()[0]
