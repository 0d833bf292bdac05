"""GPU: the hash-sharded (multi-rank) path.  All ranks run in this one process on cuda:0 through
LocalComm -- the same entry points, kernels, peer stores and in-stream barriers as one process (or thread) per
GPU, enqueued part by part for all ranks in turn -- so the multi-rank logic is exercised on a single-GPU box.
Rank r must emit exactly the reference's `<prefix>_<r>.dat`: the contigs whose start line lies in its block of
the input, in input order."""
import ctypes as C

import numpy as np
import pytest

from tools import kmergen

pytestmark = pytest.mark.gpu


def _upload(kh, sh, shards, pairs, n):
    L = kh.lib()
    world = len(shards)
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
    return blocks, bufs


def _run(k, n, c, world, seed=1, longn=0, lf=0.5, reps=2, split_insert=1):
    import cs267_hw3_b200 as kh
    from cs267_hw3_b200 import sharded as sh

    d = kmergen.Dataset(k, n, c, seed=seed, long_nodes=longn)
    pairs = d.pairs()
    n_local_max = (n + world - 1) // world
    shards = [sh.Shard(k, r, world, n_local_max, n, lf, device=0) for r in range(world)]
    comm = sh.LocalComm(shards)
    comm.connect()
    L = kh.lib()
    blocks, bufs = _upload(kh, sh, shards, pairs, n)
    pb = pairs.shape[1]
    outs = None
    for rep in range(reps):                    # second pass: begin (clear + barrier) + redo on the same handles
        comm.begin()
        for j in range(split_insert):          # the block may arrive in several insert calls (streamed ingestion)
            part = []
            for ptr, cnt in blocks:
                lo = cnt * j // split_insert
                hi = cnt * (j + 1) // split_insert
                part.append((ptr + lo * pb, hi - lo))
            sh.sharded_insert(comm, part)
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
        assert all(abs(x["n_inserted"] - share) < 40 * share ** 0.5 + 10 for x in st)   # supermers move as a unit


@pytest.mark.parametrize("k", [17, 22, 23, 29, 31, 54])
def test_sharded_other_k(k):
    """Both slot widths of the chunk table and the edges of its K range (64-bit slots up to K=22, 128-bit above)."""
    _run(k, 40000, 300, 3, seed=k, reps=1)


def test_sharded_long_contig():
    _run(51, 80000, 6, 4, seed=3, longn=60000)
    _run(19, 80000, 6, 3, seed=4, longn=60000)


def test_sharded_many_tiny_contigs():
    _run(19, 30000, 30000, 4, seed=5)
    _run(31, 50000, 10000, 2, seed=6)


def test_sharded_medium():
    _run(19, 2_000_000, 19_000, 4, seed=7)
    _run(51, 1_000_000, 9_500, 8, seed=8)


@pytest.mark.parametrize("lf", [0.3, 0.7, 0.9])
def test_sharded_load_factors(lf):
    _run(19, 150000, 500, 2, seed=11, lf=lf, reps=1)
    _run(51, 150000, 500, 3, seed=12, lf=lf, reps=1)


def test_sharded_insert_in_several_calls():
    _run(19, 90000, 700, 3, seed=21, split_insert=4, reps=1)
    _run(51, 90000, 700, 2, seed=22, split_insert=3, reps=1)


@pytest.mark.parametrize("pct", ["1", "40"])
def test_staging_overflow_paths(monkeypatch, pct):
    """Undersized staging buffers (KH_DEBUG_CAP_PCT): records that do not fit go to the owner one by one
    (ct_stage_kernel -> extra list -> ct_extra_kernel) and end up in the same chunks."""
    monkeypatch.setenv("KH_DEBUG_CAP_PCT", pct)
    _run(19, 120000, 900, 3, seed=31, reps=1)
    _run(51, 60000, 300, 2, seed=32, reps=1)


def test_sharded_large_shard():
    # 10 M k-mers per rank: many regions, chunk buffers of real size
    _run(19, 20_000_000, 190_000, 2, seed=33, reps=1)


def test_owner_function_mirror_matches_gpu():
    """cs267_hw3_b200.sharded.owner_of_slot (host mirror, used by the gloo tests) == where the GPU puts a k-mer:
    with `world` ranks, rank r's shard must hold exactly the k-mers the mirror assigns to r."""
    import cs267_hw3_b200 as kh
    from cs267_hw3_b200 import sharded as sh

    for k in (19, 51):
        world, n = 5, 4000
        d = kmergen.Dataset(k, n, 40, seed=k)
        pairs = d.pairs()
        shards = [sh.Shard(k, r, world, (n + world - 1) // world, n, 0.5, 0) for r in range(world)]
        comm = sh.LocalComm(shards)
        comm.connect()
        blocks, bufs = _upload(kh, sh, shards, pairs, n)
        comm.begin()
        sh.sharded_insert(comm, blocks)
        sh.sharded_assemble(comm)
        owners = [sh.owner_of_slot(sh.slot_from_pair(r.tobytes(), k), k, world) for r in pairs]
        assert [s.tab.stats()["n_inserted"] for s in shards] == [owners.count(w) for w in range(world)]
        for p in bufs:
            kh.lib().kh_device_free(p)
        for s in shards:
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
    blocks, bufs = _upload(kh, sh, shards, pairs, n)
    try:
        comm.begin()
        sh.sharded_insert(comm, blocks)
        sh.sharded_assemble(comm)
        return [s.result_host() for s in shards]
    finally:
        comm.close()
        for p in bufs:
            L.kh_device_free(p)
        for s in shards:
            s.close()


@pytest.mark.parametrize("world", [1, 4])
def test_sharded_missing_successor_is_reported(world):
    # kmer_hash.cpp:47-49 -- drop one interior k-mer: the chain through it cannot be completed
    from cs267_hw3_b200 import sharded as sh
    k = 19
    d = kmergen.Dataset(k, 20000, 40, seed=9)
    pairs = d.pairs()
    pl = (k + 3) // 4
    interior = np.where((pairs[:, pl] != ord("F")) & (pairs[:, pl + 1] != ord("F")))[0]
    broken = np.delete(pairs, interior[7], axis=0)
    with pytest.raises(sh.ShardedError) as e:
        _run_raw(k, broken, world, 20000)
    assert e.value.bits & 1 and "k-mer not found in Distributed HashMap" in str(e.value)


@pytest.mark.parametrize("world", [1, 4])
def test_sharded_ignores_backward_extensions_like_the_reference(world):
    """The reference never reads a backward extension except to test it for 'F' (kmer_hash.cpp:27-31): a file whose
    other backward extensions are wrong assembles to the same contigs."""
    k = 19
    d = kmergen.Dataset(k, 30000, 60, seed=10)
    pairs = d.pairs().copy()
    pl = (k + 3) // 4
    rot = {ord("A"): ord("C"), ord("C"): ord("G"), ord("G"): ord("T"), ord("T"): ord("A")}
    for i in range(pairs.shape[0]):
        if pairs[i, pl] != ord("F"):
            pairs[i, pl] = rot[int(pairs[i, pl])]
    outs = _run_raw(k, pairs, world, 30000)
    for r in range(world):
        assert outs[r][0].tobytes() == d.expected(world, r)[0]


@pytest.mark.parametrize("world", [1, 3])
def test_sharded_orphans_are_ignored(world):
    """k-mers on no start-rooted chain -- dangling chains whose successor is missing, and cycles -- are
    never visited by the reference (kmer_hash.cpp:41-53) and must not raise anything here either."""
    import oracle
    k = 19
    d = kmergen.Dataset(k, 20000, 50, seed=13)
    rng = np.random.default_rng(5)
    extra = []
    for j in range(40):                                   # orphan chains: no 'F' start, last successor missing
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, k + 12))
        for i in range(10):
            extra.append(f"{s[i:i + k]} {s[i - 1] if i else 'A'}{s[i + k]}\n")
    for j in range(10):                                   # orphan cycles of 30 k-mers
        cyc = "".join("ACGT"[i] for i in rng.integers(0, 4, 30))
        s = cyc + cyc[:k]
        for i in range(30):
            extra.append(f"{s[i:i + k]} {cyc[i - 1]}{s[i + k] if i + k < len(s) else cyc[(i + k) % 30]}\n")
    text = d.text().tobytes() + "".join(extra).encode()
    lines = [text[i:i + k + 4] for i in range(0, len(text), k + 4)]
    assert len(set(ln[:k] for ln in lines)) == len(lines)
    want = oracle.assemble_text(text, k, world)
    pairs = oracle.parse_lines(text, k)
    outs = _run_raw(k, pairs, world, pairs.shape[0])
    for r in range(world):
        assert outs[r][0].tobytes() == want[r]
