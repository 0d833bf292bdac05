"""CPU, world_size 2 over gloo: the host-side logic of the sharded path -- block partition of the input
(read_kmers.hpp:55-58), the owner function mirror, who keeps which start node, and the two things that cross the
host in the multi-process variant (cs267_hw3_b200.sharded.TorchComm): the peer-handle gather at start-up and the
OR of the error bits at the end of a step.  The kernels are replaced by their host mirrors; the data path itself
(peer stores, in-stream barriers) needs GPUs and is covered by tests/test_gpu_sharded.py and tests/test_multi_gpu.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cs267_hw3_b200 import sharded as sh
from tools import kmergen


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, n, c, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = kmergen.Dataset(k, n, c, seed=99, threads=1)
    pairs = d.pairs()
    lo, hi = sh.block_of_rank(n, world, rank)
    mine = [sh.slot_from_pair(r.tobytes(), k) for r in pairs[lo:hi]]
    owners = [sh.owner_of_slot(v, k, world) for v in mine]
    counts = torch.tensor([owners.count(w) for w in range(world)], dtype=torch.int64)       # what this rank sends where
    recv = torch.empty_like(counts)
    dist.all_to_all_single(recv, counts)
    everything = [sh.slot_from_pair(r.tobytes(), k) for r in pairs]
    want_owned = sum(1 for v in everything if sh.owner_of_slot(v, k, world) == rank)
    starts = [i for i in range(lo, hi) if pairs[i, (k + 3) // 4] == ord("F")]
    blob = bytes([rank]) * 64 * 3
    hs, ms = sh.gather_peer_handles(blob, (1234, 1000 + rank), dist, world)
    bits = sh.reduce_error_bits(4 if rank == 1 else 0, dist)
    # successive k-mers of a contig mostly share their owner (that is the point of the minimizer)
    sol = d.solution().split(b"\n")[0].decode()
    keys = [sum("ACGT".index(ch) << (2 * (k - 1 - j)) for j, ch in enumerate(sol[i:i + k])) << 6 for i in range(len(sol) - k + 1)]
    own = [sh.owner_of_slot(v, k, world) for v in keys]
    same = sum(1 for x, y in zip(own, own[1:]) if x == y)
    torch.save({"n_recv": int(recv.sum()), "want_owned": want_owned, "n_starts": len(starts),
                "expected_contigs": d.expected(world, rank)[1], "handles_ok": hs == [bytes([r]) * 192 for r in range(world)],
                "meta_ok": ms == [(1234, 1000 + r) for r in range(world)], "bits": bits,
                "stay": same / max(1, len(own) - 1)}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [19, 51])
def test_host_side_plan_world2_gloo(tmp_path, k):
    world, n, c = 2, 3000, 40
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), k, n, c, out), nprocs=world, join=True)
    res = [torch.load(out + f".{r}") for r in range(world)]
    assert all(r["n_recv"] == r["want_owned"] for r in res)            # the counts each rank announces add up to the owners' shares
    assert sum(r["n_recv"] for r in res) == n                          # a partition of the k-mers
    assert all(abs(r["n_recv"] - n / world) < 40 * (n / world) ** 0.5 for r in res)   # supermers move as a unit
    # start nodes stay with the rank that parsed them (kmer_hash.cpp:27-31)
    assert [r["n_starts"] for r in res] == [r["expected_contigs"] for r in res]
    assert sum(r["n_starts"] for r in res) == c
    assert all(r["handles_ok"] and r["meta_ok"] for r in res)
    assert all(r["bits"] == 4 for r in res)                            # rank 1's error reaches everybody
    assert all(r["stay"] > 0.6 for r in res)                           # a chain changes owner only where the minimizer does


def test_block_of_rank_matches_read_kmers():
    # read_kmers.hpp:55-58, including the ragged tail and ranks past the end
    assert [sh.block_of_rank(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [sh.block_of_rank(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sh.block_of_rank(0, 3, 1) == (0, 0)
