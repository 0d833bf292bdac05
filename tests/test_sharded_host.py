"""CPU, world_size 2 over gloo: the host-side logic of the sharded path -- block partition of the input,
owner function, per-destination grouping and the variable-size all-to-all (cs267_hw3_b200.sharded.exchange_bytes).
The kernels are replaced by their host mirrors (slot_from_pair / owner_of_slot); the oracle provides the records."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
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
    eb = sh.slot_bytes(k)
    mine = [sh.slot_from_pair(r.tobytes(), k) for r in pairs[lo:hi]]
    owners = [sh.owner_of_slot(v, k, world) for v in mine]
    order = sorted(range(len(mine)), key=lambda i: owners[i])              # what K7 does on the GPU
    send = np.frombuffer(b"".join(mine[i].to_bytes(eb, "little") for i in order), dtype=np.uint8).copy()
    counts = [owners.count(w) for w in range(world)]
    recv, rc = sh.exchange_bytes(torch.from_numpy(send), counts, eb)
    got = sorted(int.from_bytes(recv[i * eb:(i + 1) * eb].numpy().tobytes(), "little") for i in range(sum(rc)))
    # every k-mer this rank owns, from every rank's block, and nothing else
    everything = [sh.slot_from_pair(r.tobytes(), k) for r in pairs]
    want = sorted(v for v in everything if sh.owner_of_slot(v, k, world) == rank)
    starts = [i for i in range(lo, hi) if pairs[i, (k + 3) // 4] == ord("F")]
    torch.save({"ok": got == want, "n_recv": len(got), "rc": rc, "n_starts": len(starts),
                "expected_contigs": d.expected(world, rank)[1]}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [19, 51])
def test_exchange_plan_world2_gloo(tmp_path, k):
    world, n, c = 2, 3000, 40
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), k, n, c, out), nprocs=world, join=True)
    res = [torch.load(out + f".{r}") for r in range(world)]
    assert all(r["ok"] for r in res)
    assert sum(r["n_recv"] for r in res) == n                          # a partition of the k-mers
    assert all(abs(r["n_recv"] - n / world) < 30 * (n / world) ** 0.5 for r in res)   # supermers move as a unit
    # start nodes stay with the rank that parsed them (kmer_hash.cpp:27-31)
    assert [r["n_starts"] for r in res] == [r["expected_contigs"] for r in res]
    assert sum(r["n_starts"] for r in res) == c


def test_block_partition_matches_read_kmers():
    # read_kmers.hpp:55-58
    assert [sh.block_of_rank(10, 3, r) for r in range(3)] == [(0, 4), (4, 8), (8, 10)]
    assert [sh.block_of_rank(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]   # the reference underflows here
    n = 89_710_742
    blocks = [sh.block_of_rank(n, 8, r) for r in range(8)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


def test_slot_mirror_against_oracle_records():
    # slot = (key << 6) | (back << 3) | (fwd + 1); key is the packed bytes read big-endian minus the tail padding
    for k in (19, 31, 51):
        d = kmergen.Dataset(k, 200, 5, seed=k, threads=1)
        text = d.text()
        for rec, line in zip(oracle.parse_lines(text, k), text.tobytes().split(b"\n")):
            v = sh.slot_from_pair(rec.tobytes(), k)
            key = 0
            for ch in line[:k].decode():
                key = key * 4 + "ACGT".index(ch)
            assert v >> 6 == key
            assert "ACGTF"[(v >> 3) & 7] == chr(line[k + 1]) and "ACGTF"[(v & 7) - 1] == chr(line[k + 2])


def test_shard_capacity_covers_imbalance():
    for n, w in ((89_710_742, 8), (1_000_000_000, 8), (1000, 4)):
        assert sh.shard_capacity(n, w) > n / w + 6 * (n / w) ** 0.5
