"""Worker for tests/test_multi_gpu.py: one process per GPU under torch.distributed.run.

Runs the sharded path exactly as bench.py does (TorchComm: CUDA IPC peer mappings, peer stores, in-stream
barriers, one host wait per step) on a small synthetic file and compares THIS rank's bytes with the reference's
per-rank output for the same file.  Prints one line per rank; exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cs267_hw3_b200 as kh  # noqa: E402
from cs267_hw3_b200 import sharded as sh  # noqa: E402
from tools import kmergen  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # KH_TEST_ONE_DEVICE=1: every process on GPU 0 (a one-GPU box): the same CUDA IPC mappings, peer stores and in-stream
    # barriers between PROCESSES, the contexts time-sliced by the driver; NCCL refuses two ranks on one device, so the
    # start-up exchange and the error-bit reduction go over gloo
    one_device = os.environ.get("KH_TEST_ONE_DEVICE") == "1"
    if one_device:
        local = 0
    torch.cuda.set_device(local)
    if one_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red_dev = "cpu" if one_device else "cuda"
    failures = 0
    for k, n, c, longn in ((19, 400_000, 3_000, 0), (51, 300_000, 2_500, 0), (31, 200_000, 40, 150_000)):
        d = kmergen.Dataset(k, n, c, seed=500 + k, long_nodes=longn)
        lo, hi = sh.block_of_rank(n, world, rank)
        want, want_nc = d.expected(world, rank)
        pb = kh.pair_bytes(k)
        shard = sh.Shard(k, rank, world, (n + world - 1) // world, n, 0.5, device=local)
        shard.tab.set_stream(torch.cuda.current_stream().cuda_stream)
        comm = sh.TorchComm(shard)
        comm.connect()
        dev = torch.from_numpy(np.ascontiguousarray(d.pairs(lo, hi - lo)).reshape(-1)).cuda() if hi > lo else torch.empty(1, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        for rep in range(3):                                   # the same handles step after step
            comm.begin()
            sh.sharded_insert(comm, [(dev.data_ptr(), hi - lo)])
            sh.sharded_assemble(comm)                          # raises ShardedError (on every rank) if any rank flagged one
            got, nc, nn = shard.result_host()
            ok = nc == want_nc and got.tobytes() == want
            failures += 0 if ok else 1
        st = shard.tab.stats()
        tot = torch.tensor([float(st["n_inserted"]), float(nn)], device=red_dev, dtype=torch.float64)
        dist.all_reduce(tot)
        if int(tot[0]) != n or int(tot[1]) != n:
            failures += 1
        print(f"rank {rank}/{world} k={k}: {st['n_inserted']} k-mers held, {nc} contigs, {'OK' if ok else 'MISMATCH'}", flush=True)
        shard.close()
        dist.barrier()
    t = torch.tensor([float(failures)], device=red_dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
