#!/usr/bin/env python
"""Run the sharded path with all ranks in ONE process on one GPU (LocalComm) so that its kernels can be put
under ncu (a multi-rank job cannot).  python tools/probes/profile_sharded_local.py [world] [n_kmers] [k]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import cs267_hw3_b200 as kh
from cs267_hw3_b200 import sharded as sh
from tools import kmergen

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 19
c = max(1, n // 104)
d = kmergen.Dataset(k, n, c, seed=267)
pairs = d.pairs()
n_local_max = (n + world - 1) // world
shards = [sh.Shard(k, r, world, n_local_max, n, 0.5, device=0, n_starts_max=int(c / world * 1.3) + 4096) for r in range(world)]
comm = sh.LocalComm(shards)
comm.connect()
L = kh.lib()
blocks = []
for r, s in enumerate(shards):
    lo, hi = sh.block_of_rank(n, world, r)
    p = C.c_void_p()
    assert L.kh_device_alloc(C.byref(p), max(1, (hi - lo) * pairs.shape[1])) == 0
    blk = np.ascontiguousarray(pairs[lo:hi])
    s.tab._check(L.kh_copy_device(s.tab._h, p, blk.ctypes.data, blk.nbytes))
    s.tab.sync()
    blocks.append((p.value, hi - lo))
for rep in range(2):
    for s in shards:
        s.tab.clear()
    comm.barrier()
    t0 = time.perf_counter()
    sh.sharded_insert(comm, blocks)
    t1 = time.perf_counter()
    timings = {}
    rounds = sh.sharded_assemble(comm, timings=timings)
    t2 = time.perf_counter()
    print(f"rep {rep}: insert {1e3 * (t1 - t0):.2f} ms, assemble {1e3 * (t2 - t1):.2f} ms (all {world} ranks serialised on one GPU), "
          f"rounds {rounds}, phases {dict((a, round(b, 2)) for a, b in timings.items())}")
ok = all(s.result_host()[0].tobytes() == d.expected(world, r)[0] for r, s in enumerate(shards))
print("verified", ok, "stats", [(s.tab.stats()["n_inserted"], s.tab.stats()["n_segments"]) for s in shards])
