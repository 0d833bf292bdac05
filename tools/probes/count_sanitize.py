"""Small end-to-end pass over every kh_count_* kernel, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/probes/count_sanitize.py
    compute-sanitizer --tool racecheck python tools/probes/count_sanitize.py
Counts (host path with table growth, device path on an unaligned pointer), extracts records and lines, looks counters up,
and feeds the records to insert + traverse; checks the k-mer set against the generator."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cs267_hw3_b200 as kh  # noqa: E402
from tools import kmergen, readgen  # noqa: E402

L = kh.lib()
for k in (19, 51):
    d = kmergen.Dataset(k, 30000, 60, seed=9)
    reads = readgen.tile_reads(d.solution(), k, read_len=k + 40, coverage=3, seed=10)
    want = d.pairs()
    want = want[np.lexsort(want.T[::-1])]
    with kh.KmerCounter(k, 100, 0.5, device=0) as kc:                 # far too small: grows
        kc.count_reads(reads)
        got = kc.extract(2, 2)
        assert (got[np.lexsort(got.T[::-1])] == want).all()
        lines = kc.extract_lines(2, 2)
        assert lines.size == 30000 * (k + 4)
        cnt = kc.lookup(want[:1000, : (k + 3) // 4])
        assert (cnt[:, 0] >= 3).all()
        assert kc.stats()["n_grows"] > 0
    with kh.KmerCounter(k, 60000, 0.5, device=0) as kc, kh.KmerHashTable(k, 30000, 0.5, device=0) as tab:
        p = C.c_void_p()
        assert L.kh_device_alloc(C.byref(p), reads.size + 16) == 0
        tab._check(L.kh_copy_device(tab._h, p.value + 3, reads.ctypes.data, reads.size))
        tab.sync()
        kc.count_reads_device(p.value + 3, reads.size)
        ptr, n = kc.extract_device(2, 2)
        assert n == 30000
        tab.insert_pairs_device(ptr, n)
        buf, offs, nodes = tab.assemble()
        assert nodes == 30000 and b"\n".join(sorted(buf.tobytes().split(b"\n")[:-1])) + b"\n" == d.solution()
        L.kh_device_free(p)
    print(f"k={k} ok")
