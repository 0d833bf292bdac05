"""Short driver for ncu: `steps` x (clear, insert resident records, assemble) of one bench workload on one GPU."""
import sys

sys.path.insert(0, ".")
import torch

import cs267_hw3_b200 as kh
from bench import WORKLOADS
from tools import kmergen

workload = sys.argv[1] if len(sys.argv) > 1 else "chr14_k51"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k, n, c, longn = WORKLOADS[workload]
if len(sys.argv) > 3:
    c = max(1, int(int(sys.argv[3]) * c / n)); n = int(sys.argv[3])
d = kmergen.Dataset(k, n, c, seed=267, long_nodes=longn)
pb = kh.pair_bytes(k)
host = kh.PinnedBuffer(n * pb)
d.pairs_into(host.ptr, 0, n)
dev = torch.empty(n * pb, dtype=torch.uint8, device="cuda")
dev.copy_(torch.from_numpy(host.array))
torch.cuda.synchronize()
tab = kh.KmerHashTable(k, n, 0.5, device=0)
for _ in range(steps):
    tab.clear()
    tab.insert_pairs_device(dev.data_ptr(), n)
    out = tab.assemble_device()
    st = tab.stats()
    print({x: round(st[x], 3) for x in ("ms_insert", "ms_stage", "ms_build", "ms_walk", "ms_rank", "ms_emit")}, out[2:], st["n_segments"])
tab.close()
