#!/usr/bin/env python
"""Full-size digests written by the UNMODIFIED reference (oracle/_ref) -- run in the build container.

    python tests/golden/make_full_digests.py [19 51]

For each K the chr14-shape file of BASELINE.json configs[1]/[2] (89 710 742 unique k-mers, 860 329 contigs,
seed 267, tools/kmer_gen.cpp) is written to .scratch/, oracle/_ref/kmer_hash_ref_<K> runs on it in `test` mode
(one rank), and the sha256 / size / line count of the `<prefix>_0.dat` it wrote go to
tests/golden/chr14_full.json together with the reference's own timer lines from a second, untimed-output run
(`Finished inserting` / `Assembled in`, kmer_hash.cpp:144-145).  The GPU tests compare the bytes of the CUDA
path's output with that sha256 (tests/test_gpu_parity.py::test_full_chr14_shape_properties), so the full-size
shapes are pinned to the reference itself and not only to the generator's own solution.
Needs ~35 GB of RAM and ~10 GB of scratch disk per K; takes several minutes per K.
"""
import hashlib
import json
import os
import resource
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tools import kmergen  # noqa: E402

N, C, SEED = 89_710_742, 860_329, 267
OUT = os.path.join(HERE, "chr14_full.json")


def sha_file(path):
    h, size, lines = hashlib.sha256(), 0, 0
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b); size += len(b); lines += b.count(b"\n")
    return h.hexdigest(), size, lines


def main():
    ks = [int(a) for a in sys.argv[1:]] or [19, 51]
    work = os.path.join(ROOT, ".scratch")
    os.makedirs(work, exist_ok=True)
    try:
        with open(OUT) as f:
            res = json.load(f)
    except Exception:
        res = {}
    for k in ks:
        d = kmergen.Dataset(k, N, C, seed=SEED)
        path = os.path.join(work, f"chr14_k{k}.txt")
        t0 = time.time()
        with open(path, "wb") as f:
            step = 4_000_000
            for lo in range(0, N, step):
                d.text(lo, min(step, N - lo)).tofile(f)
        gen_digest = d.digest()
        want, _ = d.expected_array()
        gen_sha = hashlib.sha256(want.tobytes()).hexdigest()
        del want
        d.close()
        print(f"K={k}: text written in {time.time() - t0:.0f} s", flush=True)
        exe = oracle.ref_binary(k)
        t0 = time.time()
        r = subprocess.run([exe, path, "test", f"full{k}"], cwd=work, capture_output=True, text=True)
        wall_test = time.time() - t0
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        dat = os.path.join(work, f"full{k}_0.dat")
        sha, size, lines = sha_file(dat)
        os.unlink(dat)
        t0 = time.time()
        ins, tot = oracle.run_reference(k, path, work, test=False)
        wall_time = time.time() - t0
        rss = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
        os.unlink(path)
        res[f"chr14_k{k}"] = {
            "k": k, "n_kmers": N, "n_contigs": C, "seed": SEED,
            "dat_sha256": sha, "dat_bytes": size, "dat_lines": lines,
            "generator_expected_sha256": gen_sha, "generator_digest": list(gen_digest),
            "matches_generator": sha == gen_sha,
            "reference_insert_s": ins, "reference_total_s": tot, "reference_wall_s": wall_time,
            "reference_test_mode_metrics_line": r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "",
            "reference_max_rss_kb": rss, "host": "build container, 1 of 8 vCPUs (Intel Xeon VM), g++ -O2",
            "binary": f"oracle/_ref/kmer_hash_ref_{k} = /root/reference/kmer_hash.cpp unmodified + single-rank UPC++ stand-in",
        }
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
        print(f"K={k}: sha256 {sha} ({size} B, {lines} lines), matches generator: {sha == gen_sha}; "
              f"reference insert {ins:.1f} s total {tot:.1f} s, wall {wall_test:.0f}+{wall_time:.0f} s", flush=True)


if __name__ == "__main__":
    main()
