#!/usr/bin/env python
"""Regenerate tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile).

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tests/golden/make_golden.py

For every case it writes
    <case>.txt     k-mer file in the reference text format (input)
    <case>.dat     what oracle/_ref/kmer_hash_ref_<K> wrote to <prefix>_0.dat (contigs, start order)
    <case>.probe   first 200 lines through oracle/_ref/ref_probe_<K>: hex(kmer_pair bytes) hex(next_kmer)
and MANIFEST.json (k, n, c, sha256 of each file).  The GPU box has no /root/reference;
tests read only these files.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tools import kmergen  # noqa: E402

# README.md:27 (Figure 2): k = 3, contigs GATCTGA, AACCG, AATGC.  The eleven 3-mers and
# their extensions follow from the contigs; the line order here is arbitrary but fixed.
README_K3 = ["TCT AG", "AAC FC", "GAT FC", "CCG AF", "ATC GT", "AAT FG", "CTG TA", "TGC AF", "ACC AG", "ATG AC", "TGA CF"]

CASES = [  # name, k, n, c, seed, long_nodes
    ("k19_a", 19, 3000, 25, 19, 0),
    ("k19_singletons", 19, 64, 64, 3, 0),
    ("k19_long", 19, 4000, 3, 4, 3500),
    ("k31_a", 31, 2000, 40, 31, 0),
    ("k51_a", 51, 1500, 30, 51, 0),
]


def run_ref(k, path, case, manifest, n, c):
    with tempfile.TemporaryDirectory() as tmp:
        dat = oracle.run_reference(k, path, tmp, prefix="g")
    with open(os.path.join(HERE, case + ".dat"), "wb") as f:
        f.write(dat)
    with open(path, "rb") as f:
        head = b"".join(f.readlines()[:200])
    pr = subprocess.run([os.path.join(oracle.REF_DIR, f"ref_probe_{k}")], input=head, capture_output=True, check=True)
    with open(os.path.join(HERE, case + ".probe"), "wb") as f:
        f.write(pr.stdout)
    entry = {"k": k, "n": n, "c": c}
    for ext in ("txt", "dat", "probe"):
        with open(os.path.join(HERE, f"{case}.{ext}"), "rb") as f:
            entry[ext + "_sha256"] = hashlib.sha256(f.read()).hexdigest()
    manifest[case] = entry


def main():
    oracle.build()
    manifest = {}
    p = os.path.join(HERE, "readme_k3.txt")
    with open(p, "w") as f:
        f.write("".join(s + "\n" for s in README_K3))
    run_ref(3, p, "readme_k3", manifest, 11, 3)
    for name, k, n, c, seed, longn in CASES:
        d = kmergen.Dataset(k, n, c, seed=seed, long_nodes=longn, threads=1)
        p = os.path.join(HERE, name + ".txt")
        with open(p, "wb") as f:
            f.write(d.text().tobytes())
        run_ref(k, p, name, manifest, n, c)
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", ", ".join(sorted(manifest)))


if __name__ == "__main__":
    main()
