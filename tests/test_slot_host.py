"""CPU: the __host__ __device__ helpers of cs267_hw3_b200/csrc/slot.cuh, built host-only with nvcc (no GPU needed).
Pins the pieces the kernels share with the host -- the branch-free extension-letter table, kmer_pair <-> slot
conversion (packing.hpp:50-92 / kmer_t.hpp:43-57 semantics), next/previous k-mer, and the owner function -- against
the Python mirrors the gloo tests use (cs267_hw3_b200.sharded) and against plain arithmetic on the base string."""
import os
import shutil
import subprocess

import pytest

from cs267_hw3_b200 import sharded as sh
from tools import kmergen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "F": 4}


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("slot") / "slot_host_check")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "slot_host_check.cu")], check=True, capture_output=True)
    return exe


def _run(exe, k, world, records):
    text = f"{k} {world}\n" + "\n".join(r.hex() for r in records) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
    return [int(x) for x in out[0].split()], [ln.split() for ln in out[1:]]


def _random_records(k, n, seed):
    """kmer_pair bytes built by hand (packing.hpp:50-92: 2 bits per base, first base in the top bits of byte 0,
    last byte padded with A; kmer_t.hpp:43-45: backward then forward extension letter)."""
    import random
    rng = random.Random(seed)
    pl = (k + 3) // 4
    out = []
    for _ in range(n):
        bases = [rng.randrange(4) for _ in range(k)]
        bits = 0
        for b in bases:
            bits = (bits << 2) | b
        bits <<= 8 * pl - 2 * k
        out.append(bits.to_bytes(pl, "big") + bytes([ord(rng.choice("ACGTF")), ord(rng.choice("ACGTF"))]))
    return out


def test_extension_letter_table(checker):
    ext, _ = _run(checker, 19, 1, [])
    assert len(ext) == 256
    for c in range(256):
        assert ext[c] == CODE.get(chr(c), 7), f"ext_code({c:#x})"


@pytest.mark.parametrize("k", [2, 3, 15, 16, 19, 29, 30, 31, 51, 61])
def test_slot_conversion_and_neighbours(checker, k):
    n = 300
    pl = (k + 3) // 4
    world = 5
    records = _random_records(k, n, seed=k)
    _, rows = _run(checker, k, world, records)
    assert len(rows) == n
    mask = (1 << (2 * k)) - 1
    for raw, row in zip(records, rows):
        slot, owner, nxt, prv, ok, roundtrip = int(row[0], 16), int(row[1]), int(row[2], 16), int(row[3], 16), row[4], row[5]
        assert ok == "1" and roundtrip == "1"
        key = int.from_bytes(raw[:pl], "big") >> (8 * pl - 2 * k)              # packing.hpp:50-92: MSB-first, A-padded tail
        back, fwd = CODE[chr(raw[pl])], CODE[chr(raw[pl + 1])]
        assert slot == (key << 6) | (back << 3) | (fwd + 1)
        assert slot == sh.slot_from_pair(raw, k)
        assert owner == sh.owner_of_slot(slot, k, world)
        if fwd < 4:                                                            # kmer_t.hpp:51-53
            assert nxt == (((key << 2) | fwd) & mask) << 6
        if back < 4:                                                           # kmer_t.hpp:55-57
            assert prv == ((key >> 2) | (back << (2 * (k - 1)))) << 6


def test_bad_letters_are_flagged(checker):
    k, pl = 19, 5
    d = kmergen.Dataset(k, 4, 1, seed=3)
    recs = [bytearray(r.tobytes()) for r in d.pairs()]
    recs[0][pl] = ord("N")
    recs[1][pl + 1] = ord("a")
    recs[2][pl] = 0
    _, rows = _run(checker, k, 2, [bytes(r) for r in recs])
    assert [row[4] for row in rows] == ["0", "0", "0", "1"]
