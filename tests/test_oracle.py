"""CPU: pin the oracle (oracle/kmer_oracle.c) to the reference.

  * golden vectors produced by the UNMODIFIED reference (tests/golden/, see make_golden.py),
    including the reference's only in-repo known answer (README.md:27, k=3);
  * live differential runs against oracle/_ref when it is present (build container);
  * the generator's expected output against both.
"""
import hashlib
import os
import tempfile

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, golden_cases
from tools import kmergen

CASES = golden_cases()


def _read(case, ext):
    with open(os.path.join(GOLDEN, f"{case}.{ext}"), "rb") as f:
        return f.read()


@pytest.mark.parametrize("case", sorted(CASES))
def test_golden_files_intact(case):
    for ext in ("txt", "dat", "probe"):
        assert hashlib.sha256(_read(case, ext)).hexdigest() == CASES[case][ext + "_sha256"]


def test_readme_known_answer():
    # README.md:27: the three contigs of the k=3 example graph
    out = oracle.assemble_text(_read("readme_k3", "txt"), 3)[0]
    assert sorted(out.split()) == [b"AACCG", b"AATGC", b"GATCTGA"]
    assert out == _read("readme_k3", "dat")


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference_contigs(case):
    k = CASES[case]["k"]
    out = oracle.assemble_text(_read(case, "txt"), k)[0]
    assert out == _read(case, "dat")          # same bytes, same (start-line) order


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference_records(case):
    """kmer_pair bytes and next_kmer() bytes, record by record (oracle/ref_probe.cpp output)."""
    k = CASES[case]["k"]
    pairs = oracle.parse_lines(_read(case, "txt"), k)
    probe = _read(case, "probe").decode().split("\n")[:-1]
    assert len(probe) == min(200, pairs.shape[0])
    for rec, line in zip(pairs, probe):
        want_pair, want_next = line.split(" ")
        assert rec.tobytes().hex() == want_pair
        if want_next != "-":
            assert oracle.next_kmer(rec.tobytes(), k).hex() == want_next
        else:
            assert rec[-1] == ord("F")


def test_pack_layout():
    # packing.hpp:50-92: first base in the top two bits, A=0 C=1 G=2 T=3, tail padded with A
    assert oracle.pack_kmer("GAT") == bytes([0b10001100])
    assert oracle.pack_kmer("ACGTT") == bytes([0b00011011, 0b11000000])
    for k in (3, 19, 31, 32, 51, 64):
        rng = np.random.default_rng(k)
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, k))
        assert oracle.unpack_kmer(oracle.pack_kmer(s), k) == s
    with pytest.raises(RuntimeError):
        oracle.pack_kmer("ACGN")


def test_find_and_overwrite_semantics():
    # hash_map.hpp:33-35: map[key] = value -> the last writer wins; find: hash_map.hpp:85-92
    k = 19
    a = oracle.parse_lines(b"ACGTACGTACGTACGTACG AC\n", k)
    b = oracle.parse_lines(b"ACGTACGTACGTACGTACG GT\n", k)
    t = oracle.Table(k, 4)
    t.insert_pairs(a)
    t.insert_pairs(b)
    assert len(t) == 1
    assert t.find(a[0, :5].tobytes()) == b[0].tobytes()
    assert t.find(oracle.pack_kmer("T" * 19)) is None


def test_missing_successor_is_an_error():
    # kmer_hash.cpp:47-49
    with pytest.raises(RuntimeError, match="not found"):
        oracle.assemble_text(b"ACGTACGTACGTACGTACG FC\n", 19)


def test_empty_input():
    assert oracle.assemble_text(b"", 19) == [b""]


@pytest.mark.parametrize("k,n,c,longn", [(19, 20000, 37, 0), (31, 6000, 500, 0), (51, 5000, 50, 0),
                                         (19, 9000, 4, 8000), (21, 500, 500, 0)])
def test_generator_agrees_with_oracle(k, n, c, longn):
    d = kmergen.Dataset(k, n, c, seed=k + n, long_nodes=longn)
    text = d.text()
    assert (oracle.parse_lines(text, k) == d.pairs()).all()
    for nranks in (1, 3):
        outs = oracle.assemble_text(text, k, nranks)
        assert outs == [d.expected(nranks, r)[0] for r in range(nranks)]
    exp = d.expected()[0]
    assert sorted(exp.split(b"\n")[:-1]) == d.solution().split(b"\n")[:-1]
    assert kmergen.digest_lines(exp) == d.digest()
    if longn:
        assert d.max_contig_nodes >= longn
    # every k-mer unique
    assert len(set(text.tobytes()[i:i + k] for i in range(0, text.size, k + 4))) == n


@pytest.mark.skipif(oracle.ref_binary(19) is None, reason="oracle/_ref not built here")
@pytest.mark.parametrize("k,n,c", [(19, 50000, 60), (31, 20000, 300), (51, 20000, 150)])
def test_oracle_matches_reference_live(k, n, c):
    d = kmergen.Dataset(k, n, c, seed=1000 + k)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "in.txt")
        d.text().tofile(p)
        ref = oracle.run_reference(k, p, tmp)
    assert ref == oracle.assemble_text(d.text(), k)[0] == d.expected()[0]
