"""GPU: the kmer_hash_<K> command line end to end, the way scripts/check_it.sh uses it."""
import os
import re
import subprocess

import pytest

import oracle
from conftest import ROOT
from tools import kmergen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bins():
    from cs267_hw3_b200 import build
    return dict(zip((19, 31, 51), build.build_cli()))


@pytest.mark.parametrize("k,n,c", [(19, 200000, 300), (31, 100000, 900), (51, 150000, 1500)])
@pytest.mark.parametrize("gpu_sort", [False, True])
def test_check_script_passes(bins, tmp_path, k, n, c, gpu_sort):
    """tools/check.sh = scripts/check_it.sh:32,47-59 around our binary: sorted output == solution."""
    gen = os.path.join(ROOT, "tools", "gen_kmers")
    if not os.path.exists(gen):
        kmergen.build()
    inp = tmp_path / "synth.txt"
    subprocess.run([gen, str(k), str(n), str(c), str(inp), "--seed", "11"], check=True, capture_output=True)
    env = dict(os.environ, KH_GPU_SORT="1") if gpu_sort else dict(os.environ)
    r = subprocess.run([os.path.join(ROOT, "tools", "check.sh"), str(inp)], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"PASSED: {inp}" in r.stdout
    # the metrics line of test mode, format of kmer_hash.cpp:71-78
    assert re.search(rf"Rank 0 reconstructed {c} contigs with {n} nodes from 0 start nodes\. "
                     r"\(\d+\.\d{6} read, \d+\.\d{6} insert, \d+\.\d{6} total\)", r.stdout)


def test_cli_output_equals_reference_bytes(bins, tmp_path):
    """Same file the unmodified reference writes (start-node order, not just the sorted set)."""
    k = 19
    d = kmergen.Dataset(k, 60000, 200, seed=21)
    inp = tmp_path / "in.txt"
    d.text().tofile(inp)
    subprocess.run([bins[k], str(inp), "test", "mine"], cwd=tmp_path, check=True, capture_output=True)
    mine = (tmp_path / "mine_0.dat").read_bytes()
    assert mine == d.expected()[0]
    if oracle.ref_binary(k):
        assert mine == oracle.run_reference(k, str(inp), str(tmp_path))


def test_cli_timing_lines(bins, tmp_path):
    d = kmergen.Dataset(51, 50000, 100, seed=22)
    inp = tmp_path / "in.txt"
    d.text().tofile(inp)
    r = subprocess.run([bins[51], str(inp), "verbose"], cwd=tmp_path, capture_output=True, text=True, check=True)
    lines = r.stdout.strip().split("\n")
    assert lines[0] == "Initializing hash table of size 100000 for 50000 kmers."     # kmer_hash.cpp:115
    assert lines[1] == "Finished reading kmers."                                      # :124
    assert re.fullmatch(r"Finished inserting in \d+\.\d{6} sec", lines[2])            # :144
    assert re.fullmatch(r"Assembled in \d+\.\d{6} total", lines[3])                   # :145


def test_cli_missing_kmer_aborts_like_the_reference(bins, tmp_path):
    inp = tmp_path / "bad.txt"
    inp.write_bytes(b"ACGTACGTACGTACGTACG FC\n")
    r = subprocess.run([bins[19], str(inp), "test"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == -6 and "Error: k-mer not found in Distributed HashMap." in r.stderr   # kmer_hash.cpp:47-49


@pytest.mark.parametrize("k,ranks", [(19, 2), (19, 5), (51, 3), (31, 8)])
def test_cli_multi_rank_writes_the_reference_per_rank_files(bins, tmp_path, k, ranks):
    """KH_RANKS=P: rank r's file holds the contigs whose start line lies in its block of the input, in input order
    (kmer_hash.cpp:27-31,64-67; read_kmers.hpp:55-58) -- and check.sh accepts the concatenation."""
    d = kmergen.Dataset(k, 150000, 700, seed=31 + ranks)
    inp = tmp_path / "synth.txt"
    d.text().tofile(inp)
    (tmp_path / "synth_solution.txt").write_bytes(d.solution())
    env = dict(os.environ, KH_RANKS=str(ranks))
    r = subprocess.run([bins[k], str(inp), "test", "mr"], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    for rank in range(ranks):
        assert (tmp_path / f"mr_{rank}.dat").read_bytes() == d.expected(ranks, rank)[0]
    n0 = d.expected(ranks, 0)[1]
    assert re.search(rf"Rank 0 reconstructed {n0} contigs with \d+ nodes from 0 start nodes\.", r.stdout)
    r = subprocess.run([os.path.join(ROOT, "tools", "check.sh"), str(inp)], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0 and f"PASSED: {inp}" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("chunk_lines", ["1000", "4194304"])
def test_cli_streamed_ingest_writes_the_same_bytes(bins, tmp_path, chunk_lines):
    """KH_STREAM=1: the file is streamed through pinned chunk buffers into pack + insert (DistributedHashMap::insert_file).
    Same table, same start-node order, hence the same output bytes as the read-everything-first path."""
    k = 19
    d = kmergen.Dataset(k, 60000, 200, seed=23)
    inp = tmp_path / "in.txt"
    d.text().tofile(inp)
    env = dict(os.environ, KH_STREAM="1", KH_STREAM_CHUNK_LINES=chunk_lines)
    subprocess.run([bins[k], str(inp), "test", "st"], cwd=tmp_path, check=True, capture_output=True, env=env)
    assert (tmp_path / "st_0.dat").read_bytes() == d.expected()[0]


@pytest.mark.parametrize("case,k", [("k19_a", 19), ("k51_a", 51)])
def test_unmodified_reference_main_over_the_drop_in_headers(tmp_path, case, k):
    """the reference's kmer_hash.cpp, compiled UNMODIFIED against include/ (hash_map.hpp,
    kmer_t.hpp, read_kmers.hpp, butil.hpp and the one-process upcxx/upcxx.hpp) and libkh_b200.so -- built where the
    reference tree exists (oracle/Makefile `dropin`), shipped as oracle/_ref/kmer_hash_dropin_<K>.  Its own
    assemble_contigs loop (one DistributedHashMap::find per step, kmer_hash.cpp:38-55) over the GPU table must write the
    bytes the reference wrote with its own hash map (tests/golden/<case>.dat)."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"kmer_hash_dropin_{k}")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/kmer_hash_dropin_* is built only where /root/reference exists (make -C oracle dropin)")
    inp = os.path.join(ROOT, "tests", "golden", case + ".txt")
    r = subprocess.run([exe, inp, "test", "dd"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    with open(os.path.join(ROOT, "tests", "golden", case + ".dat"), "rb") as f:
        assert (tmp_path / "dd_0.dat").read_bytes() == f.read()
    assert re.search(r"Rank 0 reconstructed \d+ contigs with \d+ nodes from 0 start nodes\.", r.stdout)
    r = subprocess.run([exe, inp], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Finished inserting in" in r.stdout and "Assembled in" in r.stdout


@pytest.mark.parametrize("k,case", [(19, "k19_a"), (31, "k31_a"), (51, "k51_a")])
def test_header_api_find_and_hashmap(tmp_path, k, case):
    """hash_map.hpp:50-55, 83 and README.md:95-99 through the drop-in headers: insert_all(vector), find(std::string),
    the reference's own one-find-per-step walk, HashMap::insert / find(pkmer_t) -- tests/native/hash_map_check.cpp."""
    exe = tmp_path / "hm_check"
    pkg = os.path.join(ROOT, "cs267_hw3_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", f"-DKMER_LEN={k}", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "hash_map_check.cpp"), "-L" + pkg, "-lkh_b200", "-Wl,-rpath," + pkg,
                    "-o", str(exe)], check=True, capture_output=True)
    r = subprocess.run([str(exe), os.path.join(ROOT, "tests", "golden", case + ".txt")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


@pytest.mark.parametrize("k,ranks", [(19, 1), (51, 1), (51, 3), (19, 2)])
def test_cli_writes_the_sorted_solution_itself(bins, tmp_path, k, ranks):
    """Output side (scripts/check_it.sh:47-48, `cat test*.dat | sort`): with KH_SOLUTION the binary sorts every rank's
    contigs on its GPU (kh_sorted_order), merges the per-rank lists and writes the file check_it.sh would diff."""
    d = kmergen.Dataset(k, 120_000, 900, seed=40 + k + ranks)
    inp = tmp_path / "in.txt"
    d.text().tofile(inp)
    env = dict(os.environ, KH_SOLUTION=str(tmp_path / "sol.txt"), KH_RANKS=str(ranks))
    subprocess.run([bins[k], str(inp), "test", "so"], cwd=tmp_path, check=True, capture_output=True, env=env)
    lines = b"".join((tmp_path / f"so_{r}.dat").read_bytes() for r in range(ranks)).splitlines(keepends=True)
    assert (tmp_path / "sol.txt").read_bytes() == b"".join(sorted(lines)) == d.solution()      # the generator's sorted solution file


def test_sorted_order_breaks_ties_behind_the_prefix_key(tmp_path):
    """kh_sorted_order sorts a 21-character prefix key on the GPU; contigs that agree on it are finished on the host."""
    from cs267_hw3_b200 import build as b
    k = 23
    exe = b.build_cli(ks=(k,))[0]
    p = "ACGTTGCAAGGCTTACGATCGA"                                    # 22 bases shared by all three contigs
    kmers = [p + "T", p + "C", p + "G", "T" * 22 + "A", "A" * 23]
    text = "".join(f"{x} FF\n" for x in kmers)
    inp = tmp_path / "ties.txt"
    inp.write_text(text)
    env = dict(os.environ, KH_SOLUTION=str(tmp_path / "sol.txt"))
    subprocess.run([exe, str(inp), "test", "ti"], cwd=tmp_path, check=True, capture_output=True, env=env)
    assert (tmp_path / "sol.txt").read_text().split() == sorted(kmers)
    assert (tmp_path / "ti_0.dat").read_text().split() == kmers
