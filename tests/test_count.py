"""k-mer analysis (kh_count_*): reads -> unique k-mers with backward / forward extensions -- the stage whose output
the reference reads from its input file (README.md:19-21, read_kmers.hpp:54-79).

CPU: the oracle (oracle/kmer_count_oracle.c: sort-based) reproduces the generator's k-mer records from tiled reads; the
__host__ __device__ functions the kernels are made of (csrc/count_core.cuh), run serially by tests/native/
count_host_check.cu with the kernels' own loops and the host chunking, agree with the oracle on records AND counters.
GPU: the same comparisons through the C ABI, and the closed loop reads -> kh_count -> kh_insert_pairs_device ->
kh_assemble == the generator's solution (which the unmodified reference reproduces from the k-mer file)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from tools import kmergen, readgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sorted_rows(a: np.ndarray) -> np.ndarray:
    return a[np.lexsort(a.T[::-1])] if a.size else a


def _reads(k, n, c, seed, read_len=None, coverage=3, error_rate=0.0, with_n=False):
    d = kmergen.Dataset(k, n, c, seed=seed)
    reads = readgen.tile_reads(d.solution(), k, read_len=read_len or (k + 40), coverage=coverage, seed=seed + 1,
                               error_rate=error_rate)
    if with_n:                                   # an N inside a read splits it; lower case is not a base either
        rng = np.random.default_rng(seed + 2)
        pos = rng.choice(reads.size, size=max(1, reads.size // 500), replace=False)
        reads[pos] = np.where(rng.random(pos.size) < 0.5, ord("N"), ord("a")).astype(np.uint8)
    return d, reads


# ------------------------------------------------------------------------------------------------ CPU: the oracle
@pytest.mark.parametrize("k", [19, 31, 51])
def test_oracle_recovers_the_generators_kmer_records(k):
    """Error-free reads tiled over the contigs, min_count = min_ext = 2: exactly the generator's k-mers, each with the
    extensions of its line in the k-mer file ('F' at contig ends) -- the records read_kmers would parse."""
    d, reads = _reads(k, 30000, 60, seed=11)
    pairs, counts, n_occ = oracle.analyse_reads(reads, k, 2, 2)
    assert (pairs == _sorted_rows(d.pairs())).all()
    assert counts[:, 0].min() >= 3 and n_occ == int(counts[:, 0].sum())        # nothing saturates at this coverage


def test_oracle_readme_example():
    """README.md:27: contigs GATCTGA, AACCG, AATGC at k = 3 -> the 11 3-mers of Figure 2 with their extensions."""
    reads = b"GATCTGA\nGATCTGA\nAACCG\nAACCG\nAATGC\nAATGC\n"
    pairs, counts, _ = oracle.analyse_reads(reads, 3, 2, 2)
    got = {oracle.unpack_kmer(bytes(p[:1]), 3) + chr(p[1]) + chr(p[2]) for p in pairs}
    want = {"GATFC", "ATCGT", "TCTAG", "CTGTA", "TGACF", "AACFC", "ACCAG", "CCGAF", "AATFG", "ATGAC", "TGCAF"}
    assert got == want and (counts[:, 0] == 2).all()


def test_oracle_thresholds_and_forks():
    # ACGTA seen 3x followed by C, once followed by G: min_ext 2 -> 'C'; min_ext 1 -> fork -> 'F'
    reads = b"TACGTAC\nTACGTAC\nTACGTAC\nTACGTAG\n"
    for min_ext, want in ((2, "C"), (1, "F"), (4, "F")):
        pairs, _, _ = oracle.analyse_reads(reads, 5, 1, min_ext)
        rec = [p for p in pairs if oracle.unpack_kmer(bytes(p[:2]), 5) == "ACGTA"][0]
        assert chr(rec[3]) == want and chr(rec[2]) == "T"                        # T precedes it in all four reads
    pairs, _, _ = oracle.analyse_reads(reads, 5, 2, 2)                           # CGTAG occurs once: dropped by min_count 2
    assert "CGTAG" not in {oracle.unpack_kmer(bytes(p[:2]), 5) for p in pairs}


# ------------------------------------------------------------------------------- CPU: the kernels' functions, serially
@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("count") / "count_host_check")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "count_host_check.cu")], check=True, capture_output=True)
    return exe


def _run_host(exe, tmp_path, reads, k, chunk, n_slots, min_count, min_ext, misalign=0, grow_after=None):
    f = tmp_path / "reads.txt"
    np.asarray(reads, dtype=np.uint8).tofile(f)
    out = subprocess.run([exe, str(k), str(chunk), str(n_slots), str(min_count), str(min_ext), str(misalign), str(f)]
                         + ([str(grow_after)] if grow_after is not None else []),
                         capture_output=True, text=True, check=True).stdout.splitlines()
    head = [int(x) for x in out[0].split()]
    pb = (k + 3) // 4 + 2
    recs = np.array([list(bytes.fromhex(ln.split()[0])) for ln in out[1:]], dtype=np.uint8).reshape(-1, pb)
    cnts = np.array([[int(x) for x in ln.split()[1:10]] for ln in out[1:]], dtype=np.uint32).reshape(-1, 9)
    for ln in out[1:]:                                  # kc_record_to_line: the record as a line of the reference's k-mer file
        rec, text = bytes.fromhex(ln.split()[0]), ln.split()[10]
        assert text == oracle.unpack_kmer(rec[:-2], k) + "_" + chr(rec[-2]) + chr(rec[-1])
    order = np.lexsort(recs.T[::-1]) if recs.size else np.zeros(0, dtype=np.int64)
    return head, recs[order], cnts[order]


@pytest.mark.parametrize("k", [2, 3, 15, 16, 17, 19, 31, 32, 33, 47, 51, 61])
def test_kernel_functions_match_the_oracle(host_check, tmp_path, k):
    """Every K regime: 64-bit tag slots (K <= 31), 128-bit keys (K >= 32), the 2K <= 64 / > 64 window branches, K + 2 = 63
    characters per window (K = 61); small chunks so chunk and tile borders fall inside reads; substitution errors and
    N / lower-case separators; min_count and min_ext above 1."""
    n = 400 if k < 8 else 12000
    c = 4 if k < 8 else 30
    if k < 8:                                     # 4^k distinct k-mers at most: random text instead of a unique-k-mer set
        rng = np.random.default_rng(k)
        reads = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 6000)].copy()
        reads[rng.choice(reads.size, 80, replace=False)] = 10
    else:
        _, reads = _reads(k, n, c, seed=100 + k, coverage=3, error_rate=0.002, with_n=True)
    want_p, want_c, n_occ = oracle.analyse_reads(reads, k, 2, 2)
    head, recs, cnts = _run_host(host_check, tmp_path, reads, k, chunk=4096, n_slots=max(1024, 4 * len(reads) // 3), min_count=2, min_ext=2)
    assert head[0] == n_occ and head[2] == len(want_p) and head[3] == 0
    assert (recs == want_p).all()
    assert (cnts == want_c).all()


@pytest.mark.parametrize("k,misalign", [(19, 3), (51, 9), (31, 15)])
def test_kernel_functions_unaligned_buffer(host_check, tmp_path, k, misalign):
    """kh_count_reads_device takes any pointer: the byte-wise tile loader."""
    _, reads = _reads(k, 6000, 20, seed=7 + k, with_n=True)
    want_p, want_c, n_occ = oracle.analyse_reads(reads, k, 1, 1)
    head, recs, cnts = _run_host(host_check, tmp_path, reads, k, chunk=1 << 30, n_slots=1 << 16, min_count=1, min_ext=1, misalign=misalign)
    assert head[0] == n_occ and (recs == want_p).all() and (cnts == want_c).all()


def test_kernel_functions_saturate_like_the_oracle(host_check, tmp_path):
    """A homopolymer run and a 300x repeated read: occurrences stop at 255, per-base observations at 127."""
    reads = np.frombuffer(b"A" * 700 + b"\n" + b"CGTACGTTAGC\n" * 300, dtype=np.uint8).copy()
    want_p, want_c, n_occ = oracle.analyse_reads(reads, 5, 1, 1)
    head, recs, cnts = _run_host(host_check, tmp_path, reads, 5, chunk=2048, n_slots=1024, min_count=1, min_ext=1)
    assert head[0] == n_occ and (recs == want_p).all() and (cnts == want_c).all()
    assert cnts[:, 0].max() == 255 and cnts[:, 1:].max() == 127


@pytest.mark.parametrize("k", [19, 51])
def test_kernel_functions_table_growth(host_check, tmp_path, k):
    """kc_move_slot: the table doubled and rehashed in the middle of the input (twice) gives the same k-mers and counters."""
    _, reads = _reads(k, 12000, 30, seed=40 + k, coverage=3, with_n=True)
    want_p, want_c, n_occ = oracle.analyse_reads(reads, k, 1, 1)
    n_slots = 1 << 15
    for grow_after in (0, 5):
        head, recs, cnts = _run_host(host_check, tmp_path, reads, k, chunk=4096, n_slots=n_slots, min_count=1, min_ext=1, grow_after=grow_after)
        assert head[0] == n_occ and head[3] == 0 and (recs == want_p).all() and (cnts == want_c).all()


def test_kernel_functions_report_a_full_table(host_check, tmp_path):
    _, reads = _reads(19, 5000, 10, seed=3)
    head, _, _ = _run_host(host_check, tmp_path, reads, 19, chunk=4096, n_slots=1024, min_count=1, min_ext=1)
    assert head[3] == 1 and head[1] == 1024


# --------------------------------------------------------------------------------------------------- GPU: the C ABI
def _counter(k, n_distinct, lf=0.5):
    import cs267_hw3_b200 as kh
    return kh.KmerCounter(k, n_distinct, lf, device=0)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [3, 16, 19, 31, 32, 51, 61])
def test_gpu_count_matches_the_oracle(k):
    if k < 8:
        rng = np.random.default_rng(k)
        reads = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 50000)].copy()
        reads[rng.choice(reads.size, 500, replace=False)] = 10
    else:
        _, reads = _reads(k, 200000, 300, seed=200 + k, coverage=4, error_rate=0.002, with_n=True)
    want_p, want_c, n_occ = oracle.analyse_reads(reads, k, 2, 2)
    with _counter(k, max(1024, len(reads))) as kc:
        kc.count_reads(reads)
        got = _sorted_rows(kc.extract(2, 2))
        assert got.shape == want_p.shape and (got == want_p).all()
        pl = (k + 3) // 4
        assert (kc.lookup(want_p[:, :pl]) == want_c).all()
        if k >= 16:                                           # k-mers that were never seen read as all zero
            absent = want_p[:50, :pl].copy()
            absent[:, 0] ^= 0x55                              # the first four bases changed (byte 0 holds no padding)
            seen = {bytes(r) for r in oracle.analyse_reads(reads, k, 1, 1)[0][:, :pl]}
            mask = np.array([bytes(r) not in seen for r in absent])
            assert mask.any() and (kc.lookup(absent)[mask] == 0).all()
        st = kc.stats()
        assert st["n_occurrences"] == n_occ and st["n_reported"] == len(want_p) and st["n_bytes"] == reads.size
        assert st["slot_bytes"] == (16 if k <= 31 else 32)
        # other thresholds on the same table
        for mc, me in ((1, 1), (3, 2), (2, 5)):
            w, _, _ = oracle.analyse_reads(reads, k, mc, me)
            assert (_sorted_rows(kc.extract(mc, me)) == w).all()
        # counting is additive over calls; clear starts over
        kc.clear()
        half = int(np.flatnonzero(reads[: reads.size // 2] == 10)[-1]) + 1       # a read boundary
        kc.count_reads(reads[:half])
        kc.count_reads(reads[half:])
        assert (_sorted_rows(kc.extract(2, 2)) == want_p).all()


@pytest.mark.gpu
def test_gpu_count_saturation_and_contention():
    """Thousands of threads hit the same few slots at once (homopolymers, one read repeated 20 000 times): the CAS loop on
    the counter word must saturate exactly like the serial oracle."""
    reads = np.frombuffer(b"A" * 5000 + b"\n" + b"CGTACGTTAGC\n" * 20000 + b"T" * 3000 + b"\n", dtype=np.uint8).copy()
    for k in (5, 33):
        want_p, want_c, n_occ = oracle.analyse_reads(reads, k, 1, 1)
        with _counter(k, 4096) as kc:
            kc.count_reads(reads)
            got = _sorted_rows(kc.extract(1, 1))
            assert (got == want_p).all()
            assert (kc.lookup(want_p[:, : (k + 3) // 4]) == want_c).all()
            assert kc.stats()["n_occurrences"] == n_occ


@pytest.mark.gpu
def test_gpu_count_device_buffer_any_alignment_and_long_input():
    """kh_count_reads_device on a pointer that is not 16-byte aligned, and a host buffer longer than one 64 MB chunk
    (k-mers straddling the chunk border are cut once)."""
    import ctypes as C

    import cs267_hw3_b200 as kh
    k = 31
    _, reads = _reads(k, 300000, 500, seed=77, coverage=3, with_n=True)
    want_p, _, n_occ = oracle.analyse_reads(reads, k, 2, 2)
    L = kh.lib()
    with _counter(k, len(reads)) as kc:
        p = C.c_void_p()
        assert L.kh_device_alloc(C.byref(p), reads.size + 64) == 0
        with kh.KmerHashTable(k, 1024) as tab:                                  # only for its copy helper
            tab._check(L.kh_copy_device(tab._h, p.value + 5, reads.ctypes.data, reads.size))
            tab.sync()
        kc.count_reads_device(p.value + 5, reads.size)
        assert (_sorted_rows(kc.extract(2, 2)) == want_p).all()
        assert kc.stats()["n_occurrences"] == n_occ
        L.kh_device_free(p)
    # > 64 MB: the reads repeated (counts scale, the k-mer set does not), border inside a read
    reps = (70 << 20) // reads.size + 1
    big = np.tile(reads, reps)
    want_big, want_c, n_occ_big = oracle.analyse_reads(reads, k, 1, 1)
    with _counter(k, len(reads)) as kc:
        kc.count_reads(big)
        st = kc.stats()
        assert st["n_occurrences"] == n_occ * reps
        got = _sorted_rows(kc.extract(1, 1))
        assert (got == want_big).all()
        assert (kc.lookup(want_big[:, : (k + 3) // 4])[:, 0] == np.minimum(want_c[:, 0] * reps, 255)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [19, 51])
def test_gpu_count_table_grows_instead_of_filling_up(k):
    """A counter created far too small (1024 slots) for 150 000 distinct k-mers: kh_count_reads hands the kernel only as many
    positions as the table has room for and doubles it when that gets small -- same records and counters as the oracle."""
    import ctypes as C

    import cs267_hw3_b200 as kh
    _, reads = _reads(k, 150000, 200, seed=600 + k, coverage=3, error_rate=0.002)
    want_p, want_c, n_occ = oracle.analyse_reads(reads, k, 2, 2)
    with _counter(k, 100) as kc:
        kc.count_reads(reads)
        st = kc.stats()
        assert st["n_grows"] >= 7 and st["n_occurrences"] == n_occ and st["n_distinct"] <= 0.9 * st["n_slots"]
        assert (_sorted_rows(kc.extract(2, 2)) == want_p).all()
        assert (kc.lookup(want_p[:, : (k + 3) // 4]) == want_c).all()
    # the device entry point only enqueues, so it cannot grow: the overflow is reported by the next synchronising call
    L = kh.lib()
    with _counter(k, 100) as kc:
        p = C.c_void_p()
        assert L.kh_device_alloc(C.byref(p), reads.size) == 0
        with kh.KmerHashTable(k, 1024) as tab:
            tab._check(L.kh_copy_device(tab._h, p, reads.ctypes.data, reads.size))
            tab.sync()
        kc.count_reads_device(p.value, reads.size)
        with pytest.raises(kh.KhError) as e:
            kc.stats()
        assert e.value.status == kh.KH_ERR_TABLE_FULL
        L.kh_device_free(p)


@pytest.mark.gpu
def test_gpu_count_errors():
    import cs267_hw3_b200 as kh
    _, reads = _reads(19, 50000, 50, seed=5)
    os.environ["KH_COUNT_GROW"] = "0"
    try:
        with _counter(19, 1000, 1.0) as kc:                                    # 1024 slots for 50 000 k-mers, not allowed to grow
            with pytest.raises(kh.KhError) as e:
                kc.count_reads(reads)
            assert e.value.status == kh.KH_ERR_TABLE_FULL
    finally:
        os.environ.pop("KH_COUNT_GROW")
    with _counter(19, 100000) as kc:
        kc.count_reads(reads)
        for bad in ((0, 1), (256, 1), (1, 0), (1, 128)):
            with pytest.raises(kh.KhError) as e:
                kc.extract(*bad)
            assert e.value.status == kh.KH_ERR_ARG
        import ctypes as C
        n = C.c_uint64()
        buf = np.empty((10, 7), dtype=np.uint8)
        rc = kh.lib().kh_count_extract(kc._h, 2, 2, buf.ctypes.data, 10, C.byref(n))
        assert rc == kh.KH_ERR_ARG and n.value == 50000                        # too small: nothing copied, size reported
    with pytest.raises(kh.KhError):
        kh.KmerCounter(62, 1000)


@pytest.mark.gpu
@pytest.mark.parametrize("k,n,c", [(19, 400000, 800), (51, 300000, 600)])
def test_gpu_reads_to_contigs_closed_loop(k, n, c):
    """reads -> kh_count (GPU) -> records stay in device memory -> kh_insert_pairs_device -> kh_assemble: the contigs are
    the generator's solution, the file scripts/check_it.sh:47-56 diffs against -- no text k-mer file in between."""
    import cs267_hw3_b200 as kh
    # 8 passes, 0.05 % substitutions, thresholds 3: a true k-mer keeps >= 3 clean sightings and no erroneous k-mer is
    # seen three times (both with probability ~1e-4 over the whole set at these sizes; the seeds are fixed)
    d, reads = _reads(k, n, c, seed=300 + k, coverage=8, error_rate=0.0005)
    with _counter(k, reads.size // 4) as kc:
        kc.count_reads(reads)
        ptr, n_rec = kc.extract_device(3, 3)
        assert n_rec == n
        with kh.KmerHashTable(k, n_rec, 0.5, device=0) as tab:
            tab.insert_pairs_device(ptr, n_rec)
            buf, offs, nodes = tab.assemble()
        assert nodes == n and len(offs) - 1 == c
        lines = sorted(buf.tobytes().split(b"\n")[:-1])
        assert b"\n".join(lines) + b"\n" == d.solution()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [19, 51])
def test_gpu_extract_lines_is_the_reference_kmer_file(k):
    """kh_count_extract_lines writes what read_kmers parses (read_kmers.hpp:64-76): parsed back by the oracle's
    restatement of that loop, the lines are the records -- and the generator's k-mer file up to line order."""
    d, reads = _reads(k, 100000, 150, seed=400 + k)
    with _counter(k, 200000) as kc:
        kc.count_reads(reads)
        lines = kc.extract_lines(2, 2)
        recs = _sorted_rows(kc.extract(2, 2))
    assert lines.size == 100000 * (k + 4)
    assert (_sorted_rows(oracle.parse_lines(lines, k)) == recs).all()
    assert sorted(lines.tobytes().split(b"\n")) == sorted(d.text().tobytes().split(b"\n"))


def _fastq(reads: np.ndarray) -> bytes:
    """The reads as FASTQ records whose header and quality lines are made of base letters (they must not be counted)."""
    out = []
    for i, r in enumerate(bytes(reads).split(b"\n")):
        if r:
            out.append(b"@ACGT" + str(i).encode() + b"\n" + r + b"\n+\n" + b"ACGT"[i % 4: i % 4 + 1] * len(r) + b"\n")
    return b"".join(out)


@pytest.mark.gpu
@pytest.mark.parametrize("k,fmt", [(19, "fastq"), (51, "plain"), (31, "fasta")])
def test_gpu_kmer_count_cli_to_kmer_file_and_to_contigs(tmp_path, k, fmt):
    """reads file -> kmer_count -> the reference's k-mer file -> kmer_hash_<K> test (the reference's own command line) ->
    contigs == the generator's solution (scripts/check_it.sh:47-56); and kmer_count --contigs without the file in between."""
    from cs267_hw3_b200 import build
    exe = build.build_count_cli()
    hash_exe = build.build_cli(ks=(k,))[0]
    d, reads = _reads(k, 60000, 100, seed=500 + k)
    if fmt == "fastq":
        blob = _fastq(reads)
    elif fmt == "fasta":
        blob = b"".join(b">ACGTACGT read " + str(i).encode() + b"\n" + r[:30] + b"\n" + r[30:] + b"\n"      # wrapped sequence lines
                        for i, r in enumerate(bytes(reads).split(b"\n")) if r)
    else:
        blob = bytes(reads)
    (tmp_path / "reads.txt").write_bytes(blob)
    r = subprocess.run([exe, str(k), "reads.txt", "kmers.txt", "2", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"Wrote {d.n} k-mers" in r.stdout
    got = (tmp_path / "kmers.txt").read_bytes()
    assert sorted(got.split(b"\n")) == sorted(d.text().tobytes().split(b"\n"))
    r = subprocess.run([hash_exe, "kmers.txt", "test", "viafile"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    sol = d.solution()
    assert b"".join(sorted((tmp_path / "viafile_0.dat").read_bytes().splitlines(keepends=True))) == sol
    r = subprocess.run([exe, str(k), "reads.txt", "--contigs", "direct"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert b"".join(sorted((tmp_path / "direct_0.dat").read_bytes().splitlines(keepends=True))) == sol


SEQ_LINES_CHECK = r"""
#include <cstdio>
#include <iostream>
#include <iterator>
#include <string>
#include "kh/kmer_counter.hpp"
int main(int argc, char** argv) {
    std::string all((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
    const size_t piece = (size_t)atoi(argv[1]);
    char format = 0; unsigned state = 0;
    size_t pos = 0;
    if (!all.empty()) format = all[0] == '>' ? 'a' : (all[0] == '@' ? 'q' : 'p');
    std::string out;
    while (pos < all.size()) {                       // pieces cut as src/kmer_count.cpp cuts them
        size_t have = std::min(all.size() - pos, piece), cut = 0;
        for (;;) {
            cut = have == all.size() - pos ? have : kh::sequence_cut(&all[pos], have, format);
            if (cut || have == all.size() - pos) break;
            have = std::min(all.size() - pos, have * 2);
        }
        if (cut == 0) cut = have;
        size_t len = cut;
        kh::sequence_lines(&all[pos], len, format, state);
        out.append(all, pos, len);
        pos += cut;
    }
    fwrite(out.data(), 1, out.size(), stdout);
    fprintf(stderr, "%c", format);
    return 0;
}
"""


def test_sequence_lines_blanks_headers_and_qualities(tmp_path):
    """include/kh/kmer_counter.hpp: FASTA / FASTQ text -> only the sequence lines keep their letters (CPU, no library call)."""
    src = tmp_path / "seq.cpp"
    src.write_text(SEQ_LINES_CHECK)
    exe = tmp_path / "seq"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    reads = [b"ACGTTGCA", b"GGGTTTAAACCC", b"TTGACA"]
    fq = b"".join(b"@ACGT" + str(i).encode() + b"\n" + r + b"\n+\n" + b"A" * len(r) + b"\n" for i, r in enumerate(reads))
    fa = b"".join(b">GATTACA " + str(i).encode() + b"\n" + r[:4] + b"\n" + r[4:] + b"\n" for i, r in enumerate(reads))
    # FASTA sequences wrapped over two lines come out joined (k-mers span the line break)
    for blob, fmt in ((fq, "q"), (fa, "a"), (b"\n".join(reads) + b"\n", "p")):
        for piece in (7, 30, 10000):
            p = subprocess.run([str(exe), str(piece)], input=blob, capture_output=True, check=True)
            assert p.stderr.decode() == fmt
            assert [x for x in p.stdout.split(b"\n") if x] == reads


# ------------------------------------------------------------------------------------- CPU: boundary behaviour
def test_counter_has_no_cpu_fallback():
    import cs267_hw3_b200 as kh
    if kh.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(kh.KhError) as e:
        kh.KmerCounter(19, 1000)
    assert e.value.status == kh.KH_ERR_CUDA


def test_kmer_count_cli_usage_and_errors(tmp_path):
    from cs267_hw3_b200 import build
    exe = build.build_count_cli()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Usage: kmer_count K reads_file")
    r = subprocess.run([exe, "19", "reads.txt", "--contigs"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([exe, "99", "reads.txt", "out.txt"], capture_output=True, text=True)
    assert r.returncode == 1 and "K must be 2..61" in r.stderr
    r = subprocess.run([exe, "19", "/nonexistent/reads.txt", "out.txt"], capture_output=True, text=True)
    assert r.returncode == -6 and "could not open /nonexistent/reads.txt" in r.stderr      # uncaught runtime_error, like the reference's CLI


# ------------------------------------------------------------- pinned to files the unmodified reference wrote
def _golden_reads(case, k):
    """Reads tiled over the contigs the UNMODIFIED reference wrote for a golden case (tests/golden/<case>.dat, made by
    tests/golden/make_golden.py from oracle/_ref) and the k-mer file it had been given (<case>.txt)."""
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, case + ".dat"), "rb") as f:
        contigs = f.read()
    with open(os.path.join(GOLDEN, case + ".txt"), "rb") as f:
        kmer_file = f.read()
    reads = readgen.tile_reads(contigs, k, read_len=k + (2 if k < 8 else 40), coverage=2, seed=k)
    return reads, sorted(kmer_file.split(b"\n")[:-1])


def _lines_of(pairs, k):
    pl = (k + 3) // 4
    return sorted((oracle.unpack_kmer(bytes(p[:pl]), k) + " " + chr(p[pl]) + chr(p[pl + 1])).encode() for p in pairs)


GOLDEN_CASES = [("readme_k3", 3), ("k19_a", 19), ("k19_singletons", 19), ("k19_long", 19), ("k31_a", 31), ("k51_a", 51)]


@pytest.mark.parametrize("case,k", GOLDEN_CASES)
def test_oracle_gives_back_the_references_own_input_file(case, k):
    """Closing the loop on reference-written artefacts: reads cut from the contigs the reference produced, analysed, are
    line for line the k-mer file the reference had consumed (its README.md:19-21 'first preprocessing stage' undone and
    redone)."""
    reads, want = _golden_reads(case, k)
    assert _lines_of(oracle.analyse_reads(reads, k, 2, 2)[0], k) == want


@pytest.mark.gpu
@pytest.mark.parametrize("case,k", GOLDEN_CASES)
def test_gpu_gives_back_the_references_own_input_file(case, k):
    reads, want = _golden_reads(case, k)
    with _counter(k, 20000) as kc:
        kc.count_reads(reads)
        assert sorted(kc.extract_lines(2, 2).tobytes().split(b"\n")[:-1]) == want
        assert _lines_of(kc.extract(2, 2), k) == want


def _python_count(reads: bytes, k: int, min_count: int, min_ext: int):
    """The definition in oracle/kmer_count_oracle.c's header once more, as plain Python over a dict (small cases only)."""
    acc = {}
    for p in range(len(reads) - k + 1):
        kmer = reads[p:p + k]
        if any(ch not in b"ACGT" for ch in kmer):
            continue
        e = acc.setdefault(kmer, [0, {}, {}])
        e[0] += 1
        if p > 0 and reads[p - 1:p] in (b"A", b"C", b"G", b"T"):
            e[1][reads[p - 1:p]] = e[1].get(reads[p - 1:p], 0) + 1
        if p + k < len(reads) and reads[p + k:p + k + 1] in (b"A", b"C", b"G", b"T"):
            e[2][reads[p + k:p + k + 1]] = e[2].get(reads[p + k:p + k + 1], 0) + 1
    out = []
    for kmer in sorted(acc):
        n, back, fwd = acc[kmer]
        if min(n, 255) < min_count:
            continue
        ext = b""
        for side in (back, fwd):
            q = [b for b in (b"A", b"C", b"G", b"T") if min(side.get(b, 0), 127) >= min_ext]
            ext += q[0] if len(q) == 1 else b"F"
        out.append(kmer + b" " + ext)
    return out


@pytest.mark.parametrize("k,min_count,min_ext,seed", [(2, 1, 1, 1), (3, 2, 2, 2), (5, 1, 2, 3), (9, 2, 1, 4), (17, 1, 1, 5), (33, 2, 2, 6), (61, 1, 1, 7)])
def test_oracle_against_an_independent_python_statement(k, min_count, min_ext, seed):
    """Two statements of the definition that share nothing (C: sort + run scan; Python: dict): random text over
    ACGT with separators, N and lower case, short alphabets so that k-mers repeat and fork, saturating repeats."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTN\na", dtype=np.uint8) if k < 12 else np.frombuffer(b"ACGT", dtype=np.uint8)
    text = alphabet[rng.integers(0, alphabet.size, 3000)]
    if k >= 12:                                        # long k: repeat a few reads so that counts exceed 1, with point changes
        base = bytes(text[:400])
        parts = []
        for _ in range(12):
            b = bytearray(base)
            for pos in rng.integers(0, len(b), 3):
                b[pos] = b"ACGT"[int(rng.integers(0, 4))]
            parts.append(bytes(b))
        text = np.frombuffer(b"\n".join(parts) + b"\n" + b"A" * 400 + b"\n", dtype=np.uint8)
    blob = bytes(text)
    pairs, _, _ = oracle.analyse_reads(blob, k, min_count, min_ext)
    assert _lines_of(pairs, k) == _python_count(blob, k, min_count, min_ext)


def test_kernel_functions_random_inputs(host_check, tmp_path):
    """Property run: random K (every slot / window regime), random text with separators and runs, random chunk size and
    table size (one piece or many, growth in the middle or not), aligned or not -- the kernels' functions always agree
    with the oracle on records and counters."""
    rng = np.random.default_rng(20261018)
    for trial in range(40):
        k = int(rng.choice([2, 3, 4, 7, 8, 15, 16, 17, 24, 31, 32, 33, 40, 48, 55, 61]))
        n = int(rng.integers(1, 9000))
        alpha = np.frombuffer(b"ACGT" * int(rng.integers(1, 12)) + b"N\n", dtype=np.uint8)
        text = alpha[rng.integers(0, alpha.size, n)].copy()
        if rng.random() < 0.5 and n > 200:                       # repeats: counters above 1, forks, saturation of small fields
            piece = text[: int(rng.integers(k + 2, 200))].copy()
            reps = int(rng.integers(2, 160))
            text = np.concatenate([text] + [np.concatenate((piece, [10]))] * reps).astype(np.uint8)
        mc, me = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        want_p, want_c, n_occ = oracle.analyse_reads(text, k, mc, me)
        misalign = int(rng.integers(0, 16)) if rng.random() < 0.3 else 0
        chunk = int(rng.choice([2048, 4096, 8192, 1 << 20]))
        grow = int(rng.integers(0, 3)) if (misalign == 0 and rng.random() < 0.4) else None
        n_slots = max(1024, 2 * (n_occ + 1))
        head, recs, cnts = _run_host(host_check, tmp_path, text, k, chunk=chunk, n_slots=n_slots, min_count=mc, min_ext=me,
                                     misalign=misalign, grow_after=grow)
        ctx = f"trial {trial}: k={k} n={text.size} chunk={chunk} misalign={misalign} grow={grow} mc={mc} me={me}"
        assert head[0] == n_occ and head[3] == 0, ctx
        assert recs.shape == want_p.shape and (recs == want_p).all(), ctx
        assert (cnts == want_c).all(), ctx
