"""CPU: include/kh/stream_reader.hpp (streamed file ingestion) with a fake sink.  The reader must
hand over exactly the lines of the requested block, whole lines only, in file order, for any chunk size / ring size,
cut rank blocks like read_kmers.hpp:55-58, and fail cleanly (no leaked buffers, no hung thread) on a short file or a
throwing sink.  The GPU sink is one call (kh_insert_lines, see DistributedHashMap::insert_file)."""
import os
import subprocess

import pytest

from tools import kmergen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, N = 19, 10_000
LINE = K + 4


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    d = tmp_path_factory.mktemp("stream")
    exe = str(d / "stream_reader_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "stream_reader_check.cpp"), "-o", exe], check=True, capture_output=True)
    path = str(d / "kmers.txt")
    data = kmergen.Dataset(K, N, 50, seed=5).text().tobytes()
    assert len(data) == N * LINE
    open(path, "wb").write(data)
    return exe, path, data


def _run(exe, path, first, n, chunk, buffers, mode="copy"):
    return subprocess.run([exe, path, str(K), str(first), str(n), str(chunk), str(buffers), mode], capture_output=True, timeout=60)


@pytest.mark.parametrize("chunk,buffers", [(1, 2), (7, 2), (999, 3), (1000, 2), (N, 2), (10 * N, 4)])
def test_chunks_cover_the_block_in_order(env, chunk, buffers):
    exe, path, data = env
    r = _run(exe, path, 0, N, chunk, buffers)
    assert r.returncode == 0, r.stderr
    assert r.stdout == data
    lines = r.stderr.decode().splitlines()
    chunks = [tuple(map(int, ln.split(":"))) for ln in lines[:-1]]
    assert chunks[0][0] == 0 and sum(c[1] for c in chunks) == N
    assert all(a[0] + a[1] == b[0] for a, b in zip(chunks, chunks[1:]))          # consecutive, no gaps
    assert all(c[1] == min(chunk, N) for c in chunks[:-1])                      # whole chunks except the last
    assert lines[-1] == f"delivered {N} buffers {max(2, buffers)} leaked 0"


def test_sub_block_and_rank_blocks(env):
    exe, path, data = env
    r = _run(exe, path, 1234, 4321, 500, 3)
    assert r.returncode == 0 and r.stdout == data[1234 * LINE:(1234 + 4321) * LINE]
    world = 7                                                                    # read_kmers.hpp:55-58: ceil(N/P) per rank
    per = (N + world - 1) // world
    got = b""
    for rank in range(world):
        r = _run(exe, path, 0, N, 300, 2, f"rank{world}:{rank}")
        assert r.returncode == 0
        lo = min(N, per * rank)
        assert r.stdout == data[lo * LINE:min(N, lo + per) * LINE]
        got += r.stdout
    assert got == data
    r = _run(exe, path, 0, 10, 300, 2, "rank16:15")                              # a rank past the end: empty block
    assert r.returncode == 0 and r.stdout == b""


def test_short_file_and_missing_file_are_errors(env, tmp_path):
    exe, path, data = env
    cut = str(tmp_path / "cut.txt")
    open(cut, "wb").write(data[:5000 * LINE + 11])                               # ends inside line 5000
    r = _run(exe, cut, 0, N, 512, 3)
    assert r.returncode == 1
    err = r.stderr.decode()
    assert "ends inside line 5000" in err and err.rstrip().endswith("leaked 0")
    assert r.stdout == data[:4608 * LINE]                                        # the whole chunks before it were delivered
    r = _run(exe, str(tmp_path / "nope.txt"), 0, N, 512, 3)
    assert r.returncode == 1 and "could not open" in r.stderr.decode()


def test_sink_exception_stops_the_reader(env):
    exe, path, data = env
    r = _run(exe, path, 0, N, 100, 3, "throw5")
    assert r.returncode == 1
    assert "sink refused chunk 5" in r.stderr.decode() and r.stderr.decode().rstrip().endswith("leaked 0")
    assert r.stdout == data[:500 * LINE]
