"""ctypes view of tools/libkmergen.so (the synthetic dataset generator, tools/kmer_gen.cpp).

A tool for tests and bench.py; not part of the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkmergen.so")
SRC = os.path.join(HERE, "kmer_gen.cpp")


def build() -> None:
    cmds = [
        ["g++", "-O2", "-std=c++17", "-pthread", "-fPIC", "-shared", SRC, "-o", LIB_PATH],
        ["g++", "-O2", "-std=c++17", "-pthread", "-DKG_MAIN", SRC, "-o", os.path.join(HERE, "gen_kmers")],
    ]
    for c in cmds:
        r = subprocess.run(c, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("generator build failed:\n" + r.stderr)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC):
            build()
        L = C.CDLL(LIB_PATH)
        u64, vp = C.c_uint64, C.c_void_p
        L.kg_create.argtypes = [C.c_int, u64, u64, u64, u64, C.c_int, C.POINTER(vp)]
        L.kg_destroy.argtypes = [vp]
        L.kg_destroy.restype = None
        for f in ("kg_n", "kg_c", "kg_text_bytes", "kg_pair_bytes", "kg_max_contig_nodes"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u64
        L.kg_write_text.argtypes = [vp, u64, u64, vp]
        L.kg_write_text.restype = None
        L.kg_write_pairs.argtypes = [vp, u64, u64, vp]
        L.kg_write_pairs.restype = None
        L.kg_write_expected.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(u64)]
        L.kg_write_expected.restype = u64
        L.kg_write_solution.argtypes = [vp, vp]
        L.kg_write_solution.restype = u64
        L.kg_digest_lines.argtypes = [vp, u64, C.c_int, C.POINTER(u64), C.POINTER(u64)]
        L.kg_digest_lines.restype = None
        L.kg_digest_expected.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.kg_digest_expected.restype = None
        _lib = L
    return _lib


class Dataset:
    """A synthetic unique-k-mer set: `n` k-mers in `c` contigs (optionally one of `long_nodes`)."""

    def __init__(self, k: int, n: int, c: int, seed: int = 267, long_nodes: int = 0, threads: int = 0):
        self.k, self.n, self.c, self.seed = k, n, c, seed
        self._h = C.c_void_p()
        rc = lib().kg_create(k, n, c, long_nodes, seed, threads, C.byref(self._h))
        if rc != 0:
            raise ValueError(f"kg_create(k={k}, n={n}, c={c}) failed with code {rc}")

    def close(self) -> None:
        if self._h:
            lib().kg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def line_bytes(self) -> int:
        return self.k + 4

    @property
    def pair_bytes(self) -> int:
        return (self.k + 3) // 4 + 2

    @property
    def max_contig_nodes(self) -> int:
        return int(lib().kg_max_contig_nodes(self._h))

    def text(self, line0: int = 0, count: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        count = self.n - line0 if count is None else count
        if out is None:
            out = np.empty(count * self.line_bytes, dtype=np.uint8)
        assert out.nbytes >= count * self.line_bytes
        lib().kg_write_text(self._h, line0, count, out.ctypes.data)
        return out

    def pairs(self, line0: int = 0, count: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """Records in the reference's kmer_pair byte layout, shape (count, pair_bytes)."""
        count = self.n - line0 if count is None else count
        if out is None:
            out = np.empty((count, self.pair_bytes), dtype=np.uint8)
        assert out.nbytes >= count * self.pair_bytes
        lib().kg_write_pairs(self._h, line0, count, out.ctypes.data)
        return out

    def pairs_into(self, ptr: int, line0: int, count: int) -> None:
        lib().kg_write_pairs(self._h, line0, count, ptr)

    def expected(self, nranks: int = 1, rank: int = 0) -> tuple[bytes, int]:
        """What `<prefix>_<rank>.dat` must contain: (bytes, n_contigs)."""
        nc = C.c_uint64()
        size = lib().kg_write_expected(self._h, nranks, rank, None, C.byref(nc))
        buf = np.empty(size, dtype=np.uint8)
        lib().kg_write_expected(self._h, nranks, rank, buf.ctypes.data, C.byref(nc))
        return buf.tobytes(), int(nc.value)

    def expected_array(self, nranks: int = 1, rank: int = 0) -> tuple[np.ndarray, int]:
        nc = C.c_uint64()
        size = lib().kg_write_expected(self._h, nranks, rank, None, C.byref(nc))
        buf = np.empty(size, dtype=np.uint8)
        lib().kg_write_expected(self._h, nranks, rank, buf.ctypes.data, C.byref(nc))
        return buf, int(nc.value)

    def solution(self) -> bytes:
        size = lib().kg_write_solution(self._h, None)
        buf = np.empty(size, dtype=np.uint8)
        lib().kg_write_solution(self._h, buf.ctypes.data)
        return buf.tobytes()

    def digest(self) -> tuple[int, int]:
        s, n = C.c_uint64(), C.c_uint64()
        lib().kg_digest_expected(self._h, C.byref(s), C.byref(n))
        return int(s.value), int(n.value)


def digest_lines(buf: np.ndarray | bytes) -> tuple[int, int]:
    """Order-independent (sum of per-line hashes, line count) of a '\\n'-terminated buffer."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf)
    s, n = C.c_uint64(), C.c_uint64()
    lib().kg_digest_lines(a.ctypes.data, a.size, 0, C.byref(s), C.byref(n))
    return int(s.value), int(n.value)
