"""ctypes view of oracle/libkmer_oracle.so and oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs, and from nowhere else: the product path
(``cs267_hw3_b200``) must never import this package (tests/test_boundary.py greps for it).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkmer_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

ERRORS = {1: "bad argument", 2: "k-mer not found", 3: "allocation", 4: "cycle", 5: "bad base", 6: "capacity"}


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, f) for f in ("kmer_oracle.c", "kmer_count_oracle.c")]
        if not os.path.exists(LIB_PATH) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in srcs):
            build()
        L = C.CDLL(LIB_PATH)
        u8p, u64, vp = C.POINTER(C.c_uint8), C.c_uint64, C.c_void_p
        L.ko_pair_bytes.argtypes = [C.c_int]
        L.ko_pack_kmer.argtypes = [C.c_char_p, C.c_int, u8p]
        L.ko_unpack_kmer.argtypes = [u8p, C.c_int, C.c_char_p]
        L.ko_unpack_kmer.restype = None
        L.ko_parse_lines.argtypes = [vp, u64, C.c_int, vp]
        L.ko_next_kmer.argtypes = [u8p, C.c_int, u8p]
        L.ko_table_create.argtypes = [C.c_int, u64, C.POINTER(vp)]
        L.ko_table_destroy.argtypes = [vp]
        L.ko_table_destroy.restype = None
        L.ko_table_count.argtypes = [vp]
        L.ko_table_count.restype = u64
        L.ko_insert_pairs.argtypes = [vp, vp, u64]
        L.ko_find.argtypes = [vp, u8p, u8p]
        L.ko_assemble.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.kco_count_occurrences.argtypes = [vp, u64, C.c_int]
        L.kco_count_occurrences.restype = u64
        L.kco_analyse.argtypes = [vp, u64, C.c_int, C.c_uint32, C.c_uint32, vp, u64, C.POINTER(u64), vp]
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"oracle {what}: {ERRORS.get(rc, rc)}")


def pair_bytes(k: int) -> int:
    return (k + 3) // 4 + 2


def pack_kmer(kmer: str) -> bytes:
    k = len(kmer)
    out = (C.c_uint8 * ((k + 3) // 4))()
    _check(lib().ko_pack_kmer(kmer.encode(), k, out), "pack_kmer")
    return bytes(out)


def unpack_kmer(packed: bytes, k: int) -> str:
    buf = C.create_string_buffer(k)
    arr = (C.c_uint8 * len(packed)).from_buffer_copy(packed)
    lib().ko_unpack_kmer(arr, k, buf)
    return buf.raw[:k].decode()


def next_kmer(pair: bytes, k: int) -> bytes:
    arr = (C.c_uint8 * len(pair)).from_buffer_copy(pair)
    out = (C.c_uint8 * ((k + 3) // 4))()
    _check(lib().ko_next_kmer(arr, k, out), "next_kmer")
    return bytes(out)


def parse_lines(text: np.ndarray | bytes, k: int) -> np.ndarray:
    """read_kmers parse loop: text in the reference's format -> (n, pair_bytes) uint8 records."""
    t = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    assert t.size % (k + 4) == 0, "text is not a whole number of (K+4)-byte lines"
    n = t.size // (k + 4)
    out = np.empty((n, pair_bytes(k)), dtype=np.uint8)
    t = np.ascontiguousarray(t)
    _check(lib().ko_parse_lines(t.ctypes.data, n, k, out.ctypes.data), "parse_lines")
    return out


class Table:
    """DistributedHashMap semantics at one rank (hash_map.hpp:50-107)."""

    def __init__(self, k: int, n_expected: int):
        self.k = k
        self._h = C.c_void_p()
        _check(lib().ko_table_create(k, n_expected, C.byref(self._h)), "table_create")

    def close(self) -> None:
        if self._h:
            lib().ko_table_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(lib().ko_table_count(self._h))

    def insert_pairs(self, pairs: np.ndarray) -> None:
        p = np.ascontiguousarray(pairs, dtype=np.uint8)
        n = p.size // pair_bytes(self.k)
        _check(lib().ko_insert_pairs(self._h, p.ctypes.data, n), "insert_pairs")

    def find(self, packed: bytes):
        arr = (C.c_uint8 * len(packed)).from_buffer_copy(packed)
        out = (C.c_uint8 * pair_bytes(self.k))()
        return bytes(out) if lib().ko_find(self._h, arr, out) else None

    def assemble(self, pairs: np.ndarray):
        """Contigs started by the records in `pairs` (this rank's block), in order.

        Returns (buffer of '\\n'-terminated contigs as bytes, n_contigs, n_nodes)."""
        p = np.ascontiguousarray(pairs, dtype=np.uint8)
        n = p.size // pair_bytes(self.k)
        ln, nc, nn = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().ko_assemble(self._h, p.ctypes.data, n, None, 0, C.byref(ln), C.byref(nc), C.byref(nn)),
               "assemble(size)")
        buf = np.empty(ln.value, dtype=np.uint8)
        _check(lib().ko_assemble(self._h, p.ctypes.data, n, buf.ctypes.data, buf.size,
                                 C.byref(ln), C.byref(nc), C.byref(nn)), "assemble")
        return buf.tobytes(), nc.value, nn.value


def assemble_text(text: bytes | np.ndarray, k: int, nranks: int = 1):
    """Whole reference flow on a k-mer file image: per-rank contig buffers (list of bytes)."""
    pairs = parse_lines(text, k)
    n = pairs.shape[0]
    tab = Table(k, n)
    tab.insert_pairs(pairs)
    split = (n + nranks - 1) // nranks           # read_kmers.hpp:56
    outs = []
    for r in range(nranks):
        lo = min(n, split * r)
        hi = min(n, lo + split)
        outs.append(tab.assemble(pairs[lo:hi])[0])
    tab.close()
    return outs


# ---- the k-mer analysis stage that precedes the reference's path (oracle/kmer_count_oracle.c) ----------

def analyse_reads(reads: bytes | np.ndarray, k: int, min_count: int = 2, min_ext: int = 2):
    """reads (any byte outside ACGT separates them) -> (records sorted by k-mer, shape (n, pair_bytes);
    counters uint32 (n, 9): occurrences, backward A C G T, forward A C G T; number of occurrences in the reads)."""
    r = np.frombuffer(reads, dtype=np.uint8) if not isinstance(reads, np.ndarray) else np.ascontiguousarray(reads)
    n_occ = int(lib().kco_count_occurrences(r.ctypes.data, r.size, k))
    n = C.c_uint64()
    pairs = np.empty((max(n_occ, 1), pair_bytes(k)), dtype=np.uint8)        # distinct k-mers <= occurrences: one pass
    counts = np.empty((max(n_occ, 1), 9), dtype=np.uint32)
    _check(lib().kco_analyse(r.ctypes.data, r.size, k, min_count, min_ext, pairs.ctypes.data, max(n_occ, 1), C.byref(n),
                             counts.ctypes.data), "analyse")
    return pairs[: n.value].copy(), counts[: n.value].copy(), n_occ


# ---- the unmodified reference, compiled (oracle/_ref) -------------------------------------

def ref_binary(k: int) -> str | None:
    p = os.path.join(REF_DIR, f"kmer_hash_ref_{k}")
    return p if os.path.exists(p) else None


def run_reference(k: int, kmer_file: str, workdir: str, prefix: str = "ref", test: bool = True):
    """Run oracle/_ref/kmer_hash_ref_<k> on a file.

    test=True : returns the bytes of <prefix>_0.dat (contigs in start order).
    test=False: returns (insert_seconds, total_seconds) parsed from its own timer lines
                (kmer_hash.cpp:144-145) -- the reference's timed region."""
    exe = ref_binary(k)
    if exe is None:
        raise FileNotFoundError(f"oracle/_ref/kmer_hash_ref_{k} not built (run make -C oracle)")
    if test:
        r = subprocess.run([exe, kmer_file, "test", prefix], cwd=workdir, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference failed ({r.returncode}): {r.stderr}")
        with open(os.path.join(workdir, f"{prefix}_0.dat"), "rb") as f:
            return f.read()
    r = subprocess.run([exe, kmer_file], cwd=workdir, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed ({r.returncode}): {r.stderr}")
    ins = re.search(r"Finished inserting in ([0-9.]+) sec", r.stdout)
    tot = re.search(r"Assembled in ([0-9.]+) total", r.stdout)
    return float(ins.group(1)), float(tot.group(1))
