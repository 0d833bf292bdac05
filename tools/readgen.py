"""Synthetic reads for the k-mer analysis stage (kh_count_*): tilings of a set of contigs.

A tool for tests and bench.py; not part of the product path.  Every pass lays reads of `read_len` bases over every
contig so that each window of K + 2 consecutive bases lies inside at least one read of the pass -- each k-mer of a
contig is then seen, with both neighbours, at least `coverage` times, and the first / last k-mer of a contig is never
seen with a backward / forward neighbour (its extension comes out 'F', README.md:37).  Output: one read per line.
"""
from __future__ import annotations

import numpy as np

_NL = np.uint8(10)


def contigs_of(solution: bytes | np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """('\\n'-terminated contig buffer) -> (buffer as uint8, start of every contig, length of every contig)."""
    g = np.frombuffer(solution, dtype=np.uint8) if not isinstance(solution, np.ndarray) else solution
    ends = np.flatnonzero(g == _NL)
    starts = np.concatenate(([0], ends[:-1] + 1)) if ends.size else np.zeros(0, dtype=np.int64)
    return g, starts.astype(np.int64), (ends - starts).astype(np.int64)


def tile_reads(solution: bytes | np.ndarray, k: int, read_len: int = 120, coverage: int = 3, seed: int = 1,
               error_rate: float = 0.0, rows_per_block: int = 1 << 16) -> np.ndarray:
    """Reads (uint8 buffer, one per line) covering every contig `coverage` times; see the module docstring."""
    if read_len < k + 2:
        raise ValueError("read_len must be at least K + 2")
    g, cstart, clen = contigs_of(solution)
    rng = np.random.default_rng(seed)
    step = read_len - (k + 1)
    out = []
    for p in range(coverage):
        phase = 0 if p == 0 else int(rng.integers(1, step + 1))
        # reads of contig c start at 0, phase, phase + step, ... (< len - L) and at max(0, len - L)
        room = np.maximum(clen - read_len, 0)
        n_mid = np.where(room > phase, (room - phase + step - 1) // step, 0)
        per = 2 + n_mid                                    # first, the strided ones, last
        total = int(per.sum())
        owner = np.repeat(np.arange(clen.size), per)
        first_of = np.cumsum(per) - per
        j = np.arange(total) - first_of[owner]             # index of the read inside its contig
        rel = np.where(j == 0, 0, np.where(j == per[owner] - 1, room[owner], phase + (j - 1) * step))
        starts = cstart[owner] + rel
        lens = np.minimum(read_len, clen[owner])
        for lo in range(0, total, rows_per_block):
            hi = min(total, lo + rows_per_block)
            idx = starts[lo:hi, None] + np.arange(read_len)[None, :]
            rows = g[np.minimum(idx, g.size - 1)]
            rows[np.arange(read_len)[None, :] >= lens[lo:hi, None]] = _NL
            out.append(np.concatenate((rows, np.full((hi - lo, 1), _NL, dtype=np.uint8)), axis=1).reshape(-1))
    buf = np.concatenate(out) if out else np.zeros(0, dtype=np.uint8)
    if error_rate > 0:
        pos = np.flatnonzero(buf != _NL)
        hit = pos[rng.random(pos.size) < error_rate]
        code = np.zeros(256, dtype=np.uint8)
        code[[65, 67, 71, 84]] = [0, 1, 2, 3]
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        buf[hit] = letters[(code[buf[hit]] + rng.integers(1, 4, hit.size)) % 4]
    return buf
