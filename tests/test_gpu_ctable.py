"""GPU parity of the CHUNK TABLE (csrc/ctable.cuh) on small inputs.

Large tables (>= 2^20 k-mers) and every sharded handle are chunk tables; the small cases of test_gpu_parity.py
would take the plain table.  Here KH_CT=2 forces the chunk table for every supported K (17..54), and the same
tests run again: golden vectors written by the unmodified reference, the oracle, the generator's solution,
load factors, incremental inserts, find, duplicates, error behaviour.  Bit-exact, no tolerance.
"""
import numpy as np
import pytest

import oracle
import test_gpu_parity as P
from tools import kmergen

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_chunk_table(monkeypatch):
    monkeypatch.setenv("KH_CT", "2")


kh = P.kh

test_assemble_golden = P.test_assemble_golden
test_generated_vs_oracle = P.test_generated_vs_oracle
test_load_factors = P.test_load_factors
test_insert_in_chunks_keeps_start_order = P.test_insert_in_chunks_keeps_start_order
test_insert_lines_path = P.test_insert_lines_path
test_per_rank_blocks = P.test_per_rank_blocks
test_empty_and_tiny = P.test_empty_and_tiny
test_clear_gives_a_fresh_table = P.test_clear_gives_a_fresh_table
test_find_parity = P.test_find_parity
test_duplicate_keys_are_counted = P.test_duplicate_keys_are_counted
test_missing_successor = P.test_missing_successor
test_bad_extension_in_records = P.test_bad_extension_in_records
test_medium_shapes_exact = P.test_medium_shapes_exact


def test_really_a_chunk_table(kh):
    """K=19 keeps 64-bit slots, K=29 moves to 128-bit slots to make room for the segment index bits."""
    with kh.KmerHashTable(19, 1000) as a, kh.KmerHashTable(29, 1000) as b, kh.KmerHashTable(61, 1000) as c:
        assert a.stats()["slot_bits"] == 64 and b.stats()["slot_bits"] == 128 and c.stats()["slot_bits"] == 128


@pytest.mark.parametrize("k", [17, 22, 23, 30, 54])
def test_edges_of_the_k_range(kh, k):
    d = kmergen.Dataset(k, 60000, 300, seed=k)
    out, offs, nodes, st = P._assemble_text(kh, d.text(), k)
    assert out == d.expected()[0] and nodes == d.n and st["n_inserted"] == d.n


@pytest.mark.parametrize("k", [19, 51])
def test_cycle_is_reported_not_spun_on(kh, k):
    text = P._cycle_text(k, 40, 300, seed=k)
    with pytest.raises(RuntimeError, match="cycle"):
        oracle.assemble_text(text, k)
    with kh.KmerHashTable(k, 400) as tab:
        tab.insert_lines(text)
        with pytest.raises(kh.KhError) as e:
            tab.assemble()
        assert e.value.status in (kh.KH_ERR_CYCLE, kh.KH_ERR_CONVERGE)      # the chain re-enters itself: refused either way


test_orphan_chains_and_cycles_are_ignored = P.test_orphan_chains_and_cycles_are_ignored


def test_converging_chains(kh):
    """Two start nodes whose chains share a suffix (outside the input contract, README.md:33-35): the reference
    emits the shared suffix twice; this library does the same or refuses with KH_ERR_CONVERGE, never anything else."""
    k = 19
    rng = np.random.default_rng(11)
    a = "".join("ACGT"[i] for i in rng.integers(0, 4, 60))
    b = "T" + a[6:6 + k - 1] if a[5] != "T" else "G" + a[6:6 + k - 1]
    lines = [f"{a[i:i + k]} {'F' if i == 0 else a[i - 1]}{'F' if i + k == len(a) else a[i + k]}\n" for i in range(len(a) - k + 1)]
    lines.append(f"{b} F{a[6 + k - 1]}\n")
    text = "".join(lines).encode()
    want = oracle.assemble_text(text, k)[0]
    with kh.KmerHashTable(k, 64) as tab:
        tab.insert_lines(text)
        try:
            buf, _, _ = tab.assemble()
            assert buf.tobytes() == want
        except kh.KhError as e:
            assert e.status == kh.KH_ERR_CONVERGE


def test_table_full_is_reported_at_the_seal(kh):
    d = kmergen.Dataset(19, 50000, 10, seed=2)
    with kh.KmerHashTable(19, 100, 1.0) as tab:
        with pytest.raises(kh.KhError) as e:
            tab.insert_pairs(d.pairs())
            tab.assemble()
        assert e.value.status == kh.KH_ERR_TABLE_FULL


def test_assemble_find_insert_interleaved(kh):
    """A sealed table accepts more records (it re-seals from the staged records) and can be traversed repeatedly."""
    k = 31
    d = kmergen.Dataset(k, 40000, 200, seed=6)
    pairs = d.pairs()
    pl = (k + 3) // 4
    with kh.KmerHashTable(k, 40000) as tab:
        tab.insert_pairs(pairs[:25000])
        got, found = tab.find(pairs[:30000, :pl])
        assert found[:25000].all() and not found[25000:].any() and (got[:25000] == pairs[:25000]).all()
        tab.insert_pairs(pairs[25000:])
        got, found = tab.find(pairs[:, :pl])
        assert found.all() and (got == pairs).all()
        for _ in range(2):
            buf, _, nodes = tab.assemble()
            assert buf.tobytes() == d.expected()[0] and nodes == d.n
        got, found = tab.find(pairs[::7, :pl])
        assert found.all() and (got == pairs[::7]).all()
