"""CPU: the drop-in boundary -- the C ABI library loads and exports what include/kh_capi.h declares,
the host-side C++ headers mirror the reference's types byte for byte, the CLI keeps the reference's
argv/exit behaviour, and the product path never touches the oracle."""
import os
import re
import subprocess
import tempfile

import pytest

from conftest import ROOT

INCLUDE = os.path.join(ROOT, "include")


def _declared():
    with open(os.path.join(INCLUDE, "kh_capi.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(kh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import cs267_hw3_b200 as kh
    L = kh.lib()
    declared = _declared()
    assert declared == sorted(kh.ABI_SYMBOLS)
    for sym in declared:
        getattr(L, sym)                      # raises AttributeError if the .so does not export it
    nm = subprocess.run(["nm", "-D", "--defined-only", kh.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (kh_[a-z0-9_]+)", nm))
    assert exported == set(declared)          # nothing undeclared leaks out either


def test_library_is_sm100a_and_uses_wide_sector_ops():
    import cs267_hw3_b200 as kh
    kh.lib()
    r = subprocess.run(["cuobjdump", "-sass", kh.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    assert "ATOMG.E.CAS.128" in r.stdout      # 128-bit slots are claimed with one CAS
    assert "ATOMG.E.CAS.64" in r.stdout
    assert ".256" in r.stdout                 # one 32-byte bucket per load
    assert "ATOMS.CAS.128" in r.stdout        # chunk table: 128-bit slots claimed in shared memory
    assert "UBLKCP" in r.stdout               # chunk table: built chunks leave through bulk shared->global copies (cp.async.bulk)


def test_pure_helpers_need_no_gpu():
    import cs267_hw3_b200 as kh
    L = kh.lib()
    assert L.kh_abi_version() == 2
    assert [L.kh_pair_bytes(k) for k in (19, 31, 51)] == [7, 10, 15]
    assert [L.kh_packed_bytes(k) for k in (19, 31, 51)] == [5, 8, 13]
    assert b"k-mer not found in Distributed HashMap" in L.kh_status_string(kh.KH_ERR_NOT_FOUND)


def test_no_cpu_fallback_without_a_device():
    import cs267_hw3_b200 as kh
    if kh.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(kh.KhError) as e:
        kh.KmerHashTable(19, 1000)
    assert e.value.status == kh.KH_ERR_CUDA


def test_product_path_never_touches_the_oracle():
    bad = []
    for base in ("cs267_hw3_b200", "include", "src"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    with open(os.path.join(d, f)) as fh:
                        txt = fh.read()
                    if re.search(r"\bimport oracle\b|from oracle\b|oracle/|kmer_oracle|libkmer_oracle", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, f"product files reference the oracle: {bad}"


HOST_PROBE = r"""
#include <cstdio>
#include <iostream>
#include <list>
#include "kmer_t.hpp"
#include "read_kmers.hpp"
static void hex(const unsigned char* p, size_t n) { for (size_t i = 0; i < n; ++i) printf("%02x", p[i]); }
int main() {
    std::string line;
    std::list<kmer_pair> chain;
    while (std::getline(std::cin, line)) {
        kmer_pair kp(line.substr(0, KMER_LEN), line.substr(KMER_LEN + 1, 2));
        hex(reinterpret_cast<const unsigned char*>(&kp), sizeof(kp));
        printf(" ");
        if (kp.forwardExt() != 'F') { pkmer_t nx = kp.next_kmer(); hex(nx.data, sizeof(nx.data)); } else printf("-");
        printf(" %s %llu\n", kp.kmer_str().c_str(), (unsigned long long)kp.hash());
        chain.push_back(kp);
    }
    printf("%s\n", extract_contig(chain).c_str());
    return 0;
}
"""


@pytest.mark.parametrize("case,k", [("k19_a", 19), ("k31_a", 31), ("k51_a", 51), ("readme_k3", 3)])
def test_host_headers_match_reference_bytes(case, k):
    """include/kmer_t.hpp & co. build the same kmer_pair / next_kmer bytes as the reference's headers
    (golden .probe files were written by oracle/ref_probe.cpp including the reference headers)."""
    gold = os.path.join(ROOT, "tests", "golden")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "probe.cpp")
        with open(src, "w") as f:
            f.write(HOST_PROBE)
        exe = os.path.join(tmp, "probe")
        subprocess.run(["g++", "-std=c++17", "-O1", f"-DKMER_LEN={k}", "-I" + INCLUDE, src, "-o", exe], check=True)
        with open(os.path.join(gold, case + ".txt"), "rb") as f:
            head = b"".join(f.readlines()[:200])
        out = subprocess.run([exe], input=head, capture_output=True, check=True).stdout.decode().split("\n")
    with open(os.path.join(gold, case + ".probe")) as f:
        want = f.read().split("\n")[:-1]
    lines = head.decode().split("\n")[:-1]
    for got, w, line in zip(out, want, lines):
        g = got.split(" ")
        assert g[0] == w.split(" ")[0] and g[1] == w.split(" ")[1]
        assert g[2] == line[:k]                                    # get()/kmer_str() round trip
        h = 5381
        for b in bytes.fromhex(g[0])[: (k + 3) // 4]:
            h = (b + (h << 5) + h) & 0xFFFFFFFFFFFFFFFF            # pkmer_t.hpp:31-37
        assert int(g[3]) == h
    # extract_contig (read_kmers.hpp:81-92): first k-mer + every forward ext != 'F'
    assert out[len(lines)] == lines[0][:k] + "".join(l[k + 2] for l in lines if l[k + 2] != "F")


def test_cli_usage_and_k_mismatch():
    from cs267_hw3_b200 import build
    exe19 = build.build_cli(ks=(19,))[0]
    r = subprocess.run([exe19], capture_output=True, text=True)
    assert r.returncode == 1                                       # kmer_hash.cpp:87-91
    assert r.stdout == "Usage: srun -N nodes -n ranks ./kmer_hash kmer_file [verbose|test [prefix]]\n"
    p = os.path.join(ROOT, "tests", "golden", "k31_a.txt")
    r = subprocess.run([exe19, p], capture_output=True, text=True)
    assert r.returncode == -6                                      # uncaught runtime_error -> abort, as the reference
    assert (f"Error: {p} contains 31-mers, while this binary is compiled for 19-mers. "
            "Modify packing.hpp and recompile.") in r.stderr       # kmer_hash.cpp:102-104
    r = subprocess.run([exe19, "/nonexistent/file.txt"], capture_output=True, text=True)
    assert r.returncode == -6 and "kmer_size: could not open /nonexistent/file.txt" in r.stderr


def test_oracle_header_says_test_infrastructure():
    with open(os.path.join(ROOT, "oracle", "kmer_oracle.c")) as f:
        head = f.read(1500)
    assert "TEST INFRASTRUCTURE ONLY" in head and "PINNED" in head
