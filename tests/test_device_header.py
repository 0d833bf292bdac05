"""GPU: include/kh/device_table.cuh -- kh::device_find / kh::device_insert called from a caller's own kernels
(README.md:95-99's per-k-mer HashMap::insert / find as __device__ functions) agree with the batch C ABI on the same
table.  The checker (tests/native/device_view_check.cu) is compiled here with nvcc against include/ and libkh_b200.so."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    import cs267_hw3_b200 as kh
    kh.lib()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path_factory.mktemp("devhdr") / "device_view_check"
    pkg = os.path.join(ROOT, "cs267_hw3_b200")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "native", "device_view_check.cu"), "-L" + pkg, "-lkh_b200",
                        "-Xlinker", "-rpath," + pkg, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return str(exe)


@pytest.mark.parametrize("k,n", [(19, 50000), (31, 20000), (51, 50000), (13, 3000)])
def test_device_find_and_insert_agree_with_the_batch_abi(checker, k, n):
    r = subprocess.run([checker, str(k), str(n)], capture_output=True, text=True, timeout=300, env=dict(os.environ, KH_CT="0"))
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_chunk_tables_refuse_a_device_view():
    import ctypes as C

    import cs267_hw3_b200 as kh
    L = kh.lib()
    os.environ["KH_CT"] = "2"
    try:
        with kh.KmerHashTable(51, 1000) as tab:
            view = (C.c_uint8 * 64)()
            L.kh_get_device_view.argtypes = [C.c_void_p, C.c_void_p]
            assert L.kh_get_device_view(tab._h, view) == kh.KH_ERR_ARG
    finally:
        del os.environ["KH_CT"]
