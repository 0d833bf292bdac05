# A/B of KH_CT_DENSE_RETRY (ct_build_kernel phase 2: work list instead of per-lane retry loops): the K=51 bench line
# with both builds, then the chunk-table parity tests against the dense build.
mkdir -p gpurun_out
for d in 0 1; do
  KH_LIB_PATH=$PWD/tools/probes/variants/libkh_dense$d.so timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-count --also \
    > gpurun_out/dense$d.json 2> gpurun_out/dense$d.err || echo "dense$d failed"
  python -c "
import json; d=json.loads(open('gpurun_out/dense$d.json').read().strip().splitlines()[-1]); print('dense$d', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stages_ms'].items()}, d['verified'])"
done
KH_LIB_PATH=$PWD/tools/probes/variants/libkh_dense1.so timeout 240 python -m pytest tests/test_gpu_ctable.py tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/dense_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/dense_pytest.log
