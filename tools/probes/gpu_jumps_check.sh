# A/B of KH_CT_JUMPS (pointer-jumping passes per block-wide barrier in ct_build_kernel phase 4): the K=51 bench line with
# 1 / 2 / 3, then the chunk-table parity tests against the fastest of 2 / 3.  The three libraries are built beforehand with
#   for j in 1 2 3; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DKH_CT_JUMPS=$j -Xcompiler -fPIC \
#     -shared cs267_hw3_b200/csrc/capi.cu cs267_hw3_b200/csrc/count.cu -o tools/probes/variants/libkh_jumps$j.so; done
mkdir -p gpurun_out
for j in 1 2 3; do
  KH_LIB_PATH=$PWD/tools/probes/variants/libkh_jumps$j.so timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-count --also \
    > gpurun_out/jumps$j.json 2> gpurun_out/jumps$j.err || echo "jumps$j failed"
  python -c "
import json; d=json.loads(open('gpurun_out/jumps$j.json').read().strip().splitlines()[-1]); print('jumps$j', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stages_ms'].items()}, d['verified'])"
done
best=$(python -c "
import json
t={j: json.loads(open(f'gpurun_out/jumps{j}.json').read().strip().splitlines()[-1])['stages_ms']['ms_build'] for j in (2,3)}
print(min(t, key=t.get))")
echo "testing jumps$best"
KH_LIB_PATH=$PWD/tools/probes/variants/libkh_jumps$best.so timeout 200 python -m pytest tests/test_gpu_ctable.py tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/jumps_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/jumps_pytest.log
