# Tuning run for ct_build_kernel's geometry (buckets per chunk for 128-bit slots, threads per block, blocks per SM):
# builds one library per variant into tools/probes/variants/ and, with "run", times the K=51 bench line with each
# (KH_LIB_PATH selects the library).  bash tools/probes/build_variants.sh build|run
set -e
cd "$(dirname "$0")/../.."
VARIANTS="2176_512_2 1408_384_3 1408_448_3 2176_640_2"
if [ "$1" = "build" ]; then
  mkdir -p tools/probes/variants
  for v in $VARIANTS; do
    IFS=_ read b t m <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DKH_CT_BUCKETS2=$b -DKH_CT_THREADS=$t -DKH_CT_MINBLOCKS=$m \
      -Xcompiler -fPIC -shared cs267_hw3_b200/csrc/capi.cu cs267_hw3_b200/csrc/count.cu -o tools/probes/variants/libkh_v_$v.so
  done
else
  mkdir -p gpurun_out
  for v in $VARIANTS; do
    KH_LIB_PATH=$PWD/tools/probes/variants/libkh_v_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-count --also \
      > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err || echo "variant $v failed"
    python -c "
import json; d=json.loads(open('gpurun_out/variant_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stages_ms'].items()}, d['verified'])"
  done
fi
