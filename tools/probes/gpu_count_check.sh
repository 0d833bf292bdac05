# The k-mer analysis stage alone: its GPU tests (with durations), the bench for K=19/31/51 with the kernel-variant sweep.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_count.py -m gpu -x -q --durations=12 > gpurun_out/t3_pytest.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/t3_pytest.log
for k in 19 31 51; do timeout 90 python tools/count_bench.py $k 4000000 8 150 --sweep > gpurun_out/t3_count$k.json 2> gpurun_out/t3_count$k.err; echo "count$k rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/t3_count$k.json')); print(d['ms_count_device'], d['occurrences_per_s'], d['roofline']['frac'], d['variants_ms'], d['verified'])"; done
