# The k-mer analysis stage alone: its GPU tests (with durations), the bench for K=19/31/51, and the ncu captures of
# kc_count_kernel (profiles/r02_*count*).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_count.py -m gpu -x -q --durations=8 > gpurun_out/t6_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/t6_pytest.log
for k in 19 31 51; do timeout 90 python tools/count_bench.py $k 4000000 8 150 > gpurun_out/t6_count$k.json 2> gpurun_out/t6_count$k.err; echo "count$k rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/t6_count$k.json')); print(d['ms_count_device'], d['occurrences_per_s'], d['roofline']['frac'], d['ms_count_host_buffers'], d['verified'])"; done
timeout 120 ncu --set full --clock-control none --import-source on -k regex:kc_count_kernel -s 2 -c 1 -f -o gpurun_out/r02_count_k51_v2 python tools/count_bench.py 51 4000000 8 150 > gpurun_out/t6_ncu.log 2>&1; echo "ncu rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_count_launches_v2.csv python tools/count_bench.py 51 4000000 8 150 > gpurun_out/t6_ncu2.log 2>&1; echo "ncu2 rc=$?"
