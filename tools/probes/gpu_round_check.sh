# What one gpurun call of this round runs: GPU tests, smoke, the bench line, the k-mer analysis bench for K=19/31, and the
# ncu captures of kc_count_kernel (profiles/r02_count_*).  bash tools/probes/gpu_round_check.sh  (from the repo root)
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/t2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t2_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/t2_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/t2_bench.json 2> gpurun_out/t2_bench.err; echo "bench rc=$?"
for k in 19 31; do timeout 60 python tools/count_bench.py $k 4000000 8 150 > gpurun_out/t2_count$k.json 2> gpurun_out/t2_count$k.err; echo "count$k rc=$?"; done
timeout 120 ncu --set full --clock-control none --import-source on -k regex:kc_count_kernel -s 2 -c 1 -f -o gpurun_out/r02_count_k51 python tools/count_bench.py 51 4000000 8 150 > gpurun_out/t2_ncu.log 2>&1; echo "ncu rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_count_launches.csv python tools/count_bench.py 51 4000000 8 150 > gpurun_out/t2_ncu2.log 2>&1; echo "ncu2 rc=$?"
