# What one gpurun call at the end of the round runs: GPU tests, smoke, the bench line (default workload + the long-contig
# and K=31 shapes).  bash tools/probes/gpu_round_check.sh  (from the repo root)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/t9_pytest.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/t9_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t9_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/t9_smoke.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/t9_bench.json 2> gpurun_out/t9_bench.err; echo "bench rc=$?"
timeout 120 python bench.py --steps 5 --warmup 3 --workload chr14_k51_long --also --no-cpu-baseline > gpurun_out/t9_bench_long.json 2> gpurun_out/t9_bench_long.err; echo "bench long rc=$?"
timeout 120 python bench.py --steps 5 --warmup 3 --workload chr14_k31 --also --no-cpu-baseline > gpurun_out/t9_bench_k31.json 2> gpurun_out/t9_bench_k31.err; echo "bench k31 rc=$?"
