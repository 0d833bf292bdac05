# one gpurun call: the build-kernel geometry variants on the K=51 bench line, then the k-mer analysis bench
bash tools/probes/build_variants.sh run
for k in 19 51; do timeout 90 python tools/count_bench.py $k 4000000 8 150 > gpurun_out/t4_count$k.json 2> gpurun_out/t4_count$k.err; echo "count$k rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/t4_count$k.json')); print(d['ms_count_device'], d['occurrences_per_s'], d['roofline']['frac'], d['verified'])"; done
