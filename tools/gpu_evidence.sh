#!/bin/bash
# On the GPU box: the whole evidence set of one build.
#   tools/gpu_evidence.sh TAG   -> gpurun_out/TAG_*
tag=${1:-x}
o=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $o/${tag}_pytest_gpu.log 2>&1; tail -3 $o/${tag}_pytest_gpu.log
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"; cut -c1-600 $o/${tag}_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/phase_times.py > $o/${tag}_phase_times.txt 2>&1; cat $o/${tag}_phase_times.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --workload cfg2x12 --steps 2 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_launches.log 2>&1; echo "launches rc=$?"
for wl in ${NCU_WORKLOADS-cfg2x12 cfg3 cfg5}; do      # NCU_WORKLOADS="" skips the captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_count_fast -s 2 -c 1 -f -o $o/prof_${tag}_k1_${wl} python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_k1_${wl}.log 2>&1; echo "ncu $wl rc=$?"
done
