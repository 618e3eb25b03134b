#!/bin/bash
# On the GPU box: parity of the counting kernels, phase times, and one ncu capture of the counting kernel.
#   tools/k1_gpu_check.sh TAG [kernel-regex]
tag=${1:-x}
kre=${2:-k1_count_fast}
timeout 600 python -m pytest tests/test_gpu_counts.py -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python tools/phase_times.py > gpurun_out/${tag}_phase.txt 2>&1; cat gpurun_out/${tag}_phase.txt
ncu --set full --clock-control none --import-source on -k regex:${kre} -s 2 -c 1 -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page details 2>/dev/null | grep -E "Duration|Registers Per|Issue Slots Busy|Executed Instructions|Warp Cycles Per Issued|Block Limit (Reg|Shared)" 
