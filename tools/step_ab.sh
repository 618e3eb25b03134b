#!/bin/bash
# On the GPU box: the GPU suite, then value / step / K1 of the three bench workloads with the summaries on their own
# stream (default), with every counting kernel waiting for them, and with them on the compute stream (same box, same
# run).   tools/step_ab.sh TAG
tag=${1:-x}
o=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $o/${tag}_pytest_gpu.log 2>&1; tail -4 $o/${tag}_pytest_gpu.log
run() {   # label, env...
  label=$1; shift
  for wl in ${WORKLOADS:-cfg2x12 cfg3 cfg5}; do
    env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>>$o/${tag}_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); r=d['roofline']
print('$label $wl: K1 %.2f us (pipeline %.2f) frac %.3f step %.2f us value %.3g step/kernel %.2f e2e %.3g launches %d' % (1e3*r['kernel_ms'], 1e3*r['kernel_ms_in_pipeline'], r['frac'], 1e3*d['ms_per_step'], d['value'], r['step_over_kernel'], d['e2e']['value'], d['gpu_launches']))" | tee -a $o/${tag}_step_ab.txt
  done
}
run default X=1
run join_before_count BASECOUNT_B200_JOIN_K1=1
run summaries_on_compute_stream BASECOUNT_B200_SUMMARY_STREAM=0
