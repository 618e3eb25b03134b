#!/bin/bash
# On the GPU box: tools/k1_times.py for the product library and every other library given (same box, same run).
#   tools/k1_ab_quick.sh TAG [other.so ...]      WORKLOADS=cfg2x12,cfg5 to restrict
tag=${1:-x}; shift
for lib in "" "$@"; do
  BASECOUNT_B200_LIB=$lib timeout 300 python tools/k1_times.py --workloads ${WORKLOADS:-cfg2x12,cfg3,cfg5} 2>>gpurun_out/${tag}_err.log | tee -a gpurun_out/${tag}_k1_times.txt
done
