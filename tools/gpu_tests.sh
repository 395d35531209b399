#!/bin/bash
# Runs on the GPU box (via gpurun): every `-m gpu` test file in its own process (a CUDA fault in one file
# cannot poison the others), logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
rc=0
for f in ${@:-tests/test_kernels_gpu.py}; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/$b.log
  r=${PIPESTATUS[0]}
  echo "== $f rc=$r"; tail -40 gpurun_out/$b.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
