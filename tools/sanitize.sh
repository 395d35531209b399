#!/bin/bash
# compute-sanitizer over the kernels that hand-roll mbarrier / TMEM / peer-memory protocols (run through gpurun; slow):
#   tools/sanitize.sh [memcheck|racecheck|synccheck]   -> gpurun_out/sanitize_<tool>.log (summary copied to profiles/)
TOOL=${1:-memcheck}
mkdir -p gpurun_out
LOG=gpurun_out/sanitize_${TOOL}.log
: > $LOG
for T in "tests/test_gemm_tc5_gpu.py -k 'plain or conv3x3 or temporal'" "tests/test_attn_tc5_gpu.py -k '(lean-emu25 or split-mufu) and (1024 or moving)'" \
         "tests/test_kernels_gpu.py -k 'groupnorm or layernorm or temporal_attention'"; do
  echo "=== compute-sanitizer --tool $TOOL python -m pytest $T" >> $LOG
  eval timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 3 --print-limit 20 python -m pytest $T -x -q -m gpu -p no:cacheprovider >> $LOG 2>&1
  echo "rc=$?" >> $LOG
done
grep -E "^===|rc=|ERROR SUMMARY|passed|failed" $LOG
