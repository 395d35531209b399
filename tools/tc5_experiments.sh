#!/bin/bash
# bottleneck experiments for the tc5 GEMM: time the same shapes with parts of the kernel disabled
for d in 0 1 2 3 4 7 8 15; do
  echo "== HI3D_TC5_DBG=$d"
  HI3D_TC5_DBG=$d python tools/microbench.py --stage 1 --engine tc5 --only gemm 2>&1 | grep -E "C=320" | cut -c1-120
done
