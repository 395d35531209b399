#!/bin/bash
# Round-end measurement suite for ONE GPU (run through gpurun): parity tests, smoke, both bench arms, stage-2 bench,
# ncu launch list of one UNet forward and one `--set full` capture of the top kernels.  Outputs under gpurun_out/fs_*.
mkdir -p gpurun_out
T=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/fs_nvsmi.txt 2>&1
if [ "$T" = all ] || [ "$T" = tests ]; then
  timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/fs_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/fs_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fs_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/fs_smoke.log
fi
if [ "$T" = all ] || [ "$T" = bench ]; then
  timeout 900 python bench.py > gpurun_out/fs_bench_s1.json 2> gpurun_out/fs_bench_s1.err; echo "bench s1 rc=$?"; cut -c1-400 gpurun_out/fs_bench_s1.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fs_bench_ref.json 2> gpurun_out/fs_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/fs_bench_ref.json
  timeout 1200 python bench.py --stage 2 --steps 2 --warmup 3 > gpurun_out/fs_bench_s2.json 2> gpurun_out/fs_bench_s2.err; echo "bench s2 rc=$?"; cut -c1-400 gpurun_out/fs_bench_s2.json
fi
if [ "$T" = all ] || [ "$T" = ncu ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fs_launches_s1.csv python tools/one_step.py > gpurun_out/fs_ncu_list.log 2>&1; echo "ncu list rc=$?"
  # ~1.6 MB per captured kernel and gpurun_out/ travels back only below 64 MiB: capture to /tmp, copy if small
  timeout 900 ncu --set full --clock-control none -k "regex:gemm_tc5|fmha_tc5|gn_apply|gn_stats|layernorm|tattn" --launch-skip 8 --launch-count 22 -f -o /tmp/fs_full_s1 python tools/one_step.py > gpurun_out/fs_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la /tmp/fs_full_s1.ncu-rep
  if [ $(stat -c %s /tmp/fs_full_s1.ncu-rep 2>/dev/null || echo 999999999) -lt 45000000 ]; then cp /tmp/fs_full_s1.ncu-rep gpurun_out/; fi
  ncu -i /tmp/fs_full_s1.ncu-rep --page raw --csv > gpurun_out/fs_full_s1_raw.csv 2>/dev/null; ls -la gpurun_out/fs_full_s1_raw.csv
fi
