#!/bin/bash
# Round-2 end-of-round evidence on ONE GPU (run through gpurun): everything that is cited from profiles/r02_*.
#   tools/r02_final.sh [sections...]   default: tests smoke bench ref ncu sanitize
mkdir -p gpurun_out
P=gpurun_out/${R02_TAG:-r02f}
S=${@:-tests smoke bench ref ncu sanitize}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${P}_nvsmi.txt 2>&1
echo "host cores: $(nproc)" >> ${P}_nvsmi.txt
for X in $S; do
  case $X in
  tests)
    timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s -rs > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
    grep -E "passed|failed|error" ${P}_pytest_gpu.log | tail -3 ;;
  smoke)
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${P}_smoke.log ;;
  bench)
    timeout 1800 python bench.py --steps 3 --warmup 3 > ${P}_bench_stage2.json 2> ${P}_bench_stage2.err; echo "bench rc=$?"
    cut -c1-300 ${P}_bench_stage2.json ;;
  benchab)   # same box, previous FMHA kernel (split pipelines, MUFU only): A/B of the round's last kernel change
    HI3D_FMHA_VARIANT=1 HI3D_FMHA_EMU=0 HI3D_SKIP_CPU_BASELINE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-stage1 > ${P}_bench_stage2_fmha_split.json 2> ${P}_bench_stage2_fmha_split.err; echo "benchab rc=$?"
    cut -c1-200 ${P}_bench_stage2_fmha_split.json ;;
  ref)
    timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err; echo "ref rc=$?"
    cut -c1-900 ${P}_bench_reference_arm.json ;;
  ncu)
    timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches_s2.csv python tools/one_step.py --stage 2 > ${P}_ncu_list.log 2>&1; echo "ncu list rc=$?"
    timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc5|fmha_tc5|gn_|layernorm|tattn" --launch-skip 0 --launch-count 24 -f -o /tmp/r02f_full_s2 python tools/one_step.py --stage 2 > ${P}_ncu_full.log 2>&1; echo "ncu full rc=$?"
    if [ $(stat -c %s /tmp/r02f_full_s2.ncu-rep 2>/dev/null || echo 999999999) -lt 55000000 ]; then cp /tmp/r02f_full_s2.ncu-rep gpurun_out/; fi
    ncu -i /tmp/r02f_full_s2.ncu-rep --page raw --csv > ${P}_full_s2_raw.csv 2>/dev/null; ls -la ${P}_full_s2_raw.csv ;;
  sanitize)
    bash tools/sanitize.sh memcheck; cp gpurun_out/sanitize_memcheck.log ${P}_sanitize_memcheck.log ;;
  micro)
    timeout 600 python tools/microbench.py --engine tc5 --stage 2 > ${P}_microbench_s2.log 2>&1; tail -40 ${P}_microbench_s2.log ;;
  esac
done
