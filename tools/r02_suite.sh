#!/bin/bash
# Round-2 measurement suite for ONE GPU (run through gpurun).  Sections are selected by name:
#   tools/r02_suite.sh <tag> tests smoke bench2 bench1 ref ncu_list ncu_full micro
# Outputs under gpurun_out/<tag>_*.
TAG=${1:-r02}; shift
mkdir -p gpurun_out
P=gpurun_out/$TAG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${P}_nvsmi.txt 2>&1
echo "host cores: $(nproc) ; mem $(free -g | awk '/Mem/{print $2}') GB" | tee -a ${P}_nvsmi.txt
for S in "$@"; do
  case $S in
  tests)
    timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > ${P}_pytest.log 2>&1; echo "pytest rc=$?"
    grep -E "bench-shape parity|passed|failed|error" ${P}_pytest.log | tail -40 ;;
  smoke)
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 ${P}_smoke.log ;;
  bench2)
    HI3D_SKIP_CPU_BASELINE=1 timeout 1500 python bench.py --steps 2 --warmup 3 > ${P}_bench_s2.json 2> ${P}_bench_s2.err; echo "bench s2 rc=$?"
    cut -c1-600 ${P}_bench_s2.json; tail -3 ${P}_bench_s2.err ;;
  bench2full)
    timeout 2400 python bench.py --steps 3 --warmup 3 > ${P}_bench_s2_full.json 2> ${P}_bench_s2_full.err; echo "bench s2 full rc=$?"
    cut -c1-600 ${P}_bench_s2_full.json; tail -3 ${P}_bench_s2_full.err ;;
  bench1)
    HI3D_SKIP_CPU_BASELINE=1 timeout 900 python bench.py --stage 1 --steps 3 --warmup 3 > ${P}_bench_s1.json 2> ${P}_bench_s1.err; echo "bench s1 rc=$?"
    cut -c1-400 ${P}_bench_s1.json ;;
  ref1)
    timeout 900 python bench.py --impl reference --stage 1 --steps 1 --warmup 0 --ref-budget-s 300 > ${P}_bench_ref_s1.json 2> ${P}_bench_ref_s1.err; echo "ref s1 rc=$?"
    cut -c1-1200 ${P}_bench_ref_s1.json ;;
  ref2)
    timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 --ref-budget-s 600 > ${P}_bench_ref_s2.json 2> ${P}_bench_ref_s2.err; echo "ref s2 rc=$?"
    cut -c1-1200 ${P}_bench_ref_s2.json ;;
  ncu_list)
    timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches_s2.csv python tools/one_step.py --stage 2 > ${P}_ncu_list.log 2>&1; echo "ncu list rc=$?" ;;
  ncu_list1)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches_s1.csv python tools/one_step.py --stage 1 > ${P}_ncu_list1.log 2>&1; echo "ncu list s1 rc=$?" ;;
  ncu_full)
    # ~1.6 MB per captured kernel; gpurun_out/ travels back only below 64 MiB: capture to /tmp, copy if small
    timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc5|fmha_tc5|gn_|layernorm|tattn" --launch-skip 0 --launch-count ${NCU_COUNT:-28} -f -o /tmp/${TAG}_full_s2 python tools/one_step.py --stage 2 > ${P}_ncu_full.log 2>&1; echo "ncu full rc=$?"
    ls -la /tmp/${TAG}_full_s2.ncu-rep
    if [ $(stat -c %s /tmp/${TAG}_full_s2.ncu-rep 2>/dev/null || echo 999999999) -lt 50000000 ]; then cp /tmp/${TAG}_full_s2.ncu-rep gpurun_out/; fi
    ncu -i /tmp/${TAG}_full_s2.ncu-rep --page raw --csv > ${P}_full_s2_raw.csv 2>/dev/null; ls -la ${P}_full_s2_raw.csv ;;
  micro)
    timeout 600 python tools/microbench.py --engine tc5 --stage 2 --only attn > ${P}_micro_attn.log 2>&1; echo "micro attn rc=$?"; tail -8 ${P}_micro_attn.log
    timeout 600 python tools/microbench.py --engine tc5 --stage 2 --only norm > ${P}_micro_norm.log 2>&1; echo "micro norm rc=$?"; tail -12 ${P}_micro_norm.log ;;
  *) echo "unknown section $S" ;;
  esac
done
