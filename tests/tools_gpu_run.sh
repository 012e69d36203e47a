#!/bin/bash
# Diagnostic driver for gpurun calls (not a test).  usage: bash tests/tools_gpu_run.sh <mode>
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
mode=${1:-tests}
case $mode in
  tests)      # whole GPU suite
    python -m pytest tests -m gpu -q -s > gpurun_out/r2_gputest_full.log 2>&1; grep -E "passed|failed|horizon|staggered|golden:|cfg5|tiny train|full train|loss curves|FAILED|Error" gpurun_out/r2_gputest_full.log | tail -40 ;;
  sanitize)   # compute-sanitizer on the decode kernel: split K/V streams (B=3), two groups (B=32), partial last group (B=40)
    for B in 3 32 40; do
      for tool in memcheck racecheck synccheck; do
        PB=$B PT=6 PN=2 PTMAX=16 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/tools_ncu_target.py 2>&1 | tail -15 > gpurun_out/r2_sanitizer_${tool}_B$B.log
        tail -3 gpurun_out/r2_sanitizer_${tool}_B$B.log
      done
    done ;;
  bench)
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err ;;
esac
