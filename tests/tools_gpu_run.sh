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
  train)      # BASELINE metric 2: the teacher-forced training step at configs[2], plus its launch list
    python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/r2_train_bench.json 2> gpurun_out/r2_train_bench.err; cut -c1-1800 gpurun_out/r2_train_bench.json; tail -5 gpurun_out/r2_train_bench.err
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --workload train --steps 1 --warmup 3 > /dev/null 2>&1
    python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_train_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.Counter(); cnt = collections.Counter()
data = rows[hdr + 2:]
n = len(data)
# the last quarter of the launches is the single timed step (3 warm-up steps + 1)
for r in data[3 * n // 4:]:
    try:
        agg[r[kn][:70]] += float(r[mv].replace(",", "")); cnt[r[kn][:70]] += 1
    except (ValueError, IndexError):
        pass
tot = sum(agg.values())
for k, v in agg.most_common(25):
    print("%8.3f ms %5.1f%% x%4d %s" % (v / 1e6, 100 * v / tot, cnt[k], k))
print("total %.3f ms over %d launches" % (tot / 1e6, sum(cnt.values())))
PY
    ;;
  multi)      # N-GPU box: gradient equality (GradBuckets + DDP) and the train step at N GPUs ($2 = N)
    N=${2:-2}
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/tools_multi_gpu.py 2>&1 | grep -E "PASS|FAIL|Error|error" | tail -8
    for dp in buckets ddp; do
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --workload train --dp $dp > gpurun_out/r2_train_n${N}_${dp}.json 2> gpurun_out/r2_train_n${N}_${dp}.err
      python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2_train_n${N}_${dp}.json') if l.startswith('{')][-1])
print('N=${N} ${dp}: %.2f ms/step, %.0f frames/s, allreduce %s' % (d['value'], d['frames_per_s'], d.get('allreduce')))" || tail -5 gpurun_out/r2_train_n${N}_${dp}.err
    done
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cut -c1-400 gpurun_out/r2_bench_n${N}.json
    ;;
esac
