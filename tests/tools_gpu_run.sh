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
        PB=$B PT=6 PN=2 PTMAX=16 timeout 900 compute-sanitizer --tool $tool --print-limit 40 python tests/tools_ncu_target.py 2>&1 | tail -60 > gpurun_out/r2_sanitizer_${tool}_B$B.log
        tail -3 gpurun_out/r2_sanitizer_${tool}_B$B.log
      done
    done ;;
  evidence)   # round-2 evidence set: full GPU suite, default bench, launch lists, ncu --set full of the dominant kernels
    python -m pytest tests -m gpu -q -s > gpurun_out/r2_gputest_full.log 2>&1; grep -E "passed|failed|horizon|staggered|golden:|cfg5|tiny train|full train|loss curves|compaction|FAILED|Error" gpurun_out/r2_gputest_full.log | tail -40
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
    python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/r2_train_bench.json 2> gpurun_out/r2_train_bench.err; cut -c1-200 gpurun_out/r2_train_bench.json
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches_decode.csv python bench.py --steps 1 --warmup 3 --no-train --no-module-api --no-cpu-baseline > /dev/null 2>&1
    rm -f /tmp/gl.txt; TTS_GEMM_LOG=/tmp/gl.txt ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
    cp /tmp/gl.txt gpurun_out/r2_gemm_log.txt; python tests/tools_gemm_join.py gpurun_out/r2_gemm_log.txt gpurun_out/r2_train_launches.csv > gpurun_out/r2_train_gemm_by_shape.txt 2>&1
    python tests/tools_launch_summary.py gpurun_out/r2_train_launches.csv > gpurun_out/r2_train_launches_summary.txt 2>&1; head -12 gpurun_out/r2_train_launches_summary.txt
    timeout 100 python tests/tools_attn_bench.py > gpurun_out/r2_attn_bench.txt 2>&1
    timeout 100 python tests/tools_vocoder_bench.py 2> /dev/null | tail -1 > gpurun_out/r2_vocoder_bench.json; cut -c1-200 gpurun_out/r2_vocoder_bench.json
    python bench.py --workload train --cfg4 --steps 5 --warmup 3 2> /dev/null | tail -1 > gpurun_out/r2_train_cfg4.json
    PIMPL=4 PT=475 PN=50 PTMAX=640 ncu --set full --clock-control none --import-source on -k regex:pipelined_decode_kernel -s 1 -c 1 -f -o gpurun_out/r2_pipe python tests/tools_ncu_target.py > /dev/null 2>&1
    ncu -i gpurun_out/r2_pipe.ncu-rep --page raw --csv > gpurun_out/r2_pipe_ncu_raw.csv
    PB=16 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 3 -f -o gpurun_out/r2_attn python tests/tools_attn_target.py > /dev/null 2>&1
    ncu -i gpurun_out/r2_attn.ncu-rep --page raw --csv > gpurun_out/r2_attn_ncu_raw.csv
    for o in bf16 f32; do
      ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -f -o gpurun_out/r2_gemm_bf16_$o python tests/tools_gemm_bf16_one.py 64000 768 768 $o > /dev/null 2>&1
      ncu -i gpurun_out/r2_gemm_bf16_$o.ncu-rep --page raw --csv > gpurun_out/r2_gemm_bf16_${o}_ncu_raw.csv
    done
    timeout 300 python tests/tools_gemm_bf16.py 10 > gpurun_out/r2_gemm_bf16_sweep.txt 2>&1
    python tests/tools_pipe_profile.py 500 > gpurun_out/r2_pipe_phase_profile.txt 2>&1
    ls -la gpurun_out | grep r2_ | tail -30 ;;
  sweep)      # BASELINE configs[4]: long-form decode, T=2000, batch sweep
    python tests/tools_sweep.py 2000 1 2 4 8 16 32 64 128 > gpurun_out/r2_cfg5_sweep.jsonl 2> gpurun_out/r2_cfg5_sweep.err; cat gpurun_out/r2_cfg5_sweep.jsonl | cut -c1-260; tail -3 gpurun_out/r2_cfg5_sweep.err ;;
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
  scale)      # N-GPU box ($2 = N): gradient equality, then the data-parallel train step with the bucketed all-reduce
    N=${2:-8}
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/tools_multi_gpu.py 2>&1 | grep -E "PASS|FAIL|Error|error" | tail -4
    NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --workload train --dp buckets > gpurun_out/r2_train_n${N}_buckets.json 2> gpurun_out/r2_train_n${N}_buckets.err
    grep -E "NVLS|via P2P|Channel .* via" gpurun_out/r2_train_n${N}_buckets.json gpurun_out/r2_train_n${N}_buckets.err | cut -c1-160 | sort | uniq -c | sort -rn | head -6 > gpurun_out/r2_train_n${N}_nccl.txt; cat gpurun_out/r2_train_n${N}_nccl.txt
    grep "^{" gpurun_out/r2_train_n${N}_buckets.json | tail -1 | cut -c1-1200 ;;
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
