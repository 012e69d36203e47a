cd $GRAFT_REPO_ROOT
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --workload train --dp $DP 2>/dev/null | grep "^{" | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('%s %s: %.2f ms/step, allreduce %s' % ('$DP', '$*', d['value'], d.get('allreduce') and d['allreduce']['ms_median']))"; }
DP=buckets run X=1
DP=buckets run NCCL_MAX_NCHANNELS=4
DP=buckets run NCCL_MAX_NCHANNELS=8
DP=buckets-late run X=1
python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | cut -c1-120
