"""CPU checks of bench.py's host logic: algorithmic-byte model, rank aggregation over gloo
(world_size 2), and the reference arm's JSON contract on a tiny sample."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_bytes_match_survey():
    import bench
    from oracle import tts_oracle as O
    cfg = O.ModelConfig()
    assert bench.decode_step_params(cfg) == 49919746            # SURVEY.md §8d W_step
    one = bench.decode_bytes(cfg, 32, 258, 1)
    assert abs(one - 506e6) / 506e6 < 0.01                      # BASELINE.md §4, t = 0
    total = bench.decode_bytes(cfg, 32, 258, 1000)
    assert abs(total - 1095.6e9) / 1095.6e9 < 0.005             # whole 1000-frame decode


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # replicas: every rank did 1000 units; rank 1 was slower -> the job's time is the slowest rank's
    got = bench.aggregate_throughput(1000, 2.0 if rank == 0 else 4.0, world)
    q.put((rank, got))
    dist.destroy_process_group()


def test_replica_aggregation_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29613, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res[0] == res[1] == 2 * 1000 / 4.0


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--batch", "2", "--text-len", "12", "--ref-horizon", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"].startswith("mel frames/sec")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    # non-zero ranks of a torchrun launch do no work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
