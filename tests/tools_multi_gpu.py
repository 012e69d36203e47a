"""Multi-GPU check of the data-parallel training step (run under torchrun on N GPUs of one box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/tools_multi_gpu.py

Every rank holds the same tiny model (dropout 0) and its own batch.  (1) GradBuckets: after the bucketed NCCL all-reduce each
rank's gradients equal the mean of the per-rank gradients (recomputed locally on rank 0 from all batches).  (2) The same
through torch's DistributedDataParallel wrapped around the drop-in module, i.e. the reference's train.py:122-127 path.
Prints one PASS / FAIL line per check on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tts_b200 import synthetic as O
    from tts_b200.config import hparams_from
    from tts_b200.dist import GradBuckets
    from transformer import tacotron
    cfg = O.ModelConfig.tiny()
    hp = hparams_from(cfg)
    hp.transformer_dropout_rate = hp.decoder_dropout_rate = 0.0
    params = O.synth_params(cfg, seed=15)

    def model():
        m = tacotron.Tacotron(hp)
        m.load_state_dict(params, strict=True)
        return m.to(dev).train()

    def batch_of(r):
        b = O.synth_batch(cfg, batch=3, text_len=20, n_frames=30, seed=40 + r, ragged=True)
        return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}

    def grads_of(m, fwd, b):
        m.zero_grad()
        out = fwd(**b)
        tacotron.compute_loss(m, b["mel_targets"], b["target_lengths"], out, hp)["loss"].backward()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters()}

    # reference on every rank: mean over ranks of the local gradients, computed without communication
    m = model()
    per_rank = [grads_of(m, m, batch_of(r)) for r in range(world)]
    want = {n: sum(g[n] for g in per_rank) / world for n in per_rank[0]}

    def check(name, got):
        worst = max(float((got[n] - want[n]).norm() / want[n].norm().clamp_min(1e-12)) for n in want)
        same = torch.tensor([worst], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("%s %s: worst relative gradient error over ranks %.2e (world %d)" % ("PASS" if float(same) < 2e-3 else "FAIL", name, float(same), world))

    m1 = model()
    gb = GradBuckets(m1.parameters(), bucket_bytes=1 << 16)
    gb.broadcast_parameters(0)
    g = grads_of(m1, m1, batch_of(rank))
    gb.allreduce_mean()
    torch.cuda.synchronize()
    check("GradBuckets (bucketed NCCL all-reduce, %d buckets)" % len(gb.buckets), {n: p.grad for n, p in m1.named_parameters()})

    m2 = model()
    ddp = torch.nn.parallel.DistributedDataParallel(m2, device_ids=[local], output_device=local)
    grads_of(m2, ddp, batch_of(rank))
    torch.cuda.synchronize()
    check("DistributedDataParallel around the drop-in Tacotron (train.py:122-127)", {n: p.grad for n, p in m2.named_parameters()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
