"""Data-parallel gradient exchange of the training step (train.py:34-41,122-127).

The reference wraps the model in DistributedDataParallel; that keeps working on this package (plain nn.Parameters,
gradients produced by autograd.Functions).  `GradBuckets` is the explicit equivalent used by bench.py and by callers that
drive the engine without DDP: gradients are flattened into a few large fp32 buckets and all-reduced (NCCL over
NVLink / NVSwitch; gloo on CPU for the tests) on a side stream as soon as each module's backward has produced them,
then averaged and scattered back.  One process per GPU, torch.distributed for the plumbing."""
import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, params, bucket_bytes=64 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = []          # [(flat buffer, [(param, offset, numel)])]
        cur, cur_n, cap = [], 0, max(1, bucket_bytes // 4)
        for p in reversed(self.params):   # backward produces gradients roughly in reverse registration order
            if cur and cur_n + p.numel() > cap:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
            cur.append((p, cur_n, p.numel()))
            cur_n += p.numel()
        if cur:
            self._close(cur, cur_n)
        dev = self.params[0].device
        self.stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None

    def _close(self, items, n):
        dev, dt = items[0][0].device, torch.float32
        self.buckets.append((torch.zeros(n, device=dev, dtype=dt), items))

    def broadcast_parameters(self, src=0):
        """C1 of SURVEY §2.3: every rank starts from rank `src`'s weights (and BatchNorm buffers via `extra`)."""
        if self.world == 1:
            return
        for p in self.params:
            dist.broadcast(p.data, src, group=self.group)

    def allreduce_mean(self):
        """All-reduce every gradient (sum over ranks / world).  Returns once the results are back in p.grad."""
        if self.world == 1:
            return
        works = []
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            for flat, items in self.buckets:
                for p, off, n in items:
                    if p.grad is not None:
                        flat[off:off + n].copy_(p.grad.reshape(-1))
                    else:
                        flat[off:off + n].zero_()
                works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for w, (flat, items) in zip(works, self.buckets):
                w.wait()
                flat.mul_(1.0 / self.world)
                for p, off, n in items:
                    if p.grad is not None:
                        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
