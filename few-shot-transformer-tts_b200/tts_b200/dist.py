"""Data-parallel gradient exchange of the training step (train.py:34-41,122-127).

The reference wraps the model in DistributedDataParallel; that keeps working on this package (plain nn.Parameters,
gradients produced by autograd.Functions).  `GradBuckets` is the explicit equivalent used by bench.py and by callers that
drive the engine without DDP: gradients are flattened into a few large fp32 buckets and all-reduced (NCCL over
NVLink / NVSwitch; gloo on CPU for the tests) on a side stream as soon as each module's backward has produced them,
then averaged and scattered back.  One process per GPU, torch.distributed for the plumbing."""
import torch
import torch.distributed as dist


class GradBuckets:
    """params: an iterable of parameters, or a list of parameter GROUPS in the order their gradients become available in
    backward (e.g. [postnet, decoder, encoder] for Tacotron: each sub-module's hand-written backward delivers all of its
    gradients at once).  With groups, `attach_hooks()` starts a group's all-reduce on the side stream the moment its
    gradients exist, so it runs behind the rest of the backward pass; `finish()` then only waits."""

    def __init__(self, params, bucket_bytes=64 << 20, group=None):
        params = list(params)
        groups = params if params and isinstance(params[0], (list, tuple)) else [params]
        self.groups = [[p for p in g if p.requires_grad] for g in groups]
        self.params = [p for g in self.groups for p in g]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = []          # [(flat buffer, [(param, offset, numel)])]
        self.group_buckets = []    # bucket indices per parameter group
        cap = max(1, bucket_bytes // 4)
        for g in self.groups:
            first = len(self.buckets)
            cur, cur_n = [], 0
            for p in reversed(g):   # backward produces gradients roughly in reverse registration order
                if cur and cur_n + p.numel() > cap:
                    self._close(cur, cur_n)
                    cur, cur_n = [], 0
                cur.append((p, cur_n, p.numel()))
                cur_n += p.numel()
            if cur:
                self._close(cur, cur_n)
            self.group_buckets.append(list(range(first, len(self.buckets))))
        dev = self.params[0].device
        self.stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self._pending = []
        self._hooks = []
        self._done_groups = set()

    # ---- overlapped mode: one all-reduce per parameter group, started from an autograd hook ----------------------------
    def attach_hooks(self):
        """Start each group's all-reduce as soon as autograd has produced all of its gradients (once per backward)."""
        if self.world == 1 or self._hooks:
            return
        for gi, g in enumerate(self.groups):
            self._hooks.append(torch.autograd.graph.register_multi_grad_hook(tuple(g), lambda grads, gi=gi: self._launch_group(gi, grads),
                                                                             mode="all"))

    def _launch_group(self, gi, grads=None):
        """The multi-grad hook fires BEFORE autograd accumulates into p.grad: reduce the fresh gradients themselves."""
        g = self.groups[gi]
        fresh = {id(p): gr for p, gr in zip(g, grads)} if grads is not None else {}
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            for bi in self.group_buckets[gi]:
                flat, items = self.buckets[bi]
                dsts, srcs = [], []
                for p, off, n in items:
                    src = fresh.get(id(p))
                    if src is None:
                        src = p.grad
                    if src is not None:
                        dsts.append(flat[off:off + n])
                        srcs.append(src.reshape(-1))
                    else:
                        flat[off:off + n].zero_()
                if dsts:   # one multi-tensor copy per bucket instead of one small kernel per parameter
                    torch._foreach_copy_(dsts, srcs)
                self._pending.append((bi, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)))
        self._done_groups.add(gi)

    def finish(self):
        """After backward: reduce the groups whose hook did not fire, wait for everything, write the means into p.grad.
        (Hooked groups were reduced from the fresh gradients, so p.grad must not hold an older accumulation: call
        zero_grad() before backward as the training loop of train.py:173 does.)"""
        if self.world == 1:
            return
        for gi in range(len(self.groups)):
            if gi not in self._done_groups:
                self._launch_group(gi)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            for bi, work in self._pending:
                work.wait()
                flat, items = self.buckets[bi]
                flat.mul_(1.0 / self.world)
                dsts = [p.grad for p, off, n in items if p.grad is not None]
                srcs = [flat[off:off + n].view_as(p.grad) for p, off, n in items if p.grad is not None]
                if dsts:   # the scatter back is on the critical path (after backward): one multi-tensor copy per bucket
                    torch._foreach_copy_(dsts, srcs)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending = []
        self._done_groups = set()

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
