"""autograd.Function wrappers that make the hand-written forward/backward of engine_train.TrainEngine visible to
PyTorch's autograd, so that the reference's training loop (train.py:171-174: `m(**batch)`, `compute_loss`,
`loss.backward()`, then torch.optim.Adam / DistributedDataParallel) runs unchanged on plain nn.Parameters.

The parameters are passed to `apply` only to register them in the graph (and so that DDP's gradient hooks fire); the
kernels read them through the engine's live references."""
import torch

from . import train_ops as TO
from .engine_train import TrainEngine


def _ret(n_lead, names, grads):
    return (None,) * n_lead + tuple(grads.get(n) for n in names)


class EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: TrainEngine, names, train, ids, in_len, spk, lang, *params):
        mem, saved = eng.encoder_fwd(ids, in_len, spk, lang, train)
        ctx.eng, ctx.names, ctx.saved = eng, names, saved
        return mem

    @staticmethod
    def backward(ctx, d_mem):
        grads = ctx.eng.encoder_bwd(ctx.saved, d_mem)
        ctx.saved = None
        return _ret(7, ctx.names, grads)


class DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: TrainEngine, names, train, leave_one, memory, in_len, targets, tg_len, *params):
        mels, stop, saved = eng.decoder_fwd(memory, in_len, targets, tg_len, train, leave_one)
        ctx.eng, ctx.names, ctx.saved = eng, names, saved
        ctx.set_materialize_grads(False)
        return mels, stop

    @staticmethod
    def backward(ctx, d_mels, d_stop):
        grads, d_mem = ctx.eng.decoder_bwd(ctx.saved, d_mels, d_stop, need_dmem=ctx.needs_input_grad[4])
        ctx.saved = None
        return (None, None, None, None, d_mem, None, None, None) + tuple(grads.get(n) for n in ctx.names)


class PostnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: TrainEngine, names, train, add_input, mels, lengths, *params):
        out, saved = eng.postnet_fwd(mels, lengths, train, add_input)
        ctx.eng, ctx.names, ctx.saved = eng, names, saved
        return out

    @staticmethod
    def backward(ctx, d_out):
        grads, d_mels = ctx.eng.postnet_bwd(ctx.saved, d_out)
        ctx.saved = None
        return (None, None, None, None, d_mels, None) + tuple(grads.get(n) for n in ctx.names)


class LossFn(torch.autograd.Function):
    """bef_loss, aft_loss, stop_loss, aft_losses of tacotron.py:136-158 in one kernel pass, gradients included."""

    @staticmethod
    def forward(ctx, mel_bef, mel_aft, stop, targets, lengths):
        lens = lengths.to(torch.int32).contiguous()
        total = lens.sum(dtype=torch.int32)
        f = lambda t: t.detach().float().contiguous()
        sums, aft_b, d_bef, d_aft, d_stop = TO.loss_fwd(f(mel_bef), f(mel_aft), f(stop), f(targets), lens, total)
        n = total.float()
        ctx.save_for_backward(d_bef, d_aft, d_stop, lens, n)
        ctx.set_materialize_grads(False)
        return sums[0] / n, sums[1] / n, sums[2] / n, aft_b / lens.float()

    @staticmethod
    def backward(ctx, g_bef, g_aft, g_stop, g_each):
        d_bef, d_aft, d_stop, lens, n = ctx.saved_tensors
        out_bef = d_bef * g_bef if g_bef is not None else None
        out_aft = d_aft * g_aft if g_aft is not None else None
        if g_each is not None:   # per-sample after-loss (logging only in train.py:216-223): rescale the unit gradient
            extra = d_aft * (g_each * n / lens.float())[:, None, None]
            out_aft = extra if out_aft is None else out_aft + extra
        out_stop = d_stop * g_stop if g_stop is not None else None
        return out_bef, out_aft, out_stop, None, None


class L2Fn(torch.autograd.Function):
    """reg_weight * sum(||W||^2) / 2 over the tensors tacotron.py:144-146 selects, as one multi-tensor launch."""

    @staticmethod
    def forward(ctx, reg_weight, table, *params):
        out = torch.empty((), device=params[0].device, dtype=torch.float32)
        TO.sumsq_multi(table, out)
        ctx.reg, ctx.params = reg_weight, params
        return out * (0.5 * reg_weight)

    @staticmethod
    def backward(ctx, g):
        scaled = torch._foreach_mul([p.detach() for p in ctx.params], ctx.reg)
        torch._foreach_mul_(scaled, g)
        return (None, None) + tuple(scaled)
