"""Fused multi-tensor Adam for the training step: ONE launch updates every parameter, with the
reference's L2 loss term (reg_weight * sum ||W||^2 / 2 over the name-selected tensors, tacotron.py:144-146) folded in as
its exact gradient reg_weight * W.  Same update rule as the reference's torch.optim.Adam(lr, eps=hp.adam_eps)
(train.py:130): betas (0.9, 0.999), no amsgrad, no decoupled weight decay.  A torch.optim.Optimizer, so LambdaLR
(train.py:131) and state_dict() work on it."""
import torch

from . import train_ops as TO


def l2_selected_names(model):
    return {n for n, _ in model.named_parameters()
            if "weight" in n and "layer_norm" not in n and "batchnorm" not in n
            and "encoder.speaker_embed" not in n and "encoder.embed" not in n}


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, reg_weight=0.0, l2_params=()):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.reg_weight = float(reg_weight)
        self._l2_ids = {id(p) for p in l2_params}
        self._tables = {}
        self.grad_scale = 1.0   # e.g. 1 / world_size when gradients were summed, not averaged

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
            key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
            hit = self._tables.get(gi)
            if hit is None or hit[0] != key:   # grads are re-allocated by zero_grad(set_to_none=True): rebuild the pointer table
                tab = hit[1] if hit is not None else TO.MultiTable(ps[0].device)   # reused: pinned staging, async upload
                tab.build_opt([(p, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"], id(p) in self._l2_ids) for p in ps])
                hit = (key, tab)
                self._tables[gi] = hit
            step = self.state[ps[0]]["step"] + 1
            for p in ps:
                self.state[p]["step"] = step
            b1, b2 = group["betas"]
            TO.adam_multi(hit[1], float(group["lr"]), b1, b2, float(group["eps"]), step, self.reg_weight, self.grad_scale)
            # the kernel wrote through raw pointers: bump the version counters so that everything keyed on them (the bf16
            # operand copies of engine_train, the packed decode weights of engine) sees the update
            torch._C._increment_version(ps)
        return loss
