"""Diagnostics (not a test): time the training attention kernels (csrc/attn_train.cu) at the train-step shapes and, as a
yardstick only, the flash_attn 2.x library kernels (mma.sync code built for sm_80, run on the same B200) on the same problem.
usage: python tests/tools_attn_bench.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from tts_b200 import train_ops as TO  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, dh = 8, 96
D = H * dh
dev = "cuda:0"


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, Tq, Tk, causal in (("decoder self (causal)", 1000, 1000, True), ("cross", 1000, 258, False)):
    for p in (0.1, 0.0):
        q = torch.randn(B * Tq, D, device=dev).to(torch.bfloat16)
        kv = torch.randn(B * Tk, 2 * D, device=dev).to(torch.bfloat16)
        dctx = torch.randn(B * Tq, D, device=dev).to(torch.bfloat16)
        k, v = kv[:, :D], kv[:, D:]
        flops_f = 4.0 * B * H * Tq * Tk * dh * (0.5 if causal else 1.0)
        ctx, lse = TO.attn_fwd(q, k, v, B, H, Tq, Tk, dh, causal, None, p, 7, 3)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        tf = timeit(lambda: TO.attn_fwd(q, k, v, B, H, Tq, Tk, dh, causal, None, p, 7, 3))
        tb = timeit(lambda: TO.attn_bwd(q, k, v, ctx, lse, dctx, dq, dkv[:, :D], dkv[:, D:], B, H, Tq, Tk, dh, causal, None, p, 7, 3))
        line = "%-22s p=%.1f  ours fwd %.3f ms %6.1f TFLOP/s  bwd %.3f ms %6.1f TFLOP/s" % (
            name, p, tf, flops_f / tf / 1e9, tb, 2.5 * flops_f / tb / 1e9)
        try:
            from flash_attn import flash_attn_func
            q4 = q.view(B, Tq, H, dh).clone().requires_grad_(True)
            k4 = k.reshape(B, Tk, H, dh).clone().requires_grad_(True)
            v4 = v.reshape(B, Tk, H, dh).clone().requires_grad_(True)
            do4 = dctx.view(B, Tq, H, dh)
            ff = timeit(lambda: flash_attn_func(q4, k4, v4, dropout_p=p, causal=causal))
            out = flash_attn_func(q4, k4, v4, dropout_p=p, causal=causal)

            def bwd():
                q4.grad = k4.grad = v4.grad = None
                out.backward(do4, retain_graph=True)
            fb = timeit(bwd)
            line += " | flash_attn %s fwd %.3f ms %6.1f  bwd %.3f ms %6.1f" % (__import__("flash_attn").__version__, ff, flops_f / ff / 1e9, fb, 2.5 * flops_f / fb / 1e9)
        except Exception as exc:  # noqa: BLE001
            line += " | flash_attn unavailable: %r" % (exc,)
        print(line)
