"""Diagnostic (not a test): per-block SM-clock stamps of one CTA of the tcgen05 attention backward kernel (csrc/attn_bwd_tc.cuh).
usage: build with TTS_EXTRA_NVCC_FLAGS=-DTTS_ATTN_TC_TRACE_BUILD, then TTS_ATTN_TC_TRACE=1 python tests/tools_attn_trace.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from tts_b200 import _native, train_ops as TO  # noqa: E402

B, H, T, dh = int(os.environ.get("PB", "64")), 8, 1000, 96
D = H * dh
dev = "cuda:0"
qkv = torch.randn(B * T, 3 * D, device=dev).to(torch.bfloat16)
dctx = torch.randn(B * T, D, device=dev).to(torch.bfloat16)
for _ in range(2):
    ctx, lse = TO.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, T, T, dh, True, None, 0.1, 7, 3)
    dqkv = torch.empty_like(qkv)
    TO.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], ctx, lse, dctx, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B, H, T, T,
                dh, True, None, 0.1, 7, 3)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 32 * 8))()
assert _native.load().tts_attn_tc_trace(ctypes.cast(buf, ctypes.c_void_p)) == 0
t = [[[buf[(r * 32 + i) * 8 + k] for k in range(8)] for i in range(32)] for r in range(3)]
t0 = min(x for r in t for i in r for x in i if x > 0)
names = (("softmax", ("loop top", "q_full", "s_full", "tmem ld", "computed", "p_empty", "stored")),
         ("mma issuers", ("A:top", "A:q_full", "A:S issued", "B:top", "B:p_full", "B:issued")),
         ("dq / producer", ("top", "dq_full", "loaded", "P:q_empty", "P:issued", "Q complete", "P:stats in")))
for r, (name, ev) in enumerate(names):
    print("== %s (SM clocks since the first stamp; 1 us ~ 1900 clocks)" % name)
    print("blk " + " ".join("%10s" % e for e in ev))
    for i in range(17):
        if t[r][i][0] == 0:
            continue
        print("%3d " % i + " ".join("%10d" % (t[r][i][k] - t0 if t[r][i][k] else -1) for k in range(len(ev))))
