"""Diagnostics (not a test): ONE shape of the bf16 tcgen05 GEMM in a loop, as an ncu target.
usage: python tests/tools_gemm_bf16_one.py M N K out(bf16|f32) [reps] [epilogue: plain | relu_drop | res_drop]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-transformer-tts_b200"))
import torch  # noqa: E402

from tts_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
out = torch.float32 if sys.argv[4] == "f32" else torch.bfloat16
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda:0")
a = (torch.randn(M, K, device=dev) * 0.05).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
epi = sys.argv[6] if len(sys.argv) > 6 else "plain"
res = torch.randn(M, N, device=dev) if epi == "res_drop" else None
for _ in range(reps):
    if epi == "relu_drop":      # the FFN hidden layer (modules.py:16-18)
        ops.gemm_bf16(a, w, out_dtype=out, act=ops.ACT_RELU, drop_p=0.1, seed=1, rng_stream=3)
    elif epi == "res_drop":     # the output projections (modules.py:132,138)
        ops.gemm_bf16(a, w, out_dtype=out, residual=res, drop_p=0.1, seed=1, rng_stream=3)
    else:
        ops.gemm_bf16(a, w, out_dtype=out)
torch.cuda.synchronize()
print("done")
