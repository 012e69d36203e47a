"""Diagnostic (not a test): one forward + backward of the training attention at the decoder's self-attention shape, for ncu:
    ncu --set full --import-source on -k regex:attn_ -c 3 -o gpurun_out/attn python tests/tools_attn_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from tts_b200 import train_ops as TO  # noqa: E402

B, H, T, dh = int(os.environ.get("PB", "16")), 8, int(os.environ.get("PT", "1000")), 96
D = H * dh
dev = "cuda:0"
qkv = torch.randn(B * T, 3 * D, device=dev).to(torch.bfloat16)
dctx = torch.randn(B * T, D, device=dev).to(torch.bfloat16)
for _ in range(int(os.environ.get("PN", "2"))):
    ctx, lse = TO.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, T, T, dh, True, None, 0.1, 7, 3)
    dqkv = torch.empty_like(qkv)
    TO.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], ctx, lse, dctx, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B, H, T, T,
                dh, True, None, 0.1, 7, 3)
torch.cuda.synchronize()
print("done")
