import os, sys
ROOT = "/root/repo" if os.path.exists("/root/repo/few-shot-transformer-tts_b200") else os.environ["GRAFT_REPO_ROOT"]
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch
from tts_b200 import train_ops as TO
dev = "cuda:0"
R, C = 64000, 768
x = torch.randn(R, C, device=dev); g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
dy = torch.randn(R, C, device=dev).to(torch.bfloat16); dres = torch.randn(R, C, device=dev)
y, m, r = TO.ln_fwd(x, g, b)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
a = t(lambda: TO.ln_bwd(dy, x, m, r, g, dres=dres))
bb = t(lambda: TO.ln_bwd(dy, x, m, r, g, dres=dres, cast_drop=(0.1, 1, 2)))
c = t(lambda: TO.dropout_cast(dres, 0.1, 1, 2))
byts = R * C * (2 + 4 + 4 + 4)
print("ln_bwd %.1f us (%.0f GB/s)  fused+cast %.1f us (%.0f GB/s)  separate cast %.1f us" % (a, byts / a / 1e3, bb, (byts + R * C * 2) / bb / 1e3, c))
