"""Diagnostic (not a test): the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu, TTS_GEMM_TC=1) against a float64 product,
and its speed next to the FFMA2 kernel (run once with TTS_GEMM_TC=1 and once without)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from tts_b200 import ops  # noqa: E402

torch.manual_seed(0)
shapes = [(256, 128, 64), (8256, 512, 512), (8256, 2048, 512), (8256, 512, 2048), (8256, 1536, 768), (32000, 512, 2560),
          (1000, 200, 96)]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    y = ops.linear(x, w, bias=b, act=ops.ACT_RELU)
    torch.cuda.synchronize()
    want = torch.relu(x.double() @ w.double().t() + b.double())
    err = (y.double() - want).abs().max().item()
    for _ in range(3):
        ops.linear(x, w, bias=b, act=ops.ACT_RELU)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear(x, w, bias=b, act=ops.ACT_RELU)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("M=%6d N=%5d K=%5d  max|err| %.3e  %.3f ms  %.1f TFLOP/s (fp32-equivalent)  TTS_GEMM_TC=%s" % (
        M, N, K, err, ms, 2.0 * M * N * K / ms / 1e9, os.environ.get("TTS_GEMM_TC", "0")))
