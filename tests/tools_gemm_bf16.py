"""Diagnostics (not a test): throughput of the bf16 tcgen05 GEMM (csrc/gemm_bf16.cu) on the shapes of the training step at
BASELINE configs[2] (B=64 x 1000 frames, 258 tokens), next to cuBLAS (torch.matmul) on the same shapes.
usage: python tests/tools_gemm_bf16.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-transformer-tts_b200"))

import torch  # noqa: E402

from tts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(torch.bfloat16)


rows = []
MD, ME = 64000, 16512
cases = []
for (name, M) in (("dec", MD), ("enc", ME)):
    for (N, K) in ((2304, 768), (768, 768), (3072, 768), (768, 3072)):
        cases.append(("fwd  %s M=%d N=%d K=%d" % (name, M, N, K), "fwd", M, N, K))
        cases.append(("dgrad %s M=%d N=%d K=%d" % (name, M, K, N), "dgrad", M, N, K))
        cases.append(("wgrad %s N=%d K=%d over M=%d" % (name, N, K, M), "wgrad", M, N, K))
cases.append(("fwd  cross-kv M=%d N=1536 K=768" % ME, "fwd", ME, 1536, 768))
cases.append(("fwd  prenet M=%d N=256 K=256" % MD, "fwd", MD, 256, 256))
cases.append(("fwd  mel M=%d N=80 K=768" % MD, "fwd", MD, 80, 768))

for label, kind, M, N, K in cases:
    flops = 2.0 * M * N * K
    if kind == "fwd":      # C[M,N] = A[M,K] W[N,K]^T, bf16 out
        a, w = bf(M, K), bf(N, K)
        ms = timeit(lambda: ops.gemm_bf16(a, w))
        res = torch.randn(M, N, device=dev)
        ms_f32 = timeit(lambda: ops.gemm_bf16(a, w, out_dtype=torch.float32))
        ms_res = timeit(lambda: ops.gemm_bf16(a, w, out_dtype=torch.float32, residual=res, drop_p=0.1, seed=1, rng_stream=3))
        label += " [+res+drop %.3f ms %.0f]" % (ms_res, flops / ms_res / 1e9)
        del res
        ref = timeit(lambda: torch.matmul(a, w.t()))
    elif kind == "dgrad":  # dX[M,K] = dY[M,N] W[N,K]: B operand = W read MN-major
        dy, w = bf(M, N), bf(N, K)
        ms = timeit(lambda: ops.gemm_bf16(dy, w, b_mn=True))
        ms_f32 = timeit(lambda: ops.gemm_bf16(dy, w, b_mn=True, out_dtype=torch.float32))
        ref = timeit(lambda: torch.matmul(dy, w))
    else:                  # dW[N,K] = dY[M,N]^T X[M,K]: both operands MN-major, fp32 out
        dy, x = bf(M, N), bf(M, K)
        best = None
        for sk in (1, 2, 4, 8):
            t = timeit(lambda: ops.gemm_bf16(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=sk))
            if best is None or t < best[0]:
                best = (t, sk)
        ms, ms_f32 = best[0], best[0]
        label += " (best split_k=%d)" % best[1]
        ref = timeit(lambda: torch.matmul(dy.t(), x))
    print("%-78s %7.3f ms %7.1f TFLOP/s | fp32 out %7.3f ms %7.1f | cuBLAS %7.3f ms %7.1f" %
          (label, ms, flops / ms / 1e9, ms_f32, flops / ms_f32 / 1e9, ref, flops / ref / 1e9))
