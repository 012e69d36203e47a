"""Diagnostic (not a test): per-phase SM-clock breakdown of one step of the pipelined decode kernel (impl 4).
    python tests/tools_pipe_profile.py [step ...]      (PB = batch, default 32)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from oracle import tts_oracle as O  # noqa: E402
from tts_b200.engine import TtsEngine  # noqa: E402

steps = [int(x) for x in sys.argv[1:]] or [1, 500]
B = int(os.environ.get("PB", "32"))
cfg = O.ModelConfig(max_generation_frames=1024)
params = O.synth_params(cfg, seed=0)
params["decoder.stop_net.bias"] = torch.tensor([-1e4])
eng = TtsEngine.from_state_dict(params, cfg, "cuda:0")
batch = O.synth_batch(cfg, batch=B, text_len=258, n_frames=4, seed=1)
mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
sess = eng.new_session(B, 258, 1024, "encdec")
sess.begin(mem, batch["input_lengths"].cuda())
L = cfg.n_decoder_layer
names = ["pre0", "pre1", "pre2"] + ["L%d.%s" % (l, k) for l in range(L)
                                     for k in ("qkv", "self", "oproj", "cq", "cross", "coproj", "ffn1", "ffn2", "red")] + ["final"]
done = 0
for target in steps:
    if target > done:
        sess.step(target - done, impl=4)
        done = target
    sess.step(1, impl=4)
    done += 1
    torch.cuda.synchronize()
    p = sess.phase_profile(len(names)).astype(float) / 1.965e3   # us at 1965 MHz
    first = 0 if target == 0 else 1   # steps t > 0 skip pre0 (fused into the previous step's final phase): its row is stale
    print("== step t=%d  total %.1f us (first stamp to last)" % (target, p[-1, 5] - p[first, 0]))
    if first:
        p[0, :] = 0.0
    kinds = {}
    for n, d in zip(names, p):
        kinds.setdefault(n.split(".")[-1], []).append(d)
    print("   %-7s %3s %8s %8s %8s %8s %8s" % ("phase", "n", "g0 wait", "g0 work", "g1 wait", "g1 work", "span"))
    tot = [0.0] * 5
    for k, v in kinds.items():
        n = len(v)
        cols = [sum(d[1] - d[0] for d in v) / n, sum(d[2] - d[1] for d in v) / n, sum(d[4] - d[3] for d in v) / n,
                sum(d[5] - d[4] for d in v) / n, sum(d[5] - d[0] for d in v) / n]
        for i in range(5):
            tot[i] += cols[i] * n
        print("   %-7s %3d %8.2f %8.2f %8.2f %8.2f %8.2f" % ((k, n) + tuple(cols)))
    print("   %-7s %3s %8.1f %8.1f %8.1f %8.1f %8.1f" % (("sum", "") + tuple(tot)))
    print("   GEMM group 0 detail (us): frags | res-loads | W wait | products | stats + C + barrier | epilogue | end barrier")
    for k, v in kinds.items():
        if k in ("self", "cross", "red", "final"):
            continue
        n = len(v)
        f = lambda i, j: sum(d[i] - d[j] for d in v) / n
        print("   %-7s %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f" % (k, f(6, 1), f(11, 6), f(7, 11), f(8, 7), f(9, 8), f(10, 9),
                                                                   f(2, 10)))
    print("   attention group 0 detail (us): setup+q | tile loop (warp 0) | wait other warps | merge+ctx+align | unit-end barrier | group-end")
    for k in ("self", "cross"):
        v = kinds[k]
        n = len(v)
        f = lambda i, j: sum(d[i] - d[j] for d in v) / n
        print("   %-7s %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f" % (k, f(6, 1), f(7, 6), f(8, 7), f(9, 8), f(10, 9), f(2, 10)))
        print("   %-7s setup detail: x_empty arrive %5.2f | unit setup %5.2f | q wait %5.2f | q load+scale %5.2f" % (
            k, f(11, 1), f(12, 11), f(13, 12), f(6, 13)))
