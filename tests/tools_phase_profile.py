"""Diagnostic (not a test): per-phase SM-clock breakdown of one fused decode step.
    python tests/tools_phase_profile.py [step ...]"""
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
names = ["pre0", "pre1", "pre2"] + ["L%d.%s" % (l, k) for l in range(6)
                                     for k in ("qkv", "self", "oproj", "cq", "cross", "coproj", "ffn1", "ffn2")] + ["final"]
done = 0
for target in steps:
    if target > done:
        sess.step(target - done)
        done = target
    sess.step(1)
    done += 1
    torch.cuda.synchronize()
    p = sess.phase_profile().astype(float) / 1.965e3   # us at 1965 MHz
    comp, wait = p[:, 1] - p[:, 0], p[:, 2] - p[:, 1]
    print("== step t=%d  total %.1f us  (compute %.1f, barrier %.1f)" % (target, p[-1, 2] - p[0, 0], comp.sum(), wait.sum()))
    kinds = {}
    for n, c, w in zip(names, comp, wait):
        k = n.split(".")[-1]
        kinds.setdefault(k, []).append((c, w))
    for k, v in kinds.items():
        print("   %-7s n=%2d compute avg %6.2f us  barrier avg %6.2f us" % (k, len(v), sum(c for c, _ in v) / len(v),
                                                                          sum(w for _, w in v) / len(v)))
    for i in (0, 3, 5, 6, 9, 10):   # detailed stamps of a few GEMM phases
        d = p[i]
        print("   %-9s x-staged %5.2f  LN %5.2f  W-wait %5.2f  fma+reduce %5.2f  epilogue %5.2f (us, last pass)" % (
            names[i], d[3] - d[0], d[4] - d[3], d[5] - d[4], d[6] - d[5], d[7] - d[6]))
