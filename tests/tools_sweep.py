"""Diagnostic / BASELINE config 5: long-form decode (T frames) batch sweep on one GPU.
    python tests/tools_sweep.py [T] [B ...]   ->  one JSON line per batch size"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
import bench  # noqa: E402
from oracle import tts_oracle as O  # noqa: E402
from tts_b200.engine import TtsEngine  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
batches = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8, 16, 32, 64, 128]
S = 258
cfg = O.ModelConfig(max_generation_frames=T)
params = O.synth_params(cfg, seed=0)
params["decoder.stop_net.bias"] = torch.tensor([-1e4])
eng = TtsEngine.from_state_dict(params, cfg, "cuda:0")
peak, _ = bench.measured_peaks()
for B in batches:
    batch = O.synth_batch(cfg, batch=B, text_len=S, n_frames=4, seed=1)
    mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    sess = eng.new_session(B, S, T, "none")
    best = None
    for rep in range(2):
        sess.begin(mem, batch["input_lengths"].cuda())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = 0
        while done < T:
            n = min(50, T - done)
            sess.step(n)
            done += n
        e1.record()
        torch.cuda.synchronize()
        secs = e0.elapsed_time(e1) / 1e3
        best = secs if best is None else min(best, secs)
    assert int(sess.counters[1].item()) == B, "stop fired / kernel error"
    alg = bench.decode_bytes(cfg, B, S, T)
    print(json.dumps({"config": "cfg5 long-form decode", "batch": B, "frames": T, "decode_seconds": best,
                      "frames_per_s": B * T / best, "us_per_step": 1e6 * best / T,
                      "algorithmic_GB": alg / 1e9, "achieved_GBps": alg / best / 1e9,
                      "frac_of_measured_hbm": alg / best / 1e9 / peak}))
    del sess
    torch.cuda.empty_cache()
