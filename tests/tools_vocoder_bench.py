"""Diagnostics (not a test): the mel -> waveform stage (csrc/vocoder.cu, 60 Griffin-Lim iterations) on the headline synthesis
shape (B=32 utterances x 1000 frames), next to the numpy oracle of utils/audio.py on the host (one utterance, scaled).
usage: python tests/tools_vocoder_bench.py [B] [T]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from tts_b200 import vocoder as V  # noqa: E402
from oracle import audio_oracle as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
mels = np.clip(rng.standard_normal((B, T, 80)) * 0.6 + np.linspace(1.5, -2.5, 80)[None, None, :], -4, 4).astype(np.float32)
eng = V.GriffinLim(dev)
md = torch.from_numpy(mels).to(dev)
lens = [T] * B
for _ in range(2):
    eng(md, lens)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    wav, counts = eng(md, lens)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t0 = time.perf_counter()
host = V.mel2wav_batch(md, lens)          # includes the device -> host copy of the waveforms
e2e_ms = 1e3 * (time.perf_counter() - t0)
n_cpu = min(T, 400)
t0 = time.perf_counter()
A.mel2wav(mels[0, :n_cpu])
cpu_s = time.perf_counter() - t0
audio_s = B * A.HOP * (T - 1) / A.SR
# algorithmic work: 121 transforms of 2048 points per frame, 5 N log2 N flop each
flops = B * T * (2 * A.N_ITER + 1) * 5.0 * 2048 * 11
print(json.dumps({"stage": "mel2wav (Griffin-Lim, 60 iterations, n_fft 2048, hop 200)", "batch": B, "frames": T,
                  "gpu_ms": ms, "gpu_e2e_ms_with_d2h": e2e_ms, "frames_per_s": B * T / (ms / 1e3),
                  "audio_seconds": audio_s, "realtime_factor": audio_s / (ms / 1e3), "fft_gflops": flops / (ms / 1e3) / 1e9,
                  "cpu_oracle": {"frames": n_cpu, "seconds": cpu_s, "frames_per_s": n_cpu / cpu_s, "cores": 1,
                                 "kind": "port (numpy restatement of utils/audio.py + librosa 0.6.0)"}}))
