"""Diagnostic (not a test): a short decode run for ncu or compute-sanitizer.
    ncu --set full -k regex:pipelined_decode_kernel -s 1 -c 1 -o gpurun_out/pipe python tests/tools_ncu_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from oracle import tts_oracle as O  # noqa: E402
from tts_b200.engine import TtsEngine  # noqa: E402

B = int(os.environ.get("PB", "32"))
T0 = int(os.environ.get("PT", "200"))
NS = int(os.environ.get("PN", "2"))
cfg = O.ModelConfig(max_generation_frames=int(os.environ.get("PTMAX", "256")))
params = O.synth_params(cfg, seed=0)
params["decoder.stop_net.bias"] = torch.tensor([-1e4])
eng = TtsEngine.from_state_dict(params, cfg, "cuda:0")
batch = O.synth_batch(cfg, batch=B, text_len=258, n_frames=4, seed=1)
mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
sess = eng.new_session(B, 258, cfg.max_generation_frames, "encdec")
sess.begin(mem, batch["input_lengths"].cuda())
sess.step(T0, impl=int(os.environ.get("PIMPL", "0")))
torch.cuda.synchronize()
for _ in range(3):
    sess.step(NS, impl=int(os.environ.get("PIMPL", "0")))
    torch.cuda.synchronize()
print("done", sess.t)
