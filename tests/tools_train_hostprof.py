"""Diagnostics (not a test): host-side (Python) profile of the training step of bench.py --workload train.
usage: python tests/tools_train_hostprof.py"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402

from tts_b200 import synthetic as O  # noqa: E402
from tts_b200.config import hparams_from  # noqa: E402
from tts_b200.optim import FusedAdam, l2_selected_names  # noqa: E402
from transformer import tacotron  # noqa: E402

dev = torch.device("cuda:0")
cfg = O.ModelConfig()
hp = hparams_from(cfg)
hp.l2_in_optimizer = True
torch.manual_seed(0)
m = tacotron.Tacotron(hp)
m.load_state_dict(O.synth_params(cfg, seed=0), strict=True)
m.to(dev).train()
host = O.synth_batch(cfg, batch=64, text_len=258, n_frames=1000, seed=100)
keys = ("inputs", "input_lengths", "mel_targets", "target_lengths", "input_spk_ids", "input_language_vecs")
batch = {k: host[k].to(dev) for k in keys}
sel = l2_selected_names(m)
opt = FusedAdam(m.parameters(), lr=hp.max_lr, eps=hp.adam_eps, reg_weight=hp.reg_weight,
                l2_params=[p for n, p in m.named_parameters() if n in sel])


def step():
    out = m(**batch)
    losses = tacotron.compute_loss(m, batch["mel_targets"], batch["target_lengths"], out, hp)
    opt.zero_grad()
    losses["loss"].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
for what in ("forward", "loss", "backward", "opt"):
    pass
# phase timing (host issue time with an idle GPU, then GPU completion)
t0 = time.perf_counter(); out = m(**batch); t1 = time.perf_counter()
losses = tacotron.compute_loss(m, batch["mel_targets"], batch["target_lengths"], out, hp); t2 = time.perf_counter()
opt.zero_grad(); losses["loss"].backward(); t3 = time.perf_counter()
opt.step(); t4 = time.perf_counter()
torch.cuda.synchronize(); t5 = time.perf_counter()
print("host issue ms: forward %.1f loss %.1f backward %.1f opt %.1f | wait for GPU after issue %.1f | total %.1f" %
      (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t5 - t0)))
pr = cProfile.Profile()
pr.enable()
for _ in range(2):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
