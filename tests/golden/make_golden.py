"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these files —
outputs of the reference's own ``Tacotron.forward``, ``compute_loss`` and
``synthesize.eval_batch`` on seeded synthetic inputs — are what pins ``oracle/`` and,
through it, the CUDA path.  Weights are not stored: they are rebuilt from the seed by
``oracle.tts_oracle.synth_params`` (a fingerprint is stored to detect RNG drift).
"""
import copy
import json
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TTS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
for name in ("librosa", "librosa.filters", "librosa.effects", "soundfile", "fastdtw", "matplotlib",
             "matplotlib.pyplot", "editdistance"):
    sys.modules.setdefault(name, mock.MagicMock())

from hyperparams import hparams as ref_hp  # noqa: E402  (reference)
from transformer import tacotron as ref_tacotron  # noqa: E402  (reference)
import synthesize as ref_synth  # noqa: E402  (reference)
from oracle import tts_oracle as O  # noqa: E402


def hp_for(cfg: O.ModelConfig):
    hp = copy.deepcopy(ref_hp)
    for k, v in vars(cfg).items():
        hp.set_hparam(k, v)
    return hp


def ref_model(cfg, params):
    m = ref_tacotron.Tacotron(hp_for(cfg))
    m.load_state_dict(params, strict=True)  # proves the state-dict schema of synth_params
    return m


def model_inputs(batch):
    return {k: v for k, v in batch.items() if k != "names"}


def run_eval_batch(cfg, model, batch, max_frames):
    saved = {k: getattr(ref_hp, k) for k in ("max_generation_frames", "num_mels")}
    ref_hp.set_hparam("max_generation_frames", max_frames)
    ref_hp.set_hparam("num_mels", cfg.num_mels)
    try:
        data = {k: batch[k] for k in ("inputs", "input_lengths", "input_spk_ids", "input_language_vecs", "names")}
        out = ref_synth.eval_batch(model, data, use_bar=False, bar_interval=-1)
    finally:
        for k, v in saved.items():
            ref_hp.set_hparam(k, v)
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1e3))


def stop_margin(model, cfg, params, batch, max_frames):
    """Smallest |stop logit| seen on a live frame — reported so that 'bit-exact stop index'
    tests can state how close to the decision boundary the reference was."""
    out = O.eval_batch_cached(params, cfg, batch, max_frames)
    lg, ln = out["stop_logits"], out["generated_lengths"]
    live = torch.arange(lg.shape[1])[None, :] < ln[:, None]
    return float(lg.abs()[live].min()) if live.any() else float("nan")


def long_ar(n_frames=640):
    """G6: the reference's own eval_batch over a LONG horizon (B=2, text 64 tokens, stop disabled): pins the
    K/V-cached formulation (oracle and CUDA) against the reference's uncached O(T^2) loop across hundreds of
    dependent steps.  ~10 minutes of CPU; run with `python tests/golden/make_golden.py long`."""
    cfg = O.ModelConfig()
    params = O.synth_params(cfg, seed=0)
    params["decoder.stop_net.bias"] = torch.tensor([-1e4])
    model = ref_model(cfg, params).eval()
    batch = O.synth_batch(cfg, batch=2, text_len=64, n_frames=4, seed=9, ragged=True)
    res = run_eval_batch(cfg, model, batch, n_frames)
    print("G6 generated_lengths", res["generated_lengths"], res["mel_pre"].shape)
    save("full_ar_long.npz", max_frames=np.int64(n_frames), mel_pre=res["mel_pre"], mel_aft=res["mel_aft"],
         generated_lengths=np.asarray(res["generated_lengths"], dtype=np.int32))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "long":
        return long_ar()

    # ---- G1: cfg 1 — full model, B=1, 120-byte text -> 400 frames, teacher forced ------------
    cfg = O.ModelConfig()
    params = O.synth_params(cfg, seed=0)
    model = ref_model(cfg, params).eval()
    batch = O.synth_batch(cfg, batch=1, text_len=122, n_frames=400, seed=1)
    with torch.no_grad():
        out = model(**model_inputs(batch))
        loss = ref_tacotron.compute_loss(model, batch["mel_targets"], batch["target_lengths"], out, hp_for(cfg))
    save("cfg1_forward.npz",
         params_checksum=np.float64(O.params_checksum(params)),
         mel_bef=out["mel_bef"].numpy(), mel_aft=out["mel_aft"].numpy(), stop_logits=out["stop_logits"].numpy(),
         align_self_l0_h0=out["alignments"]["self"][0][0, 0, ::8, ::8].numpy(),
         align_encdec_l5_h7=out["alignments"]["encdec"][5][0, 7, :, ::8].numpy(),
         **{"loss_" + k: v.numpy() for k, v in loss.items()})

    # ---- G2: full model, ragged teacher-forced batch + loss ------------------------------------
    batch = O.synth_batch(cfg, batch=3, text_len=40, n_frames=64, seed=2, ragged=True)
    with torch.no_grad():
        out = model(**model_inputs(batch))
        loss = ref_tacotron.compute_loss(model, batch["mel_targets"], batch["target_lengths"], out, hp_for(cfg))
    save("full_ragged_forward.npz",
         mel_bef=out["mel_bef"].numpy(), mel_aft=out["mel_aft"].numpy(), stop_logits=out["stop_logits"].numpy(),
         **{"loss_" + k: v.numpy() for k, v in loss.items()})

    # ---- G3: full model, autoregressive eval_batch with staggered stops ------------------------
    params_ar = dict(params)
    params_ar["decoder.stop_net.bias"] = torch.tensor([-4.48])
    model_ar = ref_model(cfg, params_ar).eval()
    batch = O.synth_batch(cfg, batch=4, text_len=48, n_frames=4, seed=3, ragged=True)
    max_frames = 40
    res = run_eval_batch(cfg, model_ar, batch, max_frames)
    print("G3 generated_lengths", res["generated_lengths"])
    save("full_ar.npz", stop_bias=np.float32(-4.48), max_frames=np.int64(max_frames),
         mel_pre=res["mel_pre"], mel_aft=res["mel_aft"],
         generated_lengths=np.asarray(res["generated_lengths"], dtype=np.int32),
         align_encdec_l5=res["alignments"]["encdec"][5][:, :, :, -1],
         align_self_l0=res["alignments"]["self"][0][:, :, :, -1],
         stop_margin=np.float64(stop_margin(model_ar, cfg, params_ar, batch, max_frames)))

    # ---- G4: tiny model — forward, loss, gradients (dropout 0, batch-stat BN), long AR ----------
    tcfg = O.ModelConfig.tiny()
    tparams = O.synth_params(tcfg, seed=15)
    thp = hp_for(tcfg)
    thp.set_hparam("transformer_dropout_rate", 0.0)
    thp.set_hparam("decoder_dropout_rate", 0.0)
    tmodel = ref_tacotron.Tacotron(thp)
    tmodel.load_state_dict(tparams, strict=True)
    tbatch = O.synth_batch(tcfg, batch=3, text_len=20, n_frames=30, seed=6, ragged=True)
    tmodel.eval()
    with torch.no_grad():
        out = tmodel(**model_inputs(tbatch))
        loss = ref_tacotron.compute_loss(tmodel, tbatch["mel_targets"], tbatch["target_lengths"], out, thp)
    arrays = dict(mel_bef=out["mel_bef"].numpy(), mel_aft=out["mel_aft"].numpy(),
                  stop_logits=out["stop_logits"].numpy(),
                  **{"loss_" + k: v.numpy() for k, v in loss.items()})
    tmodel.train()
    out = tmodel(**model_inputs(tbatch))
    loss = ref_tacotron.compute_loss(tmodel, tbatch["mel_targets"], tbatch["target_lengths"], out, thp)
    loss["loss"].backward()
    arrays.update(train_mel_aft=out["mel_aft"].detach().numpy(), train_loss=loss["loss"].detach().numpy())
    gnames = sorted(n for n, p in tmodel.named_parameters() if p.grad is not None)
    arrays["grad_names"] = np.array(gnames)
    arrays["grad_norms"] = np.array([float(dict(tmodel.named_parameters())[n].grad.double().norm()) for n in gnames])
    arrays["grad_prenet_dense0"] = dict(tmodel.named_parameters())["decoder.prenet.dense0.weight"].grad.numpy()
    save("tiny_forward_loss_grad.npz", **arrays)

    tparams_ar = dict(tparams)
    tparams_ar["decoder.stop_net.bias"] = torch.tensor([-0.7])
    tmodel_ar = ref_model(tcfg, tparams_ar).eval()
    tb = O.synth_batch(tcfg, batch=5, text_len=24, n_frames=4, seed=7, ragged=True)
    res = run_eval_batch(tcfg, tmodel_ar, tb, 60)
    print("G4 generated_lengths", res["generated_lengths"])
    save("tiny_ar.npz", stop_bias=np.float32(-0.7), max_frames=np.int64(60),
         mel_pre=res["mel_pre"], mel_aft=res["mel_aft"],
         generated_lengths=np.asarray(res["generated_lengths"], dtype=np.int32),
         stop_margin=np.float64(stop_margin(tmodel_ar, tcfg, tparams_ar, tb, 60)))

    # ---- G5: the reference's own init under torch.manual_seed(0) (train.py:33,118-119) ----------
    torch.manual_seed(0)
    m = ref_tacotron.Tacotron(ref_hp)
    ref_tacotron.initialize_variables(m)
    sd = m.state_dict()
    fp = {k: {"shape": list(v.shape), "sum": float(v.double().sum()), "abs": float(v.double().abs().sum()),
              "head": [float(x) for x in v.flatten()[:3].double()]} for k, v in sd.items()}
    with open(os.path.join(HERE, "ref_init_seed0.json"), "w") as f:
        json.dump({"n_entries": len(sd), "n_params": sum(p.numel() for p in m.parameters()),
                   "lr_factor": {str(s): ref_tacotron.learning_rate_schedule(s, ref_hp)
                                 for s in (0, 50000, 100000, 600000, 2000000)},
                   "tensors": fp}, f, indent=0)
    print("wrote ref_init_seed0.json", len(sd), "entries")


if __name__ == "__main__":
    main()
