"""Pins oracle/tts_oracle.py to the outputs of the reference itself (tests/golden/*.npz,
produced by tests/golden/make_golden.py from /root/reference).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import tts_oracle as O

TOL = 2e-5  # fp32 CPU restatement vs fp32 CPU reference: different op order only


def _np(t):
    return t.detach().numpy()


def test_param_schema_matches_reference_state_dict(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "ref_init_seed0.json")))
    cfg = O.ModelConfig()
    mine = dict(O.param_shapes(cfg) + O.buffer_shapes(cfg))
    assert len(mine) == ref["n_entries"] == 177
    assert set(mine) == set(ref["tensors"])
    for k, shape in mine.items():
        assert list(shape) == ref["tensors"][k]["shape"], k
    n_params = sum(int(np.prod(s)) for _, s in O.param_shapes(cfg))
    assert n_params == ref["n_params"] == 83477155


def test_weights_rebuild_bit_identically(golden_dir, full_params):
    cfg, params = full_params
    z = np.load(os.path.join(golden_dir, "cfg1_forward.npz"))
    assert O.params_checksum(params) == float(z["params_checksum"])


def test_cfg1_teacher_forced_forward(golden_dir, full_params):
    cfg, params = full_params
    z = np.load(os.path.join(golden_dir, "cfg1_forward.npz"))
    batch = O.synth_batch(cfg, batch=1, text_len=122, n_frames=400, seed=1)
    with torch.no_grad():
        out = O.tacotron_forward(params, cfg, batch)
        loss = O.compute_loss(params, cfg, batch["mel_targets"], batch["target_lengths"], out)
    for k in ("mel_bef", "mel_aft", "stop_logits"):
        assert np.abs(_np(out[k]) - z[k]).max() < TOL, k
    a = out["alignments"]
    assert np.abs(_np(a["self"][0][0, 0, ::8, ::8]) - z["align_self_l0_h0"]).max() < 1e-6
    assert np.abs(_np(a["encdec"][5][0, 7, :, ::8]) - z["align_encdec_l5_h7"]).max() < 1e-6
    for k, v in loss.items():
        np.testing.assert_allclose(_np(v), z["loss_" + k], rtol=2e-5, atol=1e-7, err_msg=k)


def test_full_ragged_forward_and_loss(golden_dir, full_params):
    cfg, params = full_params
    z = np.load(os.path.join(golden_dir, "full_ragged_forward.npz"))
    batch = O.synth_batch(cfg, batch=3, text_len=40, n_frames=64, seed=2, ragged=True)
    with torch.no_grad():
        out = O.tacotron_forward(params, cfg, batch)
        loss = O.compute_loss(params, cfg, batch["mel_targets"], batch["target_lengths"], out)
    for k in ("mel_bef", "mel_aft", "stop_logits"):
        assert np.abs(_np(out[k]) - z[k]).max() < TOL, k
    for k, v in loss.items():
        np.testing.assert_allclose(_np(v), z["loss_" + k], rtol=2e-5, atol=1e-7, err_msg=k)


def test_full_autoregressive_staggered_stops(golden_dir, full_params):
    cfg, params = full_params
    z = np.load(os.path.join(golden_dir, "full_ar.npz"))
    params = dict(params)
    params["decoder.stop_net.bias"] = torch.tensor([float(z["stop_bias"])])
    batch = O.synth_batch(cfg, batch=4, text_len=48, n_frames=4, seed=3, ragged=True)
    T = int(z["max_frames"])
    with torch.no_grad():
        slow = O.eval_batch_uncached(params, cfg, batch, T, return_align=True)
        fast = O.eval_batch_cached(params, cfg, batch, T)
    assert len(set(z["generated_lengths"].tolist())) >= 3, "golden case must have staggered stops"
    for out in (slow, fast):
        assert out["generated_lengths"].dtype == torch.int32
        assert out["generated_lengths"].tolist() == z["generated_lengths"].tolist()
        assert np.abs(_np(out["mel_pre"]) - z["mel_pre"]).max() < TOL
        assert np.abs(_np(out["mel_aft"]) - z["mel_aft"]).max() < TOL
    al = slow["alignments"]
    assert np.abs(_np(al["encdec"][5][:, :, :, -1]) - z["align_encdec_l5"]).max() < 1e-6
    assert np.abs(_np(al["self"][0][:, :, :, -1]) - z["align_self_l0"]).max() < 1e-6
    assert float(z["stop_margin"]) > 1e-2  # the reference itself was not near the decision boundary


def test_tiny_forward_loss_and_gradients(golden_dir, tiny_params):
    cfg, params = tiny_params
    z = np.load(os.path.join(golden_dir, "tiny_forward_loss_grad.npz"))
    batch = O.synth_batch(cfg, batch=3, text_len=20, n_frames=30, seed=6, ragged=True)
    with torch.no_grad():
        out = O.tacotron_forward(params, cfg, batch)
        loss = O.compute_loss(params, cfg, batch["mel_targets"], batch["target_lengths"], out)
    for k in ("mel_bef", "mel_aft", "stop_logits"):
        assert np.abs(_np(out[k]) - z[k]).max() < TOL, k
    for k, v in loss.items():
        np.testing.assert_allclose(_np(v), z["loss_" + k], rtol=2e-5, atol=1e-7, err_msg=k)
    # train-mode statistics in the Postnet + gradients through the whole restatement
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in params.items()}
    out = O.tacotron_forward(leaf, cfg, batch, batch_stats=True)
    loss = O.compute_loss(leaf, cfg, batch["mel_targets"], batch["target_lengths"], out)
    loss["loss"].backward()
    assert np.abs(_np(out["mel_aft"]) - z["train_mel_aft"]).max() < 5e-5
    np.testing.assert_allclose(_np(loss["loss"]), z["train_loss"], rtol=2e-5)
    names = [str(n) for n in z["grad_names"]]
    norms = np.array([float(leaf[n].grad.double().norm()) for n in names])
    np.testing.assert_allclose(norms, z["grad_norms"], rtol=2e-3, atol=1e-7)
    assert np.abs(_np(leaf["decoder.prenet.dense0.weight"].grad) - z["grad_prenet_dense0"]).max() < 1e-5


def test_tiny_autoregressive_long(golden_dir, tiny_params):
    cfg, params = tiny_params
    z = np.load(os.path.join(golden_dir, "tiny_ar.npz"))
    params = dict(params)
    params["decoder.stop_net.bias"] = torch.tensor([float(z["stop_bias"])])
    batch = O.synth_batch(cfg, batch=5, text_len=24, n_frames=4, seed=7, ragged=True)
    with torch.no_grad():
        slow = O.eval_batch_uncached(params, cfg, batch, int(z["max_frames"]))
        fast = O.eval_batch_cached(params, cfg, batch, int(z["max_frames"]))
    for out in (slow, fast):
        assert out["generated_lengths"].tolist() == z["generated_lengths"].tolist()
        assert np.abs(_np(out["mel_pre"]) - z["mel_pre"]).max() < 1e-4
        assert np.abs(_np(out["mel_aft"]) - z["mel_aft"]).max() < 1e-4


def test_reference_lr_schedule(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "ref_init_seed0.json")))["lr_factor"]
    for step, want in ref.items():
        assert abs(O.learning_rate_factor(int(step)) - want) < 1e-12


def test_cached_oracle_vs_reference_over_640_frames(golden_dir, full_params):
    """The K/V-cached restatement against the REAL reference's uncached eval_batch over a long horizon
    (tests/golden/full_ar_long.npz: B=2, 640 dependent steps), plus the `resume` entry point used by the GPU tests."""
    cfg, params = full_params
    z = np.load(os.path.join(golden_dir, "full_ar_long.npz"))
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    batch = O.synth_batch(cfg, batch=2, text_len=64, n_frames=4, seed=9, ragged=True)
    T = int(z["max_frames"])
    with torch.no_grad():
        out = O.eval_batch_cached(p, cfg, batch, T)
    assert out["generated_lengths"].tolist() == z["generated_lengths"].tolist()
    e1 = np.abs(_np(out["mel_pre"]) - z["mel_pre"]).max()
    e2 = np.abs(_np(out["mel_aft"]) - z["mel_aft"]).max()
    print("cached oracle vs reference, 640 frames: mel_pre %.2e mel_aft %.2e" % (e1, e2))
    assert e1 < 1e-4 and e2 < 1e-4
    t0 = 600
    state = {"t": t0, "self_k": [k[:, :, :t0] for k in out["self_k"]], "self_v": [v[:, :, :t0] for v in out["self_v"]],
             "prev": out["mel_pre"][:, t0 - 1], "lengths": torch.full((2,), t0 + 1, dtype=torch.int32),
             "finished": torch.zeros(2, dtype=torch.bool)}
    with torch.no_grad():
        tail = O.eval_batch_cached(p, cfg, batch, T, resume=state, memory=out["memory"])
    assert torch.equal(tail["mel_pre"], out["mel_pre"][:, t0:])
