"""Input staging (reference: utils/__init__.py:3 dict_send_to, dataloader.py:419-439,498-508)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-transformer-tts_b200"))

from tts_b200 import staging  # noqa: E402


def _host_batch(rng, B, S, T):
    lens = rng.integers(S // 2, S + 1, size=B).astype(np.int32)
    tl = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    lens[0], tl[0] = S, T
    inputs = rng.integers(3, 256, size=(B, S)).astype(np.int64)
    mel = rng.standard_normal((B, T, 80)).astype(np.float32)
    for b in range(B):
        inputs[b, lens[b]:] = 0
        mel[b, tl[b]:] = 0
    return {"inputs": inputs, "input_lengths": lens, "mel_targets": mel, "target_lengths": tl,
            "input_spk_ids": (np.arange(B) % 572).astype(np.float32), "input_language_ids": (np.arange(B) % 38).astype(np.int64),
            "names": ["utt%d" % b for b in range(B)]}


def test_dict_send_to_cpu_semantics_match_the_reference():
    """Every direction other than host -> CUDA is the reference's code path: same keys, dtypes, values, non-tensors kept."""
    data = {"a": torch.arange(6).view(2, 3), "b": torch.ones(3, requires_grad=True) * 2, "names": ["x", "y"]}
    out = staging.dict_send_to(data, torch.device("cpu"), detach=True, as_numpy=True)
    assert set(out) == set(data) and out["names"] == ["x", "y"]
    assert isinstance(out["a"], np.ndarray) and out["a"].dtype == np.int64 and out["a"].tolist() == [[0, 1, 2], [3, 4, 5]]
    assert isinstance(out["b"], np.ndarray) and out["b"].tolist() == [2.0, 2.0, 2.0]
    out = staging.dict_send_to(data, "cpu")
    assert out["a"] is data["a"] or torch.equal(out["a"], data["a"])


def test_stager_refuses_cpu():
    with pytest.raises(RuntimeError):
        staging.BatchStager(torch.device("cpu"))


@pytest.mark.gpu
def test_stager_ragged_batches_exact_shapes_and_values():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    st = staging.BatchStager(dev, depth=2, n_languages=100)
    shapes = [(16, 258, 800), (16, 40, 240), (9, 131, 517), (16, 258, 800), (3, 32, 250)]
    pending = []
    for B, S, T in shapes:   # more batches than slots: arenas are reused, shapes differ every time
        hb = _host_batch(rng, B, S, T)
        sb = st.stage(hb)
        pending.append((hb, sb))
        if len(pending) == 2:
            hb0, sb0 = pending.pop(0)
            sb0.wait()
            assert sb0["names"] == hb0["names"]
            for k in ("inputs", "input_lengths", "mel_targets", "target_lengths", "input_spk_ids"):
                want = torch.from_numpy(hb0[k])
                got = sb0[k]
                assert got.is_cuda and got.is_contiguous() and tuple(got.shape) == tuple(want.shape), k
                assert got.dtype == staging._PROTO[k], (k, got.dtype)     # dataloader.get_input_proto
                assert torch.equal(got.cpu(), want.to(got.dtype)), k
            vec = sb0["input_language_vecs"]
            assert vec.shape == (hb0["inputs"].shape[0], 100) and vec.dtype == torch.float32
            assert torch.equal(vec.argmax(1).cpu(), torch.from_numpy(hb0["input_language_ids"]))
            assert float(vec.sum()) == hb0["inputs"].shape[0]
            sb0.release()
    assert st.h2d_bytes > 0


@pytest.mark.gpu
def test_dict_send_to_cuda_matches_plain_to():
    dev = torch.device("cuda:0")
    data = {"x": torch.randn(5, 7), "ids": torch.arange(12, dtype=torch.int32), "names": ["a"], "on_dev": torch.ones(2, device=dev)}
    out = staging.dict_send_to(data, dev)
    assert set(out) == set(data) and out["names"] == ["a"]
    for k in ("x", "ids"):
        assert out[k].is_cuda and out[k].dtype == data[k].dtype and torch.equal(out[k].cpu(), data[k])
    assert out["on_dev"].is_cuda
    again = staging.dict_send_to({"x": torch.zeros(5, 7)}, dev)     # the first result owns its memory
    assert torch.equal(out["x"].cpu(), data["x"]) and float(again["x"].abs().sum()) == 0.0
