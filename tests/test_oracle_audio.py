"""The mel -> waveform oracle (oracle/audio_oracle.py) against independent implementations: librosa 0.6.0, whose algorithms it
restates, is not in this image, so the reference's utils/audio.py cannot run here (the oracle header says PARITY UNPINNED);
what can be pinned on the CPU is pinned here."""
import os
import sys

import numpy as np
from scipy import signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "few-shot-transformer-tts_b200")]
from oracle import audio_oracle as A  # noqa: E402


def test_stft_matches_scipy_on_the_same_frames():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(A.HOP * 49).astype(np.float32)
    E = A.stft(y)
    assert E.shape == (1025, 50) and E.dtype == np.complex64          # 1 + len(y) // hop frames (librosa.stft, center=True)
    w = A.padded_window()
    assert abs(w[:624]).max() == 0 and abs(w[1424:]).max() == 0 and w[624] == 0 and abs(w[624 + 400] - 1.0) < 1e-12   # periodic Hann, centred
    yp = np.pad(y, 1024, mode="reflect")
    _, _, Z = signal.stft(yp, window=w, nperseg=2048, noverlap=2048 - A.HOP, boundary=None, padded=False)
    assert np.abs(Z * w.sum() - E).max() < 1e-5 * np.abs(E).max()


def test_istft_inverts_stft_and_matches_scipy():
    rng = np.random.default_rng(1)
    y = rng.standard_normal(A.HOP * 30).astype(np.float32)
    E = A.stft(y).astype(np.complex128)
    back = A.istft(E)
    assert back.shape == y.shape and np.abs(back - y).max() < 1e-5     # hop 200 / Hann 800: window-sum-square normalised
    w = A.padded_window()
    _, z = signal.istft(E / w.sum(), window=w, nperseg=2048, noverlap=2048 - A.HOP, input_onesided=True, boundary=False)
    assert np.abs(z[1024:1024 + len(y)] - y).max() < 1e-5


def test_mel_basis_properties_and_product_constants():
    mb = A.mel_basis()
    assert mb.shape == (80, 1025) and mb.min() >= 0
    freqs = np.linspace(0, 8000, 1025)
    centres = (mb * freqs).sum(1) / mb.sum(1)
    assert np.all(np.diff(centres) > 0)                                  # ordered triangles
    # Slaney area normalisation: every filter integrates to ~1 over frequency (bin width 7.8125 Hz)
    area = mb.sum(1) * (8000.0 / 1024)
    assert np.abs(area - 1.0).max() < 0.08
    # linear below 1 kHz: equal spacing of the first centres, 200/3 Hz per mel
    pts = A._mel_to_hz(np.linspace(A._hz_to_mel(0.0), A._hz_to_mel(8000.0), 82))
    assert np.allclose(np.diff(pts[:10]), pts[1] - pts[0]) and abs(A._hz_to_mel(1000.0) - 15.0) < 1e-9
    # the product's host-side constants (tts_b200/vocoder.py) are the same filter bank
    from tts_b200 import vocoder as V
    assert np.abs(V.mel_filterbank(16000, 2048, 80) - mb).max() < 1e-12


def test_griffin_lim_reduces_the_spectral_error():
    rng = np.random.default_rng(2)
    mel = np.clip(rng.standard_normal((40, 80)) * 1.5, -4, 4).astype(np.float32)
    S = A.linear_from_mel(mel)

    def sc(wav):
        return np.linalg.norm(np.abs(A.stft(wav)) - S) / np.linalg.norm(S)

    e0, e20 = sc(A.griffin_lim(S, 0)), sc(A.griffin_lim(S, 20))
    assert e20 < e0
    wav = A.mel2wav(mel, n_iter=3)
    assert wav.dtype == np.float32 and wav.shape == (A.HOP * 39,)


def test_product_vocoder_refuses_the_cpu_and_never_imports_the_oracle():
    """No CPU fallback in the product path (the oracle above is the only CPU implementation, and only tests import it)."""
    import pytest
    import torch
    from tts_b200 import vocoder as V
    with pytest.raises(RuntimeError):
        V.mel2wav_batch(torch.zeros(1, 30, 80), [30], device="cpu")
    src = open(V.__file__).read()
    assert "oracle" not in src.replace("Oracle", "")


def test_mel_basis_matches_an_independent_librosa_compatible_implementation():
    """transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") is a separate implementation written to
    reproduce librosa.filters.mel; the periodic Hann window likewise."""
    import pytest
    au = pytest.importorskip("transformers.audio_utils")
    fb = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=80, min_frequency=0.0, max_frequency=8000.0,
                            sampling_rate=16000, norm="slaney", mel_scale="slaney")
    assert fb.shape == (1025, 80) and np.abs(fb.T - A.mel_basis()).max() < 1e-12
    w = au.window_function(800, "hann", periodic=True)
    assert np.abs(w - A.padded_window()[624:1424]).max() < 1e-12
