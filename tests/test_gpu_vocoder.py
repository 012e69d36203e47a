"""Mel -> waveform stage (csrc/vocoder.cu) against the numpy oracle of utils/audio.py:53-99."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "few-shot-transformer-tts_b200")]
from oracle import audio_oracle as A  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    return True


def _mels(seed, lens):
    rng = np.random.default_rng(seed)
    T = max(lens)
    m = np.zeros((len(lens), T, 80), dtype=np.float32)
    for b, n in enumerate(lens):
        # smooth-ish spectra in the model's range (+-4) so that the magnitudes span several decades
        base = rng.standard_normal((n, 80)) * 0.6 + np.linspace(1.5, -2.5, 80)[None, :] + 1.5 * np.sin(np.arange(n) / 5.0)[:, None]
        m[b, :n] = np.clip(base, -4, 4)
    return m


@pytest.mark.parametrize("n_iter,tol", [(0, 2e-4), (1, 1e-3), (3, 3e-3)])
def test_griffin_lim_matches_the_oracle_iteration_by_iteration(built, n_iter, tol):
    """Ragged batch; 0 iterations = mel_to_linear + ISTFT + overlap-add + de-emphasis alone, then one and three full
    STFT -> phase -> ISTFT iterations.  Tolerance relative to the utterance's peak (fp32 here, float64 magnitudes in the oracle)."""
    from tts_b200 import vocoder as V
    lens = [37, 60, 7, 52]
    mels = _mels(3, lens)
    got = V.mel2wav_batch(torch.from_numpy(mels), lens, DEV, n_iter=n_iter)
    for b, n in enumerate(lens):
        want = A.mel2wav(mels[b, :n], n_iter=n_iter)
        assert got[b].shape == want.shape == (A.HOP * (n - 1),) and got[b].dtype == np.float32
        assert np.abs(got[b] - want).max() < tol * np.abs(want).max(), (b, n, np.abs(got[b] - want).max(), np.abs(want).max())


def test_sixty_iterations_converge_like_the_oracle(built):
    """The reference setting (hyperparams.py:17 n_iter=60).  Phase retrieval amplifies rounding differences, so the waveforms
    are compared through what Griffin-Lim optimises: the spectral convergence ||S - |STFT(wav)||| / ||S|| of the result."""
    from tts_b200 import vocoder as V
    lens = [64, 41]
    mels = _mels(5, lens)
    got = V.mel2wav_batch(torch.from_numpy(mels), lens, DEV)
    for b, n in enumerate(lens):
        S = A.linear_from_mel(mels[b, :n])
        want = A.mel2wav(mels[b, :n])

        def sc(wav):   # undo the de-emphasis (audio.py:75-76) to get back the Griffin-Lim signal
            x = np.append(wav[0], wav[1:] - A.PREEMPHASIS * wav[:-1])
            return np.linalg.norm(np.abs(A.stft(x)) - S) / np.linalg.norm(S)

        e_got, e_want = sc(got[b]), sc(want)
        rel = np.abs(got[b] - want).max() / np.abs(want).max()
        print("vocoder b=%d frames=%d: spectral convergence ours %.4f oracle %.4f, max |wav diff| / peak %.3e" % (b, n, e_got, e_want, rel))
        assert e_got < 1.03 * e_want + 1e-3
        assert np.isfinite(got[b]).all()


def test_reference_signature_and_errors(built):
    from tts_b200 import vocoder as V
    mel = _mels(9, [30])[0]
    wav = V.mel2wav(mel)                                  # utils/audio.py:60: one [T, 80] mel -> float32 waveform
    assert wav.dtype == np.float32 and wav.shape == (A.HOP * 29,)
    with pytest.raises(ValueError):
        V.mel2wav(mel[:5])                                # fewer frames than one reflection of the STFT pad needs
    with pytest.raises(RuntimeError):
        V.mel2wav_batch(mel[None], [30], device="cpu")    # no CPU fallback


def test_generate_returns_waveforms(built, tiny_params):
    """synthesize.py:56,82 end to end on the device: generate(waveform=True) = eval_batch + mel2wav of every utterance."""
    from oracle import tts_oracle as O
    from tts_b200.engine import TtsEngine
    cfg, params = tiny_params
    eng = TtsEngine.from_state_dict(dict(params), cfg, DEV)
    batch = O.synth_batch(cfg, batch=3, text_len=12, n_frames=4, seed=3)
    out = eng.generate(batch, max_frames=24, record_align="none", waveform=True)
    lens = [int(v) for v in out["generated_lengths"].tolist()]
    assert out["wav"].shape[0] == 3 and out["wav_lengths"] == [200 * max(n - 1, 0) for n in lens]
    mel = out["mel_aft"].cpu().numpy()
    for b, n in enumerate(lens):
        if n >= 7:
            want = A.mel2wav(mel[b, :n])
            got = out["wav"][b, :out["wav_lengths"][b]].cpu().numpy()
            assert np.abs(got - want).max() < 2e-3 * max(np.abs(want).max(), 1e-6)
