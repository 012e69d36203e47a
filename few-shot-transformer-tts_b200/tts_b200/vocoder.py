"""Mel -> waveform on the GPU: drop-in for ``utils/audio.py:60-79`` ``mel2wav`` of the reference, which
runs 60 Griffin-Lim iterations of librosa 0.6.0 STFT / ISTFT per utterance on the CPU (``synthesize.py:82,99``: a 4-worker
process pool).  ``mel2wav(mel)`` keeps the reference signature (one ``[T, 80]`` normalised mel -> float32 waveform of
``hop_length * (T - 1)`` samples); ``mel2wav_batch`` converts a whole synthesised batch in one call of
``tts_griffin_lim`` (csrc/vocoder.cu).  The constants the reference takes from librosa (``filters.mel``, its
pseudo-inverse, the periodic Hann window) are computed once on the host in float64 from their published definitions.
There is no CPU fallback."""
import ctypes as C
import math

import numpy as np
import torch

from . import _native as N


class AudioParams:
    """hyperparams.py:4-18 of the reference."""
    sr, n_fft, hop_length, win_length, num_mels = 16000, 2048, 200, 800, 80
    max_db, ref_db, preemphasis, max_abs_value, n_iter, power = 100.0, 20.0, 0.97, 4.0, 60, 1.5


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels):
    """The Slaney-scale, area-normalised triangular filter bank ``librosa.filters.mel(sr, n_fft, n_mels)`` builds with its
    0.6.0 defaults (fmin 0, fmax sr / 2, htk False, norm 1): ``utils/audio.py:12-15``."""
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    width = np.diff(pts)
    ramps = pts[:, None] - freqs[None, :]
    w = np.maximum(0.0, np.minimum(-ramps[:-2] / width[:-1, None], ramps[2:] / width[1:, None]))
    return w * (2.0 / (pts[2:] - pts[:-2]))[:, None]


class GriffinLim:
    """Device-resident constants + scratch of ``tts_griffin_lim`` for one device; scratch grows with the largest batch seen."""

    def __init__(self, device, hp=AudioParams):
        self.device, self.hp = torch.device(device), hp
        N.load()
        basis = mel_filterbank(hp.sr, hp.n_fft, hp.num_mels)
        inv = np.linalg.pinv(basis)                                    # [bins][mels]  (audio.py:53-56)
        self.inv_t = torch.from_numpy(np.ascontiguousarray(inv.T).astype(np.float32)).to(self.device)
        n = np.arange(hp.win_length, dtype=np.float64)
        self.window = torch.from_numpy((0.5 - 0.5 * np.cos(2.0 * np.pi * n / hp.win_length)).astype(np.float32)).to(self.device)
        # twiddle table [n_fft]: e^{-2 pi i k / n_fft} for k < n_fft / 2 (the real-transform split / merge), then the per-pass
        # twiddles of the 1024-point radix-4 Stockham transform: for Ns in (4, 16, 64, 256): for r in (1, 2, 3): e^{-2 pi i r k / (4 Ns)}
        half = hp.n_fft // 2
        k = np.arange(half, dtype=np.float64)
        ang = [2.0 * np.pi * k / hp.n_fft]
        for ns in (4, 16, 64, 256):
            kk = np.arange(ns, dtype=np.float64)
            ang += [2.0 * np.pi * r * kk / (4 * ns) for r in (1, 2, 3)]
        ang = np.concatenate(ang)
        ang = np.concatenate([ang, np.zeros(hp.n_fft - ang.size)])
        tw = np.stack([np.cos(ang), -np.sin(ang)], axis=1)
        self.twiddle = torch.from_numpy(tw.astype(np.float32)).to(self.device)
        self._scratch = None

    def __call__(self, mels, lengths, n_iter=None):
        """mels [B, T, 80] fp32 on the device (normalised, +-max_abs), lengths [B] frames -> (wav [B, hop (T - 1)] fp32 on the
        device, zero past each utterance's hop (lengths[b] - 1) samples; sample counts [B])."""
        hp = self.hp
        mels = mels.to(self.device, torch.float32).contiguous()
        B, T, M = mels.shape
        lens_host = [int(v) for v in lengths]
        if min(lens_host) < 7 or max(lens_host) > T:
            raise ValueError("mel2wav needs 7 <= frames <= %d per utterance, got %s" % (T, lens_host))
        lens = torch.tensor(lens_host, dtype=torch.int32, device=self.device)
        L = hp.hop_length * (T - 1)
        bins = hp.n_fft // 2 + 1
        need = B * T * (bins + hp.win_length) + B * L
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, device=self.device, dtype=torch.float32)
        mag = self._scratch[:B * T * bins]
        frames = self._scratch[B * T * bins:B * T * (bins + hp.win_length)]
        y = self._scratch[B * T * (bins + hp.win_length):need]
        wav = torch.zeros((B, L), device=self.device, dtype=torch.float32)
        g = N.GriffinLim()
        g.mel, g.lengths, g.inv_basis_t = mels.data_ptr(), lens.data_ptr(), self.inv_t.data_ptr()
        g.window, g.twiddle = self.window.data_ptr(), self.twiddle.data_ptr()
        g.batch, g.frames_max, g.min_frames, g.n_mels = B, T, min(lens_host), M
        g.n_fft, g.hop_length, g.win_length = hp.n_fft, hp.hop_length, hp.win_length
        g.n_iter = hp.n_iter if n_iter is None else int(n_iter)
        g.max_abs, g.max_db, g.ref_db, g.power, g.preemphasis = hp.max_abs_value, hp.max_db, hp.ref_db, hp.power, hp.preemphasis
        g.mag, g.frames, g.y, g.ldy, g.wav, g.ldw = mag.data_ptr(), frames.data_ptr(), y.data_ptr(), L, wav.data_ptr(), L
        with torch.cuda.device(self.device):
            N.check(N.load().tts_griffin_lim(C.byref(g), N.stream_ptr(self.device)), "griffin_lim")
        return wav, [hp.hop_length * (n - 1) for n in lens_host]


_ENGINES = {}


def _engine(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("tts_b200.vocoder runs on a CUDA device only (no CPU fallback)")
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _ENGINES:
        _ENGINES[key] = GriffinLim(torch.device("cuda", key[1]))
    return _ENGINES[key]


def mel2wav_batch(mels, lengths, device="cuda:0", n_iter=None):
    """A batch of normalised mels [B, T, 80] (torch tensor or array) -> list of float32 numpy waveforms (audio.py:60-79 each)."""
    mels = torch.as_tensor(np.asarray(mels) if not torch.is_tensor(mels) else mels)
    eng = _engine(mels.device if mels.is_cuda else device)
    wav, counts = eng(mels, lengths, n_iter)
    host = wav.cpu().numpy()
    return [host[b, :counts[b]].copy() for b in range(len(counts))]


def mel2wav(mel, device="cuda:0"):
    """Reference signature (utils/audio.py:60): one normalised mel [T, 80] -> float32 waveform."""
    mel = np.asarray(mel.detach().cpu() if torch.is_tensor(mel) else mel, dtype=np.float32)
    return mel2wav_batch(mel[None], [mel.shape[0]], device)[0]
