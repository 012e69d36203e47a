"""CPU oracle for the mel -> waveform stage.  TEST INFRASTRUCTURE ONLY.

numpy restatement of ``utils/audio.py:53-99`` of mutiann/few-shot-transformer-tts (``mel_to_linear``, ``mel2wav``,
``griffin_lim``, ``invert_spectrogram``) and of the third-party functions it calls.  The algorithm lives in a
dependency that is ABSENT from this image: **librosa==0.6.0** (reference ``requirements.txt``; ``librosa.stft``,
``librosa.istft``, ``librosa.filters.mel``, ``librosa.filters.window_sumsquare``, ``librosa.util.pad_center``) plus
``scipy.signal.lfilter`` (present).  Their published algorithms are restated below, each function naming the librosa
0.6.0 routine it follows.

PARITY UNPINNED against the reference itself: without librosa the reference's ``utils/audio.py`` cannot be imported
here, and the reference ships no audio golden vectors.  What IS pinned (``tests/test_oracle_audio.py``): ``stft`` /
``istft`` against ``scipy.signal.stft`` / ``istft`` (independent implementation, same frames after rescaling), the
mel filter bank against ``transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")`` (a separate
implementation written to reproduce ``librosa.filters.mel``; equal to 1e-16) and its published properties, the periodic
Hann window likewise; ``lfilter`` is scipy's own.  Only ``tests/`` may import this file; the product path (``few-shot-transformer-tts_b200/``) never does.
"""
from __future__ import annotations

import numpy as np
from scipy import signal

# hyperparams.py:4-18
SR, N_FFT, HOP, WIN, NUM_MELS = 16000, 2048, 200, 800, 80
MAX_DB, REF_DB, PREEMPHASIS, MAX_ABS, N_ITER, POWER = 100, 20, 0.97, 4.0, 60, 1.5


def _hz_to_mel(f):
    """librosa.core.time_frequency.hz_to_mel(htk=False): Slaney's Auditory Toolbox scale (linear below 1 kHz, log above)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    """librosa.core.time_frequency.mel_to_hz(htk=False)."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr=SR, n_fft=N_FFT, n_mels=NUM_MELS):
    """librosa.filters.mel(sr, n_fft, n_mels) with the 0.6.0 defaults fmin=0, fmax=sr/2, htk=False, norm=1 (audio.py:12-15):
    triangles between neighbouring mel points over the FFT bin frequencies, each scaled by 2 / (its band width in Hz)."""
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return weights * enorm[:, None]


def padded_window(win_length=WIN, n_fft=N_FFT):
    """get_window('hann', win_length, fftbins=True) centred in n_fft zeros (librosa.util.pad_center)."""
    w = signal.get_window("hann", win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad), mode="constant")


def stft(y, n_fft=N_FFT, hop=HOP, win_length=WIN):
    """librosa.stft (0.6.0: center=True, pad_mode='reflect', dtype=complex64): [1 + n_fft/2, 1 + len(y)//hop]."""
    w = padded_window(win_length, n_fft)
    yp = np.pad(np.asarray(y), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx] * w[:, None]
    return np.fft.fft(frames, axis=0)[:1 + n_fft // 2].astype(np.complex64)


def window_sumsquare(n_frames, hop=HOP, win_length=WIN, n_fft=N_FFT):
    """librosa.filters.window_sumsquare(norm=None)."""
    n = n_fft + hop * (n_frames - 1)
    x = np.zeros(n, dtype=np.float32)
    win_sq = padded_window(win_length, n_fft) ** 2
    for i in range(n_frames):
        s = i * hop
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def istft(spec, hop=HOP, win_length=WIN):
    """librosa.istft (0.6.0: window='hann', center=True, dtype=float32, length=None) = audio.py:93-99."""
    n_fft = 2 * (spec.shape[0] - 1)
    w = padded_window(win_length, n_fft)
    n_frames = spec.shape[1]
    y = np.zeros(n_fft + hop * (n_frames - 1), dtype=np.float32)
    for i in range(n_frames):
        s = spec[:, i].flatten()
        s = np.concatenate((s, s[-2:0:-1].conj()), 0)
        y[i * hop:i * hop + n_fft] += (w * np.fft.ifft(s).real).astype(np.float32)
    wss = window_sumsquare(n_frames, hop, win_length, n_fft)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def griffin_lim(spectrogram, n_iter=N_ITER):
    """audio.py:81-91."""
    x_best = spectrogram.copy()
    for _ in range(n_iter):
        x_t = istft(x_best)
        est = stft(x_t)
        phase = est / np.maximum(1e-8, np.abs(est))
        x_best = spectrogram * phase
    return np.real(istft(x_best))


def mel_to_linear(mel, inv_basis=None):
    """audio.py:53-57."""
    if inv_basis is None:
        inv_basis = np.linalg.pinv(mel_basis())
    return np.maximum(1e-10, np.dot(inv_basis, mel))


def linear_from_mel(mel):
    """The magnitude Griffin-Lim is given: audio.py:61-72 up to `mel**hp.power`.  mel: [T, 80] normalised."""
    m = (np.asarray(mel, dtype=np.float64).T + MAX_ABS) / (2 * MAX_ABS)
    m = (np.clip(m, 0, 1) * MAX_DB) - MAX_DB + REF_DB
    m = np.power(10.0, m * 0.05)
    return mel_to_linear(m) ** POWER


def mel2wav(mel, n_iter=N_ITER):
    """audio.py:60-79: mel [T, 80] (symmetric normalisation, +-4) -> waveform float32 of hop * (T - 1) samples."""
    wav = griffin_lim(linear_from_mel(mel), n_iter)
    wav = signal.lfilter([1], [1, -PREEMPHASIS], wav)
    return wav.astype(np.float32)
