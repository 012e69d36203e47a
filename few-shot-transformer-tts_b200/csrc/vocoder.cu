// Mel -> waveform stage: utils/audio.py:53-99 (mel_to_linear, mel2wav, griffin_lim, invert_spectrogram)
// of the reference, which runs 60 Griffin-Lim iterations of librosa 0.6.0 STFT / ISTFT on the CPU, one utterance per
// process-pool worker (synthesize.py:82,99).  Here a whole batch of utterances stays on the GPU:
//   mel_to_mag_kernel   de-normalise, dB -> amplitude, pseudo-inverse mel basis, clamp, ^power       (audio.py:53-57,61-72)
//   gl_frame_kernel     one CTA per STFT frame: [reflect-padded, Hann-windowed frame -> 2048-point FFT -> phase of the
//                       estimate applied to the target magnitude ->] inverse FFT -> windowed frame   (audio.py:83-88,93-99)
//                       the complex spectrogram never leaves shared memory (the reference materialises it twice per iteration)
//   gl_ola_kernel       overlap-add of the windowed frames, normalised by the window's sum of squares, centre trimmed
//                       (librosa.istft, librosa.filters.window_sumsquare); a gather over <= 4 frames per sample: deterministic
//   deemphasis_kernel   scipy.signal.lfilter([1], [1, -preemphasis])                                 (audio.py:75-76)
// FFT: radix-2 decimation-in-time in shared memory (bit-reversed load, 11 stages of 1024 butterflies, twiddles from a
// host-computed fp64 -> fp32 table), 256 threads per frame.  fp32 throughout (the reference keeps the waveform in float32 and
// the estimate in complex64; its float64 magnitudes are rounded once here).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tts {
namespace voc {

constexpr int kFft = 2048, kBins = kFft / 2 + 1, kHop = 200, kWin = 800, kLo = (kFft - kWin) / 2, kHi = kLo + kWin;
constexpr int kThreads = 256;

__device__ __forceinline__ int brev11(int i) { return (int)(__brev((unsigned)i) >> 21); }

// buf: 2048 complex values in BIT-REVERSED order -> natural-order DFT (forward: e^{-2 pi i jk / N}; inverse: conjugate twiddles,
// unscaled).  tw[k] = e^{-2 pi i k / 2048}, k < 1024.  Ends with a CTA-wide sync.
template <bool INVERSE>
__device__ __forceinline__ void fft2048(float2* buf, const float2* tw) {
#pragma unroll 1
  for (int s = 0; s < 11; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int q = 0; q < kFft / 2 / kThreads; ++q) {
      const int j = threadIdx.x + q * kThreads;
      const int pos = j & (half - 1), i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
      float2 w = tw[pos << (10 - s)];
      if (INVERSE) w.y = -w.y;
      const float2 a = buf[i0], b = buf[i1];
      const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      buf[i0] = make_float2(a.x + t.x, a.y + t.y);
      buf[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}

// mag[b][t][k] = max(1e-10, sum_m inv_t[m][k] * amp[m]) ^ power,  amp = 10 ^ (0.05 * (clip((mel + max_abs) / (2 max_abs), 0, 1) * max_db - max_db + ref_db))
__global__ void __launch_bounds__(kThreads) mel_to_mag_kernel(const float* __restrict__ mel, const int32_t* __restrict__ len,
                                                              const float* __restrict__ inv_t, int frames_max, int n_mels,
                                                              float max_abs, float max_db, float ref_db, float power,
                                                              float* __restrict__ mag) {
  extern __shared__ float amp[];
  const int t = blockIdx.x, b = blockIdx.y;
  if (t >= len[b]) return;
  const float* m = mel + ((size_t)b * frames_max + t) * n_mels;
  for (int i = threadIdx.x; i < n_mels; i += kThreads) {
    const float x = fminf(fmaxf((m[i] + max_abs) / (2.f * max_abs), 0.f), 1.f) * max_db - max_db + ref_db;
    amp[i] = exp10f(0.05f * x);
  }
  __syncthreads();
  float* out = mag + ((size_t)b * frames_max + t) * kBins;
  for (int k = threadIdx.x; k < kBins; k += kThreads) {
    float acc = 0.f;
    for (int i = 0; i < n_mels; ++i) acc = fmaf(__ldg(inv_t + (size_t)i * kBins + k), amp[i], acc);
    out[k] = powf(fmaxf(acc, 1e-10f), power);
  }
}

// One Griffin-Lim half-iteration of one frame.  FIRST: X = magnitude (zero phase).  Otherwise X = magnitude * E / max(1e-8, |E|)
// with E the STFT frame of the current waveform estimate y.  Output: the Hann-windowed inverse transform of X (its 800
// samples under the window), to be overlap-added.
template <bool FIRST>
__global__ void __launch_bounds__(kThreads) gl_frame_kernel(const float* __restrict__ mag, const int32_t* __restrict__ len,
                                                            const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                            const float* __restrict__ y, long long ldy, int frames_max,
                                                            float* __restrict__ frames) {
  __shared__ float2 buf[kFft];
  __shared__ float2 tw[kFft / 2];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int T = len[b];
  if (t >= T) return;
  for (int i = tid; i < kFft / 2; i += kThreads) tw[i] = twiddle[i];
  const float* S = mag + ((size_t)b * frames_max + t) * kBins;
  constexpr int kPer = (kBins + kThreads - 1) / kThreads;   // 5: bins tid + 256 q (the last one only for tid == 0)
  float2 X[kPer];
  if (!FIRST) {
    const int L = kHop * (T - 1);
    const float* yb = y + (size_t)b * ldy;
    for (int i = tid; i < kFft; i += kThreads) {
      float v = 0.f;
      if (i >= kLo && i < kHi) {   // np.pad(y, n_fft / 2, mode='reflect'), frame t, times the centred window
        int idx = t * kHop + i - kFft / 2;
        if (idx < 0) idx = -idx;
        if (idx >= L) idx = 2 * (L - 1) - idx;
        v = yb[idx] * window[i - kLo];
      }
      buf[brev11(i)] = make_float2(v, 0.f);
    }
    __syncthreads();
    fft2048<false>(buf, tw);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int k = tid + q * kThreads;
      if (k < kBins) {
        const float2 e = buf[k];
        const float s = S[k] / fmaxf(1e-8f, sqrtf(e.x * e.x + e.y * e.y));   // audio.py:87-88
        X[q] = make_float2(e.x * s, e.y * s);
      }
    }
    __syncthreads();
  } else {
    __syncthreads();   // twiddles loaded
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int k = tid + q * kThreads;
      if (k < kBins) X[q] = make_float2(S[k], 0.f);
    }
  }
  // Hermitian extension (librosa.istft: spec | conj(spec[-2:0:-1])); the imaginary parts of DC / Nyquist drop out of .real
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int k = tid + q * kThreads;
    if (k < kBins) {
      if (k == 0 || k == kFft / 2) {
        buf[brev11(k)] = make_float2(X[q].x, 0.f);
      } else {
        buf[brev11(k)] = X[q];
        buf[brev11(kFft - k)] = make_float2(X[q].x, -X[q].y);
      }
    }
  }
  __syncthreads();
  fft2048<true>(buf, tw);
  float* out = frames + ((size_t)b * frames_max + t) * kWin;
  for (int i = tid; i < kWin; i += kThreads) out[i] = window[i] * buf[kLo + i].x * (1.f / kFft);
}

// y[b][n] = sum_t frames[b][t][n + 1024 - 200 t - 624] / sum_t window^2[...]   (n < 200 (T - 1): the centre-trimmed ISTFT)
__global__ void __launch_bounds__(256) gl_ola_kernel(const float* __restrict__ frames, const int32_t* __restrict__ len,
                                                     const float* __restrict__ window, int frames_max, float* __restrict__ y,
                                                     long long ldy) {
  const int b = blockIdx.y, T = len[b], L = kHop * (T - 1);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= L) return;
  const int p = n + kFft / 2;
  int t_hi = (p - kLo) / kHop;
  if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = p - (kHi - 1) <= 0 ? 0 : (p - (kHi - 1) + kHop - 1) / kHop;
  float acc = 0.f, wss = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int i = p - t * kHop - kLo;
    const float w = window[i];
    acc += frames[((size_t)b * frames_max + t) * kWin + i];
    wss = fmaf(w, w, wss);
  }
  y[(size_t)b * ldy + n] = wss > 1.17549435e-38f ? acc / wss : acc;
}

// wav[n] = y[n] + c * wav[n - 1]: a first-order recurrence, one thread per utterance (0.5 ms for 200 k samples)
__global__ void deemphasis_rows_kernel(const float* __restrict__ y, long long ldy, const int32_t* __restrict__ len, int batch,
                                       float c, float* __restrict__ wav, long long ldw) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int L = kHop * (len[b] - 1);
  const float* src = y + (size_t)b * ldy;
  float* dst = wav + (size_t)b * ldw;
  float prev = 0.f;
  for (int n = 0; n < L; ++n) {
    prev = fmaf(c, prev, src[n]);
    dst[n] = prev;
  }
}

}  // namespace voc
}  // namespace tts

using namespace tts;

extern "C" int tts_griffin_lim(const TtsGriffinLim* g, void* stream) {
  TTS_REQUIRE(g && g->mel && g->lengths && g->inv_basis_t && g->window && g->twiddle && g->mag && g->frames && g->y && g->wav,
              "griffin_lim: null argument");
  TTS_REQUIRE(g->n_fft == voc::kFft && g->hop_length == voc::kHop && g->win_length == voc::kWin,
              "griffin_lim: built for n_fft 2048, hop 200, win 800 (hyperparams.py:7-15), got %d / %d / %d", g->n_fft,
              g->hop_length, g->win_length);
  TTS_REQUIRE(g->batch > 0 && g->frames_max >= 7 && g->min_frames >= 7 && g->n_mels > 0 && g->n_mels <= 1024 && g->n_iter >= 0,
              "griffin_lim: needs >= 7 frames per utterance (one reflection of the 1024-sample pad) and n_iter >= 0");
  TTS_REQUIRE(g->ldy >= (long long)voc::kHop * (g->frames_max - 1) && g->ldw >= (long long)voc::kHop * (g->frames_max - 1),
              "griffin_lim: waveform rows too short");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 fg((unsigned)g->frames_max, (unsigned)g->batch);
  voc::mel_to_mag_kernel<<<fg, voc::kThreads, g->n_mels * sizeof(float), s>>>(g->mel, g->lengths, g->inv_basis_t, g->frames_max,
                                                                             g->n_mels, g->max_abs, g->max_db, g->ref_db, g->power,
                                                                             g->mag);
  TTS_CHECK_LAUNCH();
  const dim3 og((unsigned)ceil_div(voc::kHop * (g->frames_max - 1), 256), (unsigned)g->batch);
  const float2* tw = reinterpret_cast<const float2*>(g->twiddle);
  // audio.py:81-91: X = S; n_iter x { x = istft(X); E = stft(x); X = S * E / max(1e-8, |E|) }; x = istft(X)
  voc::gl_frame_kernel<true><<<fg, voc::kThreads, 0, s>>>(g->mag, g->lengths, g->window, tw, g->y, g->ldy, g->frames_max, g->frames);
  TTS_CHECK_LAUNCH();
  voc::gl_ola_kernel<<<og, 256, 0, s>>>(g->frames, g->lengths, g->window, g->frames_max, g->y, g->ldy);
  TTS_CHECK_LAUNCH();
  for (int it = 0; it < g->n_iter; ++it) {
    voc::gl_frame_kernel<false><<<fg, voc::kThreads, 0, s>>>(g->mag, g->lengths, g->window, tw, g->y, g->ldy, g->frames_max, g->frames);
    TTS_CHECK_LAUNCH();
    voc::gl_ola_kernel<<<og, 256, 0, s>>>(g->frames, g->lengths, g->window, g->frames_max, g->y, g->ldy);
    TTS_CHECK_LAUNCH();
  }
  voc::deemphasis_rows_kernel<<<ceil_div(g->batch, 32), 32, 0, s>>>(g->y, g->ldy, g->lengths, g->batch, g->preemphasis, g->wav, g->ldw);
  TTS_CHECK_LAUNCH();
  return 0;
}
