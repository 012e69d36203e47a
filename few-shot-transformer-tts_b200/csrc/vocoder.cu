// Mel -> waveform stage: utils/audio.py:53-99 (mel_to_linear, mel2wav, griffin_lim, invert_spectrogram)
// of the reference, which runs 60 Griffin-Lim iterations of librosa 0.6.0 STFT / ISTFT on the CPU, one utterance per
// process-pool worker (synthesize.py:82,99).  Here a whole batch of utterances stays on the GPU:
//   mel_to_mag_kernel   de-normalise, dB -> amplitude, pseudo-inverse mel basis, clamp, ^power       (audio.py:53-57,61-72)
//   gl_frame_kernel     one CTA per STFT frame: [reflect-padded, Hann-windowed frame -> 2048-point FFT -> phase of the
//                       estimate applied to the target magnitude ->] inverse FFT -> windowed frame   (audio.py:83-88,93-99)
//                       the complex spectrogram never leaves shared memory (the reference materialises it twice per iteration)
//   gl_ola_kernel       overlap-add of the windowed frames, normalised by the window's sum of squares, centre trimmed
//                       (librosa.istft, librosa.filters.window_sumsquare); a gather over <= 4 frames per sample: deterministic
//   deemphasis_kernel   scipy.signal.lfilter([1], [1, -preemphasis]) as a blocked scan of affine maps   (audio.py:75-76)
// FFT: each 2048-point real transform as a 1024-point complex Stockham radix-4 FFT in shared memory (five passes, one butterfly
// per thread and pass, 256 threads per frame, twiddles from a host-computed fp64 -> fp32 table) plus the split / merge step.  fp32 throughout (the reference keeps the waveform in float32 and
// the estimate in complex64; its float64 magnitudes are rounded once here).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tts {
namespace voc {

constexpr int kFft = 2048, kBins = kFft / 2 + 1, kHop = 200, kWin = 800, kLo = (kFft - kWin) / 2, kHi = kLo + kWin;
constexpr int kThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a conj(b)

// Shared-memory index swizzle of the two FFT buffers: bits 5:4 of the index are XORed into bits 1:0 and 3:2.  Sixteen
// consecutive indices still map to sixteen different 8-byte banks, and so do the strided stores of the first two Stockham
// passes (4 j + r and 16 (j / 4) + j % 4 + 4 r), which were 4-way bank conflicts.
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 3) * 5); }

// 1024-point complex DFT, Stockham autosort, radix 4: five passes, one butterfly per thread and pass (256 threads), ping-pong
// between two shared-memory buffers, no bit reversal.  Pass with sub-transform size Ns: thread j reads a[j + 256 r], multiplies by
// e^{-+ 2 pi i r k / (4 Ns)} (k = j mod Ns), radix-4 butterfly, writes b[4 (j - k) + k + r Ns].  The twiddles of a pass are
// stored per pass as [r - 1][k] (consecutive lanes read consecutive entries; indexing one 2048-entry circle was a 16-way bank
// conflict in the middle passes): ptw = {Ns = 4 | 16 | 64 | 256}, 3 Ns entries each.  INVERSE: conjugate twiddles, unscaled.
// Returns the buffer that holds the result (natural order, swizzled).
template <bool INVERSE>
__device__ __forceinline__ float2* fft1024(float2* a, float2* b, const float2* ptw) {
  const int j = threadIdx.x;
#pragma unroll
  for (int ls = 0; ls < 10; ls += 2) {
    const int Ns = 1 << ls, k = j & (Ns - 1);
    const float2 v0 = a[swz(j)];
    float2 v1 = a[swz(j + 256)], v2 = a[swz(j + 512)], v3 = a[swz(j + 768)];
    if (ls > 0) {
      const float2* t = ptw + (Ns - 4) + k;   // 3 (4 + 16 + ...) = Ns - 4 entries precede this pass
      v1 = INVERSE ? cmulc(v1, t[0]) : cmul(v1, t[0]);
      v2 = INVERSE ? cmulc(v2, t[Ns]) : cmul(v2, t[Ns]);
      v3 = INVERSE ? cmulc(v3, t[2 * Ns]) : cmul(v3, t[2 * Ns]);
    }
    const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y), t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
    const float2 t3 = INVERSE ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);   // (v1 - v3) * (+-i)
    const int o = ((j - k) << 2) + k;
    b[swz(o)] = make_float2(t0.x + t2.x, t0.y + t2.y);
    b[swz(o + Ns)] = make_float2(t1.x + t3.x, t1.y + t3.y);
    b[swz(o + 2 * Ns)] = make_float2(t0.x - t2.x, t0.y - t2.y);
    b[swz(o + 3 * Ns)] = make_float2(t1.x - t3.x, t1.y - t3.y);
    __syncthreads();
    float2* t = a; a = b; b = t;
  }
  return a;
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
// samples under the window), to be overlap-added.  Both 2048-point REAL transforms run as 1024-point complex ones
// (z[n] = x[2n] + i x[2n+1]) with the usual O(N) split / merge step; the spectrum lives in shared memory only.
template <bool FIRST>
__global__ void __launch_bounds__(kThreads) gl_frame_kernel(const float* __restrict__ mag, const int32_t* __restrict__ len,
                                                            const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                            const float* __restrict__ y, long long ldy, int frames_max,
                                                            float* __restrict__ frames) {
  constexpr int H = kFft / 2;   // 1024
  __shared__ float2 bufA[H];
  __shared__ float2 bufB[H];
  __shared__ float2 tw[H];        // e^{-2 pi i k / 2048}, k < 1024: the split / merge step
  __shared__ float2 ptw[H];       // per-pass twiddles of the 1024-point transform (1020 entries)
  __shared__ float x_nyq;       // X[1024] (real)
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int T = len[b];
  if (t >= T) return;
  for (int i = tid; i < H; i += kThreads) {
    tw[i] = twiddle[i];
    ptw[i] = twiddle[H + i];
  }
  const float* S = mag + ((size_t)b * frames_max + t) * kBins;
  float sk[H / kThreads];   // the target magnitudes of this thread's bins, requested before the forward transform hides them
#pragma unroll
  for (int q = 0; q < H / kThreads; ++q) sk[q] = S[tid + q * kThreads];
  const float s_nyq = tid == 0 ? S[H] : 0.f;
  float2* xp = bufB;   // X'[0..1023] = target magnitude with the estimate's phase
  if (!FIRST) {
    const int L = kHop * (T - 1);
    const float* yb = y + (size_t)b * ldy;
#pragma unroll
    for (int q = 0; q < H / kThreads; ++q) {
      const int n = tid + q * kThreads;
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 2 * n + e;
        if (i >= kLo && i < kHi) {   // np.pad(y, n_fft / 2, mode='reflect'), frame t, times the centred window
          int idx = t * kHop + i - kFft / 2;
          if (idx < 0) idx = -idx;
          if (idx >= L) idx = 2 * (L - 1) - idx;
          v[e] = yb[idx] * window[i - kLo];
        }
      }
      bufA[swz(n)] = make_float2(v[0], v[1]);
    }
    __syncthreads();
    float2* Z = fft1024<false>(bufA, bufB, ptw);   // five passes: the result is in bufB
    xp = Z == bufA ? bufB : bufA;
    // split: E = (Z[k] + conj Z[-k]) / 2 (even samples), O = (Z[k] - conj Z[-k]) / 2i (odd samples), X[k] = E + e^{-2 pi i k / N} O
#pragma unroll
    for (int q = 0; q < H / kThreads; ++q) {
      const int k = tid + q * kThreads;
      const float2 zk = Z[swz(k)], zc = Z[swz((H - k) & (H - 1))];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float2 to = cmul(tw[k], o);
      float2 X = make_float2(e.x + to.x, e.y + to.y);
      if (k == 0) {
        const float xn = zk.x - zk.y;   // X[1024] = E[0] - O[0] (real); its phase is its sign
        x_nyq = s_nyq * xn / fmaxf(1e-8f, fabsf(xn));
        X.y = 0.f;
      }
      const float sc = sk[q] / fmaxf(1e-8f, sqrtf(X.x * X.x + X.y * X.y));   // audio.py:87-88
      xp[swz(k)] = make_float2(X.x * sc, X.y * sc);
    }
  } else {
#pragma unroll
    for (int q = 0; q < H / kThreads; ++q) {
      const int k = tid + q * kThreads;
      xp[swz(k)] = make_float2(sk[q], 0.f);
    }
    if (tid == 0) x_nyq = s_nyq;
  }
  __syncthreads();
  // merge (Hermitian input; the imaginary parts of DC / Nyquist drop out of librosa's ifft(...).real):
  // E = (X[k] + conj X[N/2 - k]) / 2, O = (X[k] - conj X[N/2 - k]) / 2 * e^{+2 pi i k / N}, Z[k] = E + i O
  float2* zin = xp == bufA ? bufB : bufA;
#pragma unroll
  for (int q = 0; q < H / kThreads; ++q) {
    const int k = tid + q * kThreads;
    float2 xk = xp[swz(k)];
    float2 xm;
    if (k == 0) {
      xk.y = 0.f;
      xm = make_float2(x_nyq, 0.f);
    } else {
      const float2 r = xp[swz(H - k)];
      xm = make_float2(r.x, -r.y);
    }
    const float2 e = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y + xm.y));
    const float2 o = cmulc(make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y - xm.y)), tw[k]);
    zin[swz(k)] = make_float2(e.x - o.y, e.y + o.x);
  }
  __syncthreads();
  const float2* z = fft1024<true>(zin, xp, ptw);   // z[n] = 1024 (x[2n] + i x[2n+1])
  float* out = frames + ((size_t)b * frames_max + t) * kWin;
  for (int n = tid; n < kWin / 2; n += kThreads) {
    const float2 v = z[swz(kLo / 2 + n)];
    *reinterpret_cast<float2*>(out + 2 * n) = make_float2(window[2 * n] * v.x * (1.f / H), window[2 * n + 1] * v.y * (1.f / H));
  }
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

// wav[n] = y[n] + c * wav[n - 1] (scipy.signal.lfilter([1], [1, -c])): a first-order recurrence = a scan of affine maps
// s -> a s + b.  One CTA per utterance walks tiles of 2048 samples: every thread folds 8 consecutive samples into (c^8, b),
// a shuffle scan + eight warp totals give each thread the state it starts from, and it replays its 8 samples from there.
// (A single thread per utterance took 13 ms of the stage: 200 k dependent, uncoalesced steps.)
__global__ void __launch_bounds__(256) deemphasis_kernel(const float* __restrict__ y, long long ldy, const int32_t* __restrict__ len,
                                                         float c, float* __restrict__ wav, long long ldw) {
  __shared__ float sa[8], sb[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = kHop * (len[b] - 1);
  const float* src = y + (size_t)b * ldy;
  float* dst = wav + (size_t)b * ldw;
  const float c2 = c * c, c4 = c2 * c2, c8 = c4 * c4;
  float carry = 0.f;
  for (int base = 0; base < L; base += 2048) {
    const int n0 = base + 8 * tid;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = n0 + e < L ? src[n0 + e] : 0.f;
    float bb = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) bb = fmaf(c, bb, x[e]);
    float a = c8;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {   // inclusive scan: (a, bb) o (ap, bp) = (a ap, a bp + bb)
      const float ap = __shfl_up_sync(0xffffffffu, a, d), bp = __shfl_up_sync(0xffffffffu, bb, d);
      if (lane >= d) {
        bb = fmaf(a, bp, bb);
        a *= ap;
      }
    }
    if (lane == 31) {
      sa[warp] = a;
      sb[warp] = bb;
    }
    __syncthreads();
    float st = carry, tot = carry;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) st = fmaf(sa[w], st, sb[w]);
      tot = fmaf(sa[w], tot, sb[w]);
    }
    const float ae = __shfl_up_sync(0xffffffffu, a, 1), be = __shfl_up_sync(0xffffffffu, bb, 1);
    float prev = lane == 0 ? st : fmaf(ae, st, be);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      prev = fmaf(c, prev, x[e]);
      if (n0 + e < L) dst[n0 + e] = prev;
    }
    __syncthreads();   // sa / sb are rewritten by the next tile
    carry = tot;
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
  voc::deemphasis_kernel<<<g->batch, 256, 0, s>>>(g->y, g->ldy, g->lengths, g->preemphasis, g->wav, g->ldw);
  TTS_CHECK_LAUNCH();
  return 0;
}
