// C-ABI entry points of libtts_b200.so that are not pure kernels: diagnostics, the decode
// session calls and the CUDA-graph cache that replays one decode step per launch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace tts {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_last_impl{0};   // implementation the last tts_decode_steps call ran (tts_decode_profile reads its stamps)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// decode.cu
int enqueue_step_phases(const TtsDecoderWeights* w, const TtsDecodeState* st, int update_state, cudaStream_t s);
int decode_reset(const TtsDecodeState* st, cudaStream_t s);
size_t decode_scratch_floats(const TtsDecoderWeights* w, int B);
// pipelined2.cu
int launch_pipelined2_steps(const TtsDecoderWeights* w, const TtsDecodeState* st, int n_steps, int update_state,
                            cudaStream_t s);
size_t pipelined2_scratch_floats(const TtsDecoderWeights* w, int B);
bool pipelined2_supported(const TtsDecoderWeights* w, const TtsDecodeState* st);
int pipelined2_profile(const TtsDecoderWeights* w, const TtsDecodeState* st, long long* out_host, int max_entries);
// pipelined.cu
int launch_pipelined_steps(const TtsDecoderWeights* w, const TtsDecodeState* st, int n_steps, int update_state,
                           cudaStream_t s);
size_t pipelined_scratch_floats(const TtsDecoderWeights* w, int B);
bool pipelined_supported(const TtsDecoderWeights* w, const TtsDecodeState* st);
int pipelined_profile(const TtsDecoderWeights* w, const TtsDecodeState* st, long long* out_host, int max_entries);

// ---- CUDA graph cache: one captured step per (weights, state, flags) -------------------------
struct GraphEntry {
  TtsDecoderWeights w;
  TtsDecodeState st;
  int update_state;
  cudaGraphExec_t exec;
  int kernels;
  unsigned long long stamp;
};
static std::mutex g_graph_mu;
static std::vector<GraphEntry> g_graphs;
static unsigned long long g_stamp = 0;

static int get_step_graph(const TtsDecoderWeights* w, const TtsDecodeState* st, int update_state, cudaStream_t s,
                          cudaGraphExec_t* out, int* kernels) {
  std::lock_guard<std::mutex> lock(g_graph_mu);
  for (auto& e : g_graphs)
    if (e.update_state == update_state && memcmp(&e.w, w, sizeof(*w)) == 0 && memcmp(&e.st, st, sizeof(*st)) == 0) {
      e.stamp = ++g_stamp;
      *out = e.exec;
      *kernels = e.kernels;
      return 0;
    }
  // capture on a private stream: the caller's stream may be the legacy default stream, which
  // cannot be captured; the instantiated graph is then launched on the caller's stream
  static cudaStream_t cap = nullptr;
  if (cap == nullptr) TTS_CHECK_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  (void)s;
  const long long before = g_launches.load();
  TTS_CHECK_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_step_phases(w, st, update_state, cap);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
  const int n_kernels = (int)(g_launches.load() - before);
  g_launches.store(before);  // captured launches are counted when replayed
  if (rc != 0) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  TTS_CHECK_CUDA(ce);
  cudaGraphExec_t exec = nullptr;
  TTS_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  if (g_graphs.size() >= 8) {  // evict the least recently used
    size_t victim = 0;
    for (size_t i = 1; i < g_graphs.size(); ++i)
      if (g_graphs[i].stamp < g_graphs[victim].stamp) victim = i;
    cudaGraphExecDestroy(g_graphs[victim].exec);
    g_graphs.erase(g_graphs.begin() + victim);
  }
  GraphEntry e;
  memcpy(&e.w, w, sizeof(*w));
  memcpy(&e.st, st, sizeof(*st));
  e.update_state = update_state;
  e.exec = exec;
  e.kernels = n_kernels;
  e.stamp = ++g_stamp;
  g_graphs.push_back(e);
  *out = exec;
  *kernels = n_kernels;
  return 0;
}

static int check_decode_args(const TtsDecoderWeights* w, const TtsDecodeState* st) {
  TTS_REQUIRE(w && st, "decode: null weights/state");
  TTS_REQUIRE(w->n_layers > 0 && w->n_layers <= TTS_MAX_LAYERS, "decode: n_layers=%d out of range", w->n_layers);
  TTS_REQUIRE(w->d_model % w->n_heads == 0, "decode: d_model %d not divisible by heads %d", w->d_model, w->n_heads);
  const int dh = w->d_model / w->n_heads;
  TTS_REQUIRE(dh == 32 || dh == 64 || dh == 96, "decode: head_dim %d not in {32,64,96}", dh);
  TTS_REQUIRE(w->d_model % 16 == 0 && w->d_ffn % 16 == 0 && w->prenet_hidden % 16 == 0 && w->n_mels % 16 == 0,
              "decode: widths must be multiples of 16 (D=%d F=%d P=%d M=%d)", w->d_model, w->d_ffn, w->prenet_hidden,
              w->n_mels);
  TTS_REQUIRE(w->d_model <= 768, "decode: d_model %d > 768 not supported by the LayerNorm prologue", w->d_model);
  TTS_REQUIRE(st->batch > 0 && st->mem_len > 0 && st->t_max > 0, "decode: empty state");
  TTS_REQUIRE(st->scratch && st->lengths && st->finished && st->frames && st->stop_logits && st->step_counter &&
                  st->n_unfinished && st->self_k && st->self_v && st->cross_k && st->cross_v,
              "decode: state has null buffers");
  return 0;
}

}  // namespace tts

using namespace tts;

extern "C" int tts_abi_version(void) { return TTS_ABI_VERSION; }
extern "C" const char* tts_last_error(void) { return g_error; }
extern "C" int64_t tts_launch_count(void) { return g_launches.load(); }
extern "C" void tts_launch_count_reset(void) { g_launches.store(0); }

extern "C" size_t tts_decode_scratch_bytes(const TtsDecoderWeights* w, int32_t batch, int32_t mem_len,
                                           int32_t t_max) {
  (void)mem_len;
  (void)t_max;
  if (!w || batch <= 0) return 0;
  const size_t a = decode_scratch_floats(w, batch), c = pipelined_scratch_floats(w, batch),
               e = pipelined2_scratch_floats(w, batch);
  const size_t m = a > c ? (a > e ? a : e) : (c > e ? c : e);
  return m * sizeof(float) + 256;
}

extern "C" int tts_decode_begin(const TtsDecoderWeights* w, const TtsDecodeState* st, void* stream) {
  int rc = check_decode_args(w, st);
  if (rc) return rc;
  TTS_REQUIRE(st->memory && st->input_lengths, "decode_begin: memory/input_lengths are null");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int D = w->d_model, H = w->n_heads, dh = D / H, B = st->batch, S = st->mem_len;
  for (int l = 0; l < w->n_layers; ++l) {  // attention.py:66-68, once per utterance instead of once per step
    TtsGemmEpilogue epi;
    memset(&epi, 0, sizeof(epi));
    epi.alpha = 1.f;
    epi.rows_per_batch = S;
    epi.head_dim = dh;
    epi.n_heads = H;
    epi.head_rows = S;
    const size_t off = (size_t)l * B * H * S * dh;
    epi.out_v = st->cross_v + off;
    rc = tts_gemm_nt(st->memory, D, w->layer[l].w_cross_kv, D, st->cross_k + off, 0, B * S, 2 * D, D, &epi, stream);
    if (rc) return rc;
  }
  return decode_reset(st, s);
}

extern "C" int tts_decode_profile(const TtsDecoderWeights* w, const TtsDecodeState* st, int64_t* out_host,
                                  int32_t max_entries) {
  TTS_REQUIRE(w && st && out_host && max_entries > 0, "decode_profile: bad arguments");
  TTS_REQUIRE(g_last_impl == 4 || g_last_impl == 5, "decode_profile: only the pipelined kernels (impl 4 / 5) record phase stamps");
  if (g_last_impl == 5) return pipelined2_profile(w, st, reinterpret_cast<long long*>(out_host), max_entries);
  return pipelined_profile(w, st, reinterpret_cast<long long*>(out_host), max_entries);
}

extern "C" int tts_decode_steps(const TtsDecoderWeights* w, const TtsDecodeState* st, int32_t n_steps,
                                const float* prev_mel, int64_t prev_mel_stride, int32_t update_state,
                                int32_t impl, void* stream) {
  int rc = check_decode_args(w, st);
  if (rc) return rc;
  TTS_REQUIRE(n_steps >= 0, "decode_steps: n_steps=%d", n_steps);
  if (prev_mel != nullptr) {   // frame t-1 supplied by the caller: place it where every implementation reads it from
    int t_host = 0;
    TTS_CHECK_CUDA(cudaMemcpyAsync(&t_host, st->step_counter, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    TTS_CHECK_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (t_host > 0 && t_host <= st->t_max)
      TTS_CHECK_CUDA(cudaMemcpy2DAsync(st->frames + (size_t)(t_host - 1) * w->n_mels, (size_t)st->t_max * w->n_mels * sizeof(float),
                                       prev_mel, (size_t)prev_mel_stride * sizeof(float), (size_t)w->n_mels * sizeof(float),
                                       (size_t)st->batch, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 0) {   // default: the pipelined kernel; TTS_DECODE_V2=1 opts into the two-CTAs-per-SM experiment (measured slower)
    static const bool v2 = getenv("TTS_DECODE_V2") != nullptr && atoi(getenv("TTS_DECODE_V2")) != 0;
    impl = (v2 && pipelined2_supported(w, st)) ? 5 : (pipelined_supported(w, st) ? 4 : 2);
  }
  g_last_impl = impl;
  TTS_REQUIRE(st->drop_p_prenet >= 0.f && st->drop_p_prenet < 1.f && st->drop_p_transformer >= 0.f && st->drop_p_transformer < 1.f,
              "decode_steps: dropout rates must be in [0, 1)");
  TTS_REQUIRE(impl == 4 || (st->drop_p_prenet == 0.f && st->drop_p_transformer == 0.f),
              "decode_steps: train()-mode dropout is implemented by the pipelined kernel (impl 4) only, got impl %d", impl);
  if (impl == 5) return launch_pipelined2_steps(w, st, n_steps, update_state, s);
  if (impl == 4) return launch_pipelined_steps(w, st, n_steps, update_state, s);
  if (impl == 1) {
    for (int i = 0; i < n_steps; ++i)
      if ((rc = enqueue_step_phases(w, st, update_state, s))) return rc;
    return 0;
  }
  TTS_REQUIRE(impl == 2, "decode_steps: unknown impl %d", impl);
  cudaGraphExec_t exec;
  int kernels = 0;
  if ((rc = get_step_graph(w, st, update_state, s, &exec, &kernels))) return rc;
  for (int i = 0; i < n_steps; ++i) TTS_CHECK_CUDA(cudaGraphLaunch(exec, s));
  count_launch(kernels * n_steps);
  return 0;
}
