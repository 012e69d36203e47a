// Placeholder for the fused persistent decode kernel (impl 3); filled in by a later commit.
#include "common.cuh"

namespace tts {
int launch_fused_steps(const TtsDecoderWeights*, const TtsDecodeState*, int, int, cudaStream_t) {
  set_error("fused decode kernel not built");
  return 3;
}
size_t fused_scratch_floats(const TtsDecoderWeights*, int) { return 0; }
bool fused_supported(const TtsDecoderWeights*, const TtsDecodeState*) { return false; }
}  // namespace tts
