// Tensor-core (tcgen05) MLP: host-side launchers shared with mlp_api.cu.
#pragma once
#include "common.cuh"
#include "mlp_common.cuh"

namespace nerfw {
size_t mlp_tc_packed_bytes();          // forward image only
size_t mlp_tc_packed_total_bytes();    // forward image + transposed (dgrad) image + fp16 forward image
size_t mlp_tc_packed_f16_offset();     // byte offset of the fp16 forward image
int launch_pack_weights_t(const NerfwWeights& w, void* packed, cudaStream_t stream);
int launch_pack_weights(const NerfwWeights& w, void* packed, cudaStream_t stream);
int launch_mlp_tc_fwd(const NerfwWeights& w, const void* packed, const SampleSource& src, const float* app_off,
                      int64_t n_total, int mode_flags, float* raw, void* relu_masks, cudaStream_t stream);
// rgb-logit offset of the appearance embedding, one float4 per embedding row: W_rgb (W_app e + b_app)
int launch_app_offset(const NerfwWeights& w, const float* emb, int64_t emb_rows, float* app_off, cudaStream_t stream);
}  // namespace nerfw
