// C-ABI entry points for the MLP: mode dispatch and argument validation.
#include "common.cuh"
#include "mlp_common.cuh"
#include "mlp_tc.cuh"
#include "mlp_tc_layout.cuh"

using namespace nerfw;

static int check_weights(const NerfwWeights* w, const char* who) {
  NERFW_REQUIRE(w, "%s: null weights", who);
  for (int i = 0; i < NERFW_LAYERS; ++i)
    NERFW_REQUIRE(w->pts_w[i] && w->pts_b[i], "%s: null pts_linears.%d parameter", who, i);
  NERFW_REQUIRE(w->density_w && w->density_b && w->dir_w && w->dir_b && w->rgb_w && w->rgb_b, "%s: null head parameter", who);
  NERFW_REQUIRE((w->app_w == nullptr) == (w->app_b == nullptr), "%s: appearance weight/bias must both be set or both null", who);
  return NERFW_OK;
}

// workspace = one float4 rgb-logit offset per embedding row (see app_offset_kernel)
extern "C" size_t nerfw_mlp_workspace_bytes(int64_t n_rays, int64_t emb_rows) {
  (void)n_rays;
  return 256 + 16 * (size_t)(emb_rows > 0 ? emb_rows : 0);
}

extern "C" int nerfw_mlp_fwd(const NerfwWeights* w, const void* packed, const float* pts_or_o, const float* dirs,
                             const float* z, const float* emb, int64_t emb_rows, int64_t n_rays, int n_samples,
                             int mode, float* raw, void* relu_masks, void* workspace, size_t workspace_bytes,
                             void* stream) {
  int rc = check_weights(w, "nerfw_mlp_fwd");
  if (rc) return rc;
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1, "nerfw_mlp_fwd: bad shape n_rays=%lld n_samples=%d", (long long)n_rays, n_samples);
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(z || n_samples == 1, "nerfw_mlp_fwd: n_samples must be 1 when z is NULL (per-sample inputs)");
  NERFW_REQUIRE(pts_or_o && dirs && raw, "nerfw_mlp_fwd: null input/output pointer");
  NERFW_REQUIRE(aligned16(raw), "nerfw_mlp_fwd: raw must be 16-byte aligned");
  if (emb) {
    NERFW_REQUIRE(w->app_w, "nerfw_mlp_fwd: embedding given but the model has no appearance_projection");
    NERFW_REQUIRE(emb_rows == 1 || emb_rows == n_rays, "nerfw_mlp_fwd: emb_rows=%lld must be 1 or n_rays=%lld",
                  (long long)emb_rows, (long long)n_rays);
  }
  SampleSource src;
  src.p = pts_or_o;
  src.d = dirs;
  src.z = z;
  src.emb = emb;
  src.n_per_ray = z ? n_samples : 1;
  src.emb_shared = (emb_rows == 1) ? 1 : 0;
  const int64_t total = n_rays * (z ? n_samples : 1);
  const float* app_off = nullptr;
  if (emb) {
    NERFW_REQUIRE(workspace && aligned16(workspace), "nerfw_mlp_fwd: workspace must be a 16-byte aligned device buffer");
    if (workspace_bytes < nerfw_mlp_workspace_bytes(n_rays, emb_rows)) {
      set_error("nerfw_mlp_fwd: workspace of %zu bytes, need %zu", workspace_bytes, nerfw_mlp_workspace_bytes(n_rays, emb_rows));
      return NERFW_ESIZE;
    }
    if (!(mode & NERFW_MLP_APP_CACHED)) {   // else: the offsets of these rows are already in the workspace
      rc = launch_app_offset(*w, emb, emb_rows, reinterpret_cast<float*>(workspace), as_stream(stream));
      if (rc) return rc;
    }
    app_off = reinterpret_cast<const float*>(workspace);
  }
  const int mode_flags = mode;
  mode &= 0xff;   // NERFW_MLP_SIGMA_ONLY rides in the upper bits
  switch (mode) {
    case NERFW_MLP_FP32:
      NERFW_REQUIRE(!relu_masks, "nerfw_mlp_fwd: relu_masks are produced by the tensor-core modes only");
      return launch_mlp_ffma_fwd(*w, src, app_off, total, raw, as_stream(stream));
    case NERFW_MLP_BF16X3:
    case NERFW_MLP_BF16:
    case NERFW_MLP_FP16:
      NERFW_REQUIRE(packed, "nerfw_mlp_fwd: tensor-core modes need the packed weight cache (nerfw_pack_weights)");
      NERFW_REQUIRE(!(mode_flags & NERFW_MLP_SIGMA_ONLY) || !relu_masks, "nerfw_mlp_fwd: NERFW_MLP_SIGMA_ONLY is an inference flag (no relu_masks)");
      return launch_mlp_tc_fwd(*w, packed, src, app_off, total, mode_flags, raw, relu_masks, as_stream(stream));
    default:
      set_error("nerfw_mlp_fwd: unknown mode %d", mode);
      return NERFW_EINVAL;
  }
}

extern "C" size_t nerfw_mlp_mask_bytes(int64_t n_rays, int n_samples) {
  const int64_t total = n_rays * (int64_t)(n_samples > 0 ? n_samples : 1);
  return (size_t)ceil_div64(total, tc::TM) * tc::MASK_WORDS_PER_TILE * sizeof(uint32_t);
}

extern "C" size_t nerfw_packed_bytes(void) { return mlp_tc_packed_total_bytes(); }

extern "C" int nerfw_pack_weights(const NerfwWeights* w, void* packed, size_t packed_bytes, void* stream) {
  int rc = check_weights(w, "nerfw_pack_weights");
  if (rc) return rc;
  NERFW_REQUIRE(packed, "nerfw_pack_weights: null destination");
  if (packed_bytes < mlp_tc_packed_total_bytes()) {
    set_error("nerfw_pack_weights: buffer of %zu bytes is smaller than nerfw_packed_bytes() = %zu", packed_bytes, mlp_tc_packed_total_bytes());
    return NERFW_ESIZE;
  }
  NERFW_REQUIRE(aligned16(packed), "nerfw_pack_weights: destination must be 16-byte aligned");
  rc = launch_pack_weights(*w, packed, as_stream(stream));
  if (rc) return rc;
  return launch_pack_weights_t(*w, packed, as_stream(stream));
}
