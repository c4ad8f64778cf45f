// nerfw_volume_render: the inference form of volume_render (src/render.py:5-97) as ONE call -- direction normalisation,
// stratified depths, coarse MLP pass, compositing, inverse-CDF resampling, fine MLP pass on the new depths, merge of the
// two record lists, compositing of the merged row.  It launches exactly the kernels the separate entry points launch, in
// the same order with the same arguments (so its outputs are bit-identical to the call-by-call path); what it removes
// is the host time between the launches -- nine ctypes calls and as many allocations per 4096-ray chunk in the
// reference's own chunk loop (render_aligned_spiral.py:136-155) become one call on one workspace.
#include "common.cuh"

namespace {
inline size_t up16(size_t n) { return (n + 15u) & ~(size_t)15u; }

struct Layout {
  size_t dirs, raw_c, z_new, raw_f, raw_m, app, total;
};
Layout layout(int64_t b, int n, int ni, int64_t emb_rows) {
  Layout l;
  size_t off = 0;
  l.dirs = off;  off += up16((size_t)b * 3 * sizeof(float));
  l.raw_c = off; off += (size_t)b * n * 16;
  l.z_new = off; off += up16((size_t)b * ni * sizeof(float));
  l.raw_f = off; off += (size_t)b * ni * 16;
  l.raw_m = off; off += ni > 0 ? (size_t)b * (n + ni) * 16 : 0;
  l.app = off;   off += up16(nerfw_mlp_workspace_bytes(b, emb_rows));
  l.total = off;
  return l;
}
}  // namespace

extern "C" size_t nerfw_volume_render_workspace_bytes(int64_t n_rays, int n_samples, int n_importance, int64_t emb_rows) {
  if (n_rays < 0 || n_samples < 1 || n_importance < 0) return 0;
  return layout(n_rays, n_samples, n_importance, emb_rows).total;
}

extern "C" int nerfw_volume_render(const NerfwWeights* w, const void* packed, const float* rays_o, const float* rays_d,
                                   int64_t n_rays, const float* ztab, const float* t_rand, int n_samples,
                                   const float* u_lin, const float* u_rand, int n_importance, const float* emb,
                                   int64_t emb_rows, int mode_coarse, int mode_fine, const NerfwRenderOut* out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  using namespace nerfw;
  NERFW_REQUIRE(out, "nerfw_volume_render: null output struct");
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_importance >= 0, "nerfw_volume_render: bad shape n_rays=%lld n_samples=%d n_importance=%d",
                (long long)n_rays, n_samples, n_importance);
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(rays_o && rays_d && ztab, "nerfw_volume_render: null ray / depth-table pointer");
  NERFW_REQUIRE(out->rgb && out->depth && out->acc && out->weights && out->z_vals, "nerfw_volume_render: null output pointer");
  const bool hier = n_importance > 0;
  if (hier) {
    NERFW_REQUIRE(u_lin && u_rand, "nerfw_volume_render: n_importance > 0 needs u_lin and u_rand");
    NERFW_REQUIRE(out->rgb_coarse && out->depth_coarse && out->acc_coarse && out->weights_coarse && out->z_coarse,
                  "nerfw_volume_render: n_importance > 0 needs the coarse output pointers");
  }
  NERFW_REQUIRE(((mode_coarse | mode_fine) & ~0xff) == 0, "nerfw_volume_render: mode flags are set by the call itself");
  NERFW_REQUIRE(workspace && aligned16(workspace), "nerfw_volume_render: workspace must be a 16-byte aligned device buffer");
  const Layout l = layout(n_rays, n_samples, n_importance, emb_rows);
  if (workspace_bytes < l.total) {
    set_error("nerfw_volume_render: workspace of %zu bytes, need %zu", workspace_bytes, l.total);
    return NERFW_ESIZE;
  }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* dirs = reinterpret_cast<float*>(ws + l.dirs);
  float* raw_c = reinterpret_cast<float*>(ws + l.raw_c);
  float* z_new = reinterpret_cast<float*>(ws + l.z_new);
  float* raw_f = reinterpret_cast<float*>(ws + l.raw_f);
  float* raw_m = reinterpret_cast<float*>(ws + l.raw_m);
  void* app = ws + l.app;
  const size_t app_bytes = nerfw_mlp_workspace_bytes(n_rays, emb_rows);
  float* z = hier ? out->z_coarse : out->z_vals;

  int rc = nerfw_normalize_dirs(rays_d, n_rays, dirs, stream);                                   // src/render.py:19
  if (rc) return rc;
  rc = nerfw_stratified(nullptr, nullptr, ztab, t_rand, n_rays, n_samples, z, nullptr, stream);  // src/render.py:22
  if (rc) return rc;
  rc = nerfw_mlp_fwd(w, packed, rays_o, dirs, z, emb, emb_rows, n_rays, n_samples, mode_coarse, raw_c, nullptr, app,
                     app_bytes, stream);                                                           // src/render.py:29-53
  if (rc) return rc;
  if (!hier)
    return nerfw_composite_fwd(raw_c, z, n_rays, n_samples, out->rgb, out->depth, out->acc, out->weights, stream);
  rc = nerfw_composite_fwd(raw_c, z, n_rays, n_samples, out->rgb_coarse, out->depth_coarse, out->acc_coarse,
                           out->weights_coarse, stream);                                           // src/render.py:56-80
  if (rc) return rc;
  rc = nerfw_sample_pdf(z, out->weights_coarse, u_lin, u_rand, n_rays, n_samples, n_importance, out->z_vals, nullptr,
                        z_new, nullptr, stream);                                                   // src/ray_utils.py:90-149
  if (rc) return rc;
  // the embedding rows and the weights are those of the coarse launch: their rgb-logit offsets are in `app` already
  rc = nerfw_mlp_fwd(w, packed, rays_o, dirs, z_new, emb, emb_rows, n_rays, n_importance,
                     mode_fine | (emb ? NERFW_MLP_APP_CACHED : 0), raw_f, nullptr, app, app_bytes, stream);
  if (rc) return rc;
  rc = nerfw_merge_raw(z, raw_c, z_new, raw_f, n_rays, n_samples, n_importance, raw_m, stream);
  if (rc) return rc;
  return nerfw_composite_fwd(raw_m, out->z_vals, n_rays, n_samples + n_importance, out->rgb, out->depth, out->acc,
                             out->weights, stream);
}
