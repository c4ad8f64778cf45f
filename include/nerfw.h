/*
 * nerfw.h -- C ABI of libnerfw_sm100.so, the B200 (sm_100a) NeRF-W ray-marching hot path.
 *
 * The reference (ByeongKyuPark/Depth-Aware-Shader-Effects-for-NeRF) has no FFI: its "plugin API" is the
 * set of Python callables in src/ray_utils.py, src/render.py and src/models.py.  Each
 * entry point below is what a ctypes binding for one of those callables would bind; the reference
 * file:line it replaces is cited on every declaration.  INTEGRATION.md shows the Python-side stub.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - every pointer is a DEVICE pointer unless its name ends in _host; all tensors are dense,
 *    row-major fp32 (indices int64) and 16-byte aligned.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Functions only enqueue work;
 *    they never synchronise, allocate or free device memory, and keep no pointer past the call.
 *  - return value: 0 on success, a negative NERFW_E* code on failure; nerfw_last_error() then returns a
 *    thread-local message.  There is no CPU fallback anywhere.
 */
#ifndef NERFW_H
#define NERFW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERFW_ABI_VERSION 1

enum {
  NERFW_OK = 0,
  NERFW_EINVAL = -1,   /* bad argument (shape, null pointer, unsupported architecture constant) */
  NERFW_ECUDA = -2,    /* a CUDA runtime call or launch failed */
  NERFW_EDEVICE = -3,  /* current device is not sm_100 */
  NERFW_ESIZE = -4     /* workspace / packed buffer too small */
};

/* MLP arithmetic modes (nerfw_mlp_fwd `mode`). */
enum {
  NERFW_MLP_FP32 = 0,     /* CUDA-core fp32 FFMA; closest to the reference's fp32 nn.Linear */
  NERFW_MLP_BF16X3 = 1,   /* tcgen05 kind::f16, operands split hi+lo bf16, 3 MMAs per product (~2^-16 rel) */
  NERFW_MLP_BF16 = 2,     /* tcgen05 kind::f16, single bf16 MMA (stated looser bounds) */
  NERFW_MLP_FP16 = 3      /* tcgen05 kind::f16, single fp16 MMA: 8x finer operands than bf16, activations saturate at
                             65504; the fine pass of the default hierarchical render (coarse pass in BF16X3) */
};
/* OR into `mode` (tensor-core modes, inference): skip the direction layer and the rgb head, raw = (0, 0, 0, sigma).  All the
 * coarse pass of a hierarchical render needs: its weights place the fine samples, its colour is never looked at. */
#define NERFW_MLP_SIGMA_ONLY 0x100
/* OR into `mode` of nerfw_mlp_fwd when `workspace` still holds the per-embedding-row rgb-logit offsets an earlier
 * nerfw_mlp_fwd call wrote for the SAME embedding rows and the SAME weights (the coarse and the fine launch of one render,
 * or consecutive 4096-ray chunks of one frame): the small offset kernel is not launched again. */
#define NERFW_MLP_APP_CACHED 0x200

/* Architecture constants the kernels are specialised for (config.py:10-33 defaults). */
#define NERFW_HIDDEN 256
#define NERFW_LAYERS 8
#define NERFW_SKIP 4
#define NERFW_POS_LEVELS 10
#define NERFW_DIR_LEVELS 4
#define NERFW_POS_DIM 63   /* 3 + 3*2*10, src/models.py:22-27 */
#define NERFW_DIR_DIM 27   /* 3 + 3*2*4 */
#define NERFW_DIR_HIDDEN 128
#define NERFW_APP_DIM 32

/* The 24 state_dict tensors of src/models.py:80-103 in the reference layout ([out,in] row-major fp32).
 * app_w/app_b may be NULL when use_appearance is False (src/models.py:99-101). */
typedef struct NerfwWeights {
  const float* pts_w[NERFW_LAYERS];  /* (256,63) (256,256)x3 (256,319) (256,256)x3 */
  const float* pts_b[NERFW_LAYERS];  /* (256,) */
  const float* density_w;            /* (1,256) */
  const float* density_b;            /* (1,) */
  const float* dir_w;                /* (128,283): cols 0-255 = h, 256-282 = enc_d (src/models.py:141) */
  const float* dir_b;                /* (128,) */
  const float* app_w;                /* (128,32) or NULL */
  const float* app_b;                /* (128,) or NULL */
  const float* rgb_w;                /* (3,128) */
  const float* rgb_b;                /* (3,) */
} NerfwWeights;

/* Gradients, same layout; every pointer must be valid (app_* may be NULL iff the weights' are). */
typedef struct NerfwGrads {
  float* pts_w[NERFW_LAYERS];
  float* pts_b[NERFW_LAYERS];
  float* density_w;
  float* density_b;
  float* dir_w;
  float* dir_b;
  float* app_w;
  float* app_b;
  float* rgb_w;
  float* rgb_b;
} NerfwGrads;

const char* nerfw_last_error(void);
int nerfw_abi_version(void);
/* 0 if the current CUDA device is compute capability 10.x, NERFW_EDEVICE otherwise. */
int nerfw_check_device(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches counter). */
uint64_t nerfw_launch_count(void);

/* ---- rays: get_rays(height, width, focal_length, c2w)  -- src/ray_utils.py:4-50 ------------------------
 * c2w_host: 12 floats, rows 0..2 of the camera-to-world matrix, row-major [r*4+c] (HOST memory).
 * dirs (H,W,3) unit directions, bit-identical to the reference's CPU result; origins (H,W,3) may be NULL
 * (the reference returns a stride-0 expand of c2w[:3,3], src/ray_utils.py:48). */
int nerfw_raygen(int height, int width, float focal, const float* c2w_host, float* origins, float* dirs,
                 void* stream);

/* F.normalize(rays_d, dim=-1) -- src/render.py:19 (x / max(||x||, 1e-12)). in/out may alias. */
int nerfw_normalize_dirs(const float* dirs, int64_t n_rays, float* out, void* stream);

/* ---- sample_stratified(rays_o, rays_d, near, far, n_samples, perturb) -- src/ray_utils.py:52-88 --------
 * ztab: the N-entry table near + linspace(0,1,N)*(far-near) (src/ray_utils.py:69-70), computed by the caller
 * with host torch so its bits equal the reference's.  t_rand (B,N) uniforms or NULL (perturb=False).
 * z (B,N) out; pts (B,N,3) out or NULL. */
int nerfw_stratified(const float* rays_o, const float* rays_d, const float* ztab, const float* t_rand,
                     int64_t n_rays, int n_samples, float* z, float* pts, void* stream);

/* pts = o + d*z for given depths (src/ray_utils.py:86 and :147). */
int nerfw_ray_points(const float* rays_o, const float* rays_d, const float* z, int64_t n_rays, int n_samples,
                     float* pts, void* stream);

/* ---- sample_importance(rays_o, rays_d, z_vals, weights, n_importance) -- src/ray_utils.py:90-149 -------
 * (the "sample_pdf" of the north star; z-gather index clamped to N-1)
 * u_lin: NI-entry table linspace(0,1,NI+1)[:-1] from host torch (:115); u_rand (B,NI) uniforms (:119).
 * z_out (B,N+NI) sorted ascending.  Optional outputs (may be NULL): inds (B,NI) int64 = searchsorted result
 * (:122), z_fine (B,NI) (:139), cdf (B,N+1) (:111-112). */
int nerfw_sample_pdf(const float* z_vals, const float* weights, const float* u_lin, const float* u_rand,
                     int64_t n_rays, int n_samples, int n_importance, float* z_out, int64_t* inds,
                     float* z_fine, float* cdf, void* stream);
/* Same contract through the general path only (binary searches + ranking; no sorted-input shortcut).  The default entry
 * point falls back to this path per ray whenever its assumptions fail; exported so the parity tests can compare the two
 * bit for bit. */
int nerfw_sample_pdf_general(const float* z_vals, const float* weights, const float* u_lin, const float* u_rand,
                             int64_t n_rays, int n_samples, int n_importance, float* z_out, int64_t* inds,
                             float* z_fine, float* cdf, void* stream);

/* ---- PositionalEncoding.__call__ -- src/models.py:14-46 ----------------------------------------------
 * x (n,dim) -> out (n, dim*(include_input + 2*levels)): [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]. */
int nerfw_posenc(const float* x, int64_t n, int dim, int levels, int include_input, float* out, void* stream);

/* ---- NeRF.forward(x, d, appearance_embedding) -- src/models.py:105-162 -------------------------------
 * Packed weight cache for the tensor-core modes (derived data; the state_dict stays the source of truth). */
size_t nerfw_packed_bytes(void);
int nerfw_pack_weights(const NerfwWeights* w, void* packed, size_t packed_bytes, void* stream);

/* Samples are given either per sample (pts (S,3), dirs (S,3); z == NULL, n_samples == 1, n_rays == S) or
 * per ray (rays_o (B,3), unit rays_d (B,3), z (B,N)): sample (b,i) sits at o_b + d_b*z_bi with direction
 * d_b (src/render.py:22-30), so pts is never materialised.
 * emb: NULL, or (emb_rows, 32) with emb_rows == 1 (shared, src/train.py:68) or emb_rows == n_rays
 * (per ray, src/render.py:39-44; per sample when z == NULL).
 * out: raw (S,4) = (r,g,b,sigma) after sigmoid / relu (src/models.py:138,160).
 * `packed` is required for the BF16X3/BF16 modes (else may be NULL); `workspace` must hold
 * nerfw_mlp_workspace_bytes(n_rays, emb_rows) bytes.  relu_masks (NULL, or nerfw_mlp_mask_bytes() bytes; tensor-core modes
 * only) receives the ReLU gates of every layer so that nerfw_mlp_bwd_tc gates its gradients exactly like this forward. */
size_t nerfw_mlp_workspace_bytes(int64_t n_rays, int64_t emb_rows);
size_t nerfw_mlp_mask_bytes(int64_t n_rays, int n_samples);
int nerfw_mlp_fwd(const NerfwWeights* w, const void* packed, const float* pts_or_o, const float* dirs,
                  const float* z, const float* emb, int64_t emb_rows, int64_t n_rays, int n_samples, int mode,
                  float* raw, void* relu_masks, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above (autograd of src/models.py:105-162): d_raw (S,4) in; accumulates (+=) into grads and,
 * if d_emb != NULL, into d_emb (emb_rows,32).  fp32 CUDA-core arithmetic. */
size_t nerfw_mlp_bwd_workspace_bytes(int64_t n_rays, int n_samples, int64_t emb_rows);
int nerfw_mlp_bwd(const NerfwWeights* w, const float* pts_or_o, const float* dirs, const float* z,
                  const float* emb, int64_t emb_rows, int64_t n_rays, int n_samples, const float* d_raw,
                  const NerfwGrads* grads, float* d_emb, void* workspace, size_t workspace_bytes, void* stream);

/* Tensor-core backward (bf16 operands, fp32 accumulation; stated looser bounds): same contract as nerfw_mlp_bwd, with
 * the packed weight cache of nerfw_pack_weights; emb_rows == 1 (shared) or == n_rays (one row per ray, the cross-image
 * batches of src/dataset.py:248-277 with src/render.py:39-44).  The workspace holds the per-tile activation / dZ scratch
 * (1.16 MB per 128 samples) and 544 bytes per ray for the per-ray appearance rows.  relu_masks: the gates written by nerfw_mlp_fwd
 * (recommended: gradients then follow the forward that produced the loss), or NULL to use the bf16 recompute's own. */
size_t nerfw_mlp_bwd_tc_workspace_bytes(int64_t n_rays, int n_samples);
int nerfw_mlp_bwd_tc(const NerfwWeights* w, const void* packed, const float* pts_or_o, const float* dirs, const float* z,
                     const float* emb, int64_t emb_rows, int64_t n_rays, int n_samples, const float* d_raw,
                     const void* relu_masks, const NerfwGrads* grads, float* d_emb, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Shortcut of the hierarchical render when the coarse and the fine pass use ONE network (the reference has a
 * single model, src/render.py:49): the N coarse samples were evaluated in the coarse pass, so the fine pass evaluates only
 * the NI new depths (z_fine of nerfw_sample_pdf) and this call merges both sets of (r,g,b,sigma) records into the depth
 * order of the merged row (src/ray_utils.py:142-144).  raw_out (n_rays, N+NI, 4). */
int nerfw_merge_raw(const float* z_coarse, const float* raw_coarse, const float* z_fine, const float* raw_fine,
                    int64_t n_rays, int n_samples, int n_importance, float* raw_out, void* stream);
/* Backward of nerfw_merge_raw (training with one network, so that the fine pass of a training step also evaluates only the
 * NI new depths): scatters the gradient of the merged records back to the two lists with the same slot computation --
 * d_raw_fine[k] = d_merged[slot(k)]; d_raw_coarse[i] = d_merged[slot(i)], or += when accumulate_coarse != 0 (the buffer
 * then already holds the gradient that reached the coarse records through the coarse pass's own outputs). */
int nerfw_unmerge_raw(const float* z_coarse, const float* z_fine, const float* d_raw_merged, int64_t n_rays,
                      int n_samples, int n_importance, int accumulate_coarse, float* d_raw_coarse, float* d_raw_fine,
                      void* stream);

/* ---- compositing: the tail of volume_render -- src/render.py:56-80 ------------------------------------
 * raw (B,N,4) = (r,g,b,sigma), z (B,N).  Out: rgb_map (B,3), depth (B,1), acc (B,1) = sum of weights
 * (SURVEY.md F4), weights (B,N) or NULL. */
int nerfw_composite_fwd(const float* raw, const float* z, int64_t n_rays, int n_samples, float* rgb_map,
                        float* depth, float* acc, float* weights, void* stream);
/* autograd of the above: d_rgb_map (B,3), d_depth (B,1) or NULL, d_acc (B,1) or NULL, d_weights (B,N) or NULL
 * -> d_raw (B,N,4) (overwritten). */
int nerfw_composite_bwd(const float* raw, const float* z, int64_t n_rays, int n_samples, const float* d_rgb_map,
                        const float* d_depth, const float* d_acc, const float* d_weights, float* d_raw,
                        void* stream);

/* ---- volume_render(model, rays_o, rays_d, near, far, n_samples, n_importance, appearance_embedding, ...)
 * -- src/render.py:5-97, inference (no gradients), as ONE call.  It enqueues exactly the launches of the entry points
 * above, in the order the reference's function performs the steps: F.normalize of the directions (:19),
 * sample_stratified (:22), the MLP on the N coarse depths (:29-53), compositing (:56-80) and -- when n_importance > 0,
 * with one network for both passes -- sample_importance (src/ray_utils.py:90-149), the MLP on the NI new depths,
 * nerfw_merge_raw and compositing of the merged row.  Outputs are bit-identical to the call-by-call sequence; what the
 * call saves is host time per invocation (the reference renders a frame as 157 chunks of 4096 rays,
 * render_aligned_spiral.py:136-155).
 * rays_d need not be normalised.  ztab / t_rand as for nerfw_stratified, u_lin / u_rand as for nerfw_sample_pdf (unused
 * when n_importance == 0).  emb: NULL or (emb_rows,32), emb_rows 1 or n_rays.  mode_coarse / mode_fine: NERFW_MLP_* without
 * flags (n_importance == 0: only mode_coarse is used).
 * n_importance == 0: rgb, depth, acc, weights (B,N), z_vals (B,N) are written, the *_coarse pointers are ignored.
 * n_importance  > 0: rgb, depth, acc, weights (B,N+NI), z_vals (B,N+NI) hold the final result and rgb_coarse, depth_coarse,
 * acc_coarse, weights_coarse (B,N), z_coarse (B,N) the coarse pass's. */
typedef struct NerfwRenderOut {
  float* rgb;            /* (B,3) */
  float* depth;          /* (B,1) */
  float* acc;            /* (B,1) */
  float* weights;        /* (B,N+NI) */
  float* z_vals;         /* (B,N+NI) */
  float* rgb_coarse;     /* (B,3) */
  float* depth_coarse;   /* (B,1) */
  float* acc_coarse;     /* (B,1) */
  float* weights_coarse; /* (B,N) */
  float* z_coarse;       /* (B,N) */
} NerfwRenderOut;
size_t nerfw_volume_render_workspace_bytes(int64_t n_rays, int n_samples, int n_importance, int64_t emb_rows);
int nerfw_volume_render(const NerfwWeights* w, const void* packed, const float* rays_o, const float* rays_d,
                        int64_t n_rays, const float* ztab, const float* t_rand, int n_samples, const float* u_lin,
                        const float* u_rand, int n_importance, const float* emb, int64_t emb_rows, int mode_coarse,
                        int mode_fine, const NerfwRenderOut* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- optimizer: torch.optim.Adam.step as used at src/train.py:33-39,92 (no weight decay, no amsgrad) ---
 * One fused launch over a flat parameter buffer.  step is 1-based. grad_scale multiplies g first
 * (1/world_size after the data-parallel all-reduce). */
int nerfw_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
               float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

/* mse_loss(rgb, target) forward+backward in one launch (src/train.py:87): loss_out[0] = mean((a-b)^2),
 * d_rgb = 2*(a-b)/n * loss_scale. */
int nerfw_mse(const float* rgb, const float* target, int64_t n, float loss_scale, float* loss_out, float* d_rgb,
              void* stream);

/* rgb (n,3) float in [0,1] -> uint8, exactly (x*255).astype(uint8) of render_aligned_spiral.py:161-162;
 * depth (n) -> uint8 by the min/max normalisation of :171-173 given dmin,dmax. */
int nerfw_quantize_u8(const float* rgb, int64_t n_values, uint8_t* out, void* stream);

/* ---- depth-aware effects on the device-resident fp32 depth (src/post_processor.py) -------------
 * Images are uint8 (H,W,3) like the reference's effect inputs; depth is the renderer's fp32 (H,W) buffer, normalised
 * inside the kernels exactly like the reference: depth / depth.max() when depth.max() > 1 (:64-66, :405-408,
 * :473-477).  All pointers are device pointers unless named *_host. */

/* out[0] = max(x[0..n)) (depth.max()). */
int nerfw_max_f32(const float* x, int64_t n, float* out, void* stream);

/* _effect_fog (src/post_processor.py:451-493): f = clip(max(d_norm - fog_start, 0) / (1 - fog_start), 0, 1) ** power *
 * visibility; out = clip(image * f + fog_color * (1 - f), 0, 255) truncated to uint8.  depth_max: device scalar from
 * nerfw_max_f32.  Reference constants: fog_start 0.0, power 3.0, visibility 0.3, fog_color (255,255,255). */
int nerfw_fog(const uint8_t* image, const float* depth, const float* depth_max, int64_t n_pixels, float fog_start,
              float power, float visibility, const float* fog_color_host, uint8_t* out, void* stream);

/* Depth-edge detector shared by _effect_toon (:62-78: cv2.bilateralFilter(depth_norm, 9, 75, 75) then Sobel) and
 * _effect_hologram (:402-419: Sobel on depth_norm): mag = sqrt(Sobel_x^2 + Sobel_y^2) (ksize 3, BORDER_REFLECT_101),
 * mag_max[0] = max(mag).  bilateral_d = 0: no pre-filter (`filtered` may be NULL); otherwise the filtered normalised
 * depth is written to `filtered` (H,W). */
int nerfw_depth_edges(const float* depth, const float* depth_max, int height, int width, int bilateral_d,
                      float sigma_color, float sigma_space, float* filtered, float* mag, float* mag_max, void* stream);

/* _effect_toon (:64-102) with depth: colours floor(img/255*levels)/levels*255, multiplied by (1 - edge_strength * e),
 * e = 3x3 dilation of (mag / mag_max > 0.05). */
int nerfw_toon(const uint8_t* image, const float* mag, const float* mag_max, int height, int width, int levels,
               float edge_strength, uint8_t* out, void* stream);

/* _effect_hologram (:373-449): ((img/255) * (0.8,1.0,0.2) * row_scale[y] + (mag/mag_max) * (0.1,0.6,0.3) + noise),
 * x1.5 once per interference line covering column x (col_hits[x]), clip(x*255) truncated to uint8.  row_scale (H) is
 * the scanline table of :385-393; noise (H,W,3), col_hits (W) and mag/mag_max may be NULL.  The reference draws noise
 * and line positions from numpy's global RNG; here the caller supplies them. */
int nerfw_hologram(const uint8_t* image, const float* mag, const float* mag_max, const float* row_scale,
                   const int* col_hits, const float* noise, int height, int width, uint8_t* out, void* stream);

/* Primitive self-test (tests only): D (128,n) fp32 = A (128,k) bf16 * B (n,k) bf16 ^T through one tcgen05 tile;
 * mode 0 = A from shared memory, 1 = A from tensor memory.  Pins the descriptor / swizzle / TMEM layouts. */
int nerfw_selftest_umma(const void* a_bf16, const void* b_bf16, int n, int k, int mode, float* d, void* stream);
/* Same with both operands MN-major (the wgrad form): D (128,n) = At^T Bt for At (k,128), Bt (k,n) bf16 row-major. */
int nerfw_selftest_umma_mn(const void* at_bf16, const void* bt_bf16, int n, int k, float* d, void* stream);
/* Tensor-pipe issue-rate probe: device cycles (int64 at cycles_dev) for reps x 16 MMAs of shape 128 x n x 16;
 * mode 0 = K-major smem operands, 1 = A from tensor memory, 2 = both operands MN-major (kind::f16); 3 / 4 = kind::i8
 * (s8 x s8 -> s32, 128 x n x 32 per instruction) with A from shared / tensor memory, 5 = kind::f8f6f4 (e4m3), shared memory. */
int nerfw_selftest_umma_rate(int mode, int n, int reps, long long* cycles_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFW_H */
