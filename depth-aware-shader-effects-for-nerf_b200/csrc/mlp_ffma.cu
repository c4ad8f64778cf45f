// K4 (fp32 mode): the whole NeRF-W MLP forward for a tile of 64 samples per CTA on CUDA cores (FFMA), fused with
// ray-point generation and positional encoding.  src/models.py:105-162.
//
// This is the NERFW_MLP_FP32 mode of nerfw_mlp_fwd: plain fp32 multiply-adds like the reference's nn.Linear, so it
// is the tightest-parity path (~1e-6) and the on-device cross-check for the tcgen05 kernel in mlp_tc.cu.
// Activations never leave shared memory; weights are read from L2 in their state_dict layout.
#include "common.cuh"
#include "mlp_common.cuh"
#include "ffma_dense.cuh"

namespace nerfw {

namespace ffma {
constexpr int B_STRIDE = 256;
constexpr int ENCD_STRIDE = 28;

struct Smem {
  float a[TM * A_STRIDE];
  float b[TM * B_STRIDE];
  float ws[2][KC * WS_STRIDE];
  float encd[TM * ENCD_STRIDE];
  float off[TM * 4];  // per-sample rgb logit offset W_rgb (W_app e + b_app)
};

__global__ void __launch_bounds__(THREADS, 1) mlp_ffma_fwd_kernel(NerfwWeights w, SampleSource src, const float4* __restrict__ app_off,
                                                                   int64_t n_samples_total, float4* __restrict__ raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (n_samples_total + TM - 1) / TM;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t s0 = tile * TM;
    // ---- inputs: position, direction, encodings ------------------------------------------------------------
    for (int u = tid; u < TM * (NERFW_POS_LEVELS + 1); u += THREADS) {
      int row = u / (NERFW_POS_LEVELS + 1), l = u - row * (NERFW_POS_LEVELS + 1);
      int64_t s = s0 + row;
      float x[3] = {0.f, 0.f, 0.f};
      if (s < n_samples_total) src.position(s, x);
      float* e = sm.a + row * A_STRIDE + 256;
      if (l == 0) {
        e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
        e[63] = 0.f;  // K padding
      } else {
        float f = (float)(1u << (l - 1));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sn, cs;
          sincosf(f * x[c], &sn, &cs);
          e[3 + 6 * (l - 1) + c] = sn;
          e[6 + 6 * (l - 1) + c] = cs;
        }
      }
    }
    for (int u = tid; u < TM * (NERFW_DIR_LEVELS + 1); u += THREADS) {
      int row = u / (NERFW_DIR_LEVELS + 1), l = u - row * (NERFW_DIR_LEVELS + 1);
      int64_t s = s0 + row;
      float d[3] = {0.f, 0.f, 0.f};
      if (s < n_samples_total) src.direction(s, d);
      float* e = sm.encd + row * ENCD_STRIDE;
      if (l == 0) {
        e[0] = d[0]; e[1] = d[1]; e[2] = d[2];
        e[27] = 0.f;
      } else {
        float f = (float)(1u << (l - 1));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sn, cs;
          sincosf(f * d[c], &sn, &cs);
          e[3 + 6 * (l - 1) + c] = sn;
          e[6 + 6 * (l - 1) + c] = cs;
        }
      }
    }
    // ---- appearance: rgb logit offset W_rgb (W_app e + b_app), precomputed per embedding row (mlp_tc.cu) ------
    if (tid < TM) {
      int64_t s = s0 + tid;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (app_off && s < n_samples_total) o = __ldg(app_off + src.emb_row(s));
      sm.off[tid * 4 + 0] = o.x; sm.off[tid * 4 + 1] = o.y; sm.off[tid * 4 + 2] = o.z;
    }
    __syncthreads();

    // ---- trunk (src/models.py:128-134) ----------------------------------------------------------------------
    dense<8, true, false>(sm.a + 256, A_STRIDE, NERFW_POS_DIM, w.pts_w[0], NERFW_POS_DIM, w.pts_b[0], sm.b, B_STRIDE, sm.ws);
    dense<8, true, false>(sm.b, B_STRIDE, 256, w.pts_w[1], 256, w.pts_b[1], sm.a, A_STRIDE, sm.ws);
    dense<8, true, false>(sm.a, A_STRIDE, 256, w.pts_w[2], 256, w.pts_b[2], sm.b, B_STRIDE, sm.ws);
    dense<8, true, false>(sm.b, B_STRIDE, 256, w.pts_w[3], 256, w.pts_b[3], sm.a, A_STRIDE, sm.ws);
    dense<8, true, false>(sm.a, A_STRIDE, 256 + NERFW_POS_DIM, w.pts_w[4], 256 + NERFW_POS_DIM, w.pts_b[4], sm.b, B_STRIDE, sm.ws);  // [h, enc_x]
    dense<8, true, false>(sm.b, B_STRIDE, 256, w.pts_w[5], 256, w.pts_b[5], sm.a, A_STRIDE, sm.ws);
    dense<8, true, false>(sm.a, A_STRIDE, 256, w.pts_w[6], 256, w.pts_b[6], sm.b, B_STRIDE, sm.ws);
    dense<8, true, false>(sm.b, B_STRIDE, 256, w.pts_w[7], 256, w.pts_b[7], sm.a, A_STRIDE, sm.ws);

    // ---- density head (src/models.py:137-138) + enc_d next to h for the direction layer ---------------------
    float sigma_keep = 0.f;  // lane 0 of each 8-row group keeps sigma for rows warp*8 + i (i = lane)
    {
      float wv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) wv[q] = __ldg(w.density_w + lane + 32 * q);
      float bs = __ldg(w.density_b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* h = sm.a + (warp * 8 + i) * A_STRIDE;
        float p = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) p = fmaf(h[lane + 32 * q], wv[q], p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == i) sigma_keep = fmaxf(p + bs, 0.f);
      }
    }
    for (int u = tid; u < TM * 32; u += THREADS) {
      int row = u >> 5, c = u & 31;
      sm.a[row * A_STRIDE + 256 + c] = (c < NERFW_DIR_DIM) ? sm.encd[row * ENCD_STRIDE + c] : 0.f;
    }
    __syncthreads();
    // ---- direction layer (src/models.py:141-143) ------------------------------------------------------------
    dense<4, true, false>(sm.a, A_STRIDE, 256 + NERFW_DIR_DIM, w.dir_w, 256 + NERFW_DIR_DIM, w.dir_b, sm.b, B_STRIDE, sm.ws);
    // ---- rgb head (src/models.py:159-160) -------------------------------------------------------------------
    {
      float wr[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) wr[c][q] = __ldg(w.rgb_w + c * NERFW_DIR_HIDDEN + lane + 32 * q);
      float br[3] = {__ldg(w.rgb_b), __ldg(w.rgb_b + 1), __ldg(w.rgb_b + 2)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int row = warp * 8 + i;
        const float* h = sm.b + row * B_STRIDE;
        float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float hv = h[lane + 32 * q];
#pragma unroll
          for (int c = 0; c < 3; ++c) p[c] = fmaf(hv, wr[c][q], p[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) p[c] += __shfl_xor_sync(0xffffffffu, p[c], o);
        float sg = __shfl_sync(0xffffffffu, sigma_keep, i);
        if (lane == 0 && s0 + row < n_samples_total) {
          float4 o4;
          o4.x = 1.0f / (1.0f + expf(-(p[0] + br[0] + sm.off[row * 4 + 0])));
          o4.y = 1.0f / (1.0f + expf(-(p[1] + br[1] + sm.off[row * 4 + 1])));
          o4.z = 1.0f / (1.0f + expf(-(p[2] + br[2] + sm.off[row * 4 + 2])));
          o4.w = sg;
          raw[s0 + row] = o4;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace ffma

int launch_mlp_ffma_fwd(const NerfwWeights& w, const SampleSource& src, const float* app_off, int64_t n_total, float* raw,
                        cudaStream_t stream) {
  static thread_local unsigned long long attr_mask = 0;
  const size_t smem = sizeof(ffma::Smem);
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(ffma::mlp_ffma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int64_t ntiles = ceil_div64(n_total, ffma::TM);
  int64_t grid = ntiles < sm_count() ? ntiles : sm_count();
  ffma::mlp_ffma_fwd_kernel<<<(unsigned)grid, ffma::THREADS, smem, stream>>>(w, src, reinterpret_cast<const float4*>(app_off), n_total,
                                                                             reinterpret_cast<float4*>(raw));
  NERFW_LAUNCHED();
  return NERFW_OK;
}

}  // namespace nerfw
