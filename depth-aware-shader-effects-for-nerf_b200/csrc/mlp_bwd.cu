// K5 (fp32): backward of the NeRF-W MLP (autograd of src/models.py:105-162) on CUDA cores.
//
// Per CTA and per tile of 64 samples: the forward is recomputed with every layer's activations written to a CTA-private
// scratch area (64 x 2272 floats = 568 KB per CTA, 84 MB for 148 CTAs: it stays in the 126 MB L2, so nothing but the
// inputs and the gradient atomics reaches HBM); then the chain dZ_l = dH_{l+1} * relu', dW_l += dZ_l^T X_l,
// db_l += sum dZ_l, dH_l = dZ_l W_l runs layer by layer with both operands in shared memory.
// Weight gradients are accumulated into the caller's buffers with fp32 atomics (red.global.add).
#include "common.cuh"
#include "mlp_common.cuh"
#include "ffma_dense.cuh"

namespace nerfw {
namespace ffma {

// scratch row layout (floats)
constexpr int SC_ENCX = 0;      // 64
constexpr int SC_H = 64;        // H1..H8: 8 x 256
constexpr int SC_ENCD = 2112;   // 32
constexpr int SC_HD = 2144;     // 128
constexpr int SC_ROW = 2272;
constexpr int APP_ROW = 132;    // per embedding row: G[128] (sum of d_hd) | DL[3] (sum of d_logit) | pad

struct BwdSmem {
  float p[TM * A_STRIDE];
  float q[TM * A_STRIDE];
  float ws[2][KC * WS_STRIDE];
  float dlog[TM * 4];  // d loss / d rgb logits
  float dsig[TM];      // d loss / d sigma pre-activation
};

// gW[o][k] += sum_s dZ[s][o] * X[s][k]   for o < n_out (multiple of 64), k < K.  dZ, X: smem, row stride A_STRIDE.
__device__ __forceinline__ void dw_gemm(const float* __restrict__ dz, const float* __restrict__ x, int n_out, int K,
                                        float* __restrict__ gW) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int ob = 0; ob < n_out; ob += 64) {
    const int o0 = ob + warp * 8;
    for (int kb = 0; kb < K; kb += 256) {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      const int nj = min(8, (K - kb + 31) / 32);
      if (nj == 8) {
#pragma unroll 2
        for (int s = 0; s < TM; ++s) {
          const float4 a0 = *reinterpret_cast<const float4*>(dz + s * A_STRIDE + o0);
          const float4 a1 = *reinterpret_cast<const float4*>(dz + s * A_STRIDE + o0 + 4);
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          float xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] = x[s * A_STRIDE + kb + lane + 32 * j];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
      } else {  // ragged tail of the concatenated layers (63 or 27 extra columns): two column groups at most
        for (int s = 0; s < TM; ++s) {
          const float4 a0 = *reinterpret_cast<const float4*>(dz + s * A_STRIDE + o0);
          const float4 a1 = *reinterpret_cast<const float4*>(dz + s * A_STRIDE + o0 + 4);
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          float xv[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            int k = kb + lane + 32 * j;
            xv[j] = (k < K) ? x[s * A_STRIDE + k] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int k = kb + lane + 32 * j;
          if (j < nj && k < K) atomicAdd(gW + (size_t)(o0 + i) * K + k, acc[i][j]);
        }
    }
  }
}

// gb[o] += sum_s dZ[s][o]
__device__ __forceinline__ void db_sum(const float* __restrict__ dz, int n_out, float* __restrict__ gb) {
  for (int o = threadIdx.x; o < n_out; o += THREADS) {
    float a = 0.f;
#pragma unroll 8
    for (int s = 0; s < TM; ++s) a += dz[s * A_STRIDE + o];
    atomicAdd(gb + o, a);
  }
}

__global__ void __launch_bounds__(THREADS, 1) mlp_ffma_bwd_kernel(NerfwWeights w, NerfwGrads g, SampleSource src,
                                                                   const float4* __restrict__ app_off,
                                                                   const float4* __restrict__ d_raw, int64_t n_total,
                                                                   float* __restrict__ scratch_all,
                                                                   float* __restrict__ app_acc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (n_total + TM - 1) / TM;
  float* sc = scratch_all + (size_t)blockIdx.x * TM * SC_ROW;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t s0 = tile * TM;
    // =============================== forward recompute ===============================
    for (int u = tid; u < TM * (NERFW_POS_LEVELS + 1); u += THREADS) {
      int row = u / (NERFW_POS_LEVELS + 1), l = u - row * (NERFW_POS_LEVELS + 1);
      float x[3] = {0.f, 0.f, 0.f};
      if (s0 + row < n_total) src.position(s0 + row, x);
      float* e = sm.p + row * A_STRIDE + 256;
      encode_level(e, x, l);
      if (l == 0) e[63] = 0.f;
    }
    __syncthreads();
    for (int u = tid; u < TM * 64; u += THREADS) sc[(size_t)(u >> 6) * SC_ROW + SC_ENCX + (u & 63)] = sm.p[(u >> 6) * A_STRIDE + 256 + (u & 63)];
    float* H = sc + SC_H;
    dense<8, true, false>(sm.p + 256, A_STRIDE, NERFW_POS_DIM, w.pts_w[0], NERFW_POS_DIM, w.pts_b[0], sm.q, A_STRIDE, sm.ws, H + 0 * 256, SC_ROW);
    dense<8, true, false>(sm.q, A_STRIDE, 256, w.pts_w[1], 256, w.pts_b[1], sm.p, A_STRIDE, sm.ws, H + 1 * 256, SC_ROW);
    dense<8, true, false>(sm.p, A_STRIDE, 256, w.pts_w[2], 256, w.pts_b[2], sm.q, A_STRIDE, sm.ws, H + 2 * 256, SC_ROW);
    dense<8, true, false>(sm.q, A_STRIDE, 256, w.pts_w[3], 256, w.pts_b[3], sm.p, A_STRIDE, sm.ws, H + 3 * 256, SC_ROW);
    dense<8, true, false>(sm.p, A_STRIDE, 256 + NERFW_POS_DIM, w.pts_w[4], 256 + NERFW_POS_DIM, w.pts_b[4], sm.q, A_STRIDE, sm.ws, H + 4 * 256, SC_ROW);
    dense<8, true, false>(sm.q, A_STRIDE, 256, w.pts_w[5], 256, w.pts_b[5], sm.p, A_STRIDE, sm.ws, H + 5 * 256, SC_ROW);
    dense<8, true, false>(sm.p, A_STRIDE, 256, w.pts_w[6], 256, w.pts_b[6], sm.q, A_STRIDE, sm.ws, H + 6 * 256, SC_ROW);
    dense<8, true, false>(sm.q, A_STRIDE, 256, w.pts_w[7], 256, w.pts_b[7], sm.p, A_STRIDE, sm.ws, H + 7 * 256, SC_ROW);
    // density head: d sigma_pre = d sigma * [pre > 0]   (src/models.py:137-138)
    {
      float wv[8];
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) wv[qd] = __ldg(w.density_w + lane + 32 * qd);
      const float bs = __ldg(w.density_b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 8 + i;
        const float* h = sm.p + row * A_STRIDE;
        float pr = 0.f;
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) pr = fmaf(h[lane + 32 * qd], wv[qd], pr);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
        if (lane == 0) {
          float ds = (s0 + row < n_total) ? __ldg(&d_raw[s0 + row].w) : 0.f;
          sm.dsig[row] = (pr + bs > 0.f) ? ds : 0.f;
        }
      }
    }
    // enc_d next to h8
    for (int u = tid; u < TM * (NERFW_DIR_LEVELS + 1); u += THREADS) {
      int row = u / (NERFW_DIR_LEVELS + 1), l = u - row * (NERFW_DIR_LEVELS + 1);
      float d[3] = {0.f, 0.f, 0.f};
      if (s0 + row < n_total) src.direction(s0 + row, d);
      float* e = sm.p + row * A_STRIDE + 256;
      encode_level(e, d, l);
      if (l == 0) {
#pragma unroll
        for (int k = NERFW_DIR_DIM; k < 32; ++k) e[k] = 0.f;
      }
    }
    __syncthreads();
    dense<4, true, false>(sm.p, A_STRIDE, 256 + NERFW_DIR_DIM, w.dir_w, 256 + NERFW_DIR_DIM, w.dir_b, sm.q, A_STRIDE, sm.ws, sc + SC_HD, SC_ROW);
    // rgb head: d logit_c = d rgb_c * rgb_c (1 - rgb_c)   (src/models.py:159-160)
    {
      float wr[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) wr[c][qd] = __ldg(w.rgb_w + c * NERFW_DIR_HIDDEN + lane + 32 * qd);
      const float br[3] = {__ldg(w.rgb_b), __ldg(w.rgb_b + 1), __ldg(w.rgb_b + 2)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 8 + i;
        const float* h = sm.q + row * A_STRIDE;
        float pr[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          float hv = h[lane + 32 * qd];
#pragma unroll
          for (int c = 0; c < 3; ++c) pr[c] = fmaf(hv, wr[c][qd], pr[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pr[c] += __shfl_xor_sync(0xffffffffu, pr[c], o);
        if (lane == 0) {
          const int64_t s = s0 + row;
          float4 dr = make_float4(0.f, 0.f, 0.f, 0.f), off = make_float4(0.f, 0.f, 0.f, 0.f);
          if (s < n_total) {
            dr = __ldg(d_raw + s);
            if (app_off) off = __ldg(app_off + src.emb_row(s));
          }
          float r0 = 1.0f / (1.0f + expf(-(pr[0] + br[0] + off.x)));
          float r1 = 1.0f / (1.0f + expf(-(pr[1] + br[1] + off.y)));
          float r2 = 1.0f / (1.0f + expf(-(pr[2] + br[2] + off.z)));
          sm.dlog[row * 4 + 0] = dr.x * r0 * (1.0f - r0);
          sm.dlog[row * 4 + 1] = dr.y * r1 * (1.0f - r1);
          sm.dlog[row * 4 + 2] = dr.z * r2 * (1.0f - r2);
          sm.dlog[row * 4 + 3] = 0.f;
        }
      }
    }
    __syncthreads();

    // =============================== backward ===============================
    // rgb head: dW_rgb[c][k] += sum_s dlog[s][c] hd[s][k];  db_rgb[c] += sum_s dlog[s][c]     (hd = q[:, :128])
    // per-embedding-row accumulators for the appearance branch: G[row][k] += d_hd[s][k], DL[row][c] += dlog[s][c]
    if (tid < NERFW_DIR_HIDDEN) {
      const int k = tid;
      const float w0 = __ldg(w.rgb_w + k), w1 = __ldg(w.rgb_w + 128 + k), w2 = __ldg(w.rgb_w + 256 + k);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, gacc = 0.f;
      int64_t cur = -1;
      for (int s = 0; s < TM; ++s) {
        const float d0 = sm.dlog[s * 4], d1 = sm.dlog[s * 4 + 1], d2 = sm.dlog[s * 4 + 2];
        const float hv = sm.q[s * A_STRIDE + k];
        a0 = fmaf(d0, hv, a0); a1 = fmaf(d1, hv, a1); a2 = fmaf(d2, hv, a2);
        const float dh = d0 * w0 + d1 * w1 + d2 * w2;  // d loss / d (hd + app)
        if (app_acc && s0 + s < n_total) {
          const int64_t er = src.emb_row(s0 + s);
          if (er != cur) {
            if (cur >= 0) atomicAdd(app_acc + cur * APP_ROW + k, gacc);
            cur = er; gacc = 0.f;
          }
          gacc += dh;
        }
        sm.q[s * A_STRIDE + k] = hv > 0.f ? dh : 0.f;  // dZ of the direction layer, in place
      }
      if (app_acc && cur >= 0) atomicAdd(app_acc + cur * APP_ROW + k, gacc);
      atomicAdd(g.rgb_w + k, a0);
      atomicAdd(g.rgb_w + 128 + k, a1);
      atomicAdd(g.rgb_w + 256 + k, a2);
    } else if (tid < NERFW_DIR_HIDDEN + 3) {
      const int c = tid - NERFW_DIR_HIDDEN;
      float a = 0.f, gacc = 0.f;
      int64_t cur = -1;
      for (int s = 0; s < TM; ++s) {
        const float dv = sm.dlog[s * 4 + c];
        a += dv;
        if (app_acc && s0 + s < n_total) {
          const int64_t er = src.emb_row(s0 + s);
          if (er != cur) {
            if (cur >= 0) atomicAdd(app_acc + cur * APP_ROW + 128 + c, gacc);
            cur = er; gacc = 0.f;
          }
          gacc += dv;
        }
      }
      if (app_acc && cur >= 0) atomicAdd(app_acc + cur * APP_ROW + 128 + c, gacc);
      atomicAdd(g.rgb_b + c, a);
    }
    __syncthreads();
    // direction layer: X = p = [h8 | enc_d], dZ = q[:, :128]
    db_sum(sm.q, NERFW_DIR_HIDDEN, g.dir_b);
    dw_gemm(sm.q, sm.p, NERFW_DIR_HIDDEN, 256 + NERFW_DIR_DIM, g.dir_w);
    __syncthreads();
    // density head weight grads need h8 (still in p[:, :256]) before it is overwritten
    {
      const int k = tid;  // 256 threads <-> 256 inputs
      float a = 0.f;
      for (int s = 0; s < TM; ++s) a = fmaf(sm.dsig[s], sm.p[s * A_STRIDE + k], a);
      atomicAdd(g.density_w + k, a);
      if (tid == 0) {
        float b = 0.f;
        for (int s = 0; s < TM; ++s) b += sm.dsig[s];
        atomicAdd(g.density_b, b);
      }
    }
    __syncthreads();
    // dH8 = dZd W_dir[:, :256] + dsig w_sigma^T  -> p[:, :256]
    dense<8, false, true>(sm.q, A_STRIDE, NERFW_DIR_HIDDEN, w.dir_w, 256 + NERFW_DIR_DIM, nullptr, sm.p, A_STRIDE, sm.ws);
    {
      const float wk = __ldg(w.density_w + tid);
      for (int s = 0; s < TM; ++s) sm.p[s * A_STRIDE + tid] = fmaf(sm.dsig[s], wk, sm.p[s * A_STRIDE + tid]);
    }
    __syncthreads();
    float* G = sm.p;  // gradient wrt the output of layer l (post-ReLU)
    float* O = sm.q;  // receives X_l, then dX
    for (int l = NERFW_LAYERS - 1; l >= 0; --l) {
      const int K = l == 0 ? NERFW_POS_DIM : (l == NERFW_SKIP ? 256 + NERFW_POS_DIM : 256);
      // dZ = dH * [H_{l+1} > 0] in place; X_l -> O
      for (int u = tid; u < TM * 256; u += THREADS) {
        const int row = u >> 8, c = u & 255;
        const float hv = H[(size_t)row * SC_ROW + l * 256 + c];
        if (!(hv > 0.f)) G[row * A_STRIDE + c] = 0.f;
        if (l > 0) O[row * A_STRIDE + c] = H[(size_t)row * SC_ROW + (l - 1) * 256 + c];
      }
      if (l == NERFW_SKIP || l == 0) {
        const int base = l == 0 ? 0 : 256;
        for (int u = tid; u < TM * 64; u += THREADS) O[(u >> 6) * A_STRIDE + base + (u & 63)] = sc[(size_t)(u >> 6) * SC_ROW + SC_ENCX + (u & 63)];
      }
      __syncthreads();
      db_sum(G, 256, g.pts_b[l]);
      dw_gemm(G, O, 256, K, g.pts_w[l]);
      __syncthreads();
      if (l > 0) {
        dense<8, false, true>(G, A_STRIDE, 256, w.pts_w[l], K, nullptr, O, A_STRIDE, sm.ws);
        float* t = G; G = O; O = t;
      }
    }
    __syncthreads();
  }
}

// Appearance branch finalisation, one CTA (128 threads) per embedding row:
//   a = W_app e + b_app;  dW_rgb[c][k] += DL[c] a[k];  dW_app[k][q] += G[k] e[q];  db_app[k] += G[k];
//   d_emb[row][q] += sum_k G[k] W_app[k][q]
__global__ void __launch_bounds__(128) app_bwd_kernel(NerfwWeights w, NerfwGrads g, const float* __restrict__ emb,
                                                      int64_t rows, const float* __restrict__ app_acc,
                                                      float* __restrict__ d_emb) {
  __shared__ float e[NERFW_APP_DIM];
  __shared__ float gk[NERFW_DIR_HIDDEN];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int k = threadIdx.x;
    if (k < NERFW_APP_DIM) e[k] = __ldg(emb + row * NERFW_APP_DIM + k);
    const float Gk = app_acc[row * APP_ROW + k];
    gk[k] = Gk;
    __syncthreads();
    float a = __ldg(w.app_b + k);
#pragma unroll 8
    for (int q = 0; q < NERFW_APP_DIM; ++q) a = fmaf(__ldg(w.app_w + k * NERFW_APP_DIM + q), e[q], a);
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(g.rgb_w + c * NERFW_DIR_HIDDEN + k, app_acc[row * APP_ROW + 128 + c] * a);
    for (int q = 0; q < NERFW_APP_DIM; ++q) atomicAdd(g.app_w + k * NERFW_APP_DIM + q, Gk * e[q]);
    atomicAdd(g.app_b + k, Gk);
    if (d_emb && k < NERFW_APP_DIM) {
      float de = 0.f;
      for (int kk = 0; kk < NERFW_DIR_HIDDEN; ++kk) de = fmaf(gk[kk], __ldg(w.app_w + kk * NERFW_APP_DIM + k), de);
      atomicAdd(d_emb + row * NERFW_APP_DIM + k, de);
    }
    __syncthreads();
  }
}

}  // namespace ffma

int launch_app_offset(const NerfwWeights& w, const float* emb, int64_t emb_rows, float* app_off, cudaStream_t stream);

}  // namespace nerfw

using namespace nerfw;

static size_t bwd_scratch_floats() { return (size_t)sm_count() * ffma::TM * ffma::SC_ROW; }

extern "C" size_t nerfw_mlp_bwd_workspace_bytes(int64_t n_rays, int n_samples, int64_t emb_rows) {
  (void)n_rays;
  (void)n_samples;
  size_t rows = emb_rows > 0 ? (size_t)emb_rows : 0;
  // [app_off: rows float4][app_acc: rows x 132][scratch]
  return 256 + rows * 16 + rows * ffma::APP_ROW * sizeof(float) + 256 + bwd_scratch_floats() * sizeof(float);
}

extern "C" int nerfw_mlp_bwd(const NerfwWeights* w, const float* pts_or_o, const float* dirs, const float* z,
                             const float* emb, int64_t emb_rows, int64_t n_rays, int n_samples, const float* d_raw,
                             const NerfwGrads* grads, float* d_emb, void* workspace, size_t workspace_bytes,
                             void* stream) {
  NERFW_REQUIRE(w && grads, "nerfw_mlp_bwd: null weights or grads");
  for (int i = 0; i < NERFW_LAYERS; ++i)
    NERFW_REQUIRE(w->pts_w[i] && w->pts_b[i] && grads->pts_w[i] && grads->pts_b[i], "nerfw_mlp_bwd: null pts_linears.%d parameter or gradient", i);
  NERFW_REQUIRE(w->density_w && w->density_b && w->dir_w && w->dir_b && w->rgb_w && w->rgb_b, "nerfw_mlp_bwd: null head parameter");
  NERFW_REQUIRE(grads->density_w && grads->density_b && grads->dir_w && grads->dir_b && grads->rgb_w && grads->rgb_b,
                "nerfw_mlp_bwd: null head gradient");
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1, "nerfw_mlp_bwd: bad shape");
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(z || n_samples == 1, "nerfw_mlp_bwd: n_samples must be 1 when z is NULL");
  NERFW_REQUIRE(pts_or_o && dirs && d_raw && workspace, "nerfw_mlp_bwd: null pointer");
  NERFW_REQUIRE(aligned16(d_raw) && aligned16(workspace), "nerfw_mlp_bwd: d_raw/workspace must be 16-byte aligned");
  if (emb) {
    NERFW_REQUIRE(w->app_w && w->app_b && grads->app_w && grads->app_b, "nerfw_mlp_bwd: embedding given but appearance parameters/gradients are null");
    NERFW_REQUIRE(emb_rows == 1 || emb_rows == n_rays, "nerfw_mlp_bwd: emb_rows=%lld must be 1 or n_rays=%lld", (long long)emb_rows, (long long)n_rays);
  } else {
    emb_rows = 0;
  }
  const size_t need = nerfw_mlp_bwd_workspace_bytes(n_rays, n_samples, emb_rows);
  if (workspace_bytes < need) {
    set_error("nerfw_mlp_bwd: workspace of %zu bytes, need %zu", workspace_bytes, need);
    return NERFW_ESIZE;
  }
  cudaStream_t st = as_stream(stream);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  float* app_off = nullptr;
  float* app_acc = nullptr;
  size_t off = 0;
  if (emb) {
    app_off = reinterpret_cast<float*>(base);
    off = ((size_t)emb_rows * 16 + 255) & ~(size_t)255;
    app_acc = reinterpret_cast<float*>(base + off);
    size_t acc_bytes = (size_t)emb_rows * ffma::APP_ROW * sizeof(float);
    NERFW_CUDA(cudaMemsetAsync(app_acc, 0, acc_bytes, st));
    off = (off + acc_bytes + 255) & ~(size_t)255;
    int rc = launch_app_offset(*w, emb, emb_rows, app_off, st);
    if (rc) return rc;
  }
  float* scratch = reinterpret_cast<float*>(base + off);

  SampleSource src;
  src.p = pts_or_o;
  src.d = dirs;
  src.z = z;
  src.emb = emb;
  src.n_per_ray = z ? n_samples : 1;
  src.emb_shared = (emb_rows == 1) ? 1 : 0;
  const int64_t total = n_rays * (z ? n_samples : 1);

  static thread_local unsigned long long attr_mask = 0;
  const size_t smem = sizeof(ffma::BwdSmem);
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(ffma::mlp_ffma_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int64_t ntiles = ceil_div64(total, ffma::TM);
  int64_t grid = ntiles < sm_count() ? ntiles : sm_count();
  ffma::mlp_ffma_bwd_kernel<<<(unsigned)grid, ffma::THREADS, smem, st>>>(*w, *grads, src, reinterpret_cast<const float4*>(app_off),
                                                                        reinterpret_cast<const float4*>(d_raw), total, scratch, app_acc);
  NERFW_LAUNCHED();
  if (emb) {
    int64_t blocks = emb_rows < 4096 ? emb_rows : 4096;
    ffma::app_bwd_kernel<<<(unsigned)blocks, 128, 0, st>>>(*w, *grads, emb, emb_rows, app_acc, d_emb);
    NERFW_LAUNCHED();
  }
  return NERFW_OK;
}
