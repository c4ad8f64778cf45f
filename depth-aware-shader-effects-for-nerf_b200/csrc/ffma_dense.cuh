// fp32 CUDA-core building blocks shared by the forward (mlp_ffma.cu) and backward (mlp_bwd.cu) MLP kernels:
// a 64-row tile GEMM with the weight operand staged through shared memory in 16-deep chunks.
#pragma once
#include "common.cuh"

namespace nerfw {
namespace ffma {

constexpr int TM = 64;          // samples per tile
constexpr int THREADS = 256;    // 8 warps; warp w owns rows 8w..8w+7, lane l owns columns l + 32 j
constexpr int KC = 16;          // reduction chunk staged per step
constexpr int A_STRIDE = 320;   // activation row stride: [h(256) | enc(63) | 0]
constexpr int WS_STRIDE = 258;  // staged weight row stride (conflict-free transposed stores)

// out[s][n] = act(bias[n] + sum_r in[s][r] * Wt(r,n)),  s < 64, n < 32*NJ, r < R  (in[] readable and finite up to the
// next multiple of 16).
//   TRANSPOSED == false: Wt(r,n) = W[n*ld + r]   (forward: W is [out][in], reduction over inputs)
//   TRANSPOSED == true : Wt(r,n) = W[r*ld + n]   (backward dX: reduction over outputs)
// gsave (optional): also store the outputs to global memory, row stride gsave_stride.
template <int NJ, bool RELU, bool TRANSPOSED>
__device__ __forceinline__ void dense(const float* __restrict__ in, int in_stride, int R, const float* __restrict__ W,
                                      int ld, const float* __restrict__ bias, float* __restrict__ out, int out_stride,
                                      float (*ws)[KC * WS_STRIDE], float* __restrict__ gsave = nullptr,
                                      int gsave_stride = 0) {
  constexpr int N = 32 * NJ;
  constexpr int PER_THREAD = N * KC / THREADS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc[8][NJ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

  float pre[PER_THREAD];
  const int nchunks = (R + KC - 1) / KC;
  auto fetch = [&](int c) {
    if (!TRANSPOSED) {
      // thread: fixed r offset tid % 16, outputs tid/16 + 16 i  -> 64-byte contiguous runs along r
      const int r = c * KC + (tid & (KC - 1));
      const int n0 = tid / KC;
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) pre[i] = (r < R) ? __ldg(W + (size_t)(n0 + (THREADS / KC) * i) * ld + r) : 0.f;
    } else {
      // thread: output column(s) tid (+256 i never needed: N <= 256), r = i -> fully coalesced along n
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) {
        const int e = tid + THREADS * i;
        const int r = c * KC + e / N, n = e % N;
        pre[i] = (r < R) ? __ldg(W + (size_t)r * ld + n) : 0.f;
      }
    }
  };
  auto stage = [&](int buf) {
    if (!TRANSPOSED) {
      const int rr = tid & (KC - 1), n0 = tid / KC;
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) ws[buf][rr * WS_STRIDE + n0 + (THREADS / KC) * i] = pre[i];
    } else {
#pragma unroll
      for (int i = 0; i < PER_THREAD; ++i) {
        const int e = tid + THREADS * i;
        ws[buf][(e / N) * WS_STRIDE + (e % N)] = pre[i];
      }
    }
  };
  fetch(0);
  stage(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) fetch(c + 1);
    const float* wsb = ws[buf];
    const float* inb = in + (warp * 8) * in_stride + c * KC;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4*>(inb + i * in_stride + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float wv[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wv[j] = wsb[(k4 + kk) * WS_STRIDE + lane + 32 * j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(av, wv[j], acc[i][j]);
        }
      }
    }
    if (c + 1 < nchunks) stage(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float bj = bias ? __ldg(bias + lane + 32 * j) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[i][j] + bj;
      if (RELU) v = fmaxf(v, 0.f);
      out[(warp * 8 + i) * out_stride + lane + 32 * j] = v;
      if (gsave) gsave[(size_t)(warp * 8 + i) * gsave_stride + lane + 32 * j] = v;
    }
  }
  __syncthreads();
}

// positional encoding of one row into e[0..3+6L) (src/models.py:35-44); callers zero the padding.
__device__ __forceinline__ void encode_level(float* e, const float v[3], int l) {
  if (l == 0) {
    e[0] = v[0]; e[1] = v[1]; e[2] = v[2];
  } else {
    float f = (float)(1u << (l - 1));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(f * v[c], &sn, &cs);
      e[3 + 6 * (l - 1) + c] = sn;
      e[6 + 6 * (l - 1) + c] = cs;
    }
  }
}

}  // namespace ffma
}  // namespace nerfw
