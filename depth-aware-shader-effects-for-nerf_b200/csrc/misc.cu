// K3 standalone positional encoding, K9 fused Adam, MSE loss fwd+bwd, uint8 quantisation.
#include "common.cuh"

namespace nerfw {

// src/models.py:35-44: out = [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), cos(2^1 x), ...] over a last axis of `dim`
// entries; 2^l * x is exact in fp32, and sincosf (no fast-math) is accurate over the whole argument range reached here
// (|2^9 x| ~ 3e3).  One thread per (row, identity-or-level).
__global__ void __launch_bounds__(256) posenc_kernel(const float* __restrict__ x, int64_t n, int dim, int levels,
                                                     int include_input, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int per_row = levels + 1;
  if (idx >= n * per_row) return;
  int64_t row = idx / per_row;
  int l = (int)(idx - row * per_row);
  int width = dim * ((include_input ? 1 : 0) + 2 * levels);
  const float* v = x + row * dim;
  float* o = out + row * width;
  if (l == 0) {
    if (include_input)
      for (int c = 0; c < dim; ++c) o[c] = __ldg(v + c);
  } else {
    float f = (float)(1u << (l - 1));
    float* q = o + (include_input ? dim : 0) + 2 * dim * (l - 1);
    for (int c = 0; c < dim; ++c) {
      float s, cs;
      sincosf(f * __ldg(v + c), &s, &cs);
      q[c] = s;
      q[dim + c] = cs;
    }
  }
}

// torch.optim.Adam (default flags) single-tensor update, src/train.py:39,92:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float b1,
                                                   float b2, float eps, float step_size, float inv_bc2_sqrt,
                                                   float gscale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * gscale;
    float mi = m[i] + (gi - m[i]) * (1.0f - b1);   // lerp form used by torch
    float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                  float scale, float inv_n, float* __restrict__ loss,
                                                  float* __restrict__ d_a) {
  __shared__ float part[8];
  float acc = 0.f;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = a[i] - b[i];
    acc += d * d;
    if (d_a) d_a[i] = 2.0f * d * inv_n * scale;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(loss, s * inv_n);
  }
}

// (rgb * 255).astype(np.uint8): truncation toward zero after an fp32 multiply (render_aligned_spiral.py:161-162)
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ x, int64_t n, uint8_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __fmul_rn(x[i], 255.0f);
  int q = (int)v;  // numpy float->uint8 cast truncates; values are in [0,255] after the sigmoid
  out[i] = (uint8_t)(q & 0xff);
}

}  // namespace nerfw

using namespace nerfw;

extern "C" int nerfw_posenc(const float* x, int64_t n, int dim, int levels, int include_input, float* out, void* stream) {
  NERFW_REQUIRE(n >= 0 && dim >= 1 && levels >= 0 && levels <= 24, "nerfw_posenc: bad arguments n=%lld dim=%d levels=%d",
                (long long)n, dim, levels);
  NERFW_REQUIRE(include_input || levels > 0, "nerfw_posenc: empty encoding");
  if (n == 0) return NERFW_OK;
  NERFW_REQUIRE(x && out, "nerfw_posenc: null pointer");
  int64_t total = n * (levels + 1);
  posenc_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(x, n, dim, levels, include_input, out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  NERFW_REQUIRE(n >= 0 && step >= 1, "nerfw_adam: bad arguments n=%lld step=%d", (long long)n, step);
  if (n == 0) return NERFW_OK;
  NERFW_REQUIRE(param && grad && exp_avg && exp_avg_sq, "nerfw_adam: null pointer");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  float step_size = (float)((double)lr / bc1);
  float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int64_t blocks = ceil_div64(n, 256);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                              step_size, inv_bc2_sqrt, grad_scale);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_mse(const float* rgb, const float* target, int64_t n, float loss_scale, float* loss_out,
                         float* d_rgb, void* stream) {
  NERFW_REQUIRE(n > 0, "nerfw_mse: empty input");
  NERFW_REQUIRE(rgb && target && loss_out, "nerfw_mse: null pointer");
  NERFW_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), as_stream(stream)));
  int64_t blocks = ceil_div64(n, 256);
  int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  mse_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(rgb, target, n, loss_scale, 1.0f / (float)n, loss_out, d_rgb);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_quantize_u8(const float* rgb, int64_t n_values, uint8_t* out, void* stream) {
  NERFW_REQUIRE(n_values >= 0, "nerfw_quantize_u8: negative size");
  if (n_values == 0) return NERFW_OK;
  NERFW_REQUIRE(rgb && out, "nerfw_quantize_u8: null pointer");
  quantize_kernel<<<(unsigned)ceil_div64(n_values, 256), 256, 0, as_stream(stream)>>>(rgb, n_values, out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}
