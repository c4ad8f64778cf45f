// Shared by the MLP kernels: how a sample's position / direction / embedding row is found.
#pragma once
#include "common.cuh"

namespace nerfw {

// Either per-sample inputs (z == nullptr: pts (S,3), dirs (S,3)) or per-ray (o (B,3), unit d (B,3), z (B,N)) where
// sample s = (ray s / N, index s % N) sits at o + d*z  -- src/render.py:22-30, src/ray_utils.py:86.
struct SampleSource {
  const float* p;    // pts or rays_o
  const float* d;    // dirs (per sample or per ray)
  const float* z;    // (B,N) or nullptr
  const float* emb;  // (emb_rows,32) or nullptr
  int n_per_ray;     // N (1 when z == nullptr)
  int emb_shared;    // 1: one embedding row for everything

  __device__ __forceinline__ int64_t ray_of(int64_t s) const { return z ? s / n_per_ray : s; }
  __device__ __forceinline__ void position(int64_t s, float x[3]) const {
    if (z) {
      int64_t r = s / n_per_ray;
      float zz = __ldg(z + s);
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(__ldg(p + r * 3 + c), __fmul_rn(__ldg(d + r * 3 + c), zz));
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = __ldg(p + s * 3 + c);
    }
  }
  __device__ __forceinline__ void direction(int64_t s, float v[3]) const {
    int64_t r = ray_of(s);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __ldg(d + r * 3 + c);
  }
  __device__ __forceinline__ int64_t emb_row(int64_t s) const { return emb_shared ? 0 : ray_of(s); }
};

int launch_mlp_ffma_fwd(const NerfwWeights& w, const SampleSource& src, const float* app_off, int64_t n_total, float* raw,
                        cudaStream_t stream);

}  // namespace nerfw
