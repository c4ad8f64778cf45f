// K1 ray generation, direction normalisation, K2 stratified depths, ray points.
// Every floating-point op below is written with an explicit rounding intrinsic so that nvcc cannot contract
// mul+add into FMA: the outputs of this file are compared BIT-EXACTLY with the reference's CPU PyTorch result
// (SURVEY.md section 8a-1, 8a-2).
#include "common.cuh"

namespace nerfw {

struct Cam {
  float r[3][3];
  float t[3];
};

// torch.norm / F.normalize reduce x^2+y^2+z^2 with FMA contraction on CPU (probed):
// fmaf(z,z, fmaf(y,y, x*x)).
__device__ __forceinline__ float norm3_aten(float x, float y, float z) {
  float s = __fmul_rn(x, x);
  s = __fmaf_rn(y, y, s);
  s = __fmaf_rn(z, z, s);
  return __fsqrt_rn(s);
}

// One thread per pixel.  Follows src/ray_utils.py:19-48: camera dir ((j - W/2)/f, -(i - H/2)/f, -1), rotate by
// broadcast-multiply + sum over the last axis (left-to-right), divide by the L2 norm.
__global__ void __launch_bounds__(256) raygen_kernel(int H, int W, float half_w, float half_h, float focal, Cam cam,
                                                     float* __restrict__ origins, float* __restrict__ dirs) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)H * W;
  if (p >= n) return;
  int i = (int)(p / W), j = (int)(p - (int64_t)i * W);
  float cx = __fdiv_rn(__fsub_rn((float)j, half_w), focal);
  float cy = __fdiv_rn(-__fsub_rn((float)i, half_h), focal);
  float cz = -1.0f;
  float w[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float p0 = __fmul_rn(cx, cam.r[r][0]);
    float p1 = __fmul_rn(cy, cam.r[r][1]);
    float p2 = __fmul_rn(cz, cam.r[r][2]);
    w[r] = __fadd_rn(__fadd_rn(p0, p1), p2);
  }
  float nrm = norm3_aten(w[0], w[1], w[2]);
  float* d = dirs + p * 3;
  d[0] = __fdiv_rn(w[0], nrm);
  d[1] = __fdiv_rn(w[1], nrm);
  d[2] = __fdiv_rn(w[2], nrm);
  if (origins) {
    float* o = origins + p * 3;
    o[0] = cam.t[0];
    o[1] = cam.t[1];
    o[2] = cam.t[2];
  }
}

// F.normalize(d, dim=-1): d / max(||d||, 1e-12)  (src/render.py:19)
__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float x = in[p * 3 + 0], y = in[p * 3 + 1], z = in[p * 3 + 2];
  float den = fmaxf(norm3_aten(x, y, z), 1e-12f);
  out[p * 3 + 0] = __fdiv_rn(x, den);
  out[p * 3 + 1] = __fdiv_rn(y, den);
  out[p * 3 + 2] = __fdiv_rn(z, den);
}

// One thread per (ray, sample).  src/ray_utils.py:73-86.
// perturb: mid = 0.5*(z[i+1]+z[i]); upper = [mid.., z_last]; lower = [z_0, mid..]; z = lower + (upper-lower)*t.
__global__ void __launch_bounds__(256) stratified_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                         const float* __restrict__ ztab, const float* __restrict__ trand,
                                                         int64_t B, int N, float* __restrict__ z_out,
                                                         float* __restrict__ pts) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  int64_t b = idx / N;
  int i = (int)(idx - b * N);
  float z = __ldg(ztab + i);
  if (trand) {
    float lo = z, hi = z;
    if (i > 0) lo = __fmul_rn(0.5f, __fadd_rn(z, __ldg(ztab + i - 1)));
    if (i < N - 1) hi = __fmul_rn(0.5f, __fadd_rn(__ldg(ztab + i + 1), z));
    z = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), trand[idx]));
  }
  z_out[idx] = z;
  if (pts) {
#pragma unroll
    for (int c = 0; c < 3; ++c) pts[idx * 3 + c] = __fadd_rn(__ldg(o + b * 3 + c), __fmul_rn(__ldg(d + b * 3 + c), z));
  }
}

__global__ void __launch_bounds__(256) ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                         const float* __restrict__ z, int64_t B, int N,
                                                         float* __restrict__ pts) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  int64_t b = idx / N;
  float zz = z[idx];
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[idx * 3 + c] = __fadd_rn(__ldg(o + b * 3 + c), __fmul_rn(__ldg(d + b * 3 + c), zz));
}

}  // namespace nerfw

using namespace nerfw;

extern "C" int nerfw_raygen(int height, int width, float focal, const float* c2w_host, float* origins, float* dirs,
                            void* stream) {
  NERFW_REQUIRE(height > 0 && width > 0, "nerfw_raygen: height/width must be positive (got %d x %d)", height, width);
  NERFW_REQUIRE(c2w_host && dirs, "nerfw_raygen: null c2w or dirs");
  NERFW_REQUIRE(focal != 0.0f, "nerfw_raygen: focal length is zero");
  Cam cam;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cam.r[r][c] = c2w_host[r * 4 + c];
    cam.t[r] = c2w_host[r * 4 + 3];
  }
  int64_t n = (int64_t)height * width;
  // width*0.5 and height*0.5 are exact in fp32 for any int below 2^24 (src/ray_utils.py:26-27)
  raygen_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(height, width, (float)width * 0.5f,
                                                                            (float)height * 0.5f, focal, cam, origins,
                                                                            dirs);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_normalize_dirs(const float* dirs, int64_t n_rays, float* out, void* stream) {
  NERFW_REQUIRE(n_rays >= 0, "nerfw_normalize_dirs: negative ray count");
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(dirs && out, "nerfw_normalize_dirs: null pointer");
  normalize_kernel<<<(unsigned)ceil_div64(n_rays, 256), 256, 0, as_stream(stream)>>>(dirs, n_rays, out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_stratified(const float* rays_o, const float* rays_d, const float* ztab, const float* t_rand,
                                int64_t n_rays, int n_samples, float* z, float* pts, void* stream) {
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1, "nerfw_stratified: bad shape B=%lld N=%d", (long long)n_rays, n_samples);
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(ztab && z, "nerfw_stratified: null ztab or z");
  NERFW_REQUIRE(!pts || (rays_o && rays_d), "nerfw_stratified: pts requested without rays");
  int64_t n = n_rays * n_samples;
  stratified_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, ztab, t_rand, n_rays,
                                                                                n_samples, z, pts);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_ray_points(const float* rays_o, const float* rays_d, const float* z, int64_t n_rays, int n_samples,
                                float* pts, void* stream) {
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1, "nerfw_ray_points: bad shape");
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(rays_o && rays_d && z && pts, "nerfw_ray_points: null pointer");
  int64_t n = n_rays * n_samples;
  ray_points_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, z, n_rays, n_samples, pts);
  NERFW_LAUNCHED();
  return NERFW_OK;
}
