// N3: the depth-aware parts of the reference's post-processing effects as device kernels on the renderer's
// fp32 depth buffer -- fog (src/post_processor.py:451-493), the depth-edge detector shared by the toon and hologram
// effects (bilateral filter + Sobel magnitude, :75-102 and :402-432) and the two compositing steps that consume it.
// The reference runs these with numpy / OpenCV on 8-bit PNGs read back from disk; here the depth never leaves HBM and is
// never quantised to 8 bits.  The arithmetic follows numpy's float32 evaluation order op for op (separately rounded
// operations, true divisions, truncating uint8 casts); random inputs (hologram noise, interference columns) and the
// per-row scanline table are produced by the caller, so the kernels are deterministic.
#include <math_constants.h>
#include "common.cuh"

namespace nerfw {

// ---- max over a non-negative-or-not fp32 array (depth.max(), :64-66 / :408 / :476) --------------------------------
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  // monotone mapping float -> int for signed compare: positives by int max, negatives by unsigned min
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__global__ void __launch_bounds__(256) fill_kernel(float* p, float v) { *p = v; }
__global__ void __launch_bounds__(256) max_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = -CUDART_INF_F;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, __ldg(x + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > -CUDART_INF_F) atomic_max_f32(out, m);
}

// depth_norm = depth / depth.max() if depth.max() > 1 else depth   (:64-66, :405-408, :473-477)
__device__ __forceinline__ float normalised(float d, float dmax) { return dmax > 1.0f ? __fdiv_rn(d, dmax) : d; }
__device__ __forceinline__ uint8_t to_u8(float v) {  // np.clip(v, 0, 255).astype(np.uint8): truncation
  return (uint8_t)(int)fminf(fmaxf(v, 0.0f), 255.0f);
}

// ---- fog (:451-493): f = clip(max(d - start, 0) / (1 - start), 0, 1) ** power * visibility;
//      out = clip(img * f + color * (1 - f)).astype(uint8) ----------------------------------------------------------
__global__ void __launch_bounds__(256) fog_kernel(const uint8_t* __restrict__ img, const float* __restrict__ depth,
                                                  const float* __restrict__ dmax_p, int64_t n_pix, float start,
                                                  float one_minus_start, float power, float visibility, float c0,
                                                  float c1, float c2, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const float dn = normalised(__ldg(depth + i), __ldg(dmax_p));
  float a = __fdiv_rn(fmaxf(__fsub_rn(dn, start), 0.0f), one_minus_start);
  a = fminf(fmaxf(a, 0.0f), 1.0f);
  a = (float)pow((double)a, (double)power);  // correctly rounded; np.power's vectorised powf may differ by an ulp
  const float f = __fmul_rn(a, visibility);
  const float g = __fsub_rn(1.0f, f);
  const float col[3] = {c0, c1, c2};
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[3 * i + c] = to_u8(__fadd_rn(__fmul_rn((float)img[3 * i + c], f), __fmul_rn(col[c], g)));
}

// ---- cv2.bilateralFilter(depth_norm, d, sigma_color, sigma_space) on float32 (:69): disc of radius d/2,
//      w = exp(-r^2 / (2 sigma_space^2)) * exp(-(I_q - I_p)^2 / (2 sigma_color^2)), BORDER_REFLECT_101 -----------------
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
__global__ void __launch_bounds__(256) bilateral_kernel(const float* __restrict__ depth, const float* __restrict__ dmax_p,
                                                        int H, int W, int radius, float space_coeff, float color_coeff,
                                                        float* __restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float dmax = __ldg(dmax_p);
  const float c = normalised(__ldg(depth + (size_t)y * W + x), dmax);
  float num = 0.f, den = 0.f;
  for (int dy = -radius; dy <= radius; ++dy) {
    const int yy = reflect101(y + dy, H);
    for (int dx = -radius; dx <= radius; ++dx) {
      const int r2 = dx * dx + dy * dy;
      if (r2 > radius * radius) continue;
      const float v = normalised(__ldg(depth + (size_t)yy * W + reflect101(x + dx, W)), dmax);
      const float dv = v - c;
      const float w = expf((float)r2 * space_coeff) * expf(dv * dv * color_coeff);
      num = fmaf(v, w, num);
      den += w;
    }
  }
  out[(size_t)y * W + x] = num / den;
}

// ---- cv2.Sobel(src, CV_32F, 1, 0 / 0, 1, ksize=3) and sqrt(gx^2 + gy^2) (:72-74, :414-416); the maximum of the
//      magnitude is accumulated for the normalisation that follows (:77-78, :418-419).  `normalise` != 0: src is the raw
//      depth and is normalised on the fly (hologram); 0: src is already the filtered, normalised depth (toon). ----------
__global__ void __launch_bounds__(256) sobel_kernel(const float* __restrict__ src, const float* __restrict__ dmax_p,
                                                    int normalise, int H, int W, float* __restrict__ mag,
                                                    float* __restrict__ mag_max) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  float m = 0.f;
  if (x < W && y < H) {
    const float dmax = normalise ? __ldg(dmax_p) : 0.0f;
    float v[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float s = __ldg(src + (size_t)reflect101(y + j - 1, H) * W + reflect101(x + i - 1, W));
        v[j][i] = normalise ? normalised(s, dmax) : s;
      }
    // separable form, rows first: derivative (-1, 0, 1), smoothing (1, 2, 1)
    float dx[3], sx[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dx[j] = __fsub_rn(v[j][2], v[j][0]);
      sx[j] = __fadd_rn(__fadd_rn(v[j][0], v[j][2]), __fmul_rn(v[j][1], 2.0f));
    }
    const float gx = __fadd_rn(__fadd_rn(dx[0], dx[2]), __fmul_rn(dx[1], 2.0f));
    const float gy = __fsub_rn(sx[2], sx[0]);
    m = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    mag[(size_t)y * W + x] = m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(mag_max), __float_as_int(m));
}

// ---- toon (:64-102): colours quantised to `levels`, darkened where the 3x3-dilated edge mask is set;
//      edge = (mag / mag.max() > 0.05) when mag.max() > 0 ----------------------------------------------------------
__global__ void __launch_bounds__(256) toon_kernel(const uint8_t* __restrict__ img, const float* __restrict__ mag,
                                                   const float* __restrict__ mag_max, int H, int W, float levels,
                                                   float edge_strength, float threshold, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float mm = __ldg(mag_max);
  float edge = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;  // cv2.dilate pads with the minimum
      float g = __ldg(mag + (size_t)yy * W + xx);
      if (mm > 0.f) g = __fdiv_rn(g, mm);
      if (g > threshold) edge = 1.0f;
    }
  const float keep = __fsub_rn(1.0f, __fmul_rn(edge_strength, edge));
  const size_t p = ((size_t)y * W + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // np.floor(img / 255.0 * levels) / levels * 255.0
    const float q = __fmul_rn(__fdiv_rn(floorf(__fmul_rn(__fdiv_rn((float)img[p + c], 255.0f), levels)), levels), 255.0f);
    out[p + c] = to_u8(__fmul_rn(q, keep));
  }
}

// ---- hologram (:373-449): tint (0.8, 1.0, 0.2), per-row scanline factor, depth-edge glow (0.1, 0.6, 0.3) * mag/max,
//      additive noise, x1.5 per interference line covering the column; clip(x * 255).astype(uint8) -----------------------------------
__global__ void __launch_bounds__(256) hologram_kernel(const uint8_t* __restrict__ img, const float* __restrict__ mag,
                                                       const float* __restrict__ mag_max,
                                                       const float* __restrict__ row_scale,
                                                       const int* __restrict__ col_hits,
                                                       const float* __restrict__ noise, int H, int W,
                                                       uint8_t* __restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t pix = (size_t)y * W + x;
  float e = 0.f;
  if (mag) {
    const float mm = __ldg(mag_max);
    e = __ldg(mag + pix);
    if (mm > 0.f) e = __fdiv_rn(e, mm);
  }
  const float tint[3] = {0.8f, 1.0f, 0.2f}, glow[3] = {0.1f, 0.6f, 0.3f};
  const float rs = __ldg(row_scale + y);
  const int hits = col_hits ? __ldg(col_hits + x) : 0;   // interference lines covering this column (:443-447)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = __fmul_rn(__fmul_rn(__fdiv_rn((float)img[3 * pix + c], 255.0f), tint[c]), rs);
    v = __fadd_rn(v, mag ? __fmul_rn(e, glow[c]) : 0.0f);
    v = __fadd_rn(v, noise ? __ldg(noise + 3 * pix + c) : 0.0f);
    for (int k = 0; k < hits; ++k) v = __fmul_rn(v, 1.5f);
    out[3 * pix + c] = to_u8(__fmul_rn(v, 255.0f));
  }
}

}  // namespace nerfw

using namespace nerfw;

static dim3 grid2d(int H, int W) { return dim3((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8)); }

extern "C" int nerfw_max_f32(const float* x, int64_t n, float* out, void* stream) {
  NERFW_REQUIRE(n >= 1 && x && out, "nerfw_max_f32: need n >= 1 and non-null pointers");
  fill_kernel<<<1, 1, 0, as_stream(stream)>>>(out, -INFINITY);
  int64_t blocks = ceil_div64(n, 256 * 8);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  max_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_fog(const uint8_t* image, const float* depth, const float* depth_max, int64_t n_pixels,
                         float fog_start, float power, float visibility, const float* fog_color_host, uint8_t* out,
                         void* stream) {
  NERFW_REQUIRE(n_pixels >= 0, "nerfw_fog: negative pixel count");
  if (n_pixels == 0) return NERFW_OK;
  NERFW_REQUIRE(image && depth && depth_max && fog_color_host && out, "nerfw_fog: null pointer");
  NERFW_REQUIRE(fog_start < 1.0f, "nerfw_fog: fog_start must be < 1 (got %g)", (double)fog_start);
  fog_kernel<<<(unsigned)ceil_div64(n_pixels, 256), 256, 0, as_stream(stream)>>>(
      image, depth, depth_max, n_pixels, fog_start, (float)(1.0 - (double)fog_start), power, visibility, fog_color_host[0],
      fog_color_host[1], fog_color_host[2], out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_depth_edges(const float* depth, const float* depth_max, int height, int width, int bilateral_d,
                                 float sigma_color, float sigma_space, float* filtered, float* mag, float* mag_max,
                                 void* stream) {
  NERFW_REQUIRE(height >= 1 && width >= 1, "nerfw_depth_edges: empty image");
  NERFW_REQUIRE(depth && depth_max && mag && mag_max, "nerfw_depth_edges: null pointer");
  NERFW_REQUIRE(bilateral_d == 0 || (bilateral_d >= 3 && bilateral_d <= 31 && filtered),
                "nerfw_depth_edges: bilateral_d must be 0 (off) or in [3,31] with a `filtered` buffer");
  fill_kernel<<<1, 1, 0, as_stream(stream)>>>(mag_max, 0.0f);
  if (bilateral_d) {
    NERFW_REQUIRE(sigma_color > 0.f && sigma_space > 0.f, "nerfw_depth_edges: sigmas must be positive");
    bilateral_kernel<<<grid2d(height, width), 256, 0, as_stream(stream)>>>(
        depth, depth_max, height, width, bilateral_d / 2, -0.5f / (sigma_space * sigma_space),
        -0.5f / (sigma_color * sigma_color), filtered);
    sobel_kernel<<<grid2d(height, width), 256, 0, as_stream(stream)>>>(filtered, depth_max, 0, height, width, mag, mag_max);
  } else {
    sobel_kernel<<<grid2d(height, width), 256, 0, as_stream(stream)>>>(depth, depth_max, 1, height, width, mag, mag_max);
  }
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_toon(const uint8_t* image, const float* mag, const float* mag_max, int height, int width,
                          int levels, float edge_strength, uint8_t* out, void* stream) {
  NERFW_REQUIRE(height >= 1 && width >= 1 && levels >= 1, "nerfw_toon: need a non-empty image and levels >= 1");
  NERFW_REQUIRE(image && mag && mag_max && out, "nerfw_toon: null pointer");
  toon_kernel<<<grid2d(height, width), 256, 0, as_stream(stream)>>>(image, mag, mag_max, height, width, (float)levels,
                                                                   edge_strength, 0.05f, out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_hologram(const uint8_t* image, const float* mag, const float* mag_max, const float* row_scale,
                              const int* col_hits, const float* noise, int height, int width, uint8_t* out,
                              void* stream) {
  NERFW_REQUIRE(height >= 1 && width >= 1, "nerfw_hologram: empty image");
  NERFW_REQUIRE(image && row_scale && out, "nerfw_hologram: null pointer");
  NERFW_REQUIRE((mag == nullptr) == (mag_max == nullptr), "nerfw_hologram: mag and mag_max go together");
  hologram_kernel<<<grid2d(height, width), 256, 0, as_stream(stream)>>>(image, mag, mag_max, row_scale, col_hits, noise,
                                                                       height, width, out);
  NERFW_LAUNCHED();
  return NERFW_OK;
}
