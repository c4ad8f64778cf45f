// K6/K7 front-to-back alpha compositing and its backward: src/render.py:56-80.
// One warp per ray, samples read as coalesced float4 (r,g,b,sigma) records, transmittance by a warp-level
// product scan carried across 32-sample chunks.  HBM-bound: 24 B/sample fwd, 40 B/sample bwd.
#include "common.cuh"

namespace nerfw {

constexpr int CP_WARPS = 8;
constexpr float LAST_DELTA = 1e-3f;  // src/render.py:58
constexpr float T_EPS = 1e-10f;      // src/render.py:71,80

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive product scan over the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
// inclusive suffix-sum scan (towards lane 0)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__global__ void __launch_bounds__(CP_WARPS * 32) composite_fwd_kernel(const float4* __restrict__ raw,
                                                                      const float* __restrict__ z, int64_t B, int N,
                                                                      float* __restrict__ rgb_map,
                                                                      float* __restrict__ depth, float* __restrict__ acc,
                                                                      float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * CP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * CP_WARPS;
  for (int64_t ray = warp0; ray < B; ray += nwarps) {
    const float4* rr = raw + ray * N;
    const float* zr = z + ray * N;
    float carry = 1.0f;  // transmittance entering this chunk
    float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f, swz = 0.f;
    for (int c0 = 0; c0 < N; c0 += 32) {
      int i = c0 + lane;
      bool on = i < N;
      float4 s = on ? ld_stream4(rr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float zi = on ? __ldg(zr + i) : 0.f;
      // delta_i = z_{i+1} - z_i, last = 1e-3 (:56-58)
      float zn = __shfl_down_sync(0xffffffffu, zi, 1);
      if (lane == 31 && i + 1 < N) zn = __ldg(zr + i + 1);
      float delta = (i + 1 < N) ? (zn - zi) : LAST_DELTA;
      float alpha = on ? (1.0f - expf(-s.w * delta)) : 0.f;  // (:67)
      float t = on ? ((1.0f - alpha) + T_EPS) : 1.0f;        // (:71)
      float incl = warp_scan_mul(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      float T = carry * excl;  // (:70-73) exclusive cumprod
      float w = alpha * T;     // (:76)
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      if (on) {
        if (weights) weights[ray * N + i] = w;
        sr += w * s.x; sg += w * s.y; sb += w * s.z;  // (:79)
        sw += w; swz += w * zi;                       // (:80)
      }
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sw = warp_sum(sw); swz = warp_sum(swz);
    if (lane == 0) {
      rgb_map[ray * 3 + 0] = sr;
      rgb_map[ray * 3 + 1] = sg;
      rgb_map[ray * 3 + 2] = sb;
      depth[ray] = swz / (sw + T_EPS);
      if (acc) acc[ray] = sw;
    }
  }
}

// Backward.  With t_i = 1 - alpha_i + eps, T_i = prod_{j<i} t_j, w_i = alpha_i T_i and
// g_i = dL/dw_i = <d_rgb, rgb_i> + d_depth (z_i - depth)/(W + eps) + d_acc + d_weights_i:
//   dL/dalpha_i = g_i T_i - (sum_{k>i} g_k w_k) / t_i,   dL/dsigma_i = dL/dalpha_i * delta_i * exp(-sigma_i delta_i),
//   dL/drgb_i = w_i d_rgb.
// Forward sweep: W, depth and the transmittance entering every 32-sample chunk (lane l keeps chunk l, l+32, ...);
// reverse sweep: suffix-sum scan of g_k w_k.  The second read of the ray comes from L1/L2.
__global__ void __launch_bounds__(CP_WARPS * 32) composite_bwd_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, int64_t B, int N, const float* __restrict__ d_rgb_map,
    const float* __restrict__ d_depth, const float* __restrict__ d_acc, const float* __restrict__ d_weights,
    float4* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * CP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * CP_WARPS;
  const int nchunks = (N + 31) >> 5;
  for (int64_t ray = warp0; ray < B; ray += nwarps) {
    const float4* rr = raw + ray * N;
    const float* zr = z + ray * N;
    float entry[4] = {1.0f, 1.0f, 1.0f, 1.0f};  // N <= 4096
    float carry = 1.0f, sw = 0.f, swz = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      int i = c * 32 + lane;
      bool on = i < N;
      float sig = on ? __ldg(&rr[i].w) : 0.f;
      float zi = on ? __ldg(zr + i) : 0.f;
      float zn = __shfl_down_sync(0xffffffffu, zi, 1);
      if (lane == 31 && i + 1 < N) zn = __ldg(zr + i + 1);
      float delta = (i + 1 < N) ? (zn - zi) : LAST_DELTA;
      float alpha = on ? (1.0f - expf(-sig * delta)) : 0.f;
      float t = on ? ((1.0f - alpha) + T_EPS) : 1.0f;
      float incl = warp_scan_mul(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      float w = alpha * carry * excl;
      if ((c & 31) == lane) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((c >> 5) == k) entry[k] = carry;
      }
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      sw += w; swz += w * zi;
    }
    sw = warp_sum(sw); swz = warp_sum(swz);
    const float wden = sw + T_EPS;
    const float dep = swz / wden;
    const float gr = __ldg(d_rgb_map + ray * 3 + 0), gg = __ldg(d_rgb_map + ray * 3 + 1), gb = __ldg(d_rgb_map + ray * 3 + 2);
    const float gd = d_depth ? __ldg(d_depth + ray) / wden : 0.f;
    const float ga = d_acc ? __ldg(d_acc + ray) : 0.f;

    float suffix = 0.f;  // sum of g_k w_k over later chunks
    for (int c = nchunks - 1; c >= 0; --c) {
      int i = c * 32 + lane;
      bool on = i < N;
      float4 s = on ? ld_stream4(rr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float zi = on ? __ldg(zr + i) : 0.f;
      float zn = __shfl_down_sync(0xffffffffu, zi, 1);
      if (lane == 31 && i + 1 < N) zn = __ldg(zr + i + 1);
      float delta = (i + 1 < N) ? (zn - zi) : LAST_DELTA;
      float e = on ? expf(-s.w * delta) : 1.0f;
      float alpha = on ? (1.0f - e) : 0.f;
      float t = on ? ((1.0f - alpha) + T_EPS) : 1.0f;
      float incl = warp_scan_mul(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      float ent = entry[0];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if ((c >> 5) == k) ent = entry[k];
      float cin = __shfl_sync(0xffffffffu, ent, c & 31);
      float T = cin * excl;
      float w = alpha * T;
      float g = gr * s.x + gg * s.y + gb * s.z + gd * (zi - dep) + ga;
      if (d_weights && on) g += __ldg(d_weights + ray * N + i);
      float gw = on ? g * w : 0.f;
      float rs = warp_rscan_add(gw, lane);  // inclusive suffix within the chunk
      float after = (rs - gw) + suffix;     // sum over k > i
      suffix += __shfl_sync(0xffffffffu, rs, 0);
      if (on) {
        float dalpha = g * T - after / t;
        float dsigma = dalpha * delta * e;
        d_raw[ray * N + i] = make_float4(w * gr, w * gg, w * gb, dsigma);
      }
    }
  }
}

// Register-resident variant for N <= 32 NC (NC <= 8): the ray is read ONCE -- every lane keeps its NC float4 records,
// depths and per-sample factors in registers between the forward sweep and the reverse sweep, all loads of a ray are in
// flight together, and the depth of the next chunk's first sample comes from a shuffle instead of a second load.  Same
// arithmetic in the same order as the generic kernel above, so the results are bit-identical.
template <int NC>
__global__ void __launch_bounds__(CP_WARPS * 32) composite_bwd_reg_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, int64_t B, int N, const float* __restrict__ d_rgb_map,
    const float* __restrict__ d_depth, const float* __restrict__ d_acc, const float* __restrict__ d_weights,
    float4* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * CP_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * CP_WARPS;
  for (int64_t ray = warp0; ray < B; ray += nwarps) {
    const float4* rr = raw + ray * N;
    const float* zr = z + ray * N;
    float4 s[NC];
    float zv[NC], dw[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = c * 32 + lane;
      const bool on = i < N;
      s[c] = on ? ld_stream4(rr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      zv[c] = on ? __ldg(zr + i) : 0.f;
      dw[c] = (d_weights && on) ? __ldg(d_weights + ray * N + i) : 0.f;
    }
    const float gr = __ldg(d_rgb_map + ray * 3 + 0), gg = __ldg(d_rgb_map + ray * 3 + 1), gb = __ldg(d_rgb_map + ray * 3 + 2);
    const float gd_raw = d_depth ? __ldg(d_depth + ray) : 0.f;
    const float ga = d_acc ? __ldg(d_acc + ray) : 0.f;
    float delta[NC], e[NC], T[NC];
    float carry = 1.0f, sw = 0.f, swz = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = c * 32 + lane;
      const bool on = i < N;
      float zn = __shfl_down_sync(0xffffffffu, zv[c], 1);
      const float znext = __shfl_sync(0xffffffffu, zv[c + 1 < NC ? c + 1 : c], 0);  // first depth of the next chunk
      if (lane == 31) zn = znext;
      delta[c] = (i + 1 < N) ? (zn - zv[c]) : LAST_DELTA;
      e[c] = on ? expf(-s[c].w * delta[c]) : 1.0f;
      const float alpha = on ? (1.0f - e[c]) : 0.f;
      const float t = on ? ((1.0f - alpha) + T_EPS) : 1.0f;
      const float incl = warp_scan_mul(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      T[c] = carry * excl;
      const float w = alpha * T[c];
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      sw += w; swz += w * zv[c];
    }
    sw = warp_sum(sw); swz = warp_sum(swz);
    const float wden = sw + T_EPS;
    const float dep = swz / wden;
    const float gd = d_depth ? gd_raw / wden : 0.f;
    float suffix = 0.f;  // sum of g_k w_k over later chunks
#pragma unroll
    for (int c = NC - 1; c >= 0; --c) {
      const int i = c * 32 + lane;
      const bool on = i < N;
      const float alpha = on ? (1.0f - e[c]) : 0.f;
      const float t = on ? ((1.0f - alpha) + T_EPS) : 1.0f;
      const float w = alpha * T[c];
      float g = gr * s[c].x + gg * s[c].y + gb * s[c].z + gd * (zv[c] - dep) + ga;
      if (d_weights && on) g += dw[c];
      const float gw = on ? g * w : 0.f;
      const float rs = warp_rscan_add(gw, lane);
      const float after = (rs - gw) + suffix;
      suffix += __shfl_sync(0xffffffffu, rs, 0);
      if (on) {
        const float dalpha = g * T[c] - after / t;
        const float dsigma = dalpha * delta[c] * e[c];
        d_raw[ray * N + i] = make_float4(w * gr, w * gg, w * gb, dsigma);
      }
    }
  }
}

// Merge of the per-sample MLP outputs of the coarse samples (already evaluated in the coarse pass) and of the new fine
// samples into the depth order of the merged row that sample_pdf produced: slot of coarse i = i + #{k : zf_k < z_i}, slot
// of fine k = k + #{i : z_i <= zf_k} (coarse first on ties; tied samples sit at the same point, so their order is
// immaterial).  One warp per ray, both depth lists in shared memory, fixed-length binary searches.  Used by the opt-in
// `reuse_coarse` path of volume_render: when coarse and fine pass share one network, the fine pass only has to evaluate
// the NI new samples instead of all N + NI.
//
// BACK = true is the transpose (the backward of the merge): the gradient of the merged row is scattered back to the two
// lists -- d_fine[k] = d_merged[slot(k)], d_coarse[i] (+)= d_merged[slot(i)] (ACC: the coarse list already holds the
// gradient that arrived through the coarse pass's own outputs).  Same slot computation, so forward and backward agree.
constexpr int MG_WARPS = 4;
template <bool BACK, bool ACC>
__global__ void __launch_bounds__(MG_WARPS * 32) merge_raw_kernel(const float* __restrict__ zc, float4* __restrict__ rc,
                                                                  const float* __restrict__ zf, float4* __restrict__ rf,
                                                                  int64_t B, int N, int NI, float4* __restrict__ out) {
  // one (record, slot) pair: forward copies list -> merged row, backward merged row -> list
  auto move_c = [&](int64_t ray, int i, float4* o, int slot) {
    if (!BACK) { o[slot] = __ldg(rc + ray * N + i); return; }
    float4 g = o[slot];
    if (ACC) { const float4 a = rc[ray * N + i]; g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w; }
    rc[ray * N + i] = g;
  };
  auto move_f = [&](int64_t ray, int k, float4* o, int slot) {
    if (!BACK) o[slot] = __ldg(rf + ray * NI + k);
    else rf[ray * NI + k] = o[slot];
  };
  extern __shared__ float mg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = mg_smem + (size_t)warp * (N + NI);
  float* sf = sc + N;
  for (int64_t ray = (int64_t)blockIdx.x * MG_WARPS + warp; ray < B; ray += (int64_t)gridDim.x * MG_WARPS) {
    for (int i = lane; i < N; i += 32) sc[i] = __ldg(zc + ray * N + i);
    for (int k = lane; k < NI; k += 32) sf[k] = __ldg(zf + ray * NI + k);
    __syncwarp();
    float4* o = out + ray * (N + NI);
    // both lists are sorted when they come from the renderer; a caller-injected u_rand can leave the fine list unsorted:
    // such a ray is ranked by plain counting (stable: coarse before fine, lower index first on ties)
    bool sorted = true;
    for (int i = lane; i < N; i += 32) sorted = sorted && (i == 0 || sc[i - 1] <= sc[i]);
    for (int k = lane; k < NI; k += 32) sorted = sorted && (k == 0 || sf[k - 1] <= sf[k]);
    if (!__all_sync(0xffffffffu, sorted)) {
      for (int e = lane; e < N + NI; e += 32) {
        const bool is_c = e < N;
        const float v = is_c ? sc[e] : sf[e - N];
        int rank = 0;
        for (int j = 0; j < N + NI; ++j) {
          const float w = j < N ? sc[j] : sf[j - N];
          rank += (w < v || (w == v && j < e)) ? 1 : 0;
        }
        if (is_c) move_c(ray, e, o, rank); else move_f(ray, e - N, o, rank);
      }
      __syncwarp();
      continue;
    }
    for (int i = lane; i < N; i += 32) {          // #{k : zf_k < z_i}: lower bound in the fine list
      const float v = sc[i];
      int lo = 0, hi = NI;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sf[mid] < v) lo = mid + 1; else hi = mid; }
      move_c(ray, i, o, i + lo);
    }
    for (int k = lane; k < NI; k += 32) {         // #{i : z_i <= zf_k}: upper bound in the coarse list
      const float v = sf[k];
      int lo = 0, hi = N;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sc[mid] <= v) lo = mid + 1; else hi = mid; }
      move_f(ray, k, o, k + lo);
    }
    __syncwarp();
  }
}

}  // namespace nerfw

using namespace nerfw;

static inline unsigned comp_grid(int64_t n_rays) {
  int64_t blocks = ceil_div64(n_rays, CP_WARPS);
  int64_t cap = (int64_t)sm_count() * 8 * 4;  // 8 resident CTAs of 256 threads per SM, 4 waves
  return (unsigned)(blocks < cap ? blocks : cap);
}

extern "C" int nerfw_composite_fwd(const float* raw, const float* z, int64_t n_rays, int n_samples, float* rgb_map,
                                   float* depth, float* acc, float* weights, void* stream) {
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1, "nerfw_composite_fwd: bad shape B=%lld N=%d", (long long)n_rays, n_samples);
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(raw && z && rgb_map && depth, "nerfw_composite_fwd: null pointer");
  NERFW_REQUIRE(aligned16(raw), "nerfw_composite_fwd: raw must be 16-byte aligned");
  composite_fwd_kernel<<<comp_grid(n_rays), CP_WARPS * 32, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(raw), z, n_rays, n_samples, rgb_map, depth, acc, weights);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_composite_bwd(const float* raw, const float* z, int64_t n_rays, int n_samples,
                                   const float* d_rgb_map, const float* d_depth, const float* d_acc,
                                   const float* d_weights, float* d_raw, void* stream) {
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_samples <= 4096, "nerfw_composite_bwd: bad shape B=%lld N=%d",
                (long long)n_rays, n_samples);
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(raw && z && d_rgb_map && d_raw, "nerfw_composite_bwd: null pointer");
  NERFW_REQUIRE(aligned16(raw) && aligned16(d_raw), "nerfw_composite_bwd: raw/d_raw must be 16-byte aligned");
  const int nchunks = (n_samples + 31) / 32;
  auto launch = [&](auto kernel) {
    kernel<<<comp_grid(n_rays), CP_WARPS * 32, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(raw), z, n_rays, n_samples, d_rgb_map, d_depth, d_acc, d_weights,
        reinterpret_cast<float4*>(d_raw));
  };
  if (nchunks <= 2) launch(composite_bwd_reg_kernel<2>);        // 64 coarse samples
  else if (nchunks <= 4) launch(composite_bwd_reg_kernel<4>);
  else if (nchunks <= 6) launch(composite_bwd_reg_kernel<6>);   // 64 + 128
  else if (nchunks <= 8) launch(composite_bwd_reg_kernel<8>);
  else launch(composite_bwd_kernel);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

static int merge_launch(const char* who, const float* z_coarse, float* raw_coarse, const float* z_fine, float* raw_fine,
                        int64_t n_rays, int n_samples, int n_importance, float* merged, int back, int acc, void* stream) {
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_importance >= 1 && n_samples + n_importance <= 4096,
                "%s: bad shape B=%lld N=%d NI=%d", who, (long long)n_rays, n_samples, n_importance);
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(z_coarse && raw_coarse && z_fine && raw_fine && merged, "%s: null pointer", who);
  NERFW_REQUIRE(aligned16(raw_coarse) && aligned16(raw_fine) && aligned16(merged), "%s: record buffers must be 16-byte aligned", who);
  const size_t smem = (size_t)MG_WARPS * (n_samples + n_importance) * sizeof(float);
  int64_t blocks = ceil_div64(n_rays, MG_WARPS);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  auto launch = [&](auto kernel) -> int {
    if (smem > 48 * 1024)   // up to 64 KB at N + NI = 4096
      NERFW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)blocks, MG_WARPS * 32, smem, as_stream(stream)>>>(
        z_coarse, reinterpret_cast<float4*>(raw_coarse), z_fine, reinterpret_cast<float4*>(raw_fine), n_rays, n_samples,
        n_importance, reinterpret_cast<float4*>(merged));
    return NERFW_OK;
  };
  int rc;
  if (!back) rc = launch(merge_raw_kernel<false, false>);
  else if (acc) rc = launch(merge_raw_kernel<true, true>);
  else rc = launch(merge_raw_kernel<true, false>);
  if (rc != NERFW_OK) return rc;
  NERFW_LAUNCHED();
  return NERFW_OK;
}

extern "C" int nerfw_merge_raw(const float* z_coarse, const float* raw_coarse, const float* z_fine, const float* raw_fine,
                               int64_t n_rays, int n_samples, int n_importance, float* raw_out, void* stream) {
  return merge_launch("nerfw_merge_raw", z_coarse, const_cast<float*>(raw_coarse), z_fine, const_cast<float*>(raw_fine), n_rays,
                      n_samples, n_importance, raw_out, 0, 0, stream);
}

extern "C" int nerfw_unmerge_raw(const float* z_coarse, const float* z_fine, const float* d_raw_merged, int64_t n_rays,
                                 int n_samples, int n_importance, int accumulate_coarse, float* d_raw_coarse,
                                 float* d_raw_fine, void* stream) {
  return merge_launch("nerfw_unmerge_raw", z_coarse, d_raw_coarse, z_fine, d_raw_fine, n_rays, n_samples, n_importance,
                      const_cast<float*>(d_raw_merged), 1, accumulate_coarse ? 1 : 0, stream);
}
