// K4 (tensor-core modes): the NeRF-W MLP forward as one persistent, warp-specialised tcgen05 kernel.
// src/models.py:105-162, fused with ray-point generation (src/ray_utils.py:86) and positional encoding
// (src/models.py:35-44).
//
// Per CTA (one per SM) and per tile of 128 samples:
//   * 16 epilogue warps compute the encodings into shared memory (K-major, 128B-swizzled UMMA operand tiles);
//   * one producer thread streams the pre-swizzled bf16 weight chunks (32 KB = 256 outputs x 64 inputs) from L2 into a
//     4-stage shared-memory ring with cp.async.bulk (TMA engine) + mbarrier transaction counts;
//   * one MMA thread issues tcgen05.mma kind::f16 (M=128, N=256|128, K=16) with the accumulator in TMEM columns
//     [0,256).  Layer inputs that are encodings come from shared memory (SS form); hidden activations never leave
//     tensor memory: the epilogue warps read the fp32 accumulator (tcgen05.ld), add bias, ReLU, round to bf16 and
//     write the next layer's A operand straight back into TMEM columns [256,384) (tcgen05.st; TS form);
//   * the 1-wide density head and the 3-wide rgb head are fp32 dot products on CUDA cores inside the epilogues.
//
// NERFW_MLP_BF16X3 ("fp32 parity" mode): every operand is split x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and
// each product is three MMAs  A_hi W_hi + A_lo W_hi + A_hi W_lo  (dropped term ~2^-18): ~2^-16 relative error per
// product against 2^-9 for plain bf16 and 2^-11 for TF32.  A_lo lives in TMEM columns [384,512); in the single-MMA
// modes (bf16, fp16) those columns hold the direction layer's accumulator instead, so that the next tile's layer 0 can
// be issued behind the direction layer (see DIR_ACC below).
// Every wait is an mbarrier.try_wait with a suspend-time hint (umma.cuh): a waiting warp sleeps in hardware.  Plain
// polling loops in the 16 epilogue warps cost the tensor pipe 10 % of its cycles.
#include <cuda_fp16.h>
#include "common.cuh"
#include "mlp_common.cuh"
#include "mlp_tc.cuh"
#include "umma.cuh"
#include "mlp_tc_layout.cuh"
#include <stdlib.h>

namespace nerfw {
namespace tc {

// ---------------------------------------------------------------------------------------------------------------
// weight packing: state_dict fp32 [out,in] -> bf16 hi/lo chunks in the exact shared-memory image (128B swizzle)
__global__ void __launch_bounds__(256) pack_weights_kernel(NerfwWeights w, uint8_t* __restrict__ packed,
                                                           uint8_t* __restrict__ packed_f16) {
  // one thread per (chunk, output row, 8-wide k group)
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t big_items = (int64_t)N_BIG * 256 * 8;
  const int64_t total = big_items + (int64_t)N_SMALL * 128 * 8;
  if (idx < total) {
    int chunk, n, g;
    if (idx < big_items) {
      chunk = (int)(idx / (256 * 8));
      int r = (int)(idx % (256 * 8));
      n = r / 8; g = r % 8;
    } else {
      int64_t j = idx - big_items;
      chunk = N_BIG + (int)(j / (128 * 8));
      int r = (int)(j % (128 * 8));
      n = r / 8; g = r % 8;
    }
    ChunkSrc cs = chunk_source(chunk);
    const float* W;
    int K;
    if (cs.layer == 8) { W = w.dir_w; K = 256 + NERFW_DIR_DIM; }
    else { W = w.pts_w[cs.layer]; K = cs.layer == 0 ? NERFW_POS_DIM : (cs.layer == NERFW_SKIP ? 256 + NERFW_POS_DIM : 256); }
    __align__(16) __nv_bfloat16 hi[8], lo[8];
    __align__(16) __half hf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int col = cs.col0 + g * 8 + e;
      if (cs.layer >= 1 && cs.col0 < 256) col = kperm_feature(cs.col0 >> 6, g * 8 + e);  // hidden-feature K blocks
      float v = col < K ? __ldg(W + (size_t)n * K + col) : 0.f;
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[e] = h;
      lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
      hf[e] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
    }
    const uint32_t sz = chunk < N_BIG ? BIG_CHUNK : SMALL_CHUNK;
    uint8_t* base = packed + chunk_offset(chunk);
    uint32_t off = sw128_offset((uint32_t)n, (uint32_t)g * 8);
    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + sz + off) = *reinterpret_cast<const uint4*>(lo);
    if (packed_f16) *reinterpret_cast<uint4*>(packed_f16 + chunk_offset_f16(chunk) + off) = *reinterpret_cast<const uint4*>(hf);
  }
  // vector block
  float* vec = reinterpret_cast<float*>(packed + W_BYTES);
  for (int64_t v = idx; v < V_FLOATS; v += (int64_t)gridDim.x * blockDim.x) {
    float x = 0.f;
    if (v < V_DIRB) x = __ldg(w.pts_b[v / 256] + (v % 256));
    else if (v < V_DENW) x = __ldg(w.dir_b + (v - V_DIRB));
    else if (v < V_RGBW) x = __ldg(w.density_w + (v - V_DENW));
    else if (v < V_DENB) x = __ldg(w.rgb_w + (v - V_RGBW));
    else if (v == V_DENB) x = __ldg(w.density_b);
    else if (v < V_RGBB + 3) x = __ldg(w.rgb_b + (v - V_RGBB));
    vec[v] = x;
  }
}

// appearance: off[row] = W_rgb (W_app e_row + b_app)  (src/models.py:146-160; the projection is added after the ReLU of
// the direction layer, so its effect on the rgb logits is this per-embedding 3-vector).  One 128-thread CTA per row:
// thread k forms a_k (32 MACs), then three warps reduce a against the three rows of W_rgb.
__global__ void __launch_bounds__(128) app_offset_kernel(NerfwWeights w, const float* __restrict__ emb, int64_t rows,
                                                         float4* __restrict__ out) {
  __shared__ float a_s[NERFW_DIR_HIDDEN];
  __shared__ float e_s[NERFW_APP_DIM];
  __shared__ float o_s[3];
  const int k = threadIdx.x, lane = k & 31, warp = k >> 5;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    if (k < NERFW_APP_DIM) e_s[k] = __ldg(emb + row * NERFW_APP_DIM + k);
    __syncthreads();
    float a = __ldg(w.app_b + k);
#pragma unroll 8
    for (int q = 0; q < NERFW_APP_DIM; ++q) a = fmaf(__ldg(w.app_w + k * NERFW_APP_DIM + q), e_s[q], a);
    a_s[k] = a;
    __syncthreads();
    if (warp < 3) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) p = fmaf(__ldg(w.rgb_w + warp * NERFW_DIR_HIDDEN + lane + 32 * j), a_s[lane + 32 * j], p);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (lane == 0) o_s[warp] = p;
    }
    __syncthreads();
    if (k == 0) out[row] = make_float4(o_s[0], o_s[1], o_s[2], 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward kernel warp roles: 16 epilogue warps (lane quadrant = warp % 4, column quarter = warp / 4: four resident
// warps per scheduler hide the TMEM / shared-memory latencies of the epilogue), one producer warp, one MMA warp.
constexpr int F_EPI_WARPS = 16;
constexpr int F_PRODUCER_WARP = 16;
constexpr int F_MMA_WARP = 17;
constexpr int F_THREADS = 576;
constexpr int F_EPI_THREADS = F_EPI_WARPS * 32;

// profiling: event stamps (clock64) of CTA 0, third tile: slot -> time.  Only when a timeline buffer is passed.
#define NERFW_STAMP(slot) do { if (timeline && tile == stamp_tile) timeline[slot] = clock64(); } while (0)

// X3: split (three-MMA) arithmetic; F16 (with X3 = false): operands in fp16 instead of bf16 (11-bit significand, activations
// saturate at 65504) from the fp16 weight image at packed + f16_offset.
template <bool X3, bool F16 = false, bool SIGMA = false>
__global__ void __launch_bounds__(F_THREADS, 1) mlp_tc_fwd_kernel(const uint8_t* __restrict__ packed, SampleSource src,
                                                                 const float4* __restrict__ app_off, int64_t n_total,
                                                                 float4* __restrict__ raw, uint32_t* __restrict__ masks, int flags,
                                                                 long long* __restrict__ timeline, size_t f16_offset) {
  static_assert(!(X3 && F16), "the fp16 mode is single pass");
  // flags: 1 = profiling, reuse whatever the ring holds after the first tile (wrong results).  SIGMA (sigma only): the
  // direction layer and the rgb head are skipped and raw = (0, 0, 0, sigma) -- all the coarse pass of a hierarchical
  // inference render needs (its weights place the fine samples; its colour is never looked at)
  const int debug_skip_weights = flags & 1;
  constexpr bool sigma_only = SIGMA;
  constexpr int n_chunks = sigma_only ? N_BIG : N_CHUNKS;
  // Hand-off scheduling (the tensor pipe idles between the last MMA of a layer and the first of the next; these shorten it):
  //  * the K blocks whose A operand is an encoding tile in shared memory (skip part of layer 4, direction part of the
  //    direction layer) are issued FIRST in their layer -- they need the accumulator but no epilogue output, so they run
  //    while the previous layer's epilogue is still producing its first operand granule;
  //  * DIR_ACC (single-MMA modes): the direction layer accumulates into TMEM columns [384,512) (the A_lo region, unused
  //    without the split), so layer 0 of the NEXT tile is issued right behind it and the epilogue warps finish that layer 0
  //    -- i.e. restart the tensor pipe on layer 1 -- BEFORE they turn to the direction-layer accumulator and the rgb head of
  //    the previous tile;
  //  * split mode (no free TMEM columns): the direction accumulator is reduced to the three rgb partial sums at once, and the
  //    rest of that epilogue (exchange, sigmoid, output store) waits until the next tile's layer 0 has been handed on.
  constexpr bool DIR_ACC = !X3 && !SIGMA;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + SM_BAR);
  // bf16 mode has no lo operand tiles: that space is a fifth ring stage (deeper weight prefetch across the epilogue gaps)
  constexpr int STAGES = X3 ? NSTAGES : NSTAGES + 1;
  constexpr uint32_t RING = X3 ? SM_RING : SM_RING - BIG_CHUNK;
  constexpr uint32_t PED_HI = X3 ? SM_PED_HI : SM_PEX_LO;
  struct RingPipe {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  };
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  // per-K-block operand barriers, accumulator-free and encodings-ready barriers (one arrival per epilogue warp)
  uint64_t* a_kb = acc_full + 2;
  uint64_t* acc_free = a_kb + 4;
  uint64_t* pe_ready = acc_free + 1;
  uint64_t* acc2_full = acc_full + 1;   // direction-layer accumulator (DIR_ACC)
  uint64_t* acc2_free = pe_ready + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + SM_TMEMPTR);
  float* vec = reinterpret_cast<float*>(sm + SM_VEC);
  float* sig_part = reinterpret_cast<float*>(sm + SM_SIG);
  float4* rgb_part = reinterpret_cast<float4*>(sm + SM_RGB);

  if (warp == F_PRODUCER_WARP && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, F_EPI_WARPS);
    for (int i = 0; i < 4; ++i) mbar_init(&a_kb[i], F_EPI_WARPS);
    mbar_init(acc_free, F_EPI_WARPS);
    mbar_init(pe_ready, F_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == F_MMA_WARP) tmem_alloc<512>(tmem_ptr);
  if (warp < F_EPI_WARPS) {
    const float* gv = reinterpret_cast<const float*>(packed + W_BYTES);
    for (int i = tid; i < V_FLOATS; i += F_EPI_THREADS) vec[i] = __ldg(gv + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int64_t ntiles = (n_total + TM - 1) / TM;
  // profiling: SM cycles and nanoseconds of CTA 0 over the whole launch (average SM clock, cycles per tile)
  if (timeline && blockIdx.x == 0 && tid == 0) { timeline[120] = clock64(); timeline[121] = (long long)global_timer_ns(); }
  const int64_t stamp_tile = blockIdx.x == 0 ? 2 * (int64_t)gridDim.x : -1;   // profiling: CTA 0, third tile
  // Split (3-MMA) arithmetic for the direction layer only in training mode: its ReLU gates then match the fp32 forward.
  // For inference the layer runs as a single bf16 MMA -- it only feeds the rgb sigmoid (sigma, hence depth, acc and the
  // resampling, come from the trunk); measured effect on rendered rgb <= 5e-5.
  const bool dir_split = X3 && masks != nullptr;

  if (warp == F_PRODUCER_WARP) {
    // ===================== weight producer =====================
    if (lane == 0) {
      RingPipe p;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int j = 0; j < n_chunks; ++j) {
          // consumption order of the MMA thread: the encoding chunk of layer 4 / of the direction layer leads its layer
          int i = j;
          if (j >= 13 && j < 18) i = j == 13 ? 17 : j - 1;
          if (j >= N_BIG) i = j == N_BIG ? N_CHUNKS - 1 : j - 1;
          const uint32_t sz = i < N_BIG ? BIG_CHUNK : SMALL_CHUNK;
          if (!sigma_only && !dir_split && j > N_BIG) {
            // unsplit direction layer: its four 16 KB hidden K blocks travel two per ring stage (one full / empty hand-shake
            // per eight N = 128 MMAs: at four, the hand-shake, not the tensor pipe, sets the pace -- umma_rate mode 6)
            if ((j - N_BIG) & 1) {   // j = N_BIG + 1, N_BIG + 3: chunks (30, 31), (32, 33)
              mbar_wait(&empty[p.stage], p.phase ^ 1);
              if (debug_skip_weights && tile != (int64_t)blockIdx.x) {
                mbar_arrive(&full[p.stage]);
              } else {
                mbar_arrive_expect_tx(&full[p.stage], 2 * SMALL_CHUNK);
                for (int q = 0; q < 2; ++q) {
                  const uint8_t* srcw = F16 ? packed + f16_offset + chunk_offset_f16(i + q) : packed + chunk_offset(i + q);
                  bulk_g2s(sm + RING + p.stage * BIG_CHUNK + q * SMALL_CHUNK, srcw, SMALL_CHUNK, &full[p.stage]);
                }
              }
              p.advance();
            }
            continue;
          }
          const int copies = (X3 && (i < N_BIG || dir_split)) ? 2 : 1;
          for (int v = 0; v < copies; ++v) {
            mbar_wait(&empty[p.stage], p.phase ^ 1);
            if (debug_skip_weights && tile != (int64_t)blockIdx.x) {
              mbar_arrive(&full[p.stage]);   // profiling only: reuse whatever the stage holds (results are wrong)
            } else {
              mbar_arrive_expect_tx(&full[p.stage], sz);
              const uint8_t* srcw = F16 ? packed + f16_offset + chunk_offset_f16(i) : packed + chunk_offset(i) + (size_t)v * sz;
              bulk_g2s(sm + RING + p.stage * BIG_CHUNK, srcw, sz, &full[p.stage]);
            }
            p.advance();
          }
        }
      }
    }
  } else if (warp == F_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      RingPipe p;
      const uint32_t idesc256 = F16 ? idesc_f16(128, 256) : idesc_bf16(128, 256);
      const uint32_t idesc128 = F16 ? idesc_f16(128, 128) : idesc_bf16(128, 128);
      const uint32_t ring = smem_u32(sm + RING);
      const uint32_t d_acc = tmem + COL_ACC;
      const uint32_t d_dir = DIR_ACC ? tmem + COL_ALO : d_acc;
      // one 64-wide K block: A (hi[,lo]) x W chunk (hi[,lo]); a_* are either TMEM addresses (TS) or smem descs (SS)
      auto kblock = [&](uint32_t d, bool from_tmem, uint64_t a_hi, uint64_t a_lo, uint32_t idesc, int ksteps, bool first, bool split = X3) {
        mbar_wait(&full[p.stage], p.phase);
        tc_fence_after();
        uint64_t b = smem_desc_sw128(ring + p.stage * BIG_CHUNK);
        for (int k = 0; k < ksteps; ++k) {
          uint32_t accf = (first && k == 0) ? 0u : 1u;
          if (from_tmem) mma_ts(d, (uint32_t)a_hi + 8 * k, b + 2 * k, idesc, accf);
          else mma_ss(d, a_hi + 2 * k, b + 2 * k, idesc, accf);
        }
        if (split) {
          for (int k = 0; k < ksteps; ++k) {
            if (from_tmem) mma_ts(d, (uint32_t)a_lo + 8 * k, b + 2 * k, idesc, 1u);
            else mma_ss(d, a_lo + 2 * k, b + 2 * k, idesc, 1u);
          }
        }
        mma_commit(&empty[p.stage]);
        p.advance();
        if (split) {
          mbar_wait(&full[p.stage], p.phase);
          tc_fence_after();
          uint64_t bl = smem_desc_sw128(ring + p.stage * BIG_CHUNK);
          for (int k = 0; k < ksteps; ++k) {
            if (from_tmem) mma_ts(d, (uint32_t)a_hi + 8 * k, bl + 2 * k, idesc, 1u);
            else mma_ss(d, a_hi + 2 * k, bl + 2 * k, idesc, 1u);
          }
          mma_commit(&empty[p.stage]);
          p.advance();
        }
      };
      const uint64_t pex_hi = smem_desc_sw128(smem_u32(sm + SM_PEX_HI)), pex_lo = smem_desc_sw128(smem_u32(sm + SM_PEX_LO));
      const uint64_t ped_hi = smem_desc_sw128(smem_u32(sm + PED_HI)), ped_lo = smem_desc_sw128(smem_u32(sm + SM_PED_LO));
      // The epilogue frees the accumulator as soon as it sits in registers and publishes the next A operand one 64-wide
      // K block at a time, so the MMAs of layer l+1 start while most of epilogue l is still running.
      uint32_t ph_free = 0, ph_pe = 0, ph_kb = 0, ph_free2 = 0;
      auto wait_bar = [&](uint64_t* bar, uint32_t phase) {
        mbar_wait(bar, phase);
        tc_fence_after();
      };
      int64_t tile = blockIdx.x;
      // layer 0 of `tile`: position encodings (shared memory) x chunk 0
      auto layer0 = [&]() {
        wait_bar(pe_ready, ph_pe);
        ph_pe ^= 1;
        NERFW_STAMP(0);
        wait_bar(acc_free, ph_free);
        ph_free ^= 1;
        NERFW_STAMP(10);                     // accumulator free seen by the MMA thread
        kblock(d_acc, false, pex_hi, pex_lo, idesc256, 4, true);
        mma_commit(acc_full);
        NERFW_STAMP(13);                     // all MMAs of the layer issued
      };
      for (; tile < ntiles; tile += gridDim.x) {
        // DIR_ACC: layer 0 of every tile but the first was issued behind the previous tile's direction layer
        if (!DIR_ACC || tile == (int64_t)blockIdx.x) layer0();
        for (int layer = 1; layer < NERFW_LAYERS; ++layer) {
          wait_bar(acc_free, ph_free);
          ph_free ^= 1;
          NERFW_STAMP(10 + layer * 8);
          if (layer == NERFW_SKIP) kblock(d_acc, false, pex_hi, pex_lo, idesc256, 4, true);
          for (int kb = 0; kb < 4; ++kb) {
            wait_bar(&a_kb[kb], ph_kb);
            if (kb == 0) NERFW_STAMP(11 + layer * 8);   // first operand K block seen
            if (kb == 3) NERFW_STAMP(12 + layer * 8);   // last operand K block seen
            kblock(d_acc, true, tmem + COL_AHI + 32 * kb, tmem + COL_ALO + 32 * kb, idesc256, 4, kb == 0 && layer != NERFW_SKIP);
          }
          ph_kb ^= 1;
          mma_commit(acc_full);
          NERFW_STAMP(13 + layer * 8);
        }
        if (sigma_only) continue;
        // direction layer (N = 128): [h (256) | direction encoding (27 -> 32)]
        if (DIR_ACC) {
          wait_bar(acc2_free, ph_free2);
          ph_free2 ^= 1;
        } else {
          wait_bar(acc_free, ph_free);
          ph_free ^= 1;
        }
        kblock(d_dir, false, ped_hi, ped_lo, idesc128, 2, true, dir_split);
        if (dir_split) {
          for (int kb = 0; kb < 4; ++kb) {
            wait_bar(&a_kb[kb], ph_kb);
            kblock(d_dir, true, tmem + COL_AHI + 32 * kb, tmem + COL_ALO + 32 * kb, idesc128, 4, false, true);
          }
        } else {
          // two hidden K blocks per ring stage (see the producer)
          for (int pr = 0; pr < 2; ++pr) {
            wait_bar(&a_kb[2 * pr], ph_kb);
            mbar_wait(&full[p.stage], p.phase);
            tc_fence_after();
            const uint64_t b0 = smem_desc_sw128(ring + p.stage * BIG_CHUNK);
            for (int k = 0; k < 4; ++k) mma_ts(d_dir, tmem + COL_AHI + 32 * (2 * pr) + 8 * k, b0 + 2 * k, idesc128, 1u);
            wait_bar(&a_kb[2 * pr + 1], ph_kb);
            const uint64_t b1 = smem_desc_sw128(ring + p.stage * BIG_CHUNK + SMALL_CHUNK);
            for (int k = 0; k < 4; ++k) mma_ts(d_dir, tmem + COL_AHI + 32 * (2 * pr + 1) + 8 * k, b1 + 2 * k, idesc128, 1u);
            mma_commit(&empty[p.stage]);
            p.advance();
          }
        }
        ph_kb ^= 1;
        mma_commit(DIR_ACC ? acc2_full : acc_full);
        NERFW_STAMP(94);                     // direction layer issued
        if (DIR_ACC && tile + gridDim.x < ntiles) {
          tile += gridDim.x;                 // (stamps of layer 0 belong to the tile it is part of)
          layer0();
          tile -= gridDim.x;
        }
      }
    }
  } else {
    // ===================== encoders + epilogues (16 warps, thread <-> sample row x column quarter) ==========
    const uint32_t quad = warp & 3, cq = warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t tlane = tmem + ((quad * 32) << 16);
    uint32_t acc_phase = 0;
    uint8_t* pex_hi = sm + SM_PEX_HI;
    uint8_t* pex_lo = sm + SM_PEX_LO;
    uint8_t* ped_hi = sm + PED_HI;
    mbar_arrive_warp(acc_free);   // the accumulator starts out free
    if (DIR_ACC) mbar_arrive_warp(acc2_free);
    // ---- encodings (src/models.py:35-44).  Column quarters 0 / 1 write position features 0..31 / 32..63 of a tile,
    // quarter 2 its direction tile.  They are not on the MMA critical path: the position tile of the NEXT sample tile and
    // the direction tile of the current one are filled in between two trunk epilogues, once the skip layer has consumed
    // the position tile, while the tensor pipe works on layer 6.
    // `pre`: the sample's position / direction fetched ahead of time (top of the tile loop), or nullptr to fetch here
    auto encode_pos = [&](int64_t t, const float* pre) {
      if (cq < 2) {
        const int64_t sr = t * TM + row;
        float x[3] = {0.f, 0.f, 0.f};
        if (pre) { x[0] = pre[0]; x[1] = pre[1]; x[2] = pre[2]; }
        else if (sr < n_total) src.position(sr, x);
        float v[32];
        if (cq == 0) {
          pos_features32<0, !X3>(x, v);
          store_features32<X3, F16>(pex_hi, pex_lo, row, 0, v);
        } else {
          pos_features32<1, !X3>(x, v);
          store_features32<X3, F16>(pex_hi, pex_lo, row, 32, v);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(pe_ready);
    };
    auto encode_dir = [&](int64_t t, const float* pre) {
      if (cq == 2) {
        const int64_t sr = t * TM + row;
        float d[3] = {0.f, 0.f, 0.f};
        if (pre) { d[0] = pre[0]; d[1] = pre[1]; d[2] = pre[2]; }
        else if (sr < n_total) src.direction(sr, d);
        float v[32];
        dir_features32<!X3>(d, v);
        if (dir_split) store_features32<true>(ped_hi, sm + SM_PED_LO, row, 0, v);
        else store_features32<false, F16>(ped_hi, ped_hi, row, 0, v);
        fence_proxy_async_smem();   // ordered before this warp's later a_kb arrivals, which the MMA thread waits on
      }
    };
    // ---- direction-layer epilogue + rgb head of one tile (src/models.py:141-160), in two parts.  Part A pulls the accumulator
    // and reduces it to this thread's three rgb partial sums; part B exchanges them, applies the sigmoid and stores the
    // sample.  Part B of tile t always runs inside tile t+1, after its layer-0 epilogue, i.e. with the tensor pipe already
    // restarted on layer 1; DIR_ACC (own accumulator and barriers) moves part A there as well.
    uint32_t acc2_phase = 0;
    float p3[3] = {0.f, 0.f, 0.f};
    auto dir_part_a = [&](int64_t tile) {
      if (DIR_ACC) {
        mbar_wait(acc2_full, acc2_phase);
        acc2_phase ^= 1;
      } else {
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1;
      }
      tc_fence_after();
      if (tid == 0) NERFW_STAMP(91);   // direction-layer accumulator complete seen
      p3[0] = p3[1] = p3[2] = 0.f;
      const uint32_t col = cq * 32;
      uint32_t r[32];
      tmem_ld32(tlane + (DIR_ACC ? COL_ALO : COL_ACC) + col, r);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_warp(DIR_ACC ? acc2_free : acc_free);   // the next direction layer / the next tile's layer 0 may start
      if (tid == 0) NERFW_STAMP(92);
      uint32_t bits = 0;
      // bias and rgb-head rows as 16-byte shared-memory loads (a quarter of the wavefronts of scalar loads: this epilogue
      // runs next to MMAs of the following tile, which read their B operand through the same pipe)
      const float4* b4 = reinterpret_cast<const float4*>(vec + V_DIRB + col);
      const float4* w4 = reinterpret_cast<const float4*>(vec + V_RGBW + col);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 bb = b4[j4], wr = w4[j4], wg = w4[32 + j4], wb = w4[64 + j4];
        const float bj[4] = {bb.x, bb.y, bb.z, bb.w};
        const float wj[3][4] = {{wr.x, wr.y, wr.z, wr.w}, {wg.x, wg.y, wg.z, wg.w}, {wb.x, wb.y, wb.z, wb.w}};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * j4 + e;
          float hv = fmaxf(__uint_as_float(r[j]) + bj[e], 0.f);
          bits |= (hv > 0.f ? 1u : 0u) << j;
#pragma unroll
          for (int c = 0; c < 3; ++c) p3[c] = fmaf(hv, wj[c][e], p3[c]);
        }
      }
      if (masks) masks[mask_index(tile, NERFW_LAYERS, row, cq >> 1, (int)(cq & 1))] = bits;
    };
    auto dir_part_b = [&](int64_t tile) {
      const int64_t s = tile * TM + row;
      const bool live = s < n_total;
      if (cq != 0) rgb_part[cq * TM + row] = make_float4(p3[0], p3[1], p3[2], 0.f);
      named_bar_sync(1, F_EPI_THREADS);
      if (cq == 0 && live) {
        const float4 o1 = rgb_part[TM + row], o2 = rgb_part[2 * TM + row], o3 = rgb_part[3 * TM + row];
        float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
        if (app_off) off = __ldg(app_off + src.emb_row(s));
        float sg = (sig_part[row] + sig_part[TM + row]) + (sig_part[2 * TM + row] + sig_part[3 * TM + row]) + vec[V_DENB];
        float4 o;
        o.x = 1.0f / (1.0f + expf(-(p3[0] + o1.x + o2.x + o3.x + vec[V_RGBB + 0] + off.x)));
        o.y = 1.0f / (1.0f + expf(-(p3[1] + o1.y + o2.y + o3.y + vec[V_RGBB + 1] + off.y)));
        o.z = 1.0f / (1.0f + expf(-(p3[2] + o1.z + o2.z + o3.z + vec[V_RGBB + 2] + off.z)));
        o.w = fmaxf(sg, 0.f);
        raw[s] = o;
      }
      if (tid == 0) NERFW_STAMP(93);   // tile written
    };
    int64_t pending = -1;
    if ((int64_t)blockIdx.x < ntiles) encode_pos(blockIdx.x, nullptr);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t s = tile * TM + row;
      const bool live = s < n_total;
      // Inputs of the encodings written after layer 5 (next tile's positions: quarters 0 / 1; this tile's directions:
      // quarter 2), fetched now: the 64-bit ray-index division and the global loads are off the path by then -- with the
      // MMA phases at the pipe's rate the encodings no longer hide behind layer 6 and were delaying its epilogue.
      // Single-MMA modes only: the split mode's MMA phases are three times as long and still cover the encodings, and
      // its epilogue has no registers to spare.
      float pre_v[3] = {0.f, 0.f, 0.f};
      const float* pre = nullptr;
      if constexpr (!X3) {
        if (cq < 2) {
          const int64_t sr = (tile + gridDim.x) * TM + row;
          if (tile + gridDim.x < ntiles && sr < n_total) src.position(sr, pre_v);
        } else if (cq == 2 && !sigma_only) {
          const int64_t sr = tile * TM + row;
          if (sr < n_total) src.direction(sr, pre_v);
        }
        pre = pre_v;
      }
      // ---- trunk epilogues: acc -> bias, ReLU -> bf16 (hi[,lo]) -> next layer's A operand in TMEM ----
      float sig = 0.f;
      // Thread <-> (row, accumulator columns 64 cq .. 64 cq + 63).  All 64 values are pulled into registers first and the
      // accumulator is released; the four 16-column granules are then finished and published one at a time -- granule
      // kb of the four column quarters is K block kb of the next layer (K order: kperm_feature in mlp_tc_layout.cuh) --
      // so the next layer's MMAs overlap with three quarters of this epilogue.
      for (int layer = 0; layer < NERFW_LAYERS; ++layer) {
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (tid == 0) NERFW_STAMP(14 + layer * 8);   // accumulator complete seen by the epilogue
        const float* bias = vec + V_PTSB + layer * 256;
        // inference: the direction layer consumes A_hi only
        const bool want_lo = X3 && (layer != NERFW_LAYERS - 1 || dir_split);
        uint32_t r[4][16];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tmem_ld16(tlane + COL_ACC + cq * 64 + kb * 16, r[kb]);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive_warp(acc_free);
        if (tid == 0) NERFW_STAMP(15 + layer * 8);   // accumulator in registers
        if (sigma_only && layer == NERFW_LAYERS - 1) {
          // nobody consumes layer 7's activations as an operand: only the density head's dot product
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t col = cq * 64 + kb * 16;
#pragma unroll
            for (int e = 0; e < 16; ++e)
              sig = fmaf(fmaxf(__fadd_rn(__uint_as_float(r[kb][e]), bias[col + e]), 0.f), vec[V_DENW + col + e], sig);
          }
          continue;
        }
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t col = cq * 64 + kb * 16;         // accumulator column = output feature
          const uint32_t apos = (kb * 64 + cq * 16) >> 1;  // operand position: granule kb of quarter cq (kperm_feature)
          const float4* b4 = reinterpret_cast<const float4*>(bias + col);
          uint32_t ph[8];
          float a[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = b4[j4];
            unpack2f(add2(pack2(r[kb][4 * j4], r[kb][4 * j4 + 1]), pack2f(bb.x, bb.y)), a[4 * j4], a[4 * j4 + 1]);
            unpack2f(add2(pack2(r[kb][4 * j4 + 2], r[kb][4 * j4 + 3]), pack2f(bb.z, bb.w)), a[4 * j4 + 2], a[4 * j4 + 3]);
            ph[2 * j4] = F16 ? relu_pack_f16x2(a[4 * j4], a[4 * j4 + 1]) : relu_pack_bf16x2(a[4 * j4], a[4 * j4 + 1]);
            ph[2 * j4 + 1] = F16 ? relu_pack_f16x2(a[4 * j4 + 2], a[4 * j4 + 3]) : relu_pack_bf16x2(a[4 * j4 + 2], a[4 * j4 + 3]);
          }
          tmem_st8(tlane + COL_AHI + apos, ph);
          if (want_lo) {  // lo = bf16(relu(a) - hi)
            uint32_t pl[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float l0v, l1v;
              unpack2f(sub2(pack2f(fmaxf(a[2 * j], 0.f), fmaxf(a[2 * j + 1], 0.f)), pack2(ph[j] << 16, ph[j] & 0xffff0000u)), l0v, l1v);
              pl[j] = pack_bf16x2(l0v, l1v);
            }
            tmem_st8(tlane + COL_ALO + apos, pl);
          }
          if (masks) {  // 16 ReLU gates of this slice = one half of gate word col / 32
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) bits |= (a[j] > 0.f ? 1u : 0u) << j;
            reinterpret_cast<unsigned short*>(masks)[2 * mask_index(tile, layer, row, 0, (int)(col >> 5)) + ((col >> 4) & 1)] = (unsigned short)bits;
          }
          if (layer == NERFW_LAYERS - 1) {
            const float4* w4 = reinterpret_cast<const float4*>(vec + V_DENW + col);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 ww = w4[j4];
              sig = fmaf(fmaxf(a[4 * j4], 0.f), ww.x, sig); sig = fmaf(fmaxf(a[4 * j4 + 1], 0.f), ww.y, sig);
              sig = fmaf(fmaxf(a[4 * j4 + 2], 0.f), ww.z, sig); sig = fmaf(fmaxf(a[4 * j4 + 3], 0.f), ww.w, sig);
            }
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_warp(&a_kb[kb]);
          if (tid == 0 && kb == 0) NERFW_STAMP(16 + layer * 8);
          if (tid == 0 && kb == 3) NERFW_STAMP(17 + layer * 8);
        }
        // layer 1's operand is out and the tensor pipe busy again: now finish the PREVIOUS tile (its sig_part, rgb partial
        // sums and -- DIR_ACC -- direction accumulator are untouched until this tile's layer 7 / direction layer)
        if (layer == 0 && pending >= 0) {
          if (DIR_ACC) dir_part_a(pending);
          dir_part_b(pending);
          pending = -1;
        }
        if (layer == NERFW_SKIP + 1) {
          if (!sigma_only) encode_dir(tile, pre);
          if (tile + gridDim.x < ntiles) encode_pos(tile + gridDim.x, pre);
          if (tid == 0) NERFW_STAMP(90);   // encodings written
        }
      }
      sig_part[cq * TM + row] = sig;
      if (sigma_only) {
        named_bar_sync(1, F_EPI_THREADS);
        if (cq == 0 && live) {
          const float sg = (sig_part[row] + sig_part[TM + row]) + (sig_part[2 * TM + row] + sig_part[3 * TM + row]) + vec[V_DENB];
          raw[s] = make_float4(0.f, 0.f, 0.f, fmaxf(sg, 0.f));
        }
        continue;
      }

      if (!DIR_ACC) dir_part_a(tile);   // shared accumulator: free it for the next tile's layer 0 right away
      pending = tile;                   // the rest inside the next tile (or after the loop)
    }
    if (pending >= 0) {
      if (DIR_ACC) dir_part_a(pending);
      dir_part_b(pending);
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (timeline && blockIdx.x == 0 && tid == 0) { timeline[122] = clock64(); timeline[123] = (long long)global_timer_ns(); }
  if (warp == F_MMA_WARP) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Self-test of the primitives: D (128 x N fp32) = A (128 x K bf16) * B (N x K bf16)^T for one CTA.
// mode 0: A from shared memory (SS); mode 1: A from tensor memory (TS).  K multiple of 64 (<= 256), N multiple of 16.
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __nv_bfloat16* __restrict__ A,
                                                               const __nv_bfloat16* __restrict__ B, int N, int K, int mode,
                                                               float* __restrict__ D) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint8_t* sA = sm;                  // K/64 tiles of 128 x 128 B
  uint8_t* sB = sm + 65536;          // K/64 tiles of N x 128 B (256*128 stride)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536 + 131072);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = K / 64;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  // stage operands
  for (int idx = tid; idx < N * K; idx += 128) {
    int n = idx / K, k = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + (k / 64) * BIG_CHUNK + sw128_offset(n, k % 64)) = B[idx];
  }
  if (mode == 0) {
    for (int idx = tid; idx < 128 * K; idx += 128) {
      int m = idx / K, k = idx % K;
      *reinterpret_cast<__nv_bfloat16*>(sA + (k / 64) * 16384 + sw128_offset(m, k % 64)) = A[idx];
    }
  } else {
    const uint32_t tl = tmem + ((warp * 32) << 16);
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t pk[16];
      for (int j = 0; j < 16; ++j) {
        const __nv_bfloat16* a = A + (size_t)tid * K + 2 * (c0 + j);
        pk[j] = (uint32_t)__bfloat16_as_ushort(a[0]) | ((uint32_t)__bfloat16_as_ushort(a[1]) << 16);
      }
      tmem_st16(tl + COL_AHI + c0, pk);
    }
    tmem_wait_st();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, (uint32_t)N);
    for (int kb = 0; kb < nkb; ++kb) {
      uint64_t b = smem_desc_sw128(smem_u32(sB + kb * BIG_CHUNK));
      uint64_t a = smem_desc_sw128(smem_u32(sA + kb * 16384));
      for (int k = 0; k < 4; ++k) {
        uint32_t accf = (kb | k) ? 1u : 0u;
        if (mode == 0) mma_ss(tmem + COL_ACC, a + 2 * k, b + 2 * k, idesc, accf);
        else mma_ts(tmem + COL_ACC, tmem + COL_AHI + 32 * kb + 8 * k, b + 2 * k, idesc, accf);
      }
    }
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  {
    const uint32_t tl = tmem + ((warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tl + COL_ACC + c0, r);
      tmem_wait_ld();
      for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace tc

size_t mlp_tc_packed_bytes() { return tc::PACKED_BYTES; }

int launch_pack_weights(const NerfwWeights& w, void* packed, cudaStream_t stream) {
  const int64_t total = (int64_t)tc::N_BIG * 256 * 8 + (int64_t)tc::N_SMALL * 128 * 8;
  tc::pack_weights_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(
      w, reinterpret_cast<uint8_t*>(packed), reinterpret_cast<uint8_t*>(packed) + mlp_tc_packed_f16_offset());
  NERFW_LAUNCHED();
  return NERFW_OK;
}

int launch_app_offset(const NerfwWeights& w, const float* emb, int64_t emb_rows, float* app_off, cudaStream_t stream) {
  tc::app_offset_kernel<<<(unsigned)(emb_rows < 4096 ? emb_rows : 4096), 128, 0, stream>>>(w, emb, emb_rows, reinterpret_cast<float4*>(app_off));
  NERFW_LAUNCHED();
  return NERFW_OK;
}

int launch_mlp_tc_fwd(const NerfwWeights& w, const void* packed, const SampleSource& src, const float* app_off,
                      int64_t n_total, int mode_flags, float* raw, void* relu_masks, cudaStream_t stream) {
  const int mode = mode_flags & 0xff;
  const bool x3 = mode == NERFW_MLP_BF16X3, f16 = mode == NERFW_MLP_FP16;
  const bool sigma_only = (mode_flags & NERFW_MLP_SIGMA_ONLY) != 0;
  (void)w;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    NERFW_CUDA(cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    NERFW_CUDA(cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    NERFW_CUDA(cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    NERFW_CUDA(cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    NERFW_CUDA(cudaFuncSetAttribute(tc::mlp_tc_fwd_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
  }
  int64_t ntiles = ceil_div64(n_total, tc::TM);
  int64_t grid = ntiles < sm_count() ? ntiles : sm_count();
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
  int dbg = 0;                     // kernel flags: bit 0 = profiling switch (reuse the ring contents; wrong results)
  long long* timeline = nullptr;   // profiling: device pointer to 128 int64 clock64 slots
#ifdef NERFW_PROFILE
  // Only in the separate profiling build (make PROFILE=1 -> libnerfw_sm100_profile.so, used by scripts/): the product
  // library never reads the environment.
  if (getenv("NERFW_FWD_SKIP_WEIGHTS")) dbg |= 1;
  if (const char* t = getenv("NERFW_FWD_TIMELINE")) timeline = reinterpret_cast<long long*>(strtoull(t, nullptr, 10));
#endif
  const float4* ao = reinterpret_cast<const float4*>(app_off);
  const size_t f16_off = mlp_tc_packed_f16_offset();
  float4* out = reinterpret_cast<float4*>(raw);
  uint32_t* mk = reinterpret_cast<uint32_t*>(relu_masks);
  auto launch = [&](auto kernel) {
    kernel<<<(unsigned)grid, tc::F_THREADS, tc::SMEM_BYTES, stream>>>(pk, src, ao, n_total, out, mk, dbg, timeline, f16_off);
  };
  if (sigma_only) {
    if (x3) launch(tc::mlp_tc_fwd_kernel<true, false, true>);
    else if (f16) launch(tc::mlp_tc_fwd_kernel<false, true, true>);
    else launch(tc::mlp_tc_fwd_kernel<false, false, true>);
  } else {
    if (x3) launch(tc::mlp_tc_fwd_kernel<true>);
    else if (f16) launch(tc::mlp_tc_fwd_kernel<false, true>);
    else launch(tc::mlp_tc_fwd_kernel<false>);
  }
  NERFW_LAUNCHED();
  return NERFW_OK;
}

}  // namespace nerfw

// D = A B^T through tcgen05 for one 128-row tile; used by tests to pin descriptor / swizzle / TMEM layouts.
extern "C" int nerfw_selftest_umma(const void* a_bf16, const void* b_bf16, int n, int k, int mode, float* d, void* stream) {
  using namespace nerfw;
  NERFW_REQUIRE(a_bf16 && b_bf16 && d, "nerfw_selftest_umma: null pointer");
  NERFW_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0, "nerfw_selftest_umma: N must be a multiple of 16 in [16,256]");
  NERFW_REQUIRE(k >= 64 && k <= 256 && k % 64 == 0, "nerfw_selftest_umma: K must be a multiple of 64 in [64,256]");
  NERFW_REQUIRE(mode == 0 || mode == 1, "nerfw_selftest_umma: mode must be 0 (SS) or 1 (TS)");
  const size_t smem = 65536 + 131072 + 64 + 1024;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(tc::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  tc::umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(a_bf16),
                                                               reinterpret_cast<const __nv_bfloat16*>(b_bf16), n, k, mode, d);
  NERFW_LAUNCHED();
  return NERFW_OK;
}
