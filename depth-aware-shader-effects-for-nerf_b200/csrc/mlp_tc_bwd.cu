// K5 (tensor-core mode): backward of the NeRF-W MLP (autograd of src/models.py:105-162) on tcgen05, bf16 operands with
// fp32 accumulation.  Two kernels:
//
//  pass 1  `mlp_tc_bwd_pass1_kernel` -- per 128-sample tile, same warp-specialised skeleton as the forward kernel:
//          forward recompute (activations X_l written to an HBM scratch tile as bf16, ReLU masks kept in shared
//          memory), then the dgrad chain dZ_l = dH_{l+1} * relu', dH_l = dZ_l W_l with dZ as the TMEM A operand and the
//          TRANSPOSED weight image as B; every dZ_l is written to the scratch tile as well.
//  pass 2  `mlp_tc_wgrad_kernel` -- dW_l = dZ_l^T X_l as UMMA with both operands MN-major straight from the scratch
//          tiles (they are stored as [sample][feature] blocks of 128 x 64 in the 128B-swizzled shared-memory image, so a
//          tile is both a valid K-major forward operand and a valid MN-major wgrad operand); each CTA owns one weight
//          block and a contiguous range of tiles, keeps the fp32 accumulator resident in TMEM for the whole range and
//          flushes it once with atomics.  Bias / head gradients are column reductions of the same shared-memory tiles
//          on CUDA cores while the MMAs run.
//
// Scratch tile = 71 blocks of 16 KB per 128 samples (1.16 MB): pass 2 is HBM-bound by construction (128 FLOP/B).
#include "common.cuh"
#include "mlp_common.cuh"
#include "mlp_tc.cuh"
#include "mlp_tc_layout.cuh"
#include <stdlib.h>
#include <string.h>

namespace nerfw {
namespace tcb {

using namespace umma;
using namespace tc;

// ---- transposed weight image (dgrad B operands), consumption order: dir (2 chunks), L7..L1 (4 chunks each) ----------
constexpr int NT_CHUNKS = 30;
constexpr size_t WT_BYTES = (size_t)NT_CHUNKS * BIG_CHUNK;
constexpr size_t PACKED_T_OFFSET = (PACKED_BYTES + 1023) & ~(size_t)1023;
constexpr size_t PACKED_F16_OFFSET = PACKED_T_OFFSET + WT_BYTES;
constexpr size_t PACKED_TOTAL = PACKED_F16_OFFSET + F16_BYTES;

// ---- scratch tile layout: 16 KB blocks [128 samples][64 features] bf16, 128B swizzle ------------------------------------
constexpr uint32_t BLK = 16384;
constexpr int XB_ENCX = 0;
__host__ __device__ constexpr int XB_H(int l) { return 1 + 4 * (l - 1); }  // l = 1..8: output of trunk layer l-1
constexpr int XB_ENCD = 33;
constexpr int XB_HDT = 34;                                                   // relu(dir) + appearance feature, 2 blocks
__host__ __device__ constexpr int ZB(int l) { return 36 + 4 * l; }         // l = 0..7: dZ of trunk layer l
constexpr int ZB_DIR = 68;                                                   // 2 blocks
constexpr int DLS_BLOCK = 70;                                                // 128 float4: (dlogit r,g,b, dsigma_pre)
constexpr size_t TILE_BYTES = 71ull * BLK;

// pass-1 shared memory: forward map, with the *_LO tiles reused for the ReLU masks and two small vectors appended
constexpr uint32_t SM1_MASK_A = SM_PEX_LO;  // layers 0..3: [layer][row][ch][4 words]
constexpr uint32_t SM1_MASK_B = SM_PED_LO;  // layers 4..7
constexpr uint32_t SM1_DSIG = SM_RGB + 4 * TM * 16;  // after the [4][128] float4 rgb partial sums
constexpr uint32_t SM1_APPV = SM1_DSIG + TM * 4;
constexpr uint32_t SM1_BAR = SM1_APPV + 128 * 4;
constexpr uint32_t SM1_TMEMPTR = SM1_BAR + (2 * NSTAGES + 2 + 5 + 2) * 8;
// With the forward's gate masks the mask regions are free, and with a 3-stage weight ring so is the fourth ring stage:
// together a 4 x 16 KB staging image of the 64 KB every epilogue step writes to the scratch tile.  The epilogue warps only
// write shared memory (16-byte stores in the block image, conflict free); a dedicated warp ships each finished image with
// four cp.async.bulk stores (TMA engine), so no st.global is issued by the epilogue warps.
constexpr int P1_STAGES = 3;
__device__ __forceinline__ uint32_t sm1_staging(int b) {
  return b == 0 ? SM1_MASK_A : b == 1 ? SM1_MASK_B : SM_RING + (uint32_t)P1_STAGES * BIG_CHUNK + (uint32_t)(b - 2) * BLK;
}
struct Pipe3 {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == P1_STAGES) { stage = 0; phase ^= 1; }
  }
};
constexpr size_t SMEM1_BYTES = SM1_TMEMPTR + 16 + 1024;

__global__ void __launch_bounds__(256) pack_weights_t_kernel(NerfwWeights w, uint8_t* __restrict__ packed_t) {
  // one thread per (chunk, input feature n (row), 8-wide output group g)
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)NT_CHUNKS * 256 * 8) return;
  int chunk = (int)(idx / (256 * 8));
  int r = (int)(idx % (256 * 8));
  int n = r / 8, g = r % 8;
  const float* W;
  int ld, ob;
  if (chunk < 2) { W = w.dir_w; ld = 256 + NERFW_DIR_DIM; ob = chunk; }
  else {
    int l = 7 - (chunk - 2) / 4;
    ob = (chunk - 2) % 4;
    W = w.pts_w[l];
    ld = (l == NERFW_SKIP) ? 256 + NERFW_POS_DIM : 256;
  }
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    // K = the layer's output features: in operand order (kperm_feature) when dZ comes from a trunk epilogue
    const int o = chunk < 2 ? ob * 64 + g * 8 + e : kperm_feature(ob, g * 8 + e);
    v[e] = __float2bfloat16_rn(__ldg(W + (size_t)o * ld + n));
  }
  *reinterpret_cast<uint4*>(packed_t + (size_t)chunk * BIG_CHUNK + sw128_offset((uint32_t)n, (uint32_t)g * 8)) =
      *reinterpret_cast<const uint4*>(v);
}

// pass-1 warp roles (same split as the forward kernel): 16 epilogue warps, one producer, one MMA issuer
constexpr int P1_EPI_WARPS = 16;
constexpr int P1_PRODUCER_WARP = 16;
constexpr int P1_MMA_WARP = 17;
constexpr int P1_STORE_WARP = 18;
constexpr int P1_THREADS = 608;
constexpr int P1_EPI_THREADS = P1_EPI_WARPS * 32;

// own ReLU gate words (used only when the caller passes no forward masks): [layer][row][8 words], word = column / 32
__device__ __forceinline__ uint32_t* mask_words(uint8_t* sm, int layer, uint32_t row) {
  uint8_t* base = sm + (layer < 4 ? SM1_MASK_A : SM1_MASK_B);
  return reinterpret_cast<uint32_t*>(base) + ((layer & 3) * TM + row) * 8;
}
// 0xffffffff if bit `pos` of x is set, else 0 (one BFE.S32)
__device__ __forceinline__ uint32_t bit_mask(uint32_t x, int pos) {
  int r;
  asm("bfe.s32 %0, %1, %2, 1;" : "=r"(r) : "r"((int)x), "r"(pos));
  return (uint32_t)r;
}
// packed bf16x2 (g_lo, g_hi) gated by bits 2j and 2j+1 of `bits`
__device__ __forceinline__ uint32_t gate_pack(float g_lo, float g_hi, uint32_t bits, int j) {
  const uint32_t p = pack_bf16x2(g_lo, g_hi);
  return p & __byte_perm(bit_mask(bits, 2 * j), bit_mask(bits, 2 * j + 1), 0x5410);
}

// 32 consecutive features (16 packed bf16x2 words) of one sample row -> scratch block in the swizzled image
__device__ __forceinline__ void store_row32(uint8_t* tile, int block0, uint32_t row, uint32_t col, const uint32_t (&p)[16]) {
  uint8_t* blk = tile + (size_t)(block0 + (col >> 6)) * BLK;
  const uint32_t k0 = col & 63;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v = make_uint4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
    *reinterpret_cast<uint4*>(blk + sw128_offset(row, k0 + 8 * c)) = v;
  }
}

// ====================================================================================================================
// profiling: clock64 stamps of CTA 0, third tile (slot = 16 + 8 * step + event); only when a timeline buffer is passed
#define P1_STAMP(slot) do { if (timeline && blockIdx.x == 0 && tile == (int64_t)(2 * gridDim.x)) timeline[slot] = clock64(); } while (0)

__global__ void __launch_bounds__(P1_THREADS, 1) mlp_tc_bwd_pass1_kernel(const uint8_t* __restrict__ packed, SampleSource src,
                                                                       const float4* __restrict__ app_off,
                                                                       const float* __restrict__ app_vec,
                                                                       const float4* __restrict__ d_raw, int64_t n_total,
                                                                       uint8_t* __restrict__ scratch,
                                                                       float* __restrict__ dl_acc,
                                                                       const uint32_t* __restrict__ fwd_masks, int debug,
                                                                       long long* __restrict__ timeline) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + SM1_BAR);
  uint64_t* empty = full + NSTAGES;
  uint64_t* acc_full = empty + NSTAGES;
  uint64_t* a_ready = acc_full + 1;   // per tile: encodings in place, operand columns free
  uint64_t* acc_free = a_ready + 1;   // accumulator copied to registers
  uint64_t* a_kb = acc_free + 1;      // [4]: K block kb of the next operand written
  uint64_t* st_full = a_kb + 4;       // staging image of this step complete (one arrival per epilogue warp)
  uint64_t* st_free = st_full + 1;    // staging image read out by the bulk stores
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + SM1_TMEMPTR);
  float* vec = reinterpret_cast<float*>(sm + SM_VEC);
  float* sig_part = reinterpret_cast<float*>(sm + SM_SIG);
  float4* rgb_part = reinterpret_cast<float4*>(sm + SM_RGB);  // [4][128]
  float* dsig_s = reinterpret_cast<float*>(sm + SM1_DSIG);
  float* appv = reinterpret_cast<float*>(sm + SM1_APPV);

  if (warp == P1_PRODUCER_WARP && lane == 0) {
    for (int i = 0; i < NSTAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(a_ready, P1_EPI_WARPS);   // one arrival per epilogue warp
    mbar_init(acc_free, P1_EPI_WARPS);
    for (int i = 0; i < 4; ++i) mbar_init(&a_kb[i], P1_EPI_WARPS);
    mbar_init(st_full, P1_EPI_WARPS);
    mbar_init(st_free, 1);
    fence_mbar_init();
  }
  if (warp == P1_MMA_WARP) tmem_alloc<512>(tmem_ptr);
  if (warp < P1_EPI_WARPS) {
    const float* gv = reinterpret_cast<const float*>(packed + W_BYTES);
    for (int i = tid; i < V_FLOATS; i += P1_EPI_THREADS) vec[i] = __ldg(gv + i);
    if (tid < 128) appv[tid] = app_vec ? __ldg(app_vec + tid) : 0.f;
    // the unused half of the direction-encoding tile must be finite: it is a (discarded) wgrad operand column
    for (int i = tid; i < 16384 / 16; i += P1_EPI_THREADS) reinterpret_cast<uint4*>(sm + SM_PED_HI)[i] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int64_t ntiles = (n_total + TM - 1) / TM;
  const uint8_t* packed_t = packed + PACKED_T_OFFSET;

  if (warp == P1_PRODUCER_WARP) {
    if (lane == 0) {
      Pipe3 p;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = 0; i < N_CHUNKS; ++i) {  // forward stream (hi copies only)
          const uint32_t sz = i < N_BIG ? BIG_CHUNK : SMALL_CHUNK;
          mbar_wait(&empty[p.stage], p.phase ^ 1);
          mbar_arrive_expect_tx(&full[p.stage], sz);
          bulk_g2s(sm + SM_RING + p.stage * BIG_CHUNK, packed + chunk_offset(i), sz, &full[p.stage]);
          p.advance();
        }
        for (int i = 0; i < NT_CHUNKS; ++i) {  // dgrad stream (transposed weights)
          mbar_wait(&empty[p.stage], p.phase ^ 1);
          mbar_arrive_expect_tx(&full[p.stage], BIG_CHUNK);
          bulk_g2s(sm + SM_RING + p.stage * BIG_CHUNK, packed_t + (size_t)i * BIG_CHUNK, BIG_CHUNK, &full[p.stage]);
          p.advance();
        }
      }
    }
  } else if (warp == P1_STORE_WARP) {
    // ===================== scratch-tile stores: one bulk copy per 16 KB block of every finished staging image =========
    if (lane == 0 && fwd_masks && !(debug & 4)) {
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint8_t* tsc = scratch + (size_t)tile * TILE_BYTES;
        for (int step = 0; step < 17; ++step) {   // 8 forward layers, direction layer, 8 dgrad layers (7 .. 0)
          mbar_wait(st_full, ph);
          ph ^= 1;
          if (step == 8) {
            bulk_s2g(tsc + (size_t)ZB_DIR * BLK, sm + sm1_staging(0), BLK);
            bulk_s2g(tsc + (size_t)(ZB_DIR + 1) * BLK, sm + sm1_staging(1), BLK);
            bulk_s2g(tsc + (size_t)XB_HDT * BLK, sm + sm1_staging(2), 2 * BLK);
          } else {
            const int b0 = step < 8 ? XB_H(step + 1) : ZB(16 - step);
            bulk_s2g(tsc + (size_t)b0 * BLK, sm + sm1_staging(0), BLK);
            bulk_s2g(tsc + (size_t)(b0 + 1) * BLK, sm + sm1_staging(1), BLK);
            bulk_s2g(tsc + (size_t)(b0 + 2) * BLK, sm + sm1_staging(2), 2 * BLK);
          }
          bulk_commit();
          bulk_wait_read_all();
          mbar_arrive(st_free);
        }
      }
      bulk_wait_all();
    }
  } else if (warp == P1_MMA_WARP) {
    if (lane == 0) {
      Pipe3 p;
      uint32_t ar_phase = 0;
      const uint32_t idesc256 = idesc_bf16(128, 256), idesc128 = idesc_bf16(128, 128);
      const uint32_t ring = smem_u32(sm + SM_RING);
      const uint32_t d_acc = tmem + COL_ACC;
      auto kblock = [&](bool from_tmem, uint64_t a, uint32_t idesc, int ksteps, bool first) {
        mbar_wait(&full[p.stage], p.phase);
        tc_fence_after();
        uint64_t b = smem_desc_sw128(ring + p.stage * BIG_CHUNK);
        for (int k = 0; k < ksteps; ++k) {
          uint32_t accf = (first && k == 0) ? 0u : 1u;
          if (from_tmem) mma_ts(d_acc, (uint32_t)a + 8 * k, b + 2 * k, idesc, accf);
          else mma_ss(d_acc, a + 2 * k, b + 2 * k, idesc, accf);
        }
        mma_commit(&empty[p.stage]);
        p.advance();
      };
      const uint64_t pex = smem_desc_sw128(smem_u32(sm + SM_PEX_HI));
      const uint64_t ped = smem_desc_sw128(smem_u32(sm + SM_PED_HI));
      // Handshakes per MMA step: acc_free (the previous accumulator sits in registers), then one barrier per 64-wide K
      // block of the operand -- the epilogue publishes its four 16-column granules one at a time (K order: kperm_feature),
      // so the MMAs of a step start while three quarters of the previous epilogue are still running.
      uint32_t ph_free = 0, ph_kb[4] = {0, 0, 0, 0};
      auto wait_bar = [&](uint64_t* bar, uint32_t& phase) {
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
      };
      int64_t tile = 0;
      int step = 0;
      auto from_tmem = [&](int nkb, uint32_t idesc) {
        for (int kb = 0; kb < nkb; ++kb) {
          wait_bar(&a_kb[kb], ph_kb[kb]);
          if (kb == 0) P1_STAMP(16 + 8 * step + 1);
          if (kb == nkb - 1) P1_STAMP(16 + 8 * step + 2);
          kblock(true, tmem + COL_AHI + 32 * kb, idesc, 4, kb == 0);
        }
      };
      for (tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        wait_bar(a_ready, ar_phase);
        P1_STAMP(0);
        // ---- forward recompute ----
        for (int layer = 0; layer < NERFW_LAYERS; ++layer) {
          step = layer;
          wait_bar(acc_free, ph_free);
          P1_STAMP(16 + 8 * step + 0);
          if (layer == 0) {
            kblock(false, pex, idesc256, 4, true);
          } else {
            from_tmem(4, idesc256);
            if (layer == NERFW_SKIP) kblock(false, pex, idesc256, 4, false);
          }
          mma_commit(acc_full);
          P1_STAMP(16 + 8 * step + 3);
        }
        step = 8;
        wait_bar(acc_free, ph_free);
        P1_STAMP(16 + 8 * step + 0);
        from_tmem(4, idesc128);
        kblock(false, ped, idesc128, 2, false);
        mma_commit(acc_full);
        P1_STAMP(16 + 8 * step + 3);
        // ---- dgrad chain: dH8 = dZdir W_dir[:, :256], then dH_l = dZ_l W_l[:, :256] for l = 7..1 ----
        step = 9;
        wait_bar(acc_free, ph_free);
        P1_STAMP(16 + 8 * step + 0);
        from_tmem(2, idesc256);
        mma_commit(acc_full);
        P1_STAMP(16 + 8 * step + 3);
        for (int l = NERFW_LAYERS - 1; l >= 1; --l) {
          step = 10 + (NERFW_LAYERS - 1 - l);
          wait_bar(acc_free, ph_free);
          P1_STAMP(16 + 8 * step + 0);
          from_tmem(4, idesc256);
          mma_commit(acc_full);
          P1_STAMP(16 + 8 * step + 3);
        }
      }
    }
  } else {
    // ===================== encoders + epilogues: thread <-> (sample row, column quarter) =====================
    const uint32_t quad = warp & 3, cq = warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t tlane = tmem + ((quad * 32) << 16);
    uint32_t acc_phase = 0;
    uint8_t* pex = sm + SM_PEX_HI;
    uint8_t* ped = sm + SM_PED_HI;
    // with the forward's gates the mask region of shared memory is free: 16 x 2 KB staging buffers for coalesced stores
    // with the forward's gates: 32-column pieces go to block `sblock` of the staging image (shipped by the store warp);
    // without them the mask regions are in use and the pieces are stored directly (slow path, API completeness only)
    const bool staged = fwd_masks != nullptr;
    uint32_t st_phase = 0;
    auto put = [&](uint8_t* tile_base, int block0, int sblock, uint32_t col, const uint32_t (&p)[16]) {
      if (debug & 4) return;   // profiling only (NERFW_WGRAD_DEBUG bit 2): no scratch stores
      if (staged) store_row32(sm + sm1_staging(sblock) - (size_t)(col >> 6) * BLK, 0, row, col, p);
      else store_row32(tile_base, block0, row, col, p);
    };
    auto staging_acquire = [&]() {   // the previous image has been read out
      if (staged && !(debug & 4)) { mbar_wait(st_free, st_phase ^ 1); st_phase ^= 1; }
    };
    auto staging_release = [&]() {
      if (staged && !(debug & 4)) { fence_proxy_async_smem(); mbar_arrive_warp(st_full); }
    };
    // ---- encodings (as in the bf16 forward kernel) of tile t into shared memory and, as wgrad operands of layer 0, of the
    // skip part of layer 4 and of the direction layer, into its scratch tile.  Off the critical path: the encodings of
    // the NEXT tile are produced during the dgrad phase of the current one, when the forward MMAs that read the shared
    // tiles are long complete; the closing arrival on a_ready doubles as "operand TMEM columns free".
    auto encode_tile = [&](int64_t t) {
      const int64_t sr = t * TM + row;
      float x[3] = {0.f, 0.f, 0.f};
      if (sr < n_total && cq < 2) src.position(sr, x);
      float v[32];
      if (cq == 0) {
        pos_features32<0, true>(x, v);
        store_features32<false>(pex, pex, row, 0, v);
      } else if (cq == 1) {
        pos_features32<1, true>(x, v);
        store_features32<false>(pex, pex, row, 32, v);
      } else if (cq == 2) {
        float d[3] = {0.f, 0.f, 0.f};
        if (sr < n_total) src.direction(sr, d);
        dir_features32<true>(d, v);
        store_features32<false>(ped, ped, row, 0, v);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, P1_EPI_THREADS);
      uint8_t* dst = scratch + (size_t)t * TILE_BYTES;
      const int tt = warp * 32 + lane;  // 0..511
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int o = (tt + 512 * i) * 16;
        *reinterpret_cast<uint4*>(dst + (size_t)XB_ENCX * BLK + o) = *reinterpret_cast<const uint4*>(pex + o);
        *reinterpret_cast<uint4*>(dst + (size_t)XB_ENCD * BLK + o) = *reinterpret_cast<const uint4*>(ped + o);
      }
    };
    mbar_arrive_warp(acc_free);   // the accumulator starts out free
    if ((int64_t)blockIdx.x < ntiles) {
      encode_tile(blockIdx.x);
      mbar_arrive_warp(a_ready);
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t s = tile * TM + row;
      const bool live = s < n_total;
      uint8_t* tsc = scratch + (size_t)tile * TILE_BYTES;

      // ---- forward trunk epilogues: next A operand + activation tile (+ own ReLU gates when none were passed) ----
      float sig = 0.f;
      for (int layer = 0; layer < NERFW_LAYERS; ++layer) {
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const float* bias = vec + V_PTSB + layer * 256;
        const bool plain = fwd_masks != nullptr && layer != NERFW_LAYERS - 1;  // warp-uniform
        if (tid == 0) P1_STAMP(16 + 8 * (layer + 1) + 4);   // accumulator of step `layer` complete (slot of the next step)
        // thread <-> accumulator columns 64 cq .. 64 cq + 63: into registers, accumulator released, then granule by granule
        uint32_t r[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld16(tlane + COL_ACC + cq * 64 + j * 16, r[j]);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive_warp(acc_free);
        uint32_t ph2[2][16];   // the same values as two 32-column pieces for the scratch tile
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t col = cq * 64 + j * 16;
          const float4* b4 = reinterpret_cast<const float4*>(bias + col);
          uint32_t* ph = &ph2[j >> 1][(j & 1) * 8];
          float a[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = b4[j4];
            unpack2f(add2(pack2(r[j][4 * j4], r[j][4 * j4 + 1]), pack2f(bb.x, bb.y)), a[4 * j4], a[4 * j4 + 1]);
            unpack2f(add2(pack2(r[j][4 * j4 + 2], r[j][4 * j4 + 3]), pack2f(bb.z, bb.w)), a[4 * j4 + 2], a[4 * j4 + 3]);
            ph[2 * j4] = relu_pack_bf16x2(a[4 * j4], a[4 * j4 + 1]);
            ph[2 * j4 + 1] = relu_pack_bf16x2(a[4 * j4 + 2], a[4 * j4 + 3]);
          }
          if (!plain) {
            if (!fwd_masks) {
              uint32_t bits = 0;
#pragma unroll
              for (int e = 0; e < 16; ++e) bits |= (a[e] > 0.f ? 1u : 0u) << e;
              reinterpret_cast<unsigned short*>(mask_words(sm, layer, row))[col >> 4] = (unsigned short)bits;
            }
            if (layer == NERFW_LAYERS - 1) {
              const float4* w4 = reinterpret_cast<const float4*>(vec + V_DENW + col);
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 ww = w4[e4];
                sig = fmaf(fmaxf(a[4 * e4], 0.f), ww.x, sig); sig = fmaf(fmaxf(a[4 * e4 + 1], 0.f), ww.y, sig);
                sig = fmaf(fmaxf(a[4 * e4 + 2], 0.f), ww.z, sig); sig = fmaf(fmaxf(a[4 * e4 + 3], 0.f), ww.w, sig);
              }
            }
          }
          const uint32_t p8[8] = {ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]};
          tmem_st8(tlane + COL_AHI + ((j * 64 + cq * 16) >> 1), p8);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_warp(&a_kb[j]);
          if (tid == 0 && j == 0) P1_STAMP(16 + 8 * (layer + 1) + 5);
          if (tid == 0 && j == 3) P1_STAMP(16 + 8 * (layer + 1) + 6);
        }
        staging_acquire();
#pragma unroll
        for (int q = 0; q < 2; ++q) put(tsc, XB_H(layer + 1), (int)cq, cq * 64 + q * 32, ph2[q]);
        staging_release();
        if (tid == 0) P1_STAMP(16 + 8 * (layer + 1) + 7);
      }
      sig_part[cq * TM + row] = sig;

      // ---- direction layer epilogue: rgb, d logits, d sigma_pre, dZ of the direction layer (32 columns per thread) ----
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      float p3[3] = {0.f, 0.f, 0.f};
      uint32_t hmask;
      const uint32_t dcol = cq * 32;
      uint32_t ph_hdt[16];
      {
        uint32_t r[32];
        tmem_ld32(tlane + COL_ACC + dcol, r);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive_warp(acc_free);
        uint32_t bits = 0;
        float hv[32];
        // bias, rgb-head rows and the appearance vector as 16-byte shared-memory loads: this step is on the critical path of
        // the tile (the tensor pipe waits for its dZ operand) and scalar loads made it LSU-bound (224 wavefronts per warp)
        const float4* b4 = reinterpret_cast<const float4*>(vec + V_DIRB + dcol);
        const float4* w4 = reinterpret_cast<const float4*>(vec + V_RGBW + dcol);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = b4[j4], wr = w4[j4], wg = w4[32 + j4], wb = w4[64 + j4];
          const float bj[4] = {bb.x, bb.y, bb.z, bb.w};
          const float wj[3][4] = {{wr.x, wr.y, wr.z, wr.w}, {wg.x, wg.y, wg.z, wg.w}, {wb.x, wb.y, wb.z, wb.w}};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            hv[j] = fmaxf(__uint_as_float(r[j]) + bj[e], 0.f);
            bits |= (hv[j] > 0.f ? 1u : 0u) << j;
#pragma unroll
            for (int c = 0; c < 3; ++c) p3[c] = fmaf(hv[j], wj[c][e], p3[c]);
          }
        }
        if (src.emb_shared || !app_vec) {
          const float4* a4p = reinterpret_cast<const float4*>(appv + dcol);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 a4 = a4p[j4];
            ph_hdt[2 * j4] = pack_bf16x2(hv[4 * j4] + a4.x, hv[4 * j4 + 1] + a4.y);
            ph_hdt[2 * j4 + 1] = pack_bf16x2(hv[4 * j4 + 2] + a4.z, hv[4 * j4 + 3] + a4.w);
          }
        } else {  // per-ray embeddings: a = W_app e_ray + b_app, one 128-float row per ray (app_vec_kernel)
          const float4* av = reinterpret_cast<const float4*>(app_vec + (live ? src.emb_row(s) : 0) * NERFW_DIR_HIDDEN + dcol);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 a4 = __ldg(av + j4);
            ph_hdt[2 * j4] = pack_bf16x2(hv[4 * j4] + a4.x, hv[4 * j4 + 1] + a4.y);
            ph_hdt[2 * j4 + 1] = pack_bf16x2(hv[4 * j4 + 2] + a4.z, hv[4 * j4 + 3] + a4.w);
          }
        }
        hmask = fwd_masks ? __ldg(fwd_masks + mask_index(tile, NERFW_LAYERS, row, cq >> 1, (int)(cq & 1))) : bits;
      }
      tc_fence_before();
      rgb_part[cq * TM + row] = make_float4(p3[0], p3[1], p3[2], 0.f);
      named_bar_sync(1, P1_EPI_THREADS);
      float dlog[3];
      {
        const float4 q0 = rgb_part[row], q1 = rgb_part[TM + row], q2 = rgb_part[2 * TM + row], q3 = rgb_part[3 * TM + row];
        float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
        if (app_off && live) off = __ldg(app_off + src.emb_row(s));
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) dr = __ldg(d_raw + s);
        const float lg[3] = {(q0.x + q1.x) + (q2.x + q3.x) + vec[V_RGBB + 0] + off.x,
                             (q0.y + q1.y) + (q2.y + q3.y) + vec[V_RGBB + 1] + off.y,
                             (q0.z + q1.z) + (q2.z + q3.z) + vec[V_RGBB + 2] + off.z};
        const float dd[3] = {dr.x, dr.y, dr.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float rgb = 1.0f / (1.0f + expf(-lg[c]));
          dlog[c] = dd[c] * rgb * (1.0f - rgb);
        }
        const float pre = (sig_part[row] + sig_part[TM + row]) + (sig_part[2 * TM + row] + sig_part[3 * TM + row]) + vec[V_DENB];
        const float ds = pre > 0.f ? dr.w : 0.f;
        if (cq == 0) {
          dsig_s[row] = ds;
          *reinterpret_cast<float4*>(tsc + (size_t)DLS_BLOCK * BLK + row * 16) = make_float4(dlog[0], dlog[1], dlog[2], ds);
          if (dl_acc && !src.emb_shared) {  // per-ray embeddings: sum of d logits per ray
            // the 32 samples of a warp are consecutive: with >= 32 samples per ray they belong to one ray (one warp reduction,
            // one atomic triple) unless the warp straddles a ray boundary or the end of the batch (per-lane atomics)
            const int64_t er = live ? src.emb_row(s) : -1;
            const int64_t er0 = __shfl_sync(0xffffffffu, er, 0);
            if (__all_sync(0xffffffffu, er == er0)) {
              float a0 = dlog[0], a1 = dlog[1], a2 = dlog[2];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
              }
              if (lane == 0 && er0 >= 0) { atomicAdd(dl_acc + 4 * er0, a0); atomicAdd(dl_acc + 4 * er0 + 1, a1); atomicAdd(dl_acc + 4 * er0 + 2, a2); }
            } else if (live) {
              float* acc = dl_acc + 4 * er;
              atomicAdd(acc + 0, dlog[0]); atomicAdd(acc + 1, dlog[1]); atomicAdd(acc + 2, dlog[2]);
            }
          } else if (dl_acc) {  // sum of d logits of the shared embedding row: appearance gradients are finished from it
            float a0 = dlog[0], a1 = dlog[1], a2 = dlog[2];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              a0 += __shfl_xor_sync(0xffffffffu, a0, o);
              a1 += __shfl_xor_sync(0xffffffffu, a1, o);
              a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (lane == 0) { atomicAdd(dl_acc + 0, a0); atomicAdd(dl_acc + 1, a1); atomicAdd(dl_acc + 2, a2); }
          }
        }
      }
      {
        uint32_t ph[16];
        const float4* w4 = reinterpret_cast<const float4*>(vec + V_RGBW + dcol);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 wr = w4[j4], wg = w4[32 + j4], wb = w4[64 + j4];
          const float g0 = dlog[0] * wr.x + dlog[1] * wg.x + dlog[2] * wb.x;
          const float g1 = dlog[0] * wr.y + dlog[1] * wg.y + dlog[2] * wb.y;
          const float g2 = dlog[0] * wr.z + dlog[1] * wg.z + dlog[2] * wb.z;
          const float g3 = dlog[0] * wr.w + dlog[1] * wg.w + dlog[2] * wb.w;
          ph[2 * j4] = gate_pack(g0, g1, hmask, 2 * j4);
          ph[2 * j4 + 1] = gate_pack(g2, g3, hmask, 2 * j4 + 1);
        }
        tmem_st16(tlane + COL_AHI + (dcol >> 1), ph);   // K = direction-layer feature, natural order (2 K blocks)
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_warp(&a_kb[0]);
        mbar_arrive_warp(&a_kb[1]);
        if (tid == 0) P1_STAMP(16 + 8 * 9 + 5);
        staging_acquire();
        put(tsc, ZB_DIR, (int)(dcol >> 6), dcol, ph);
        put(tsc, XB_HDT, 2 + (int)(dcol >> 6), dcol, ph_hdt);
        staging_release();
      }
      named_bar_sync(1, P1_EPI_THREADS);  // dsig_s visible to every column quarter

      // ---- dgrad epilogues, layer 7 down to 0: dZ_l = dH_{l+1} * gate_l ----
      for (int l = NERFW_LAYERS - 1; l >= 0; --l) {
        // gate words for this layer, fetched before the accumulator wait so their latency is hidden behind the MMAs
        uint32_t gate[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)
          gate[q] = fwd_masks ? __ldg(fwd_masks + mask_index(tile, l, row, 0, (int)(cq * 2 + q))) : 0u;
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (tid == 0) P1_STAMP(16 + 8 * (10 + (NERFW_LAYERS - 1 - l)) + 4);
        const float ds = dsig_s[row];
        const uint64_t ds2 = pack2f(ds, ds);
        uint32_t r[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld16(tlane + COL_ACC + cq * 64 + j * 16, r[j]);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive_warp(acc_free);
        uint32_t ph2[2][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t col = cq * 64 + j * 16;
          const uint32_t word = fwd_masks ? gate[j >> 1] : mask_words(sm, l, row)[col >> 5];
          const uint32_t bits = word >> ((j & 1) * 16);
          uint32_t* ph = &ph2[j >> 1][(j & 1) * 8];
          if (l == NERFW_LAYERS - 1) {  // + density head: d sigma_pre * w_sigma
            const float2* w2 = reinterpret_cast<const float2*>(vec + V_DENW + col);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 ww = w2[e];
              float g0, g1;
              unpack2f(fma2(ds2, pack2f(ww.x, ww.y), pack2(r[j][2 * e], r[j][2 * e + 1])), g0, g1);
              ph[e] = gate_pack(g0, g1, bits, e);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) ph[e] = gate_pack(__uint_as_float(r[j][2 * e]), __uint_as_float(r[j][2 * e + 1]), bits, e);
          }
          if (l > 0) {
            const uint32_t p8[8] = {ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]};
            tmem_st8(tlane + COL_AHI + ((j * 64 + cq * 16) >> 1), p8);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&a_kb[j]);
            if (tid == 0 && j == 0) P1_STAMP(16 + 8 * (10 + (NERFW_LAYERS - 1 - l)) + 5);
            if (tid == 0 && j == 3) P1_STAMP(16 + 8 * (10 + (NERFW_LAYERS - 1 - l)) + 6);
          }
        }
        // accumulator and operand columns are free and the next tile's encodings are in place (written during the dgrad
        // phase): its layer 0 may start while this tile's last dZ block is still being stored
        if (l == 0 && tile + gridDim.x < ntiles) mbar_arrive_warp(a_ready);
        staging_acquire();
#pragma unroll
        for (int q = 0; q < 2; ++q) put(tsc, ZB(l), (int)cq, cq * 64 + q * 32, ph2[q]);
        staging_release();
        if (tid == 0) P1_STAMP(16 + 8 * (10 + (NERFW_LAYERS - 1 - l)) + 7);
        if (l == NERFW_LAYERS - 2 && tile + gridDim.x < ntiles) encode_tile(tile + gridDim.x);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P1_MMA_WARP) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}

// ====================================================================================================================
// pass 2: weight gradients
struct WgradBlock {
  int dz_block;     // first dZ block in the scratch tile
  int n_mhalves;    // outputs / 128 (1 or 2)
  int x_block;      // first X block (-1: no MMA, rgb-head reduction only)
  int x_nblocks;    // input features / 64
  int n_valid;      // valid input features (<= 64 * x_nblocks)
  int ld;           // row pitch of the destination weight matrix
  int k0;           // first destination column
  int flags;        // 1: bias column sums, 2: density head (needs dls + X = H8), 4: rgb head block
  float* dW;        // [outputs][ld]
  float* db;        // [outputs] or null
};
constexpr int MAX_UNITS = 148;
struct WgradPlan {
  int n_units;
  int unit_block[MAX_UNITS];
  int unit_t0[MAX_UNITS];
  int unit_t1[MAX_UNITS];
  WgradBlock blocks[13];
  float* d_density_w;
  float* d_density_b;
  float* d_rgb_w;
  float* d_rgb_b;
  int debug;  // bit 0: skip MMAs, bit 1: skip CUDA-core reductions (profiling only)
};

constexpr int W2_THREADS = 320;          // warp 0 producer, warp 1 MMA, warps 2..9 reducers / flush
constexpr int W2_STAGES = 3;
constexpr uint32_t W2_PIECE = 8192;      // 64 samples x 64 features
constexpr uint32_t W2_DZ = 0;            // up to 4 pieces
constexpr uint32_t W2_X = 4 * W2_PIECE;  // up to 4 pieces
constexpr uint32_t W2_DLS = 8 * W2_PIECE;            // 64 float4
constexpr uint32_t W2_STAGE = 8 * W2_PIECE + 2048;   // 67 584 B, 1024-aligned
constexpr uint32_t W2_BAR = W2_STAGES * W2_STAGE;
constexpr size_t SMEM2_BYTES = W2_BAR + 128 + 1024;

// MN-major operand tile: rows = K (samples), 64 features (128 B) per row, 8-row groups 1024 B apart (SBO), consecutive
// 64-feature blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __host__ constexpr uint32_t idesc_bf16_mn(uint32_t M, uint32_t N) { return idesc_bf16(M, N) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

__global__ void __launch_bounds__(W2_THREADS, 1) mlp_tc_wgrad_kernel(const __grid_constant__ WgradPlan plan,
                                                                     const uint8_t* __restrict__ scratch) {
  if ((int)blockIdx.x >= plan.n_units) return;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + W2_BAR);
  uint64_t* empty = full + W2_STAGES;
  uint64_t* done = empty + W2_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const WgradBlock& wb = plan.blocks[plan.unit_block[blockIdx.x]];
  const int t0 = plan.unit_t0[blockIdx.x], t1 = plan.unit_t1[blockIdx.x];
  const bool has_mma = wb.x_block >= 0 && !(wb.flags & 4) && !(plan.debug & 1);
  const int rflags = (plan.debug & 2) ? 0 : wb.flags;
  const int n_dz = 2 * wb.n_mhalves;
  const int N = 64 * wb.x_nblocks;
  const bool need_dls = (wb.flags & 6) != 0;

  if (tid == 0) {
    for (int i = 0; i < W2_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + 8); }  // MMA commit + one arrival per reducer warp
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const int n_stages_total = (t1 - t0) * 2;  // two 64-sample halves per tile

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_stages_total; ++it) {
        const int tile = t0 + (it >> 1), half = it & 1;
        const uint8_t* tsc = scratch + (size_t)tile * TILE_BYTES + (size_t)half * W2_PIECE;
        uint8_t* st = sm + stage * W2_STAGE;
        mbar_wait(&empty[stage], phase ^ 1);
        uint32_t bytes = (uint32_t)(n_dz + wb.x_nblocks) * W2_PIECE + (need_dls ? 1024u : 0u);
        mbar_arrive_expect_tx(&full[stage], bytes);
        for (int j = 0; j < n_dz; ++j) bulk_g2s(st + W2_DZ + j * W2_PIECE, tsc + (size_t)(wb.dz_block + j) * BLK, W2_PIECE, &full[stage]);
        const int xb = wb.x_block >= 0 ? wb.x_block : 0;
        for (int j = 0; j < wb.x_nblocks; ++j) bulk_g2s(st + W2_X + j * W2_PIECE, tsc + (size_t)(xb + j) * BLK, W2_PIECE, &full[stage]);
        if (need_dls)
          bulk_g2s(st + W2_DLS, scratch + (size_t)tile * TILE_BYTES + (size_t)DLS_BLOCK * BLK + (size_t)half * 1024, 1024, &full[stage]);
        if (++stage == W2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = idesc_bf16_mn(128, (uint32_t)N);
      for (int it = 0; it < n_stages_total; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (has_mma) {
          const uint32_t st = smem_u32(sm + stage * W2_STAGE);
          const uint64_t b = smem_desc_sw128_mn(st + W2_X, W2_PIECE);
          for (int h = 0; h < wb.n_mhalves; ++h) {
            const uint64_t a = smem_desc_sw128_mn(st + W2_DZ + (uint32_t)h * 2 * W2_PIECE, W2_PIECE);
            for (int ks = 0; ks < 4; ++ks)  // 64 samples = 4 K steps of 16 rows (2 KB each)
              mma_ss(tmem + (uint32_t)(h * N), a + 128 * ks, b + 128 * ks, idesc, (it | ks) ? 1u : 0u);
          }
          mma_commit(&empty[stage]);
        } else {
          mbar_arrive(&empty[stage]);
        }
        if (++stage == W2_STAGES) { stage = 0; phase ^= 1; }
      }
      if (has_mma) mma_commit(done); else mbar_arrive(done);
    }
  } else {
    // ---- reducers (8 warps): bias / head gradients as column reductions of the shared-memory tiles, 16 bytes
    // (8 bf16 columns) per load, 8 rows of every 64-row stage per thread; then the accumulator flush ----
    const int rt = tid - 64;           // 0..255
    const int grp = rt & 31;           // 8-column group: columns 8 grp .. 8 grp + 7  (256 columns = 4 pieces x 8 chunks)
    const int rset = rt >> 5;          // rows 8 rset .. 8 rset + 7 of the stage
    float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // bias sums, or density-head sums (flag 2 uses acc8b)
    float acc8b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float rg[3][8];                                                // rgb head: 3 channels x 8 columns
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) rg[c][j] = 0.f;
    float hb[4] = {0.f, 0.f, 0.f, 0.f};  // sums of d logits (x,y,z) and d sigma_pre over this thread's rows (grp == 0 only)
    int stage = 0;
    uint32_t phase = 0;
    const int ncols_out = 128 * wb.n_mhalves;
    auto unpack8 = [](const uint4& v, float (&f)[8]) {
      f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
      f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
      f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
      f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
    };
    for (int it = 0; it < n_stages_total; ++it) {
      mbar_wait(&full[stage], phase);
      const uint8_t* st = sm + stage * W2_STAGE;
      const float4* dls = reinterpret_cast<const float4*>(st + W2_DLS);
      if (rflags & 1) {
        // 256 output columns: lane <-> 8-column group, 8 rows each; 128 columns: the two half-warps split the rows
        const int g = ncols_out == 256 ? grp : (grp & 15);
        const int r0 = ncols_out == 256 ? 0 : 4 * (grp >> 4), nr = ncols_out == 256 ? 8 : 4;
        const uint8_t* piece = st + W2_DZ + (g >> 3) * W2_PIECE;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nr) {
            const uint32_t r = rset * 8 + r0 + i;
            const uint4 v = *reinterpret_cast<const uint4*>(piece + sw128_offset(r, (uint32_t)(g & 7) * 8));
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc8[j] += f[j];
          }
        }
      }
      if (rflags & 2) {  // d density_w[k] += sum_s dsig[s] h8[s][k]   (X = H8, 256 columns)
        const uint8_t* piece = st + W2_X + (grp >> 3) * W2_PIECE;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t r = rset * 8 + i;
          const uint4 v = *reinterpret_cast<const uint4*>(piece + sw128_offset(r, (uint32_t)(grp & 7) * 8));
          const float ds = dls[r].w;
          float f[8];
          unpack8(v, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc8b[j] = fmaf(ds, f[j], acc8b[j]);
          if (grp == 0) hb[3] += ds;
        }
      }
      if (rflags & 4) {  // d rgb_w[c][k] += sum_s dlog[s][c] hdt[s][k]   (X = hd + appearance, 128 columns = 16 groups;
                         // the two half-warps take rows 0..3 / 4..7 of the row set)
        const int g = grp & 15;
        const uint8_t* piece = st + W2_X + (g >> 3) * W2_PIECE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t r = rset * 8 + 4 * (grp >> 4) + i;
          const uint4 v = *reinterpret_cast<const uint4*>(piece + sw128_offset(r, (uint32_t)(g & 7) * 8));
          const float4 dl = dls[r];
          float f[8];
          unpack8(v, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            rg[0][j] = fmaf(dl.x, f[j], rg[0][j]);
            rg[1][j] = fmaf(dl.y, f[j], rg[1][j]);
            rg[2][j] = fmaf(dl.z, f[j], rg[2][j]);
          }
          if (g == 0) { hb[0] += dl.x; hb[1] += dl.y; hb[2] += dl.z; }
        }
      }
      mbar_arrive_warp(&empty[stage]);
      if (++stage == W2_STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- flush the CUDA-core partial sums (each of the 8 row sets holds a partial of the same columns) ----
    if (rflags & 1) {
      const int g = ncols_out == 256 ? grp : (grp & 15);
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(wb.db + 8 * g + j, acc8[j]);
    }
    if (rflags & 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(plan.d_density_w + 8 * grp + j, acc8b[j]);
      if (grp == 0) atomicAdd(plan.d_density_b, hb[3]);
    }
    if (rflags & 4) {
      const int g = grp & 15;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(plan.d_rgb_w + c * 128 + 8 * g + j, rg[c][j]);
      if (g == 0) { atomicAdd(plan.d_rgb_b + 0, hb[0]); atomicAdd(plan.d_rgb_b + 1, hb[1]); atomicAdd(plan.d_rgb_b + 2, hb[2]); }
    }
    mbar_wait(done, 0);
    tc_fence_after();
    if (has_mma && n_stages_total > 0) {
      const uint32_t quad = warp & 3;
      const int chalf = (warp - 2) >> 2;  // the two warps of a lane quadrant take alternate 32-column groups
      const uint32_t tl = tmem + ((quad * 32) << 16);
      for (int h = 0; h < wb.n_mhalves; ++h) {
        const int o = h * 128 + quad * 32 + lane;
        float* dst = wb.dW + (size_t)o * wb.ld + wb.k0;
        for (int c0 = 32 * chalf; c0 < N; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(tl + (uint32_t)(h * N + c0), r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < wb.n_valid) atomicAdd(dst + c0 + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}

// appearance branch, shared embedding: G = W_rgb^T DL; dW_app[k][q] += G[k] e[q]; db_app[k] += G[k];
// d_emb[q] += sum_k G[k] W_app[k][q].  (dW_rgb already contains the appearance term: pass 2 reduces against hd + a.)
__global__ void __launch_bounds__(128) app_bwd_shared_kernel(NerfwWeights w, NerfwGrads g, const float* __restrict__ emb,
                                                             const float* __restrict__ app_vec,
                                                             const float* __restrict__ dl, float* __restrict__ d_emb) {
  __shared__ float gk[NERFW_DIR_HIDDEN];
  const int k = threadIdx.x;
  const float d0 = dl[0], d1 = dl[1], d2 = dl[2];
  const float G = d0 * __ldg(w.rgb_w + k) + d1 * __ldg(w.rgb_w + 128 + k) + d2 * __ldg(w.rgb_w + 256 + k);
  gk[k] = G;
  (void)app_vec;
  for (int q = 0; q < NERFW_APP_DIM; ++q) atomicAdd(g.app_w + k * NERFW_APP_DIM + q, G * __ldg(emb + q));
  atomicAdd(g.app_b + k, G);
  __syncthreads();
  if (d_emb && k < NERFW_APP_DIM) {
    float de = 0.f;
    for (int kk = 0; kk < NERFW_DIR_HIDDEN; ++kk) de = fmaf(gk[kk], __ldg(w.app_w + kk * NERFW_APP_DIM + k), de);
    atomicAdd(d_emb + k, de);
  }
}

// a[row][k] = W_app e_row + b_app (128 floats per embedding row)
__global__ void __launch_bounds__(128) app_vec_kernel(NerfwWeights w, const float* __restrict__ emb, int64_t rows,
                                                      float* __restrict__ out) {
  const int k = threadIdx.x;
  float wk[NERFW_APP_DIM];
#pragma unroll
  for (int q = 0; q < NERFW_APP_DIM; ++q) wk[q] = __ldg(w.app_w + k * NERFW_APP_DIM + q);
  const float bk = __ldg(w.app_b + k);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    float a = bk;
#pragma unroll
    for (int q = 0; q < NERFW_APP_DIM; ++q) a = fmaf(wk[q], __ldg(emb + row * NERFW_APP_DIM + q), a);
    out[row * NERFW_DIR_HIDDEN + k] = a;
  }
}

// appearance branch, one embedding row per ray: per row r, G_r = W_rgb^T DL_r (DL_r = sum of the ray's d logits);
// dW_app[k][q] += sum_r G_r[k] e_r[q]; db_app[k] += sum_r G_r[k]; d_emb[r][q] += sum_k G_r[k] W_app[k][q].
// A CTA walks a contiguous block of rows with thread k keeping its row of dW_app in registers: one flush per CTA.
__global__ void __launch_bounds__(128) app_bwd_rows_kernel(NerfwWeights w, NerfwGrads g, const float* __restrict__ emb,
                                                           int64_t rows, const float* __restrict__ dl,
                                                           float* __restrict__ d_emb) {
  __shared__ float gk[NERFW_DIR_HIDDEN];
  __shared__ float e[NERFW_APP_DIM];
  const int k = threadIdx.x;
  const float r0 = __ldg(w.rgb_w + k), r1 = __ldg(w.rgb_w + 128 + k), r2 = __ldg(w.rgb_w + 256 + k);
  float acc[NERFW_APP_DIM];
#pragma unroll
  for (int q = 0; q < NERFW_APP_DIM; ++q) acc[q] = 0.f;
  float bsum = 0.f;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t rbeg = blockIdx.x * per, rend = rbeg + per < rows ? rbeg + per : rows;
  for (int64_t row = rbeg; row < rend; ++row) {
    __syncthreads();   // previous row's gk / e consumed
    if (k < NERFW_APP_DIM) e[k] = __ldg(emb + row * NERFW_APP_DIM + k);
    const float G = dl[4 * row] * r0 + dl[4 * row + 1] * r1 + dl[4 * row + 2] * r2;
    gk[k] = G;
    __syncthreads();
    bsum += G;
#pragma unroll
    for (int q = 0; q < NERFW_APP_DIM; ++q) acc[q] = fmaf(G, e[q], acc[q]);
    if (d_emb && k < NERFW_APP_DIM) {
      float de = 0.f;
      for (int kk = 0; kk < NERFW_DIR_HIDDEN; ++kk) de = fmaf(gk[kk], __ldg(w.app_w + kk * NERFW_APP_DIM + k), de);
      atomicAdd(d_emb + row * NERFW_APP_DIM + k, de);
    }
  }
  if (rend > rbeg) {
#pragma unroll
    for (int q = 0; q < NERFW_APP_DIM; ++q) atomicAdd(g.app_w + k * NERFW_APP_DIM + q, acc[q]);
    atomicAdd(g.app_b + k, bsum);
  }
}

// MN-major self-test: D (128 x N) = At^T Bt with At (K x 128) and Bt (K x N) row-major bf16 (K, N multiples of 64)
__global__ void __launch_bounds__(128, 1) umma_selftest_mn_kernel(const __nv_bfloat16* __restrict__ At,
                                                                  const __nv_bfloat16* __restrict__ Bt, int N, int K,
                                                                  float* __restrict__ D) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint8_t* sA = sm;            // 2 blocks of K rows x 128 B
  uint8_t* sB = sm + 65536;    // N/64 blocks of K rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536 + 131072);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t blk = (uint32_t)K * 128u;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  for (int idx = tid; idx < K * 128; idx += 128) {
    int k = idx / 128, m = idx % 128;
    *reinterpret_cast<__nv_bfloat16*>(sA + (m / 64) * blk + sw128_offset(k, m % 64)) = At[idx];
  }
  for (int idx = tid; idx < K * N; idx += 128) {
    int k = idx / N, n = idx % N;
    *reinterpret_cast<__nv_bfloat16*>(sB + (n / 64) * blk + sw128_offset(k, n % 64)) = Bt[idx];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16_mn(128, (uint32_t)N);
    const uint64_t a = smem_desc_sw128_mn(smem_u32(sA), blk);
    const uint64_t b = smem_desc_sw128_mn(smem_u32(sB), blk);
    for (int ks = 0; ks < K / 16; ++ks) mma_ss(tmem, a + 128 * ks, b + 128 * ks, idesc, ks ? 1u : 0u);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  {
    const uint32_t tl = tmem + ((warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tl + c0, r);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Issue-rate probe: `reps` x 16 MMAs (M=128, N, K=16) over resident operand tiles, cycles measured with clock64.
// mode 0: K-major SS, 1: K-major TS (A in TMEM), 2: MN-major SS; 3 / 4: kind::i8 (s8 x s8 -> s32, K = 32) SS / TS; 5: kind::f8f6f4
// (e4m3, K = 32) SS.  Operand contents are irrelevant (zeros).
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int mode, int N, int reps, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536 + 131072);   // [0] end of run, [2] a completed phase, [3] commit sink
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 131072) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 2, 1); mbar_init(bar + 3, 1); fence_mbar_init(); mbar_arrive(bar + 2); }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (tid == 0) {
    const uint32_t sa = smem_u32(sm), sb = smem_u32(sm + 65536);
    // the mode is dispatched OUTSIDE the issue loop: a six-way branch per instruction would itself bound the measured rate
    const uint32_t n32 = (uint32_t)N;
    auto run = [&](auto issue) {
      for (int r = 0; r < reps; ++r)
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) issue(kb, k);
    };
    long long t0 = clock64();
    switch (mode) {
      case 0: run([&](int kb, int k) { mma_ss(tmem, smem_desc_sw128(sa + kb * 16384) + 2 * k, smem_desc_sw128(sb + kb * BIG_CHUNK) + 2 * k, idesc_bf16(128, n32), 1u); }); break;
      case 1: run([&](int kb, int k) { mma_ts(tmem, tmem + COL_AHI + 32 * kb + 8 * k, smem_desc_sw128(sb + kb * BIG_CHUNK) + 2 * k, idesc_bf16(128, n32), 1u); }); break;
      case 2: run([&](int kb, int k) { mma_ss(tmem, smem_desc_sw128_mn(sa, 8192) + 128 * (kb * 4 + k) % 512, smem_desc_sw128_mn(sb, 8192) + 128 * ((kb * 4 + k) % 4), idesc_bf16_mn(128, n32), 1u); }); break;
      // 8-bit kinds: 128 x N x 32 per instruction (twice the K of the 16-bit kinds over the same 32 operand bytes per row)
      case 3: run([&](int kb, int k) { mma_ss_i8(tmem, smem_desc_sw128(sa + kb * 16384) + 2 * k, smem_desc_sw128(sb + kb * BIG_CHUNK) + 2 * k, idesc_s8(128, n32), 1u); }); break;
      case 4: run([&](int kb, int k) { mma_ts_i8(tmem, tmem + COL_AHI + 32 * kb + 8 * k, smem_desc_sw128(sb + kb * BIG_CHUNK) + 2 * k, idesc_s8(128, n32), 1u); }); break;
      case 6:   // A from tensor memory with the production kernels' per-K-block protocol around every four instructions:
                // wait on an (already complete) full barrier, tcgen05 fence, four MMAs, commit to a ring-stage barrier
        for (int r = 0; r < reps; ++r)
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar + 2, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_ts(tmem, tmem + COL_AHI + 32 * kb + 8 * k, smem_desc_sw128(sb + kb * BIG_CHUNK) + 2 * k, idesc_bf16(128, n32), 1u);
            mma_commit(bar + 3);
          }
        break;
      default: run([&](int kb, int k) { mma_ss_f8(tmem, smem_desc_sw128(sa + kb * 16384) + 2 * k, smem_desc_sw128(sb + kb * BIG_CHUNK) + 2 * k, idesc_f16(128, n32), 1u); }); break;
    }
    mma_commit(bar);
    mbar_wait(bar, 0);
    cycles[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace tcb

size_t mlp_tc_packed_total_bytes() { return tcb::PACKED_TOTAL; }
size_t mlp_tc_packed_f16_offset() { return tcb::PACKED_F16_OFFSET; }

int launch_pack_weights_t(const NerfwWeights& w, void* packed, cudaStream_t stream) {
  const int64_t total = (int64_t)tcb::NT_CHUNKS * 256 * 8;
  tcb::pack_weights_t_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(
      w, reinterpret_cast<uint8_t*>(packed) + tcb::PACKED_T_OFFSET);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

}  // namespace nerfw

using namespace nerfw;

// workspace: [4096-byte header: shared-embedding app_off / dl_acc / app_vec][scratch tiles][per-ray appearance rows:
// app_off float4 | dl_acc float4 | app_vec 128 floats, for n_rays rows (used when emb_rows == n_rays)]
static size_t bwd_tc_rows_bytes(int64_t n_rays) { return (size_t)n_rays * (16 + 16 + NERFW_DIR_HIDDEN * sizeof(float)); }
extern "C" size_t nerfw_mlp_bwd_tc_workspace_bytes(int64_t n_rays, int n_samples) {
  const int64_t total = n_rays * (int64_t)(n_samples > 0 ? n_samples : 1);
  const int64_t ntiles = ceil_div64(total, tc::TM);
  return 4096 + (size_t)ntiles * tcb::TILE_BYTES + bwd_tc_rows_bytes(n_rays);
}

// Same contract as nerfw_mlp_bwd (include/nerfw.h), tensor-core arithmetic; no, one shared or one embedding row per ray.
extern "C" int nerfw_mlp_bwd_tc(const NerfwWeights* w, const void* packed, const float* pts_or_o, const float* dirs,
                                const float* z, const float* emb, int64_t emb_rows, int64_t n_rays, int n_samples,
                                const float* d_raw, const void* relu_masks, const NerfwGrads* grads, float* d_emb,
                                void* workspace, size_t workspace_bytes, void* stream) {
  NERFW_REQUIRE(w && grads && packed, "nerfw_mlp_bwd_tc: null weights, grads or packed weights");
  for (int i = 0; i < NERFW_LAYERS; ++i)
    NERFW_REQUIRE(w->pts_w[i] && w->pts_b[i] && grads->pts_w[i] && grads->pts_b[i], "nerfw_mlp_bwd_tc: null pts_linears.%d parameter or gradient", i);
  NERFW_REQUIRE(grads->density_w && grads->density_b && grads->dir_w && grads->dir_b && grads->rgb_w && grads->rgb_b,
                "nerfw_mlp_bwd_tc: null head gradient");
  NERFW_REQUIRE(n_rays >= 0 && n_samples >= 1, "nerfw_mlp_bwd_tc: bad shape");
  if (n_rays == 0) return NERFW_OK;
  NERFW_REQUIRE(z || n_samples == 1, "nerfw_mlp_bwd_tc: n_samples must be 1 when z is NULL");
  NERFW_REQUIRE(pts_or_o && dirs && d_raw && workspace, "nerfw_mlp_bwd_tc: null pointer");
  NERFW_REQUIRE(aligned16(d_raw) && aligned16(workspace), "nerfw_mlp_bwd_tc: d_raw and workspace must be 16-byte aligned");
  if (emb) {
    NERFW_REQUIRE(emb_rows == 1 || emb_rows == n_rays, "nerfw_mlp_bwd_tc: emb_rows=%lld must be 1 or n_rays=%lld",
                  (long long)emb_rows, (long long)n_rays);
    NERFW_REQUIRE(w->app_w && w->app_b && grads->app_w && grads->app_b, "nerfw_mlp_bwd_tc: embedding given but appearance parameters/gradients are null");
  }
  const size_t need = nerfw_mlp_bwd_tc_workspace_bytes(n_rays, z ? n_samples : 1);
  if (workspace_bytes < need) {
    set_error("nerfw_mlp_bwd_tc: workspace of %zu bytes, need %zu", workspace_bytes, need);
    return NERFW_ESIZE;
  }
  cudaStream_t st = as_stream(stream);
  // header: [app_off float4][dl_acc 4 floats][app_vec 128 floats] ... tiles from +4096
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  float* app_off = reinterpret_cast<float*>(base);
  float* dl_acc = reinterpret_cast<float*>(base + 16);
  float* app_vec = reinterpret_cast<float*>(base + 64);
  uint8_t* scratch = base + 4096;
  NERFW_CUDA(cudaMemsetAsync(base, 0, 4096, st));
  const bool per_ray = emb && emb_rows > 1;
  if (per_ray) {   // one (app_off, dl_acc, app_vec) row per ray, behind the scratch tiles
    uint8_t* rows_base = base + (need - bwd_tc_rows_bytes(n_rays));
    app_off = reinterpret_cast<float*>(rows_base);
    dl_acc = reinterpret_cast<float*>(rows_base + (size_t)n_rays * 16);
    app_vec = reinterpret_cast<float*>(rows_base + (size_t)n_rays * 32);
    NERFW_CUDA(cudaMemsetAsync(dl_acc, 0, (size_t)n_rays * 16, st));
  }
  if (emb) {
    int rc = launch_app_offset(*w, emb, emb_rows, app_off, st);
    if (rc) return rc;
    tcb::app_vec_kernel<<<(unsigned)(emb_rows < 1184 ? emb_rows : 1184), 128, 0, st>>>(*w, emb, emb_rows, app_vec);
    NERFW_LAUNCHED();
  }
  SampleSource src;
  src.p = pts_or_o;
  src.d = dirs;
  src.z = z;
  src.emb = emb;
  src.n_per_ray = z ? n_samples : 1;
  src.emb_shared = per_ray ? 0 : 1;
  const int64_t total = n_rays * (z ? n_samples : 1);
  const int64_t ntiles = ceil_div64(total, tc::TM);
  NERFW_REQUIRE(ntiles < (1ll << 30), "nerfw_mlp_bwd_tc: too many samples");

  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(tcb::mlp_tc_bwd_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcb::SMEM1_BYTES));
    NERFW_CUDA(cudaFuncSetAttribute(tcb::mlp_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcb::SMEM2_BYTES));
  }
  const int sms = sm_count();
  int64_t grid1 = ntiles < sms ? ntiles : sms;
  int dbg_all = 0;                 // profiling switches (results are wrong when set)
  long long* timeline = nullptr;   // profiling: device pointer to clock64 slots
#ifdef NERFW_PROFILE
  // Only in the separate profiling build (make PROFILE=1 -> libnerfw_sm100_profile.so): the product library never reads
  // the environment.
  if (const char* e = getenv("NERFW_WGRAD_DEBUG")) dbg_all = atoi(e);
  if (const char* e = getenv("NERFW_BWD_TIMELINE")) timeline = reinterpret_cast<long long*>(strtoull(e, nullptr, 10));
#endif
  if (!(dbg_all & 16))   // bit 4: wgrad kernel only
  tcb::mlp_tc_bwd_pass1_kernel<<<(unsigned)grid1, tcb::P1_THREADS, tcb::SMEM1_BYTES, st>>>(
      reinterpret_cast<const uint8_t*>(packed), src, emb ? reinterpret_cast<const float4*>(app_off) : nullptr,
      emb ? app_vec : nullptr, reinterpret_cast<const float4*>(d_raw), total, scratch, emb ? dl_acc : nullptr,
      reinterpret_cast<const uint32_t*>(relu_masks), dbg_all, timeline);
  NERFW_LAUNCHED();

  // ---- pass-2 plan: one weight block and one contiguous tile range per CTA, CTAs shared out by bytes per tile ----
  tcb::WgradPlan plan;
  memset(&plan, 0, sizeof(plan));
  int nb = 0;
  auto add = [&](int dzb, int mh, int xb, int xn, int nvalid, int ld, int k0, int flags, float* dW, float* db) {
    tcb::WgradBlock& b = plan.blocks[nb++];
    b.dz_block = dzb; b.n_mhalves = mh; b.x_block = xb; b.x_nblocks = xn; b.n_valid = nvalid; b.ld = ld; b.k0 = k0;
    b.flags = flags; b.dW = dW; b.db = db;
  };
  add(tcb::ZB(0), 2, tcb::XB_ENCX, 1, NERFW_POS_DIM, NERFW_POS_DIM, 0, 1, grads->pts_w[0], grads->pts_b[0]);
  for (int l = 1; l < NERFW_LAYERS; ++l) {
    const int ld = (l == NERFW_SKIP) ? 256 + NERFW_POS_DIM : 256;
    add(tcb::ZB(l), 2, tcb::XB_H(l), 4, 256, ld, 0, 1, grads->pts_w[l], grads->pts_b[l]);
    if (l == NERFW_SKIP) add(tcb::ZB(l), 2, tcb::XB_ENCX, 1, NERFW_POS_DIM, ld, 256, 0, grads->pts_w[l], nullptr);
  }
  add(tcb::ZB_DIR, 1, tcb::XB_H(8), 4, 256, 256 + NERFW_DIR_DIM, 0, 1 | 2, grads->dir_w, grads->dir_b);
  add(tcb::ZB_DIR, 1, tcb::XB_ENCD, 1, NERFW_DIR_DIM, 256 + NERFW_DIR_DIM, 256, 0, grads->dir_w, nullptr);
  add(tcb::ZB_DIR, 0, tcb::XB_HDT, 2, 128, 128, 0, 4, grads->rgb_w, grads->rgb_b);
  plan.d_density_w = grads->density_w;
  plan.d_density_b = grads->density_b;
  plan.d_rgb_w = grads->rgb_w;
  plan.d_rgb_b = grads->rgb_b;
  plan.debug = dbg_all;
  double cost[13], csum = 0;
  for (int b = 0; b < nb; ++b) {
    const tcb::WgradBlock& wb = plan.blocks[b];
    // units are handed out in proportion to the time of one 64-sample stage: HBM bytes (8 KB pieces; the kernel streams
    // at ~80 % of HBM peak) or, for the rgb-head block, its CUDA-core reduction (24 FMAs per element; measured ~2x its bytes)
    cost[b] = 2.0 * wb.n_mhalves + wb.x_nblocks + 0.25;
    if (wb.flags & 4) cost[b] *= 2.0;
    csum += cost[b];
  }
  const int max_units = sms < tcb::MAX_UNITS ? sms : tcb::MAX_UNITS;
  int alloc[13], used = 0;
  for (int b = 0; b < nb; ++b) {
    int c = (int)(cost[b] / csum * max_units);
    if (c < 1) c = 1;
    if (c > ntiles) c = (int)ntiles;
    alloc[b] = c;
    used += c;
  }
  while (used > max_units) {  // trim the largest allocations
    int bi = 0;
    for (int b = 1; b < nb; ++b) if (alloc[b] > alloc[bi]) bi = b;
    if (alloc[bi] <= 1) break;
    --alloc[bi]; --used;
  }
  for (bool grew = true; grew && used < max_units;) {  // hand out the remainder to the most loaded blocks
    grew = false;
    int bi = -1;
    double worst = 0;
    for (int b = 0; b < nb; ++b) {
      if (alloc[b] >= ntiles) continue;
      double load = cost[b] / alloc[b];
      if (load > worst) { worst = load; bi = b; }
    }
    if (bi >= 0) { ++alloc[bi]; ++used; grew = true; }
  }
  int u = 0;
  for (int b = 0; b < nb; ++b) {
    for (int c = 0; c < alloc[b]; ++c) {
      int64_t t0 = ntiles * c / alloc[b], t1 = ntiles * (c + 1) / alloc[b];
      if (t1 <= t0) continue;
      plan.unit_block[u] = b;
      plan.unit_t0[u] = (int)t0;
      plan.unit_t1[u] = (int)t1;
      ++u;
    }
  }
  plan.n_units = u;
  if (!(dbg_all & 8))    // bit 3: pass 1 only
  tcb::mlp_tc_wgrad_kernel<<<(unsigned)u, tcb::W2_THREADS, tcb::SMEM2_BYTES, st>>>(plan, scratch);
  NERFW_LAUNCHED();
  if (per_ray) {
    const int64_t blocks = ceil_div64(emb_rows, 32);
    tcb::app_bwd_rows_kernel<<<(unsigned)(blocks < 592 ? blocks : 592), 128, 0, st>>>(*w, *grads, emb, emb_rows, dl_acc, d_emb);
    NERFW_LAUNCHED();
  } else if (emb) {
    tcb::app_bwd_shared_kernel<<<1, 128, 0, st>>>(*w, *grads, emb, app_vec, dl_acc, d_emb);
    NERFW_LAUNCHED();
  }
  return NERFW_OK;
}

// MN-major primitive self-test: D (128,n) fp32 = At^T Bt, At (k,128), Bt (k,n) bf16 row-major.
extern "C" int nerfw_selftest_umma_mn(const void* at_bf16, const void* bt_bf16, int n, int k, float* d, void* stream) {
  NERFW_REQUIRE(at_bf16 && bt_bf16 && d, "nerfw_selftest_umma_mn: null pointer");
  NERFW_REQUIRE(n >= 64 && n <= 256 && n % 64 == 0, "nerfw_selftest_umma_mn: N must be a multiple of 64 in [64,256]");
  NERFW_REQUIRE(k >= 16 && k <= 256 && k % 16 == 0, "nerfw_selftest_umma_mn: K must be a multiple of 16 in [16,256]");
  const size_t smem = 65536 + 131072 + 64 + 1024;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(tcb::umma_selftest_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  tcb::umma_selftest_mn_kernel<<<1, 128, smem, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(at_bf16),
                                                                    reinterpret_cast<const __nv_bfloat16*>(bt_bf16), n, k, d);
  NERFW_LAUNCHED();
  return NERFW_OK;
}

// Tensor-pipe issue-rate probe (tests / DESIGN.md numbers): cycles for reps x 16 MMAs of shape 128 x n x 16.
extern "C" int nerfw_selftest_umma_rate(int mode, int n, int reps, long long* cycles_dev, void* stream) {
  NERFW_REQUIRE(cycles_dev && mode >= 0 && mode <= 6 && n >= 16 && n <= 256 && n % 16 == 0 && reps >= 1, "nerfw_selftest_umma_rate: bad arguments");
  const size_t smem = 65536 + 131072 + 64 + 1024;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    NERFW_CUDA(cudaFuncSetAttribute(tcb::umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  tcb::umma_rate_kernel<<<1, 128, smem, as_stream(stream)>>>(mode, n, reps, cycles_dev);
  NERFW_LAUNCHED();
  return NERFW_OK;
}
