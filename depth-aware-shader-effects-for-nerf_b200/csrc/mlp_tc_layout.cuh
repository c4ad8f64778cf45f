// Layout constants shared by the tensor-core MLP kernels (forward: mlp_tc.cu, backward: mlp_tc_bwd.cu): tile size,
// packed weight stream, TMEM columns, shared-memory map of the forward pipeline.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace nerfw {
namespace tc {

using namespace umma;

constexpr int TM = 128;
constexpr int EPI_WARPS = 8;
constexpr int PRODUCER_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int THREADS = 320;
constexpr int EPI_THREADS = EPI_WARPS * 32;

constexpr uint32_t BIG_CHUNK = 256 * 128;    // 256 outputs x 64 bf16
constexpr uint32_t SMALL_CHUNK = 128 * 128;  // 128 outputs x 64 bf16
constexpr int N_BIG = 30;                    // L0:1, L1-3:12, L4:4+1, L5-7:12
constexpr int N_SMALL = 5;                   // dir layer: 4 + 1
constexpr int N_CHUNKS = N_BIG + N_SMALL;
constexpr size_t W_BYTES = 2ull * (N_BIG * (size_t)BIG_CHUNK + N_SMALL * (size_t)SMALL_CHUNK);  // hi and lo

// fp32 vector block behind the weight chunks
constexpr int V_PTSB = 0;       // 8 x 256
constexpr int V_DIRB = 2048;    // 128
constexpr int V_DENW = 2176;    // 256
constexpr int V_RGBW = 2432;    // 3 x 128
constexpr int V_DENB = 2816;    // 1
constexpr int V_RGBB = 2817;    // 3
constexpr int V_FLOATS = 2824;  // padded to 16 B
constexpr size_t PACKED_BYTES = W_BYTES + V_FLOATS * sizeof(float);
// fp16 copy of the forward image (single-pass fp16 mode), one copy per chunk, appended after the transposed image
constexpr size_t F16_BYTES = N_BIG * (size_t)BIG_CHUNK + N_SMALL * (size_t)SMALL_CHUNK;
__host__ __device__ inline size_t chunk_offset_f16(int i) {
  return i < N_BIG ? (size_t)i * BIG_CHUNK : (size_t)N_BIG * BIG_CHUNK + (size_t)(i - N_BIG) * SMALL_CHUNK;
}

// TMEM columns
constexpr uint32_t COL_ACC = 0;
constexpr uint32_t COL_AHI = 256;
constexpr uint32_t COL_ALO = 384;

// shared memory map (offsets from a 1024-aligned base)
constexpr int NSTAGES = 4;
constexpr uint32_t SM_PEX_HI = 0;
constexpr uint32_t SM_PEX_LO = 16384;
constexpr uint32_t SM_PED_HI = 32768;
constexpr uint32_t SM_PED_LO = 49152;
constexpr uint32_t SM_RING = 65536;
constexpr uint32_t SM_VEC = SM_RING + NSTAGES * BIG_CHUNK;        // 196608
constexpr uint32_t SM_SIG = SM_VEC + V_FLOATS * 4;                // [4][128] floats (density partial sums)
constexpr uint32_t SM_RGB = SM_SIG + 4 * TM * 4;                  // [4][128] float4 (rgb logit partial sums)
constexpr uint32_t SM_BAR = SM_RGB + 4 * TM * 16;                     // full[4], empty[4], acc_full, a_ready
constexpr uint32_t SM_TMEMPTR = SM_BAR + 64 * 8;   // room for 64 mbarriers (forward: 2 x 5 ring + 2 + 4 + 2)
constexpr uint32_t SM_TOTAL = SM_TMEMPTR + 16;
constexpr size_t SMEM_BYTES = SM_TOTAL + 1024;  // slack for the manual 1024-byte alignment

// Chunk i of the stream -> (source matrix, first input column).  Order = consumption order of the MMA thread.
struct ChunkSrc {
  int layer;  // 0..7 trunk, 8 = dir layer
  int col0;   // first input column of this 64-wide K block
};
__host__ __device__ inline ChunkSrc chunk_source(int i) {
  if (i == 0) return {0, 0};
  if (i < 13) return {1 + (i - 1) / 4, ((i - 1) % 4) * 64};
  if (i < 18) return {4, (i - 13) * 64};  // i == 17: columns 256.. = enc_x part of the skip layer
  if (i < 30) return {5 + (i - 18) / 4, ((i - 18) % 4) * 64};
  return {8, (i - 30) * 64};              // i == 34: columns 256.. = enc_d part
}
// K order inside the 256 hidden features of an operand produced by a trunk epilogue.  An epilogue thread owns 64
// CONTIGUOUS accumulator columns (column quarter cq: columns 64 cq .. 64 cq + 63, so its activation / gradient rows are
// stored to the scratch tiles in 64-byte pieces) but hands them to the MMA thread in four 16-column granules j; granule j
// of all four quarters forms K block j of the next MMA.  Position p = 16 cq + i of K block j therefore holds feature
// 64 cq + 16 j + i, and every weight image consumed with such an operand is packed in the same order.
__host__ __device__ inline int kperm_feature(int kblock, int pos) { return 64 * (pos >> 4) + 16 * kblock + (pos & 15); }
__host__ __device__ inline size_t chunk_offset(int i) {  // byte offset of the hi copy; lo follows at +size
  return i < N_BIG ? 2ull * i * BIG_CHUNK : 2ull * N_BIG * BIG_CHUNK + 2ull * (i - N_BIG) * SMALL_CHUNK;
}

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
  }
};

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FFMA2) and fused ReLU + bf16x2 conversion ------------------------
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) { return pack2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a);
  hi = __uint_as_float(b);
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// {bf16(max(lo_elem,0)), bf16(max(hi_elem,0))} packed, lo_elem in the low half (one F2FP.RELU)
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// the same to fp16, saturating (F2FP.SATFINITE.RELU.F16)
__device__ __forceinline__ uint32_t relu_pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// ---- positional encodings of one sample row, 32 consecutive features at a time (src/models.py:35-44) ---------------
// Feature k of the 64-wide position tile: k < 3: x_k; k = 3 + 6 l + c: sin(2^l x_c); k = 6 + 6 l + c: cos(2^l x_c); 63: 0.
// FAST (bf16 mode): level 0 by sincosf, higher levels by angle doubling (abs. error ~2^l ulp <= 1e-4, far below the bf16
// rounding of the operand); otherwise every level by sincosf on the exactly scaled argument, like the reference.
template <int HALF, bool FAST>
__device__ __forceinline__ void pos_features32(const float (&x)[3], float (&v)[32]) {
  constexpr int K0 = 32 * HALF;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  if (HALF == 0) { v[0] = x[0]; v[1] = x[1]; v[2] = x[2]; }
  float sn[3], cs[3];
#pragma unroll
  for (int l = 0; l < NERFW_POS_LEVELS; ++l) {
    const bool needed = (8 + 6 * l >= K0) && (3 + 6 * l < K0 + 32);
    if (FAST) {
      if (l == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) sincosf(x[c], &sn[c], &cs[c]);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float s2 = 2.0f * sn[c] * cs[c];
          cs[c] = fmaf(-2.0f * sn[c], sn[c], 1.0f);
          sn[c] = s2;
        }
      }
    } else if (needed) {
      const float f = (float)(1u << l);
#pragma unroll
      for (int c = 0; c < 3; ++c) sincosf(f * x[c], &sn[c], &cs[c]);
    }
    if (needed) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int ks = 3 + 6 * l + c - K0, kc = 6 + 6 * l + c - K0;
        if (ks >= 0 && ks < 32) v[ks] = sn[c];
        if (kc >= 0 && kc < 32) v[kc] = cs[c];
      }
    }
  }
}
// direction tile: 27 features, zero padded to 32
template <bool FAST>
__device__ __forceinline__ void dir_features32(const float (&d)[3], float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  v[0] = d[0]; v[1] = d[1]; v[2] = d[2];
  float sn[3], cs[3];
#pragma unroll
  for (int l = 0; l < NERFW_DIR_LEVELS; ++l) {
    if (FAST && l > 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float s2 = 2.0f * sn[c] * cs[c];
        cs[c] = fmaf(-2.0f * sn[c], sn[c], 1.0f);
        sn[c] = s2;
      }
    } else {
      const float f = (float)(1u << l);
#pragma unroll
      for (int c = 0; c < 3; ++c) sincosf(f * d[c], &sn[c], &cs[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[3 + 6 * l + c] = sn[c];
      v[6 + 6 * l + c] = cs[c];
    }
  }
}
// 32 features of one row -> K-major swizzled operand tile(s): four 16-byte stores per copy
template <bool X3, bool F16 = false>
__device__ __forceinline__ void store_features32(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t row, uint32_t k0, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[8 * c + 2 * j], b = v[8 * c + 2 * j + 1];
      h[j] = F16 ? pack_f16x2(a, b) : pack_bf16x2(a, b);
      if (X3) l[j] = pack_bf16x2(a - __uint_as_float(h[j] << 16), b - __uint_as_float(h[j] & 0xffff0000u));
    }
    const uint32_t off = sw128_offset(row, k0 + 8 * c);
    *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (X3) *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ReLU gate words written by the forward kernel in training mode and read by the backward kernel:
// [tile][layer 0..8][row][column half][4 words of 32 gates]; layer 8 = direction layer (2 words per half used).
constexpr size_t MASK_WORDS_PER_TILE = 9ull * TM * 8;
__host__ __device__ inline size_t mask_index(int64_t tile, int layer, uint32_t row, uint32_t ch, int q) {
  return ((size_t)tile * 9 + layer) * (TM * 8) + row * 8 + ch * 4 + q;
}

}  // namespace tc
}  // namespace nerfw
