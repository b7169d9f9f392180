// Implicit-GEMM convolution / linear kernel for sm_100a.
//
//   * one persistent CTA per SM, 12 warps: warp 0 = TMA producer (A/W tiles), warp 1 = tcgen05.mma issuer + TMEM
//     owner, warp 2 = epilogue loader (TMA-prefetches residual / accumulate tiles into a shared-memory ring so
//     their latency hides behind the mainloop), warps 3..10 = two groups of epilogue math warps (TMEM -> registers
//     -> swizzled shared memory -> TMA store: every global access of the epilogue is a full-line bulk transfer),
//     warp 11 = activation-box loader of the stream mainloop
//   * A (activations, channels-last) is never im2col'ed: for filter tap j the producer issues a tiled TMA load
//     of the 128-pixel box shifted by (dh_j, dw_j); TMA's out-of-bounds zero fill IS the conv zero padding.
//     The landed box is 128 rows x 64 channels x 16 bit = rows of 128 B with the 128-byte swizzle, i.e. the
//     canonical K-major UMMA operand layout.
//   * W is [N, taps * c_pad] (K-major) and is loaded the same way with a {64, BLOCK_N} box.
//   * accumulators: 2 TMEM stages of BLOCK_N fp32 columns (MMA of tile i+1 overlaps the epilogue of tile i)
//   * epilogue fusions: bias, per-image row add (time embedding), activation (SiLU / GEGLU / tanh / LeakyReLU),
//     residual add, accumulate-into-output, scale, strided/offset output rows (transposed-conv phases),
//     optional second 16-bit output with its own activation (next layer's operand).
//
// Reference call sites replaced: see ctta_gemm in include/ctta.h.
#include <cuda.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include "ctta_internal.h"
#include "ctta_ptx.cuh"

namespace ctta {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // 64 x 16-bit = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kThreads = 384;                        // 3 control warps + 8 epilogue warps + the stream-mode A loader
constexpr int kSmemMaxDynamic = 232448 - 1024;     // 227 KiB minus the static barriers
constexpr int kSmemBudget = kSmemMaxDynamic - 1024;  // minus alignment slack
constexpr int kEpiWarp0 = 3;                         // first of the 4 epilogue math warps
constexpr int kChunkCols = 32;                       // accumulator columns per epilogue chunk
constexpr int kRingSlotBytes = kBlockM * kChunkCols * 4;  // 16 KiB: 128 rows x 32 fp32 columns (128-B swizzle)
constexpr int kMaxRing = 8;
constexpr int kStage16Bytes = 32 * kChunkCols * 2;   // 2 KiB: 32 rows x 32 16-bit columns per warp per buffer
constexpr int kBiasBytes = 2 * 256 * 4;

struct GemmKParams {
  // tiling
  int n_tiles_m, n_tiles_n, block_n, acc_stride, tmem_cols;
  int k_chunks, ntaps, n_stages, stage_bytes;
  int a_mode, is_bf16;
  // tile -> coordinates
  int tiles_per_img;            // CONV1D: tiles along t per batch row
  int box_w, box_h, box_n;      // CONV2D
  int tiles_w, tiles_h;         // CONV2D
  int H, W, n_img, rows_per_img;
  short tap_d0[CTTA_MAX_TAPS], tap_d1[CTTA_MAX_TAPS];
  // epilogue
  int N;
  const float* bias;
  const float* rowadd;
  int rowadd_ld, rowadd_rows;
  int act;
  float act_slope;
  const void* residual;
  int res_dtype, res_ld;
  int accumulate;
  float out_scale;
  void* out;
  int out_dtype, out_ld;
  void* out2;
  int out2_ld, act2;
  float act2_slope;
  int out_rows_per_img, out_stride, out_off;
  int vec_ok;  // all pointers 16-B aligned and all lds multiples of 8 -> 8-column vector path allowed
  // TMA-staged epilogue
  int epi_tma;             // 1: outputs / residual go through shared memory + TMA; 0: direct per-thread global access
  int ring_slots;          // number of 16-KiB fp32 ring slots (0 when the ring is unused)
  int ring_in;             // fp32 input tiles per chunk loaded by the loader warp (residual, previous out): 0..2
  int ring_per_chunk;      // ring slots consumed per chunk (max(ring_in, 1) when the ring is used)
  int out_rows_tile_img;   // rows of one image inside a 128-row tile (128 / box_n)
  int row_coord_shift;     // logical row -> TMA row coordinate (transposed-conv phases start at q_start)
  int ring_off, stage16_off, bias_off;  // byte offsets from the 1024-aligned dynamic smem base
  int kk_last;             // 16-wide K slices actually needed in the last channel chunk of a tap (1..4)
  // halo mode (CONV1D, c <= 64, single N tile): all taps' weights stay resident in shared memory and every tile
  // loads ONE halo'd activation box; tap j is a row-shifted UMMA view of it (no per-tap re-load from L2)
  int halo, halo_rows, halo_min_shift, w_bytes, tiles_off;
  // M super-tiles (ROWS / CONV1D with narrow N): one tile = sub_tiles x 128 rows sharing one accumulator stage,
  // which amortises the per-tile hand-offs when a 128 x N tile is only a few hundred cycles of work
  int sub_tiles;
  int epi_groups;          // 1 or 2 groups of 4 epilogue warps (2 = alternate chunks between the groups)
  // fused GroupNorm moments of the OUTPUT (sum, sum of squares per (image, group)), accumulated with atomics
  float* stats;
  int stats_groups, stats_cpg, stats_rows;   // channels per group (power of two >= 4); logical rows per image
  // 16-bit residual (same dtype as the operands): ring slots are 128 rows x 64 B; negative residual values are scaled by
  // res_neg_scale before the add (inverse LeakyReLU when the residual is given as its LeakyReLU'ed copy)
  int out4d;               // upsample-phase output: 4-D TMA map {cols, w, h, img}; store coordinates (col, r % W, r / W, img)
  int w_batched;           // the W operand has one [n, K] matrix per image (batched GEMM: attention scores / P.V)
  int res16, ring_slot_bytes;
  float res_neg_scale;
  int cw, pair, st16_bufs, st16_bytes;   // wide / paired 16-bit epilogue I/O (see the TMA-staged epilogue)
  // stream mode in CTA pairs (cluster of 2): the two CTAs work on neighbouring M tiles of the SAME N tile in lock step;
  // each loads half of every weight chunk and TMA-multicasts it to both, halving the weight bytes an SM pulls from L2
  int w_mcast, total_tiles_mc;   // w_mcast = CTAs per cluster (0: independent CTAs, 2 or 4)
  int acc_single;          // 1: one accumulator stage (sub_tiles * block_n * 2 > 512 TMEM columns), else two
  // stream mode (multi-tap convolutions): per (channel chunk, tap group) ONE halo'd activation box is loaded into
  // the A ring and every tap of the group is a row-shifted UMMA view of it; weight chunks stream through their own
  // ring and are shared by the sub_tiles of an M super-tile.  Cuts the L2 -> shared-memory traffic per FLOP by the
  // number of taps (A) and by sub_tiles (W), which is what bounds the tap-by-tap mainloop (~14 TB/s chip-wide).
  int stream;
  int n_groups;                       // tap groups: CONV1D 1, CONV2D one per distinct horizontal shift
  short grp_first[4], grp_count[4], grp_shift[4];   // taps [first, first+count) of a group; box shift along w
  short tap_id[CTTA_MAX_TAPS];        // original tap index (column block of the packed weights)
  short tap_rows[CTTA_MAX_TAPS];      // row offset of the tap's view inside the box
  int a_loads, a_box_rows, a_stage_bytes, n_a_stages, a_row0_shift;
  int w_stage_bytes, w_ring_off;
};
constexpr int kMaxAStages = 3;

struct TileCoord {
  int n0;          // first output channel of the tile
  int c1, c2, c3;  // A box origin (mode dependent)
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile) {
  TileCoord tc;
  int m_tile = tile;
  tc.n0 = 0;
  if (p.w_mcast) {
    // tile = 2 * (pair index q) + half: the CTAs of a pair (consecutive block indices) share q, hence the N tile, and
    // take M tiles 2 * (q / n_tiles_n) + half.  M tiles past the end decode to an out-of-range image: their loads are
    // zero-filled and their stores clipped by the TMA unit, so a pair always runs the same number of tiles.
    const int q = tile / p.w_mcast, half = tile - q * p.w_mcast;
    const int qm = q / p.n_tiles_n;
    tc.n0 = (q - qm * p.n_tiles_n) * p.block_n;
    m_tile = p.w_mcast * qm + half;
  } else if (p.n_tiles_n != 1) {   // (one N tile: no division on the per-tile path of the narrow convolutions)
    m_tile = tile / p.n_tiles_n;
    tc.n0 = (tile - m_tile * p.n_tiles_n) * p.block_n;
  }
  if (p.a_mode == CTTA_A_ROWS) {
    tc.c1 = m_tile * kBlockM * p.sub_tiles;
    tc.c2 = 0;
    tc.c3 = 0;
  } else if (p.a_mode == CTTA_A_CONV1D) {
    const int b = m_tile / p.tiles_per_img;
    tc.c1 = (m_tile - b * p.tiles_per_img) * kBlockM * p.sub_tiles;
    tc.c2 = b;
    tc.c3 = 0;
  } else {
    const int tw = m_tile % p.tiles_w;
    const int r = m_tile / p.tiles_w;
    const int th = r % p.tiles_h;
    const int ng = r / p.tiles_h;
    tc.c1 = tw * p.box_w;
    tc.c2 = th * p.box_h;
    tc.c3 = ng * p.box_n;
  }
  return tc;
}

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  switch (act) {
    case CTTA_ACT_SILU: return v / (1.f + __expf(-v));
    case CTTA_ACT_TANH: return tanhf(v);
    case CTTA_ACT_LRELU: return v > 0.f ? v : v * slope;
    default: return v;
  }
}
__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }
// Exact-erf GELU at half the instructions of erff(): erfc(a) = 2^q(a) on a = |g| / sqrt(2) in [0, 4.4] with a degree-6
// polynomial q (least squares on Chebyshev nodes of log2 erfc), Phi(g) = 1 - erfc/2 (g >= 0) or erfc/2 (g < 0).
// Against the float64 erf GELU: max abs error 8.6e-6, max relative error 6e-5 — an order of magnitude under the
// 16-bit rounding of the value it produces.  Used by the GEGLU epilogue, where erff() made the epilogue ALU-bound.
__device__ __forceinline__ float gelu_fast(float g) {
  float a = fminf(fabsf(g) * 0.70710678f, 4.4f);
  float q = 1.7291631e-04f;
  q = fmaf(q, a, -3.3274852e-03f);
  q = fmaf(q, a, 2.8219042e-02f);
  q = fmaf(q, a, -1.4366551e-01f);
  q = fmaf(q, a, -9.2350090e-01f);
  q = fmaf(q, a, -1.6263064e+00f);
  q = fmaf(q, a, -7.6538774e-05f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  const float h = 0.5f * e;
  return g * (g >= 0.f ? 1.f - h : h);
}

__device__ __forceinline__ float ld16(const void* p, int is_bf16, long long i) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                 : __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ unsigned short cvt16(float v, int is_bf16) {
  if (is_bf16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __half_as_ushort(__float2half_rn(v));
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int is_bf16) {
  return static_cast<uint32_t>(cvt16(a, is_bf16)) | (static_cast<uint32_t>(cvt16(b, is_bf16)) << 16);
}
__device__ __forceinline__ void unpack16x8(const uint4& u, int is_bf16, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (is_bf16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

// Epilogue for 8 consecutive accumulator columns [col, col + 8) of one output row.
__device__ __forceinline__ void epilogue8(const GemmKParams& p, float* v, int col, long long orow, long long ra_row) {
  const bool full = p.vec_ok && (col + 8 <= p.N);
  if (p.act == CTTA_ACT_GEGLU) {
    // interleaved (value, gate) pairs -> 4 outputs at columns col/2 .. col/2+3; host guarantees N % 8 == 0
    if (col >= p.N) return;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float h = v[2 * i], g = v[2 * i + 1];
      if (p.bias) {
        h += p.bias[col + 2 * i];
        g += p.bias[col + 2 * i + 1];
      }
      o[i] = h * gelu_erf(g) * p.out_scale;
    }
    const long long off = orow * p.out_ld + (col >> 1);
    if (p.out_dtype == CTTA_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      const int bf = p.out_dtype == CTTA_BF16;
      uint2 u;
      u.x = pack16(o[0], o[1], bf);
      u.y = pack16(o[2], o[3], bf);
      *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(p.out) + off) = u;
    }
    return;
  }
  if (full) {
    if (p.bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col);
      const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (p.rowadd) {
      const float* ra = p.rowadd + ra_row * p.rowadd_ld + col;
      const float4 b0 = *reinterpret_cast<const float4*>(ra);
      const float4 b1 = *reinterpret_cast<const float4*>(ra + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (p.act != CTTA_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = act_apply(v[i], p.act, p.act_slope);
    }
    if (p.residual) {
      const long long off = orow * p.res_ld + col;
      if (p.res_dtype == CTTA_F32) {
        const float* r = reinterpret_cast<const float*>(p.residual) + off;
        const float4 b0 = *reinterpret_cast<const float4*>(r);
        const float4 b1 = *reinterpret_cast<const float4*>(r + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(p.residual) + off);
        float f[8];
        unpack16x8(u, p.res_dtype == CTTA_BF16, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += f[i] < 0.f ? f[i] * p.res_neg_scale : f[i];
      }
    }
    const long long ooff = orow * p.out_ld + col;
    if (p.out_scale != 1.f) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= p.out_scale;
    }
    if (p.accumulate) {
      if (p.out_dtype == CTTA_F32) {
        const float* r = reinterpret_cast<const float*>(p.out) + ooff;
        const float4 b0 = *reinterpret_cast<const float4*>(r);
        const float4 b1 = *reinterpret_cast<const float4*>(r + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(p.out) + ooff);
        float f[8];
        unpack16x8(u, p.out_dtype == CTTA_BF16, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += f[i];
      }
    }
    if (p.out) {
      if (p.out_dtype == CTTA_F32) {
        float* o = reinterpret_cast<float*>(p.out) + ooff;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        const int bf = p.out_dtype == CTTA_BF16;
        uint4 u;
        u.x = pack16(v[0], v[1], bf); u.y = pack16(v[2], v[3], bf);
        u.z = pack16(v[4], v[5], bf); u.w = pack16(v[6], v[7], bf);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(p.out) + ooff) = u;
      }
    }
    if (p.out2) {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = act_apply(v[i], p.act2, p.act2_slope);
      uint4 u;
      u.x = pack16(w[0], w[1], p.is_bf16); u.y = pack16(w[2], w[3], p.is_bf16);
      u.z = pack16(w[4], w[5], p.is_bf16); u.w = pack16(w[6], w[7], p.is_bf16);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(p.out2) + orow * p.out2_ld + col) = u;
    }
    return;
  }
  // scalar tail / unaligned path
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = col + i;
    if (c >= p.N) break;
    float x = v[i];
    if (p.bias) x += p.bias[c];
    if (p.rowadd) x += p.rowadd[ra_row * p.rowadd_ld + c];
    x = act_apply(x, p.act, p.act_slope);
    if (p.residual) {
      const long long off = orow * p.res_ld + c;
      float r = p.res_dtype == CTTA_F32 ? reinterpret_cast<const float*>(p.residual)[off]
                                        : ld16(p.residual, p.res_dtype == CTTA_BF16, off);
      if (p.res_dtype != CTTA_F32 && r < 0.f) r *= p.res_neg_scale;
      x += r;
    }
    const long long ooff = orow * p.out_ld + c;
    x *= p.out_scale;
    if (p.accumulate) {
      x += p.out_dtype == CTTA_F32 ? reinterpret_cast<const float*>(p.out)[ooff]
                                   : ld16(p.out, p.out_dtype == CTTA_BF16, ooff);
    }
    if (p.out) {
      if (p.out_dtype == CTTA_F32) reinterpret_cast<float*>(p.out)[ooff] = x;
      else reinterpret_cast<unsigned short*>(p.out)[ooff] = cvt16(x, p.out_dtype == CTTA_BF16);
    }
    if (p.out2) {
      reinterpret_cast<unsigned short*>(p.out2)[orow * p.out2_ld + c] =
          cvt16(act_apply(x, p.act2, p.act2_slope), p.is_bf16);
    }
  }
}


// Compile-time specialisation of the TMA-staged epilogue.  -1 = decided at run time (generic instance).
template <int ACT, int ROWADD, int RING_IN, int OUT_F32, int OUT_16, int OUT2, int ACT2>
struct EpiCfg {
  static constexpr int kAct = ACT, kRowadd = ROWADD, kRingIn = RING_IN, kOutF32 = OUT_F32, kOut16 = OUT_16,
                       kOut2 = OUT2, kAct2 = ACT2;
};
using EpiGeneric = EpiCfg<-1, -1, -1, -1, -1, -1, -1>;
#define CTTA_CFG(field, runtime) (Cfg::field >= 0 ? Cfg::field : (runtime))

// two fp32 -> packed f16x2 with saturation to the finite range, one instruction
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// Packed fp32 pairs (FADD2 / FMUL2 on sm_100): IEEE round-to-nearest like the scalar forms, half the issue slots.  The
// TMA-staged epilogue is issue-bound on the narrow convolutions (8 warps x ~500 instructions per 32 x 32 chunk).
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
__device__ __forceinline__ void mul2s(float& o0, float& o1, float a0, float a1, float s) {
  uint64_t a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(s));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(a));
}
// LeakyReLU of a pair: max(v, slope * v) for slope in [0, 1] (min for slope > 1) == v > 0 ? v : slope * v
__device__ __forceinline__ void lrelu2(float& v0, float& v1, float slope) {
  float m0, m1;
  mul2s(m0, m1, v0, v1, slope);
  if (slope <= 1.f) {
    v0 = fmaxf(v0, m0);
    v1 = fmaxf(v1, m1);
  } else {
    v0 = fminf(v0, m0);
    v1 = fminf(v1, m1);
  }
}
// x = x * a + b on both halves (FFMA2), a and b pairs
__device__ __forceinline__ void fma2(float& x0, float& x1, float a0, float a1, float b) {
  uint64_t x, av, bv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(av), "l"(bv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}
// gelu_fast() of two values with the polynomial on packed pairs (same operations and roundings, half the FFMA slots)
__device__ __forceinline__ void gelu_fast2(float g0, float g1, float& r0, float& r1) {
  const float a0 = fminf(fabsf(g0) * 0.70710678f, 4.4f), a1 = fminf(fabsf(g1) * 0.70710678f, 4.4f);
  float q0 = 1.7291631e-04f, q1 = 1.7291631e-04f;
  fma2(q0, q1, a0, a1, -3.3274852e-03f);
  fma2(q0, q1, a0, a1, 2.8219042e-02f);
  fma2(q0, q1, a0, a1, -1.4366551e-01f);
  fma2(q0, q1, a0, a1, -9.2350090e-01f);
  fma2(q0, q1, a0, a1, -1.6263064e+00f);
  fma2(q0, q1, a0, a1, -7.6538774e-05f);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const float h0 = 0.5f * e0, h1 = 0.5f * e1;
  r0 = g0 * (g0 >= 0.f ? 1.f - h0 : h0);
  r1 = g1 * (g1 >= 0.f ? 1.f - h1 : h1);
}

// in-place activation of an even-length register array
template <int N>
__device__ __forceinline__ void act_apply_n(float* v, int act, float slope) {
  if (act == CTTA_ACT_LRELU) {
#pragma unroll
    for (int i = 0; i < N; i += 2) lrelu2(v[i], v[i + 1], slope);
  } else if (act != CTTA_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = act_apply(v[i], act, slope);
  }
}

// GroupNorm moments of one 32-column chunk: v[32] = this thread's row, NG = groups inside the chunk (32 / cpg, or 1
// when a group spans the whole chunk).  Rows are summed across the warp with a halving butterfly (NG * 2 values ->
// one value per lane after log2(2 NG) exchange steps), then one red.global per (group, moment).
template <int NG>
__device__ __forceinline__ void stats_chunk(const float* v, bool row_valid, int lane, float* dst /* [groups][2] */,
                                            int g0, int groups) {
  constexpr int CPG = 32 / NG;
  constexpr int NV = 2 * NG;
  float a[NV];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; ++i) {
      const float x = row_valid ? v[g * CPG + i] : 0.f;
      s += x;
      q = fmaf(x, x, q);
    }
    a[2 * g] = s;
    a[2 * g + 1] = q;
  }
  int idx = 0;       // index of the value this lane ends up owning
  int mask = 16;
#pragma unroll
  for (int step = NV / 2; step >= 1; step /= 2, mask /= 2) {
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = upper ? a[i] : a[i + step];
      const float keep = upper ? a[i + step] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
    if (upper) idx += step;
  }
  // a[0] now holds value `idx` summed over the lanes that differ in the bits already consumed
  float r = a[0];
  for (; mask >= 1; mask /= 2) r += __shfl_xor_sync(0xffffffffu, r, mask);
  constexpr int kUsedBits = (NV == 2) ? 1 : (NV == 4) ? 2 : (NV == 8) ? 3 : 4;
  const int low = lane & ((16 >> (kUsedBits - 1)) - 1);   // lanes sharing one value: only low == 0 writes
  const int g = g0 + (idx >> 1);
  if (low == 0 && g < groups) atomicAdd(dst + 2 * g + (idx & 1), r);
}

// KK tcgen05.mma of one 64-wide K chunk; descriptors advance by 2 (x16 B) per 16-element K slice.
template <int KK>
__device__ __forceinline__ void umma_chunk(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int kk = 0; kk < KK; ++kk)
    umma_f16(d_tmem, umma_desc_from_lo(a_lo + 2 * kk), umma_desc_from_lo(b_lo + 2 * kk), idesc, kk == 0 ? acc_first : 1u);
}
__device__ __forceinline__ void umma_chunk_n(int nkk, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t acc_first) {
  switch (nkk) {
    case 1: umma_chunk<1>(d_tmem, a_lo, b_lo, idesc, acc_first); break;
    case 2: umma_chunk<2>(d_tmem, a_lo, b_lo, idesc, acc_first); break;
    case 3: umma_chunk<3>(d_tmem, a_lo, b_lo, idesc, acc_first); break;
    default: umma_chunk<4>(d_tmem, a_lo, b_lo, idesc, acc_first); break;
  }
}
// all taps of one resident-weight (halo mode) sub-tile: tap j = row-shifted view of the activation box
template <int KK>
__device__ __forceinline__ void umma_halo_taps(const GemmKParams& p, uint32_t d_tmem, uint32_t a_lo0, uint32_t b_lo0,
                                               uint32_t b_step, uint32_t idesc) {
  for (int j = 0; j < p.ntaps; ++j) {
    const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(p.tap_rows[j]) * 8u;
    umma_chunk<KK>(d_tmem, a_lo, b_lo0 + j * b_step, idesc, j != 0 ? 1u : 0u);
  }
}

struct GemmBars {
  uint64_t full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2], ring_full[kMaxRing], ring_empty[kMaxRing],
      weights[1], a_full[kMaxAStages], a_empty[kMaxAStages];
};
#define BAR(field, idx) (bar_base + static_cast<uint32_t>(offsetof(GemmBars, field)) + 8u * static_cast<uint32_t>(idx))

// DIRECT = 1: the instance for problems whose epilogue cannot go through TMA (direct per-thread global access); it is
// compiled apart so that the TMA-staged instances do not carry its code (the epilogue warps of the narrow convolutions
// lost ~9 % of their issue slots to instruction-cache misses in the combined 17 k-instruction kernel).
template <class Cfg, int NTHREADS = kThreads, int DIRECT = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
               const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // every mbarrier lives in ONE static block and is addressed as (pinned base register + constant offset): taking the
  // shared-window address of each array separately makes ptxas re-derive it with an S2UR SR_CgaCtaId (~30 cycles of
  // uniform-datapath latency) at every use inside the single-thread issue loops
  __shared__ __align__(8) GemmBars bars_s;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles_base = pin_u32((smem_u32(smem_raw) + 1023u) & ~1023u);  // SWIZZLE_128B needs 1024-B alignment
  const uint32_t bar_base = pin_u32(smem_u32(&bars_s));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(BAR(full, s), 1);
      mbar_init(BAR(empty, s), p.w_mcast ? p.w_mcast : 1);   // cluster mode: a weight stage is released by every CTA's MMA warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(tmem_full, s), 1);
      mbar_init(BAR(tmem_empty, s), p.epi_tma ? 4 * p.epi_groups : 4);  // one arrive per epilogue warp
    }
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(BAR(ring_full, s), 1);
      mbar_init(BAR(ring_empty, s), 4);
    }
    mbar_init(BAR(weights, 0), 1);
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(BAR(a_full, s), 1);
      mbar_init(BAR(a_empty, s), 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.epi_tma) {
      if (p.out) tma_prefetch_desc(&tmap_out);
      if (p.out2) tma_prefetch_desc(&tmap_out2);
      if (p.residual) tma_prefetch_desc(&tmap_res);
    }
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.w_mcast) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast into them
  const uint32_t tmem_base = tmem_base_slot;

  const int total_tiles = p.w_mcast ? p.total_tiles_mc : p.n_tiles_m * p.n_tiles_n;
  const int k_iters = p.ntaps * p.k_chunks;

  const uint32_t stages_base = tiles_base + static_cast<uint32_t>(p.tiles_off);
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    const bool leader = elect_one_sync();
    if (leader && p.stream) {
      // weight chunks only; the halo'd activation boxes come from their own loader warp (the last warp) so that a box is
      // requested as soon as its ring slot frees up, independent of the weight ring's back-pressure
      int w_stage = 0;
      uint32_t w_phase = 0;
      const uint32_t w_base = tiles_base + static_cast<uint32_t>(p.w_ring_off);
      const uint32_t mc_rank = p.w_mcast ? cluster_ctarank() : 0u;
      const uint16_t mc_mask = static_cast<uint16_t>((1u << p.w_mcast) - 1u);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int j = 0; j < p.ntaps; ++j) {   // taps are stored group by group
            mbar_wait_ool(BAR(empty, w_stage), w_phase ^ 1u);
            const uint32_t wfull = BAR(full, w_stage);
            mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(p.w_stage_bytes));
            if (p.w_mcast) {
              // this CTA's half of the chunk (block_n / 2 weight rows) goes to both CTAs of the pair
              tma_load_2d_mcast(w_base + w_stage * p.w_stage_bytes + mc_rank * (p.w_stage_bytes / p.w_mcast), &tmap_b, wfull,
                                (p.tap_id[j] * p.k_chunks + kc) * kBlockK,
                                tc.n0 + static_cast<int>(mc_rank) * (p.block_n / p.w_mcast), mc_mask);
            } else {
              tma_load_2d(w_base + w_stage * p.w_stage_bytes, &tmap_b, wfull, (p.tap_id[j] * p.k_chunks + kc) * kBlockK,
                          tc.n0);
            }
            if (++w_stage == p.n_stages) {
              w_stage = 0;
              w_phase ^= 1u;
            }
          }
        }
      }
    } else if (leader && p.halo) {
      // resident weights: one {64 x block_n} box per tap, loaded once per CTA
      const uint32_t wbar = BAR(weights, 0);
      mbar_arrive_expect_tx(wbar, static_cast<uint32_t>(p.w_bytes));
      for (int j = 0; j < p.ntaps; ++j)
        tma_load_2d(tiles_base + static_cast<uint32_t>(j * p.block_n * 128), &tmap_b, wbar, j * kBlockK, 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.halo_rows * 128);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        for (int sub = 0; sub < p.sub_tiles; ++sub) {
          mbar_wait_ool(BAR(empty, stage), phase ^ 1u);
          const uint32_t full = BAR(full, stage);
          mbar_arrive_expect_tx(full, tx_bytes);
          tma_load_3d(stages_base + static_cast<uint32_t>(stage * p.stage_bytes), &tmap_a, full, 0,
                      tc.c1 + sub * kBlockM + p.halo_min_shift, tc.c2);
          if (++stage == p.n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.stage_bytes);
      const uint32_t mc_rank = p.w_mcast ? cluster_ctarank() : 0u;
      const uint16_t mc_mask = static_cast<uint16_t>((1u << p.w_mcast) - 1u);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        for (int sub = 0; sub < p.sub_tiles; ++sub) {
          const int r0 = tc.c1 + sub * kBlockM;  // sub_tiles > 1 only in ROWS / CONV1D mode
          int kb = 0;
          for (int j = 0; j < p.ntaps; ++j) {
            const int d0 = p.tap_d0[j], d1 = p.tap_d1[j];
            for (int kc = 0; kc < p.k_chunks; ++kc, ++kb) {
              mbar_wait_ool(BAR(empty, stage), phase ^ 1u);
              const uint32_t full = BAR(full, stage);
              mbar_arrive_expect_tx(full, tx_bytes);
              const uint32_t a_dst = stages_base + static_cast<uint32_t>(stage * p.stage_bytes);
              const uint32_t b_dst = a_dst + kATileBytes;
              if (p.a_mode == CTTA_A_ROWS) {
                tma_load_2d(a_dst, &tmap_a, full, kc * kBlockK, r0);
              } else if (p.a_mode == CTTA_A_CONV1D) {
                tma_load_3d(a_dst, &tmap_a, full, kc * kBlockK, r0 + d0, tc.c2);
              } else {
                tma_load_4d(a_dst, &tmap_a, full, kc * kBlockK, tc.c1 + d0, tc.c2 + d1, tc.c3);
              }
              if (p.w_mcast)   // CTA pair: this CTA's half of the weight tile goes to both CTAs' stages
                tma_load_2d_mcast(b_dst + mc_rank * static_cast<uint32_t>(p.block_n * 128 / p.w_mcast), &tmap_b, full, kb * kBlockK,
                                  tc.n0 + static_cast<int>(mc_rank) * (p.block_n / p.w_mcast), mc_mask);
              else if (p.w_batched) tma_load_3d(b_dst, &tmap_b, full, kb * kBlockK, tc.n0, tc.c2);
              else tma_load_2d(b_dst, &tmap_b, full, kb * kBlockK, tc.n0);
              if (++stage == p.n_stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one lane)
    const bool leader = elect_one_sync();
    if (leader && p.stream) {
      const uint32_t idesc = umma_idesc(kBlockM, p.block_n, p.is_bf16);
      int a_stage = 0, w_stage = 0;
      uint32_t a_phase = 0, w_phase = 0;
      const uint32_t w_base = tiles_base + static_cast<uint32_t>(p.w_ring_off);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = p.acc_single ? 0 : (it & 1);
        const uint32_t acc_phase = p.acc_single ? (it & 1) : ((it >> 1) & 1);
        mbar_wait_ool(BAR(tmem_empty, acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
        uint32_t started = 0;   // 0 until the first MMA of every sub-tile has been issued
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const int nkk = (kc == p.k_chunks - 1) ? p.kk_last : kBlockK / 16;
          for (int g = 0; g < p.n_groups; ++g) {
            mbar_wait_ool(BAR(a_full, a_stage), a_phase);
            const uint32_t a_base = tiles_base + static_cast<uint32_t>(a_stage * p.a_stage_bytes);
            const int j1 = p.grp_first[g] + p.grp_count[g];
            for (int j = p.grp_first[g]; j < j1; ++j) {
              mbar_wait_ool(BAR(full, w_stage), w_phase);
              tc_fence_after();
              const uint32_t b_lo = umma_desc_lo(w_base + static_cast<uint32_t>(w_stage * p.w_stage_bytes));
              const uint32_t a_lo = umma_desc_lo(a_base) + static_cast<uint32_t>(p.tap_rows[j]) * 8u;
              if (nkk == kBlockK / 16) {
                for (int sub = 0; sub < p.sub_tiles; ++sub)
                  umma_chunk<4>(d_tmem + sub * p.block_n, a_lo + sub * (kBlockM * 8), b_lo, idesc, started);
              } else {
                for (int sub = 0; sub < p.sub_tiles; ++sub)
                  umma_chunk_n(nkk, d_tmem + sub * p.block_n, a_lo + sub * (kBlockM * 8), b_lo, idesc, started);
              }
              started = 1;
              if (p.w_mcast) umma_commit_mcast(BAR(empty, w_stage), static_cast<uint16_t>((1u << p.w_mcast) - 1u));
              else umma_commit(BAR(empty, w_stage));
              if (++w_stage == p.n_stages) {
                w_stage = 0;
                w_phase ^= 1u;
              }
            }
            umma_commit(BAR(a_empty, a_stage));
            if (++a_stage == p.n_a_stages) {
              a_stage = 0;
              a_phase ^= 1u;
            }
          }
        }
        umma_commit(BAR(tmem_full, acc));
      }
    } else if (leader && p.halo) {
      const uint32_t idesc = umma_idesc(kBlockM, p.block_n, p.is_bf16);
      mbar_wait_ool(BAR(weights, 0), 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = p.acc_single ? 0 : (it & 1);
        const uint32_t acc_phase = p.acc_single ? (it & 1) : ((it >> 1) & 1);
        mbar_wait_ool(BAR(tmem_empty, acc), acc_phase ^ 1u);
        for (int sub = 0; sub < p.sub_tiles; ++sub) {
          mbar_wait_ool(BAR(full, stage), phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride + sub * p.block_n);
          const uint32_t a_base = stages_base + static_cast<uint32_t>(stage * p.stage_bytes);
          // tap j = the resident box shifted by tap_rows[j] rows of 128 bytes.  The swizzle phase is taken from the
          // shared-memory address bits, so a row-shifted start needs no descriptor base offset.
          {
            const uint32_t a_lo0 = umma_desc_lo(a_base), b_lo0 = umma_desc_lo(tiles_base);
            const uint32_t b_step = static_cast<uint32_t>(p.block_n) * 8u;
            switch (p.kk_last) {
              case 1: umma_halo_taps<1>(p, d_tmem, a_lo0, b_lo0, b_step, idesc); break;
              case 2: umma_halo_taps<2>(p, d_tmem, a_lo0, b_lo0, b_step, idesc); break;
              case 3: umma_halo_taps<3>(p, d_tmem, a_lo0, b_lo0, b_step, idesc); break;
              default: umma_halo_taps<4>(p, d_tmem, a_lo0, b_lo0, b_step, idesc); break;
            }
          }
          umma_commit(BAR(empty, stage));
          if (++stage == p.n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(BAR(tmem_full, acc));
      }
    } else if (leader) {
      const uint32_t idesc = umma_idesc(kBlockM, p.block_n, p.is_bf16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = p.acc_single ? 0 : (it & 1);
        const uint32_t acc_phase = p.acc_single ? (it & 1) : ((it >> 1) & 1);
        mbar_wait_ool(BAR(tmem_empty, acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int sub = 0; sub < p.sub_tiles; ++sub) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride + sub * p.block_n);
        int kc = 0;
        for (int kb = 0; kb < k_iters; ++kb) {
          mbar_wait_ool(BAR(full, stage), phase);
          tc_fence_after();
          const uint32_t a_lo = umma_desc_lo(stages_base + static_cast<uint32_t>(stage * p.stage_bytes));
          const uint32_t b_lo = a_lo + (kATileBytes >> 4);
          if (kc != p.k_chunks - 1 || p.kk_last == kBlockK / 16)
            umma_chunk<4>(d_tmem, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);
          else
            umma_chunk_n(p.kk_last, d_tmem, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);  // skip all-zero K slices
          if (p.w_mcast) umma_commit_mcast(BAR(empty, stage), static_cast<uint16_t>((1u << p.w_mcast) - 1u));   // every CTA of the cluster
          else umma_commit(BAR(empty, stage));  // frees the smem stage once these MMAs retire
          if (++kc == p.k_chunks) kc = 0;
          if (++stage == p.n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        }
        umma_commit(BAR(tmem_full, acc));  // accumulator complete -> epilogue
      }
    }
  } else if (warp == NTHREADS / 32 - 1) {
    // ------------------------------------------------------------------ stream mode: activation box loader (one lane)
    if (elect_one_sync() && p.stream) {
      int a_stage = 0;
      uint32_t a_phase = 0;
      const uint32_t a_box_bytes = static_cast<uint32_t>(p.a_box_rows * 128);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int g = 0; g < p.n_groups; ++g) {
            mbar_wait_ool(BAR(a_empty, a_stage), a_phase ^ 1u);
            const uint32_t afull = BAR(a_full, a_stage);
            mbar_arrive_expect_tx(afull, a_box_bytes * p.a_loads);
            const uint32_t a_dst = tiles_base + static_cast<uint32_t>(a_stage * p.a_stage_bytes);
            if (p.a_mode == CTTA_A_CONV1D) {
              for (int l = 0; l < p.a_loads; ++l)
                tma_load_3d(a_dst + l * a_box_bytes, &tmap_a, afull, kc * kBlockK,
                            tc.c1 + p.a_row0_shift + l * p.a_box_rows, tc.c2);
            } else {
              tma_load_4d(a_dst, &tmap_a, afull, kc * kBlockK, p.grp_shift[g], tc.c2 + p.a_row0_shift, tc.c3);
            }
            if (++a_stage == p.n_a_stages) {
              a_stage = 0;
              a_phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ epilogue loader (one lane)
    // Streams the fp32 residual / previous-output tiles of every chunk into the ring, running ahead of the
    // epilogue math by up to ring_slots chunks.  With no inputs it only hands out free slots (pure staging).
    if (elect_one_sync() && p.epi_tma && p.ring_slots > 0) {
      const int io_cols = kChunkCols * p.cw;
      const int n_chunks = p.block_n / io_cols;
      const uint32_t ring_base = tiles_base + static_cast<uint32_t>(p.ring_off);
      // shared-memory pitch of one tile row inside a slot: fp32 128 B; 16-bit 64 B (narrow / paired) or 128 B (wide)
      const int row_pitch = p.res16 ? (p.cw == 2 ? 128 : 64) : 128;
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        int row0, img0;
        if (p.a_mode == CTTA_A_ROWS) {
          row0 = tc.c1; img0 = 0;
        } else if (p.a_mode == CTTA_A_CONV1D) {
          row0 = tc.c1; img0 = tc.c2;
        } else {
          row0 = tc.c2 * p.W + tc.c1; img0 = tc.c3;
        }
        row0 -= p.row_coord_shift;
        for (int cc = 0; cc < n_chunks * p.sub_tiles; ++cc) {
          const int sub = cc / n_chunks;
          const int ch = cc - sub * n_chunks;
          const int col = tc.n0 + ch * io_cols;
          const int srow0 = row0 + sub * kBlockM;
          for (int k = 0; k < p.ring_per_chunk; ++k) {
            mbar_wait_ool(BAR(ring_empty, slot), phase ^ 1u);
            const uint32_t full = BAR(ring_full, slot);
            if (k < p.ring_in) {
              mbar_arrive_expect_tx(full, static_cast<uint32_t>(p.ring_slot_bytes));
              // the residual tile: one box per epilogue warp quarter (the box shape the stores use)
              const void* map = static_cast<const void*>(&tmap_res);
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                const int wr = qq * 32;
                const uint32_t dst = ring_base + slot * p.ring_slot_bytes + wr * row_pitch;
                if (p.pair) tma_load_3d(dst, map, full, 0, (srow0 + wr) >> 1, img0);
                else tma_load_3d(dst, map, full, col, srow0 + (wr % p.out_rows_tile_img), img0 + wr / p.out_rows_tile_img);
              }
            } else {
              mbar_arrive(full);
            }
            if (++slot == p.ring_slots) {
              slot = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (DIRECT) {
   if (warp < kEpiWarp0 + 4) {
    // ------------------------------------------------------------------ epilogue warps, direct global access
    // (tiny / unaligned outputs: N <= 16 or row pitch not a multiple of 16 bytes)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int lr = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = p.acc_single ? 0 : (it & 1);
      const uint32_t acc_phase = p.acc_single ? (it & 1) : ((it >> 1) & 1);
      const TileCoord tc = decode_tile(p, tile);
      // logical row -> (image, row in image) -> output row
      int img, r;
      bool valid;
      if (p.a_mode == CTTA_A_ROWS) {
        img = 0;
        r = tc.c1 + lr;
        valid = r < p.rows_per_img;
      } else if (p.a_mode == CTTA_A_CONV1D) {
        img = tc.c2;
        r = tc.c1 + lr;
        valid = r < p.rows_per_img;
      } else {
        const int iw = lr % p.box_w;
        const int t = lr / p.box_w;
        const int ih = t % p.box_h;
        img = tc.c3 + t / p.box_h;
        r = (tc.c2 + ih) * p.W + tc.c1 + iw;
        valid = img < p.n_img;
      }
      const long long o = static_cast<long long>(r) * p.out_stride + p.out_off;
      valid = valid && o >= 0 && o < p.out_rows_per_img;
      const long long orow = static_cast<long long>(img) * p.out_rows_per_img + o;
      const long long ra_row =
          p.rowadd ? (static_cast<long long>(img) * p.rows_per_img + r) / p.rowadd_rows : 0;

      mbar_wait_ool(BAR(tmem_full, acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc * p.acc_stride) + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t u[32];
        const bool wide = (c0 + 32 <= p.block_n);
        if (wide) tmem_ld_x32(t_addr + c0, u);
        else tmem_ld_x16(t_addr + c0, u);
        tmem_ld_wait();
        if (c0 + 32 >= p.block_n) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(tmem_empty, acc));
        }
        if (valid) {
          const int ncols = wide ? 32 : 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g * 8 < ncols) {
              const int col = tc.n0 + c0 + g * 8;
              if (col < p.N) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(u[g * 8 + i]);
                epilogue8(p, v, col, orow, ra_row);
              }
            }
          }
        }
      }
    }
   }
  } else if (!DIRECT && warp < kEpiWarp0 + 4 * p.epi_groups) {
    // ------------------------------------------------------------------ epilogue warps, TMA-staged
    // thread = accumulator row.  Two warps per TMEM lane quarter (group 0 / 1) take alternate 32-column chunks.
    // Per chunk: TMEM -> registers, + bias (smem broadcast) + rowadd, act, + residual / previous out (ring slot,
    // prefetched by warp 2), scale; the fp32 result is written back in place into the ring slot, 16-bit results
    // into the per-warp staging buffer; one lane issues the TMA stores.
    const int has_rowadd = CTTA_CFG(kRowadd, p.rowadd != nullptr ? 1 : 0);
    const int ring_in = CTTA_CFG(kRingIn, p.ring_in);
    const int out_f32 = CTTA_CFG(kOutF32, (p.out != nullptr && p.out_dtype == CTTA_F32) ? 1 : 0);
    const int out_16 = CTTA_CFG(kOut16, (p.out != nullptr && p.out_dtype != CTTA_F32) ? 1 : 0);
    const int has_out2 = CTTA_CFG(kOut2, p.out2 != nullptr ? 1 : 0);
    const int act = CTTA_CFG(kAct, p.act);
    const int act2 = CTTA_CFG(kAct2, p.act2);
    const bool generic = Cfg::kAct < 0;       // only the generic instance supports bf16 outputs
    const int ring_per_chunk = (out_f32 || ring_in > 0) ? (ring_in > 1 ? ring_in : 1) : 0;

    const bool leader = elect_one_sync();     // the lane that owns this warp's bulk-store groups
    const int q = warp & 3;
    const int ew = warp - kEpiWarp0;          // 0..7, owner of staging buffers
    const int grp = ew >> 2;                  // chunk phase handled by this warp
    const int ngrp = p.epi_groups;
    const int lr = q * 32 + lane;             // row inside the tile
    // 16-bit-only epilogues move their tiles in 128-byte rows: cw = 2 pairs two 32-column halves into one I/O chunk
    // (one ring slot, one staging buffer, one store box of 64 columns); `pair` views an N = 32 tensor as
    // [rows / 2, 64].  The TMA unit's cost is per box ROW, so 64-byte rows halve its throughput.
    const int cw = p.cw, pair = p.pair, st16_bufs = p.st16_bufs;
    const int io_cols = kChunkCols * cw;
    const int n_chunks = p.block_n / io_cols;
    const int tot_chunks = n_chunks * p.sub_tiles;
    const uint32_t ring_base = tiles_base + static_cast<uint32_t>(p.ring_off);
    const uint32_t st16_base = tiles_base + static_cast<uint32_t>(p.stage16_off) + ew * st16_bufs * p.st16_bytes;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (tiles_base - smem_u32(smem_raw)) + p.bias_off);
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    const int sw7 = lane & 7;
    const int wrow = q * 32;
    const int wrow_in_img = wrow % p.out_rows_tile_img;
    const int wimg = wrow / p.out_rows_tile_img;
    uint32_t my_ctr = 0;    // chunks processed by this warp (staging buffer parity)
    const bool bias_once = p.n_tiles_n == 1;
    // ring position of global chunk (it * tot_chunks): slot index and phase advance incrementally (no 64-bit div / mod
    // per chunk); tile_adv_q / tile_adv_r = (tot_chunks * ring_per_chunk) div / mod ring_slots
    int tile_slot = 0;
    uint32_t tile_phase = 0;
    int tile_adv_q = 0, tile_adv_r = 0;
    // chunk -> warp group by GLOBAL chunk index (it * tot_chunks + cc) % ngrp, so that a chunk count that is not a
    // multiple of the group count (4 chunks over 3 groups) still balances over consecutive tiles
    int rot = 0;
    const int rot_step = tot_chunks % ngrp;
    if (ring_per_chunk > 0) {
      tile_adv_q = (tot_chunks * ring_per_chunk) / p.ring_slots;
      tile_adv_r = (tot_chunks * ring_per_chunk) - tile_adv_q * p.ring_slots;
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = p.acc_single ? 0 : (it & 1);
      const uint32_t acc_phase = p.acc_single ? (it & 1) : ((it >> 1) & 1);
      const TileCoord tc = decode_tile(p, tile);
      int row0, img0;
      if (p.a_mode == CTTA_A_ROWS) {
        row0 = tc.c1; img0 = 0;
      } else if (p.a_mode == CTTA_A_CONV1D) {
        row0 = tc.c1; img0 = tc.c2;
      } else {
        row0 = tc.c2 * p.W + tc.c1; img0 = tc.c3;
      }
      // this group's first chunk of the tile, its ring position, and the ring position of the next tile
      int cc0 = grp - rot;
      if (cc0 < 0) cc0 += ngrp;
      rot += rot_step;
      if (rot >= ngrp) rot -= ngrp;
      int c_slot = tile_slot + cc0 * ring_per_chunk;
      uint32_t c_phase = tile_phase;
      while (ring_per_chunk > 0 && c_slot >= p.ring_slots) { c_slot -= p.ring_slots; c_phase ^= 1u; }
      tile_slot += tile_adv_r;
      tile_phase ^= static_cast<uint32_t>(tile_adv_q & 1);
      if (ring_per_chunk > 0 && tile_slot >= p.ring_slots) { tile_slot -= p.ring_slots; tile_phase ^= 1u; }
      // this warp's 32 rows: coordinates of the store box (of sub-tile 0)
      const int st_row0 = row0 + wrow_in_img - p.row_coord_shift;
      const int st_img = img0 + wimg;
      long long ra_gr0 = 0, ra_gmax = 0;   // global logical row of this thread (sub-tile 0) for the per-image row add
      if (has_rowadd) {
        const int r_in_img = (p.a_mode == CTTA_A_CONV2D) ? (row0 + (lr % p.out_rows_tile_img)) : (row0 + lr);
        const int img = (p.a_mode == CTTA_A_CONV2D) ? (img0 + lr / p.out_rows_tile_img) : img0;
        ra_gr0 = static_cast<long long>(img) * p.rows_per_img + r_in_img;
        ra_gmax = static_cast<long long>(p.n_img) * p.rows_per_img - 1;  // rows past the end are clipped by the TMA store
      }
      // bias of this tile -> shared memory (double buffered by tile parity); with a single N tile it never changes
      float* bs = bias_s + (bias_once ? 0 : (it & 1) * 256);
      if (!bias_once || it == 0) {
        for (int i = et; i < p.block_n; i += ngrp * 128) {
          const int c = tc.n0 + i;
          bs[i] = (p.bias != nullptr && c < p.N) ? p.bias[c] : 0.f;
        }
        named_barrier_sync(1, ngrp * 128);
      }

      mbar_wait_ool(BAR(tmem_full, acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc * p.acc_stride) + (static_cast<uint32_t>(q * 32) << 16);
      const int last_cc = cc0 < tot_chunks ? ((tot_chunks - 1 - cc0) / ngrp) * ngrp + cc0 : -1;  // last chunk of this group
      if (cc0 >= tot_chunks) {
        // nothing to read from this accumulator stage
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(tmem_empty, acc));
      }
      int sub = 0, ch = cc0;
      while (ch >= n_chunks) { ch -= n_chunks; ++sub; }
      for (int cc = cc0; cc < tot_chunks; cc += ngrp, ++my_ctr) {
        const int col_io = tc.n0 + ch * io_cols;          // first column of this I/O chunk (cw halves of 32 columns)
        const int st_row = st_row0 + sub * kBlockM;
        const uint32_t st16 = st16_base + (st16_bufs == 2 ? (my_ctr & 1) : 0) * p.st16_bytes;
        if (st16_bufs == 1 && (out_16 || has_out2)) {
          // single staging buffer (wide 16-bit chunks): the previous chunk's store must have read it
          if (leader) tma_store_wait_read<0>();
          __syncwarp();
        }
        uint32_t slot_addr = 0;
        int first_slot = -1;
#pragma unroll 1
        for (int hf = 0; hf < cw; ++hf) {
        const int c0 = ch * io_cols + hf * kChunkCols;
        const int col = tc.n0 + c0;
        uint32_t u[32];
        tmem_ld_x32(t_addr + sub * p.block_n + c0, u);
        // per-row additive term (time embedding): issue the loads before waiting on TMEM / the ring
        float4 ra[8];
        if (has_rowadd) {
          long long gr = ra_gr0 + sub * kBlockM;
          if (gr > ra_gmax) gr = ra_gmax;   // keep the load in bounds
          const float* rp = p.rowadd + (gr / p.rowadd_rows) * p.rowadd_ld + col;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (col + 4 * g + 3 < p.N) ra[g] = __ldg(reinterpret_cast<const float4*>(rp + 4 * g));
            else ra[g] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tmem_ld_wait();
        if (cc == last_cc && hf == cw - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(tmem_empty, acc));
        }
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = *reinterpret_cast<const float4*>(bs + c0 + 4 * g);
          v[4 * g + 0] = __uint_as_float(u[4 * g + 0]);
          v[4 * g + 1] = __uint_as_float(u[4 * g + 1]);
          v[4 * g + 2] = __uint_as_float(u[4 * g + 2]);
          v[4 * g + 3] = __uint_as_float(u[4 * g + 3]);
          add2(v[4 * g + 0], v[4 * g + 1], b4.x, b4.y);
          add2(v[4 * g + 2], v[4 * g + 3], b4.z, b4.w);
        }
        if (has_rowadd) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            add2(v[4 * g + 0], v[4 * g + 1], ra[g].x, ra[g].y);
            add2(v[4 * g + 2], v[4 * g + 3], ra[g].z, ra[g].w);
          }
        }
        if (act == CTTA_ACT_GEGLU) {
          // 32 accumulator columns = 16 (value, gate) pairs -> 16 outputs = 32 bytes per row (unswizzled staging)
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float ge0, ge1;
            gelu_fast2(v[4 * i + 1], v[4 * i + 3], ge0, ge1);
            const float o0 = v[4 * i + 0] * ge0 * p.out_scale;
            const float o1 = v[4 * i + 2] * ge1 * p.out_scale;
            w[i] = (generic && p.out_dtype == CTTA_BF16) ? pack16(o0, o1, 1) : pack_f16_sat(o0, o1);
          }
          if (cw > 1) {
            // wide GEGLU chunk: cw halves of 16 outputs (32 B) fill one 128-byte staging row, 128-byte swizzle
            const uint32_t dst = st16 + lane * 128;
            const int sx = lane & 7;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2 * hf) ^ sx) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2 * hf + 1) ^ sx) << 4)), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
          } else {
            const uint32_t dst = st16 + lane * 32;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
          }
        } else {
          act_apply_n<32>(v, act, p.act_slope);
          // ---- inputs from the ring (residual, previous out)
          if (ring_per_chunk > 0) {
            int slot = c_slot;
            uint32_t ring_phase = c_phase;
            for (int k = 0; k < ring_per_chunk; ++k) {
              if (k > 0 && ++slot == p.ring_slots) { slot = 0; ring_phase ^= 1u; }
              if (hf == 0) mbar_wait_ool(BAR(ring_full, slot), ring_phase);
              const uint32_t sa = ring_base + slot * p.ring_slot_bytes + lr * 128;
              if (k == 0) {
                slot_addr = sa;
                first_slot = slot;
              }
              if (k < ring_in && p.res16) {
                // 16-bit residual tile as the TMA boxes wrote it.  Narrow chunk: rows of 64 B, 64-byte swizzle.  Wide
                // chunk (cw = 2): rows of 128 B holding both halves, 128-byte swizzle.  Paired rows (N = 32 viewed as
                // [rows / 2, 64]): row r lives in half (r & 1) of 128-byte row r / 2.
                uint32_t row_addr;
                int pc0, sx;
                if (cw == 2) {
                  row_addr = ring_base + slot * p.ring_slot_bytes + lr * 128;
                  pc0 = hf * 4;
                  sx = lane & 7;
                } else if (pair) {
                  row_addr = ring_base + slot * p.ring_slot_bytes + (lr >> 1) * 128;
                  pc0 = (lane & 1) * 4;
                  sx = (lr >> 1) & 7;
                } else {
                  row_addr = ring_base + slot * p.ring_slot_bytes + lr * 64;
                  pc0 = 0;
                  sx = (lane >> 1) & 3;
                }
                const float ns = p.res_neg_scale;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  uint4 u4;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(u4.x), "=r"(u4.y), "=r"(u4.z), "=r"(u4.w)
                               : "r"(row_addr + (((pc0 + g) ^ sx) << 4)));
                  float f8[8];
                  unpack16x8(u4, p.is_bf16, f8);
                  // inverse LeakyReLU: f < 0 ? ns * f : f  ==  min(f, ns * f) for ns >= 1 (max for ns < 1)
#pragma unroll
                  for (int i = 0; i < 8; i += 2) {
                    float m0, m1;
                    mul2s(m0, m1, f8[i], f8[i + 1], ns);
                    if (ns >= 1.f) add2(v[8 * g + i], v[8 * g + i + 1], fminf(f8[i], m0), fminf(f8[i + 1], m1));
                    else add2(v[8 * g + i], v[8 * g + i + 1], fmaxf(f8[i], m0), fmaxf(f8[i + 1], m1));
                  }
                }
              } else if (k < ring_in) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  float4 r4;
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                               : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w)
                               : "r"(sa + ((g ^ sw7) << 4)));
                  add2(v[4 * g + 0], v[4 * g + 1], r4.x, r4.y);
                  add2(v[4 * g + 2], v[4 * g + 3], r4.z, r4.w);
                }
              }
              if ((k > 0 || !out_f32) && hf == cw - 1) {
                // slot only read: release it right away
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(ring_empty, slot));
              }
            }
          }
          if (p.out_scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) mul2s(v[i], v[i + 1], v[i], v[i + 1], p.out_scale);
          }
          if (p.stats != nullptr) {
            // image of this warp's 32 rows (warp-uniform: the host checks stats_rows % 32 == 0) and row validity
            int s_img;
            bool rv;
            if (p.a_mode == CTTA_A_CONV2D) {
              s_img = img0 + wimg;
              rv = s_img < p.n_img;
            } else {
              const int r = row0 + sub * kBlockM + lr;
              rv = r < p.rows_per_img;
              s_img = (p.a_mode == CTTA_A_ROWS) ? (row0 + sub * kBlockM + wrow) / p.stats_rows : img0;
              if (p.a_mode == CTTA_A_ROWS && !(row0 + sub * kBlockM + wrow < p.rows_per_img)) s_img = 0;  // all lanes invalid
            }
            if (s_img >= p.n_img && p.a_mode != CTTA_A_ROWS) {   // padding tile of a CTA pair
              rv = false;
              s_img = 0;
            }
            float* sdst = p.stats + static_cast<long long>(s_img) * p.stats_groups * 2;
            const int g0 = col / p.stats_cpg;
            switch (p.stats_cpg) {
              case 4: stats_chunk<8>(v, rv, lane, sdst, g0, p.stats_groups); break;
              case 8: stats_chunk<4>(v, rv, lane, sdst, g0, p.stats_groups); break;
              case 16: stats_chunk<2>(v, rv, lane, sdst, g0, p.stats_groups); break;
              default: stats_chunk<1>(v, rv, lane, sdst, g0, p.stats_groups); break;   // cpg >= 32
            }
          }
          if (out_f32) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot_addr + ((g ^ sw7) << 4)),
                           "f"(v[4 * g + 0]), "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3])
                           : "memory");
            }
          }
          if (out_16 || has_out2) {
            const bool bf = generic && (out_16 ? (p.out_dtype == CTTA_BF16) : (p.is_bf16 != 0));
            // staging layout = the store box: see the residual tile above
            uint32_t dst;
            int pc0, sx;
            if (cw == 2) {
              dst = st16 + lane * 128;
              pc0 = hf * 4;
              sx = lane & 7;
            } else if (pair) {
              dst = st16 + (lane >> 1) * 128;
              pc0 = (lane & 1) * 4;
              sx = (lane >> 1) & 7;
            } else {
              dst = st16 + lane * 64;
              pc0 = 0;
              sx = (lane >> 1) & 3;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float w8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w8[i] = v[8 * g + i];
              if (!out_16) act_apply_n<8>(w8, act2, p.act2_slope);
              uint32_t w0, w1, w2, w3;
              if (bf) {
                w0 = pack16(w8[0], w8[1], 1); w1 = pack16(w8[2], w8[3], 1);
                w2 = pack16(w8[4], w8[5], 1); w3 = pack16(w8[6], w8[7], 1);
              } else {
                w0 = pack_f16_sat(w8[0], w8[1]); w1 = pack_f16_sat(w8[2], w8[3]);
                w2 = pack_f16_sat(w8[4], w8[5]); w3 = pack_f16_sat(w8[6], w8[7]);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((pc0 + g) ^ sx) << 4)), "r"(w0), "r"(w1),
                           "r"(w2), "r"(w3)
                           : "memory");
            }
          }
        }
        }  // hf
        {
          fence_proxy_async();
          __syncwarp();
          if (leader) {
            // the fp32 store is its own bulk group so that its ring slot can be handed back to the loader warp as soon
            // as the TMA unit has read it (a slot held until the NEXT chunk leaves the loader no room to run ahead
            // and exposes the residual's load latency once per chunk); the 16-bit staging stays double buffered
            if (out_f32) {
              const uint32_t src = ring_base + first_slot * p.ring_slot_bytes + wrow * 128;
              if (p.accumulate) tma_reduce_add_3d(&tmap_out, src, col_io, st_row, st_img);  // out += tile
              else if (p.out4d) tma_store_4d(&tmap_out, src, col_io, st_row % p.W, st_row / p.W, st_img);
              else tma_store_3d(&tmap_out, src, col_io, st_row, st_img);
              tma_store_commit();
            }
            if (out_16 || has_out2) {
              const int sc = pair ? 0 : (act == CTTA_ACT_GEGLU ? (col_io >> 1) : col_io), sr = pair ? (st_row >> 1) : st_row;
              if (out_16) tma_store_3d(&tmap_out, st16, sc, sr, st_img);
              if (has_out2) tma_store_3d(&tmap_out2, st16, sc, sr, st_img);
              tma_store_commit();
              if (st16_bufs == 2) tma_store_wait_read<1>();  // all but the 16-bit group just committed have read smem
            } else {
              tma_store_wait_read<0>();
            }
            if (out_f32) mbar_arrive(BAR(ring_empty, first_slot));
          }
          __syncwarp();
        }
        // advance (sub, ch) and the ring position to this group's next chunk
        ch += ngrp;
        while (ch >= n_chunks) { ch -= n_chunks; ++sub; }
        if (ring_per_chunk > 0) {
          c_slot += ngrp * ring_per_chunk;
          while (c_slot >= p.ring_slots) { c_slot -= p.ring_slots; c_phase ^= 1u; }
        }
      }
    }
    if (leader) tma_store_wait_all();  // global writes complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (p.w_mcast) cluster_sync_all();   // the peer may still multicast into this CTA's shared memory / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_tmap_ex(CUtensorMap* m, int dtype, CUtensorMapSwizzle swz, const void* base, int rank,
                        const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(CTTA_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == CTTA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == CTTA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(m, dt, rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(CTTA_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                     static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                     rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return 0;
}

static int make_tmap(CUtensorMap* m, int is_bf16, const void* base, int rank, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  return make_tmap_ex(m, is_bf16 ? CTTA_BF16 : CTTA_F16, CU_TENSOR_MAP_SWIZZLE_128B, base, rank, dims, strides_bytes,
                      box);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Row-addressed epilogue tensor (out / out2 / residual) as a 3-D map {columns, rows of one image, images} with a
// {box_cols x 32 rows x 1} box.  Transposed-conv phases write every out_stride-th row starting at first_row.
static int make_row_tmap(CUtensorMap* m, int dtype, const void* base, long long ld, int ncols, long long first_row,
                         long long row_step, long long n_rows, long long img_rows, int n_img, int box_cols,
                         CUtensorMapSwizzle swz, int box_rows = 32) {
  const int esz = dtype == CTTA_F32 ? 4 : 2;
  const char* b = reinterpret_cast<const char*>(base) + first_row * ld * esz;
  cuuint64_t dims[3] = {(cuuint64_t)ncols, (cuuint64_t)(n_rows > 0 ? n_rows : 1), (cuuint64_t)n_img};
  cuuint64_t strides[2] = {(cuuint64_t)(row_step * ld * esz), (cuuint64_t)(img_rows * ld * esz)};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  return make_tmap_ex(m, dtype, swz, b, 3, dims, strides, box);
}

}  // namespace ctta

using namespace ctta;

extern "C" int ctta_gemm(const ctta_gemm_desc* d, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(d != nullptr, "ctta_gemm: null descriptor");
  CTTA_REQUIRE(d->a && d->wgt, "ctta_gemm: null operand");
  CTTA_REQUIRE(d->out || d->out2, "ctta_gemm: no output");
  CTTA_REQUIRE(d->ab_dtype == CTTA_F16 || d->ab_dtype == CTTA_BF16, "ctta_gemm: operands must be 16-bit");
  CTTA_REQUIRE(d->a_mode >= 0 && d->a_mode <= 2, "ctta_gemm: bad a_mode");
  CTTA_REQUIRE(d->c > 0 && d->n > 0 && d->a_ld >= d->c && d->a_ld % 8 == 0,
               "ctta_gemm: need c > 0, n > 0, a_ld >= c and a_ld %% 8 == 0 (c=%d a_ld=%d)", d->c, d->a_ld);
  CTTA_REQUIRE(d->ntaps >= 1 && d->ntaps <= CTTA_MAX_TAPS, "ctta_gemm: ntaps out of range");
  CTTA_REQUIRE(aligned16(d->a) && aligned16(d->wgt), "ctta_gemm: operands must be 16-byte aligned");
  CTTA_REQUIRE(d->n_img >= 1 && d->rows_per_img >= 1 && d->out_rows_per_img >= 1, "ctta_gemm: bad row counts");
  if (d->a_mode == CTTA_A_ROWS) CTTA_REQUIRE(d->n_img == 1 && d->ntaps == 1, "ctta_gemm: ROWS mode is 1 image, 1 tap");
  if (d->act == CTTA_ACT_GEGLU)
    CTTA_REQUIRE(d->n % 8 == 0 && !d->residual && !d->accumulate && !d->out2 && !d->rowadd && d->out,
                 "ctta_gemm: GEGLU epilogue supports bias only and needs n %% 8 == 0");
  if (d->accumulate) CTTA_REQUIRE(d->out != nullptr && d->out2 == nullptr, "ctta_gemm: accumulate needs out and excludes out2");
  if (d->rowadd) CTTA_REQUIRE(d->rowadd_rows >= 1, "ctta_gemm: rowadd_rows must be >= 1");

  GemmKParams p{};
  p.a_mode = d->a_mode;
  p.is_bf16 = d->ab_dtype == CTTA_BF16;
  p.N = d->n;
  int block_n;
  if (d->n <= 16) {
    block_n = 16;
  } else {
    // multiple of 32 (epilogue chunk) minimising the padded width; ties -> the wider tile
    long long best = -1;
    block_n = 32;
    for (int bn = 256; bn >= 32; bn -= 32) {
      const long long padded = static_cast<long long>((d->n + bn - 1) / bn) * bn;
      if (best < 0 || padded < best) {
        best = padded;
        block_n = bn;
      }
    }
    // small-M problems (the UNet's 32x2 / 64x4 levels): a 128 x 256 tiling leaves most SMs idle; halve the tile width
    // while the tile count still fits one wave
    const long long m_tiles = (static_cast<long long>(d->n_img) * d->rows_per_img + kBlockM - 1) / kBlockM;
    while (block_n >= 128 && block_n % 64 == 0 && d->n % (block_n / 2) == 0 &&
           2 * m_tiles * ((d->n + block_n - 1) / block_n) <= sm_count() && getenv("CTTA_NO_NARROW_TILES") == nullptr)
      block_n /= 2;
  }
  p.block_n = block_n;
  p.n_tiles_n = (d->n + block_n - 1) / block_n;
  int acc_stride = 32;
  while (acc_stride < block_n) acc_stride *= 2;
  p.acc_stride = acc_stride;
  p.tmem_cols = 2 * acc_stride;
  p.k_chunks = (d->c + kBlockK - 1) / kBlockK;
  p.ntaps = d->ntaps;
  p.kk_last = (d->c - (p.k_chunks - 1) * kBlockK + 15) / 16;
  p.stage_bytes = kATileBytes + block_n * kBlockK * 2;
  p.H = d->h;
  p.W = d->w;
  p.n_img = d->n_img;
  // Output rows r * out_stride + out_off < 0 are dropped by definition: skip them up front by starting the logical
  // rows at q_start (the A taps shift by the same amount), so no tile ever maps to a negative output row.
  int q_start = 0;
  if (d->out_off < 0) {
    CTTA_REQUIRE(d->out_stride >= 1 && d->a_mode == CTTA_A_CONV1D, "ctta_gemm: negative out_off needs CONV1D mode");
    q_start = (-d->out_off + d->out_stride - 1) / d->out_stride;
  }
  CTTA_REQUIRE(d->out_stride >= 1, "ctta_gemm: out_stride must be >= 1");
  const int eff_out_off = q_start * d->out_stride + d->out_off;
  int eff_rows = d->rows_per_img - q_start;
  {
    // never compute rows whose output row lies past the end of the image
    const long long max_rows = eff_out_off < d->out_rows_per_img
                                   ? (static_cast<long long>(d->out_rows_per_img) - 1 - eff_out_off) / d->out_stride + 1
                                   : 0;
    if (eff_rows > max_rows) eff_rows = static_cast<int>(max_rows);
  }
  if (eff_rows <= 0) return 0;  // nothing to write
  if (d->out_up_phase)
    CTTA_REQUIRE(d->out_stride == 1 && d->out_off == 0, "ctta_gemm: upsample-phase output excludes out_stride / out_off");
  p.rows_per_img = eff_rows;
  for (int j = 0; j < d->ntaps; ++j) {
    p.tap_d0[j] = static_cast<short>(d->tap_d0[j] + q_start);
    p.tap_d1[j] = d->tap_d1[j];
  }

  const int c_pad = p.k_chunks * kBlockK;
  CUtensorMap tmap_a, tmap_b;
  const int esz = 2;
  if (d->a_mode == CTTA_A_ROWS) {
    p.n_tiles_m = (p.rows_per_img + kBlockM - 1) / kBlockM;
    cuuint64_t dims[2] = {(cuuint64_t)d->c, (cuuint64_t)d->rows_per_img};
    cuuint64_t strides[1] = {(cuuint64_t)d->a_ld * esz};
    cuuint32_t box[2] = {kBlockK, kBlockM};
    int rc = make_tmap(&tmap_a, p.is_bf16, d->a, 2, dims, strides, box);
    if (rc) return rc;
  } else if (d->a_mode == CTTA_A_CONV1D) {
    p.tiles_per_img = (p.rows_per_img + kBlockM - 1) / kBlockM;
    p.n_tiles_m = p.tiles_per_img * d->n_img;
    // halo mode: one activation box with all the rows any tap needs, weights of all taps resident in smem
    int min_shift = p.tap_d0[0], max_shift = p.tap_d0[0];
    for (int j = 1; j < d->ntaps; ++j) {
      if (p.tap_d0[j] < min_shift) min_shift = p.tap_d0[j];
      if (p.tap_d0[j] > max_shift) max_shift = p.tap_d0[j];
    }
    const int halo_rows = (kBlockM + (max_shift - min_shift) + 7) / 8 * 8;
    const int w_bytes = d->ntaps * block_n * 128;
    const bool halo_env = getenv("CTTA_NO_HALO") == nullptr;
    if (halo_env && p.k_chunks == 1 && p.n_tiles_n == 1 && halo_rows <= 256 && d->ntaps > 1 && w_bytes <= 96 * 1024) {
      p.halo = 1;
      p.halo_rows = halo_rows;
      p.halo_min_shift = min_shift;
      for (int j = 0; j < d->ntaps; ++j) p.tap_rows[j] = static_cast<short>(p.tap_d0[j] - min_shift);
      p.w_bytes = w_bytes;
      p.tiles_off = (w_bytes + 1023) / 1024 * 1024;
      p.stage_bytes = (halo_rows * 128 + 1023) / 1024 * 1024;
    }
    cuuint64_t dims[3] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->n_img};
    cuuint64_t strides[2] = {(cuuint64_t)d->a_ld * esz, (cuuint64_t)d->a_ld * esz * (cuuint64_t)d->w};
    cuuint32_t box[3] = {kBlockK, (cuuint32_t)(p.halo ? halo_rows : kBlockM), 1};
    int rc = make_tmap(&tmap_a, p.is_bf16, d->a, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    CTTA_REQUIRE(d->rows_per_img == d->h * d->w, "ctta_gemm: CONV2D needs rows_per_img == h*w");
    int bw = d->w < kBlockM ? d->w : kBlockM;
    CTTA_REQUIRE(kBlockM % bw == 0 && d->w % bw == 0, "ctta_gemm: CONV2D width %d must divide / be a multiple of 128", d->w);
    int bh = kBlockM / bw;
    if (bh > d->h) bh = d->h;
    CTTA_REQUIRE(d->h % bh == 0 && kBlockM % (bw * bh) == 0, "ctta_gemm: CONV2D height %d incompatible with 128-row tiles", d->h);
    const int bn = kBlockM / (bw * bh);
    p.box_w = bw;
    p.box_h = bh;
    p.box_n = bn;
    p.tiles_w = d->w / bw;
    p.tiles_h = d->h / bh;
    p.n_tiles_m = p.tiles_w * p.tiles_h * ((d->n_img + bn - 1) / bn);
    cuuint64_t dims[4] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n_img};
    cuuint64_t s1 = (cuuint64_t)d->a_ld * esz;
    cuuint64_t strides[3] = {s1, s1 * (cuuint64_t)d->w, s1 * (cuuint64_t)d->w * (cuuint64_t)d->h};
    cuuint32_t box[4] = {kBlockK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    int rc = make_tmap(&tmap_a, p.is_bf16, d->a, 4, dims, strides, box);
    if (rc) return rc;
  }
  if (d->wgt_img_stride != 0) {
    // batched GEMM: image i multiplies its own [n, c_pad] matrix (attention scores q.k^T and P.V per sample)
    CTTA_REQUIRE(d->a_mode == CTTA_A_CONV1D && d->ntaps == 1 && d->wgt_img_stride % 8 == 0 &&
                     d->wgt_img_stride >= static_cast<long long>(d->n) * c_pad,
                 "ctta_gemm: per-image weights need CONV1D mode with one tap and a 16-byte aligned image stride");
    cuuint64_t dims[3] = {(cuuint64_t)c_pad, (cuuint64_t)d->n, (cuuint64_t)d->n_img};
    cuuint64_t strides[2] = {(cuuint64_t)c_pad * esz, (cuuint64_t)d->wgt_img_stride * esz};
    cuuint32_t box[3] = {kBlockK, (cuuint32_t)block_n, 1};
    int rc = make_tmap(&tmap_b, p.is_bf16, d->wgt, 3, dims, strides, box);
    if (rc) return rc;
    p.w_batched = 1;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->ntaps * c_pad, (cuuint64_t)d->n};
    cuuint64_t strides[1] = {(cuuint64_t)d->ntaps * c_pad * esz};
    cuuint32_t box[2] = {kBlockK, (cuuint32_t)block_n};
    int rc = make_tmap(&tmap_b, p.is_bf16, d->wgt, 2, dims, strides, box);
    if (rc) return rc;
  }

  p.bias = d->bias;
  p.rowadd = d->rowadd;
  p.rowadd_ld = d->rowadd_ld;
  p.rowadd_rows = d->rowadd ? d->rowadd_rows : 1;
  p.act = d->act;
  p.act_slope = d->act_slope;
  p.residual = d->residual;
  p.res_neg_scale = d->res_neg_scale != 0.f ? d->res_neg_scale : 1.f;
  p.ring_slot_bytes = kRingSlotBytes;
  p.res_dtype = d->res_dtype;
  p.res_ld = d->res_ld;
  p.accumulate = d->accumulate;
  p.out_scale = d->out_scale;
  p.out = d->out;
  p.out_dtype = d->out_dtype;
  p.out_ld = d->out_ld;
  p.out2 = d->out2;
  p.out2_ld = d->out2_ld;
  p.act2 = d->act2;
  p.act2_slope = d->act2_slope;
  p.out_rows_per_img = d->out_rows_per_img;
  p.out_stride = d->out_stride;
  p.out_off = eff_out_off;

  bool vec = true;
  if (d->bias && !aligned16(d->bias)) vec = false;
  if (d->rowadd && (!aligned16(d->rowadd) || d->rowadd_ld % 4)) vec = false;
  if (d->residual && (!aligned16(d->residual) || d->res_ld % 8)) vec = false;
  if (d->out && (!aligned16(d->out) || d->out_ld % 8)) vec = false;
  if (d->out2 && (!aligned16(d->out2) || d->out2_ld % 8)) vec = false;
  p.vec_ok = vec ? 1 : 0;
  if (d->act == CTTA_ACT_GEGLU)
    CTTA_REQUIRE(aligned16(d->out) && d->out_ld % 4 == 0 && (!d->bias || aligned16(d->bias)),
                 "ctta_gemm: GEGLU output must be 16-byte aligned with out_ld %% 4 == 0");

  // ---- epilogue plan: TMA-staged whenever every row pitch is a multiple of 16 bytes
  const int out_esz = d->out ? (d->out_dtype == CTTA_F32 ? 4 : 2) : 0;
  bool tma_ok = d->n >= 32 && block_n % kChunkCols == 0;
  if (d->out && ((static_cast<long long>(d->out_ld) * out_esz) % 16 || !aligned16(d->out))) tma_ok = false;
  if (d->out2 && ((static_cast<long long>(d->out2_ld) * 2) % 16 || !aligned16(d->out2))) tma_ok = false;
  const bool res16 = d->residual && d->res_dtype != CTTA_F32;
  if (d->residual && !res16 && ((static_cast<long long>(d->res_ld) * 4) % 16 || !aligned16(d->residual))) tma_ok = false;
  if (res16 && (d->res_dtype != d->ab_dtype || (static_cast<long long>(d->res_ld) * 2) % 16 || !aligned16(d->residual) ||
                (d->out && d->out_dtype == CTTA_F32)))
    tma_ok = false;  // a 16-bit residual tile cannot double as the fp32 staging slot
  if (d->out && d->out_dtype != CTTA_F32 && d->out2) tma_ok = false;
  if (d->accumulate && d->out_dtype != CTTA_F32) tma_ok = false;
  if (d->act == CTTA_ACT_GEGLU && d->out_dtype == CTTA_F32) tma_ok = false;
  if (d->bias && !((reinterpret_cast<uintptr_t>(d->bias) & 3) == 0)) tma_ok = false;
  if (d->rowadd && (!aligned16(d->rowadd) || d->rowadd_ld % 4)) tma_ok = false;
  // ---- stream mode (see GemmKParams::stream): multi-tap convolutions whose epilogue can go through TMA
  p.sub_tiles = 1;
  if (tma_ok && !p.halo && d->ntaps > 1 && d->a_mode != CTTA_A_ROWS && getenv("CTTA_NO_STREAM") == nullptr) {
    // sub-tiles sharing one weight chunk: two accumulator stages of st * block_n columns must fit the 512 TMEM
    // columns so that the epilogue of one tile overlaps the MMAs of the next (N = 256 keeps st = 1)
    int st = 256 / block_n;
    if (st > 4) st = 4;
    if (st < 1) st = 1;
    if (getenv("CTTA_STREAM_ST")) st = atoi(getenv("CTTA_STREAM_ST"));
    bool ok = false;
    int box_rows = 0, a_loads = 1, a_stage = 0, box_h = 0;
    if (d->a_mode == CTTA_A_CONV1D) {
      int min_shift = p.tap_d0[0], max_shift = p.tap_d0[0];
      for (int j = 1; j < d->ntaps; ++j) {
        if (p.tap_d0[j] < min_shift) min_shift = p.tap_d0[j];
        if (p.tap_d0[j] > max_shift) max_shift = p.tap_d0[j];
      }
      const long long row_tiles = (p.rows_per_img + kBlockM - 1) / kBlockM;
      while (st > 1 && ((row_tiles + st - 1) / st) * d->n_img * p.n_tiles_n < 2LL * sm_count()) st /= 2;
      const int need = st * kBlockM + (max_shift - min_shift);
      a_loads = (need + 255) / 256;
      box_rows = ((need + a_loads - 1) / a_loads + 7) / 8 * 8;
      a_stage = (a_loads * box_rows * 128 + 1023) / 1024 * 1024;
      p.n_groups = 1;
      p.grp_first[0] = 0;
      p.grp_count[0] = static_cast<short>(d->ntaps);
      p.grp_shift[0] = 0;
      for (int j = 0; j < d->ntaps; ++j) {
        p.tap_id[j] = static_cast<short>(j);
        p.tap_rows[j] = static_cast<short>(p.tap_d0[j] - min_shift);
      }
      p.a_row0_shift = min_shift;
      ok = true;
    } else {
      // CONV2D: a tile is box_h full-width image rows = st * 128 pixels of ONE image; one box per horizontal shift
      // holds rows [h0 + dh_min, h0 + box_h + dh_max) so the taps sharing that shift are row-shifted views of it
      int dh_min = 0, dh_max = 0, dws[4], ndw = 0;
      bool taps_ok = true;
      for (int j = 0; j < d->ntaps; ++j) {
        if (p.tap_d1[j] < dh_min) dh_min = p.tap_d1[j];
        if (p.tap_d1[j] > dh_max) dh_max = p.tap_d1[j];
        int k = 0;
        while (k < ndw && dws[k] != p.tap_d0[j]) ++k;
        if (k == ndw) {
          if (ndw == 4) { taps_ok = false; break; }
          dws[ndw++] = p.tap_d0[j];
        }
      }
      for (; st >= 1 && taps_ok; st /= 2) {
        const int pix = st * kBlockM;
        if (d->w <= kBlockM && pix % d->w == 0 && d->h % (pix / d->w) == 0 &&
            static_cast<long long>(d->h / (pix / d->w)) * d->n_img * p.n_tiles_n >= (st > 1 ? 2LL * sm_count() : 1)) {
          box_h = pix / d->w;
          ok = box_h + (dh_max - dh_min) <= 256;
          break;
        }
      }
      if (ok) {
        box_rows = (box_h + dh_max - dh_min) * d->w;
        a_stage = (box_rows * 128 + 1023) / 1024 * 1024;
        p.n_groups = ndw;
        int t = 0;
        for (int g = 0; g < ndw; ++g) {
          p.grp_first[g] = static_cast<short>(t);
          p.grp_shift[g] = static_cast<short>(dws[g]);
          for (int j = 0; j < d->ntaps; ++j) {
            if (p.tap_d0[j] != dws[g]) continue;
            p.tap_id[t] = static_cast<short>(j);
            p.tap_rows[t] = static_cast<short>((p.tap_d1[j] - dh_min) * d->w);
            ++t;
          }
          p.grp_count[g] = static_cast<short>(t - p.grp_first[g]);
        }
        p.a_row0_shift = dh_min;
      }
    }
    // shared-memory feasibility: 2 A stages + 2 weight stages + the smallest epilogue plan
    const int w_stage = block_n * 128;
    const bool f32o = d->out && d->out_dtype == CTTA_F32;
    const int min_epi = ((f32o || d->residual) ? 3 * kRingSlotBytes : 0) +
                        (((d->out && !f32o) || d->out2) ? 4 * 2 * kStage16Bytes : 0) + kBiasBytes;
    if (ok && 2 * a_stage + 2 * w_stage + min_epi <= kSmemBudget) {
      p.stream = 1;
      p.sub_tiles = st;
      p.a_loads = a_loads;
      p.a_box_rows = box_rows;
      p.a_stage_bytes = a_stage;
      p.n_a_stages = 2;
      if (3 * a_stage + 4 * w_stage + min_epi + 2 * kRingSlotBytes + 4 * 2 * kStage16Bytes <= kSmemBudget) p.n_a_stages = 3;
      p.w_stage_bytes = w_stage;
      p.stage_bytes = w_stage;                       // the generic planner below sizes the weight ring
      p.tiles_off = p.n_a_stages * a_stage;          // ... after the A ring
      p.w_ring_off = p.tiles_off;
      int acc = 32;
      while (acc < st * block_n) acc *= 2;
      p.acc_stride = acc;
      p.acc_single = 2 * acc > 512 ? 1 : 0;
      p.tmem_cols = p.acc_single ? acc : 2 * acc;
      if (d->a_mode == CTTA_A_CONV1D) {
        const long long row_tiles = (p.rows_per_img + kBlockM - 1) / kBlockM;
        p.tiles_per_img = static_cast<int>((row_tiles + st - 1) / st);
        p.n_tiles_m = p.tiles_per_img * d->n_img;
        cuuint64_t dims[3] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->n_img};
        cuuint64_t strides[2] = {(cuuint64_t)d->a_ld * esz, (cuuint64_t)d->a_ld * esz * (cuuint64_t)d->w};
        cuuint32_t box[3] = {kBlockK, (cuuint32_t)box_rows, 1};
        int rc = make_tmap(&tmap_a, p.is_bf16, d->a, 3, dims, strides, box);
        if (rc) return rc;
      } else {
        p.box_w = d->w;
        p.box_h = box_h;
        p.box_n = 1;
        p.tiles_w = 1;
        p.tiles_h = d->h / box_h;
        p.n_tiles_m = p.tiles_h * d->n_img;
        cuuint64_t dims[4] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n_img};
        cuuint64_t s1 = (cuuint64_t)d->a_ld * esz;
        cuuint64_t strides[3] = {s1, s1 * (cuuint64_t)d->w, s1 * (cuuint64_t)d->w * (cuuint64_t)d->h};
        cuuint32_t box[4] = {kBlockK, (cuuint32_t)d->w, (cuuint32_t)(box_rows / d->w), 1};
        int rc = make_tmap(&tmap_a, p.is_bf16, d->a, 4, dims, strides, box);
        if (rc) return rc;
      }
    }
  }
  int rows_tile_img = kBlockM;
  if (d->a_mode == CTTA_A_CONV2D && !p.stream) {
    rows_tile_img = kBlockM / p.box_n;
    if (p.tiles_w != 1 || rows_tile_img % 32 != 0) tma_ok = false;
  }
  p.out_rows_tile_img = rows_tile_img;
  p.epi_tma = tma_ok ? 1 : 0;
  p.epi_groups = 1;
  if (d->stats) {
    const int cpg = d->stats_groups > 0 ? d->n / d->stats_groups : 0;
    const bool pow2 = cpg >= 4 && (cpg & (cpg - 1)) == 0;
    int n_stat_img = d->n_img;
    bool ok = tma_ok && d->act != CTTA_ACT_GEGLU && d->out_stride == 1 && d->out_off == 0 && !d->accumulate && pow2 &&
              d->stats_groups * cpg == d->n;
    if (d->a_mode == CTTA_A_ROWS) {
      ok = ok && d->stats_rows_per_img > 0 && d->stats_rows_per_img % 32 == 0 && d->rows_per_img % d->stats_rows_per_img == 0;
      if (ok) n_stat_img = d->rows_per_img / d->stats_rows_per_img;
    }
    if (!ok)
      return set_error(CTTA_ERR_UNSUPPORTED, "ctta_gemm: fused GroupNorm moments unsupported for this problem (n=%d groups=%d)",
                       d->n, d->stats_groups);
    if (!d->stats_keep) CTTA_CUDA(cudaMemsetAsync(d->stats, 0, sizeof(float) * 2 * n_stat_img * d->stats_groups, stream));
    p.stats = d->stats;
    p.stats_groups = d->stats_groups;
    p.stats_cpg = cpg;
    p.stats_rows = d->a_mode == CTTA_A_ROWS ? d->stats_rows_per_img : 1;
  }
  // M super-tiles for narrow N (see GemmKParams::sub_tiles)
  if (!p.stream && tma_ok && d->a_mode != CTTA_A_CONV2D && !d->rowadd && block_n <= 128 && getenv("CTTA_NO_SUPERTILE") == nullptr) {
    int st = 256 / block_n;
    if (st > 4) st = 4;
    const long long row_tiles = (p.rows_per_img + kBlockM - 1) / kBlockM;
    while (st > 1 && ((row_tiles + st - 1) / st) * d->n_img * p.n_tiles_n < 2LL * sm_count()) st /= 2;
    p.sub_tiles = st;
    if (d->a_mode == CTTA_A_ROWS) {
      p.n_tiles_m = static_cast<int>((row_tiles + st - 1) / st);
    } else {
      p.tiles_per_img = static_cast<int>((row_tiles + st - 1) / st);
      p.n_tiles_m = p.tiles_per_img * d->n_img;
    }
    int acc = 32;
    while (acc < st * block_n) acc *= 2;
    p.acc_stride = acc;
    p.tmem_cols = 2 * acc;
  }
  CUtensorMap tmap_out, tmap_out2, tmap_res;
  memset(&tmap_out, 0, sizeof(tmap_out));
  memset(&tmap_out2, 0, sizeof(tmap_out2));
  memset(&tmap_res, 0, sizeof(tmap_res));
  int ring_bytes = 0, st16_bytes = 0, bias_bytes = 0;
  if (d->out_up_phase && !tma_ok)
    return set_error(CTTA_ERR_UNSUPPORTED, "ctta_gemm: upsample-phase output needs the TMA epilogue (n >= 32, aligned pitches)");
  if (tma_ok) {
    // geometry of the row maps (see ctta_gemm_desc: out row = r * out_stride + out_off, dropped outside the image)
    const long long first_row = eff_out_off;
    const long long n_rows = p.rows_per_img;
    p.row_coord_shift = 0;
    const bool geglu = d->act == CTTA_ACT_GEGLU;
    // ---- 16-bit-only epilogues: 128-byte TMA rows (the TMA unit's cost is per box row, ~3 cycles whatever its width)
    const bool f32_out = d->out && d->out_dtype == CTTA_F32;
    const bool io16 = !f32_out && !geglu && (!d->residual || res16) && !d->accumulate && !d->out_up_phase &&
                      getenv("CTTA_NO_WIDE16") == nullptr;
    p.cw = (io16 && block_n % 64 == 0) ? 2 : 1;
    const bool geglu_wide = geglu && block_n % 128 == 0 && getenv("CTTA_NO_WIDE16") == nullptr;
    if (geglu_wide) p.cw = 4;   // 4 x 32 accumulator columns = 64 outputs = one 128-byte row
    {
      // N = 32 tensors (HiFi-GAN last stage) viewed as [rows / 2, 64]: every pitch must be exactly 32 columns
      const void* o16 = d->out ? d->out : d->out2;
      const int o16_ld = d->out ? d->out_ld : d->out2_ld;
      p.pair = (io16 && p.cw == 1 && block_n == 32 && d->n == 32 && o16 != nullptr && o16_ld == 32 &&
                (!d->residual || d->res_ld == 32) && d->a_mode != CTTA_A_CONV2D && d->out_stride == 1 && first_row % 2 == 0 &&
                n_rows % 2 == 0 && d->out_rows_per_img % 2 == 0)
                   ? 1 : 0;
    }
    p.st16_bufs = p.cw >= 2 ? 1 : 2;
    p.st16_bytes = p.cw >= 2 ? 2 * kStage16Bytes : kStage16Bytes;
    // row map of a 16-bit epilogue tensor in the layout chosen above
    auto make_map16 = [&](CUtensorMap* m, int dtype, const void* base, long long ld) -> int {
      if (p.pair)
        return make_row_tmap(m, dtype, base, 64, 64, first_row / 2, 1, n_rows / 2, d->out_rows_per_img / 2, d->n_img, 64,
                             CU_TENSOR_MAP_SWIZZLE_128B, 16);
      if (p.cw == 2)
        return make_row_tmap(m, dtype, base, ld, d->n, first_row, d->out_stride, n_rows, d->out_rows_per_img, d->n_img, 64,
                             CU_TENSOR_MAP_SWIZZLE_128B);
      return make_row_tmap(m, dtype, base, ld, d->n, first_row, d->out_stride, n_rows, d->out_rows_per_img, d->n_img,
                           kChunkCols, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    if (d->out_up_phase) {
      // logical pixel (h, w) -> output pixel (2h + ph, 2w + pw): a warp's 32 consecutive pixels are a {wb x 32 / wb} box
      const int ph = (d->out_up_phase - 1) >> 1, pw = (d->out_up_phase - 1) & 1;
      const bool ok = d->out_up_phase >= 1 && d->out_up_phase <= 4 && d->a_mode == CTTA_A_CONV2D && p.stream && d->out &&
                      d->out_dtype == CTTA_F32 && !d->out2 && !d->residual && !d->accumulate &&
                      (d->w >= 32 ? d->w % 32 == 0 : 32 % d->w == 0);
      if (!ok) return set_error(CTTA_ERR_UNSUPPORTED, "ctta_gemm: upsample-phase output unsupported for this problem (h=%d w=%d)", d->h, d->w);
      const int wb = d->w >= 32 ? 32 : d->w;
      const long long pix = static_cast<long long>(d->out_ld) * 4;   // bytes per output pixel
      const char* base = reinterpret_cast<const char*>(d->out) + (static_cast<long long>(ph) * 2 * d->w + pw) * pix;
      cuuint64_t dims[4] = {(cuuint64_t)d->n, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n_img};
      cuuint64_t strides[3] = {(cuuint64_t)(2 * pix), (cuuint64_t)(4LL * d->w * pix), (cuuint64_t)(4LL * d->w * d->h * pix)};
      cuuint32_t box[4] = {(cuuint32_t)kChunkCols, (cuuint32_t)wb, (cuuint32_t)(32 / wb), 1};
      int rc = make_tmap_ex(&tmap_out, CTTA_F32, CU_TENSOR_MAP_SWIZZLE_128B, base, 4, dims, strides, box);
      if (rc) return rc;
      p.out4d = 1;
    } else if (d->out) {
      const bool f32 = d->out_dtype == CTTA_F32;
      int rc;
      if (f32 || geglu)
        rc = make_row_tmap(&tmap_out, d->out_dtype, d->out, d->out_ld, geglu ? d->n / 2 : d->n, first_row, d->out_stride,
                           n_rows, d->out_rows_per_img, d->n_img, geglu ? (geglu_wide ? 64 : 16) : kChunkCols,
                           (f32 || geglu_wide) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
      else
        rc = make_map16(&tmap_out, d->out_dtype, d->out, d->out_ld);
      if (rc) return rc;
    }
    if (d->out2) {
      int rc = make_map16(&tmap_out2, d->ab_dtype, d->out2, d->out2_ld);
      if (rc) return rc;
    }
    if (d->residual) {
      int rc = res16 ? make_map16(&tmap_res, d->res_dtype, d->residual, d->res_ld)
                     : make_row_tmap(&tmap_res, CTTA_F32, d->residual, d->res_ld, d->n, first_row, d->out_stride, n_rows,
                                     d->out_rows_per_img, d->n_img, kChunkCols, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    p.res16 = res16 ? 1 : 0;
    p.res_neg_scale = d->res_neg_scale != 0.f ? d->res_neg_scale : 1.f;
    const int slot_b = res16 ? (p.cw == 2 ? kRingSlotBytes : kRingSlotBytes / 2) : kRingSlotBytes;
    p.ring_slot_bytes = slot_b;
    p.ring_in = d->residual ? 1 : 0;
    const bool out_f32 = d->out && d->out_dtype == CTTA_F32;
    const bool use_ring = out_f32 || p.ring_in > 0;
    p.ring_per_chunk = use_ring ? (p.ring_in > 1 ? p.ring_in : 1) : 0;
    const bool need16 = (d->out && !out_f32) || d->out2;
    bias_bytes = kBiasBytes;
    int stages = 0, ring = 0, groups = 2;
    // two groups of epilogue warps double the epilogue throughput but cost staging memory and ring depth; fall
    // back to one group when that would leave fewer than 3 mainloop stages (wide, MMA-bound tiles).  (A third group in
    // a 512-thread instance was measured on the HiFi-GAN C <= 128 convolutions: <= 5 %, they are bound by the TMA
    // unit's request rate, not by epilogue warps.)
    for (groups = 2; groups >= 1 && stages == 0; --groups) {
      st16_bytes = need16 ? groups * 4 * 2 * kStage16Bytes : 0;
      const int budget = kSmemBudget - p.tiles_off - st16_bytes - bias_bytes;
      if (!use_ring) {
        const int st = budget / p.stage_bytes;
        if (st >= 3 || groups == 1) stages = st;
      } else {
        // every warp group holds one slot while it works on a chunk; the rest is the loader's prefetch distance
        const int need = (groups > 2 ? groups : 2) * p.ring_per_chunk;
        const int prefs[3] = {need + 3, need + 2, need + 1};
        for (int i = 0; i < 3 && stages == 0; ++i) {
          const int st = (budget - prefs[i] * slot_b) / p.stage_bytes;
          if (st >= 3 || (groups == 1 && i == 2 && st >= 2)) {
            stages = st;
            ring = prefs[i];
          }
        }
      }
      if (stages > 0) break;
    }
    CTTA_REQUIRE(stages > 0, "ctta_gemm: shared-memory plan failed (block_n=%d)", block_n);
    p.epi_groups = groups;
    const int budget = kSmemBudget - p.tiles_off - st16_bytes - bias_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    CTTA_REQUIRE(stages >= 2, "ctta_gemm: shared-memory plan failed (block_n=%d, stages=%d)", block_n, stages);
    if (use_ring) {
      // leftover room extends the ring
      int extra = (budget - stages * p.stage_bytes - ring * slot_b) / slot_b;
      ring += extra;
      if (ring > kMaxRing) ring = kMaxRing;
    }
    p.n_stages = stages;
    p.ring_slots = ring;
    ring_bytes = ring * slot_b;
    p.ring_off = p.tiles_off + stages * p.stage_bytes;
    p.stage16_off = p.ring_off + ring_bytes;
    p.bias_off = p.stage16_off + st16_bytes;
  } else {
    int n_stages = (kSmemBudget - p.tiles_off) / p.stage_bytes;
    if (n_stages > kMaxStages) n_stages = kMaxStages;
    p.n_stages = n_stages;
  }
  const int smem_bytes = p.tiles_off + p.n_stages * p.stage_bytes + ring_bytes + st16_bytes + bias_bytes + 1024;
  int total_tiles = p.n_tiles_m * p.n_tiles_n;
  int grid = sm_count();
  // ---- CTA pairs with multicast weight chunks (see GemmKParams::w_mcast): stream mode, enough tiles for every pair
  const bool mc_generic = !p.stream && !p.halo && p.epi_tma && getenv("CTTA_NO_MCAST_GENERIC") == nullptr;
  int cs = getenv("CTTA_MCAST_CS") ? atoi(getenv("CTTA_MCAST_CS")) : 2;   // CTAs per cluster
  if (cs != 2 && cs != 4) cs = 2;
  if ((p.stream || mc_generic) && !p.w_batched && block_n >= 128 && block_n % (8 * cs) == 0 && p.n_tiles_m >= cs &&
      (grid % cs) == 0 && total_tiles >= 2 * grid && getenv("CTTA_NO_MCAST") == nullptr) {
    p.w_mcast = cs;
    p.total_tiles_mc = cs * ((p.n_tiles_m + cs - 1) / cs) * p.n_tiles_n;
    total_tiles = p.total_tiles_mc;
    cuuint64_t dims[2] = {(cuuint64_t)d->ntaps * c_pad, (cuuint64_t)d->n};
    cuuint64_t strides[1] = {(cuuint64_t)d->ntaps * c_pad * esz};
    cuuint32_t box[2] = {kBlockK, (cuuint32_t)(block_n / cs)};
    int rc = make_tmap(&tmap_b, p.is_bf16, d->wgt, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (grid > total_tiles) grid = total_tiles;

  // ---- pick the compile-time specialised epilogue instance (fp16 operands / outputs only), else the generic one
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const CUtensorMap, const GemmKParams);
  struct Variant {
    int act, rowadd, ring_in, out_f32, out_16, out2, act2;
    KernelFn fn;
  };
  constexpr int N_ = CTTA_ACT_NONE, L_ = CTTA_ACT_LRELU, G_ = CTTA_ACT_GEGLU;
  static const Variant variants[] = {
      {N_, 0, 0, 0, 0, 1, L_, gemm_tc_kernel<EpiCfg<N_, 0, 0, 0, 0, 1, L_>>},  // conv -> lrelu'ed 16-bit operand
      {N_, 0, 1, 1, 0, 1, L_, gemm_tc_kernel<EpiCfg<N_, 0, 1, 1, 0, 1, L_>>},  // + residual, fp32 stream + operand
      {N_, 0, 1, 0, 0, 1, L_, gemm_tc_kernel<EpiCfg<N_, 0, 1, 0, 0, 1, L_>>},  // + 16-bit residual -> lrelu'ed operand
      {N_, 0, 0, 1, 0, 1, L_, gemm_tc_kernel<EpiCfg<N_, 0, 0, 1, 0, 1, L_>>},  // transposed conv
      {N_, 0, 1, 1, 0, 0, N_, gemm_tc_kernel<EpiCfg<N_, 0, 1, 1, 0, 0, N_>>},  // fp32 out + residual
      {N_, 1, 0, 1, 0, 0, N_, gemm_tc_kernel<EpiCfg<N_, 1, 0, 1, 0, 0, N_>>},  // fp32 out + time embedding
      {N_, 0, 0, 1, 0, 0, N_, gemm_tc_kernel<EpiCfg<N_, 0, 0, 1, 0, 0, N_>>},  // fp32 out
      {N_, 0, 0, 0, 1, 0, N_, gemm_tc_kernel<EpiCfg<N_, 0, 0, 0, 1, 0, N_>>},  // 16-bit out
      {N_, 1, 0, 0, 1, 0, N_, gemm_tc_kernel<EpiCfg<N_, 1, 0, 0, 1, 0, N_>>},  // 16-bit out + time embedding
      {N_, 0, 1, 0, 1, 0, N_, gemm_tc_kernel<EpiCfg<N_, 0, 1, 0, 1, 0, N_>>},  // 16-bit out + fp32 residual
      {G_, 0, 0, 0, 1, 0, N_, gemm_tc_kernel<EpiCfg<G_, 0, 0, 0, 1, 0, N_>>},  // GEGLU
  };
  KernelFn fn = p.epi_tma ? gemm_tc_kernel<EpiGeneric> : gemm_tc_kernel<EpiGeneric, kThreads, 1>;
  if (p.epi_tma && !p.is_bf16 && d->out_dtype != CTTA_BF16 && getenv("CTTA_GENERIC_EPILOGUE") == nullptr) {
    const int k_act = d->act, k_rowadd = d->rowadd ? 1 : 0, k_ring = p.ring_in;
    const int k_f32 = (d->out && d->out_dtype == CTTA_F32) ? 1 : 0, k_16 = (d->out && d->out_dtype != CTTA_F32) ? 1 : 0;
    const int k_out2 = d->out2 ? 1 : 0, k_act2 = d->out2 ? d->act2 : CTTA_ACT_NONE;
    for (const Variant& v : variants) {
      if (v.act == k_act && v.rowadd == k_rowadd && v.ring_in == k_ring && v.out_f32 == k_f32 && v.out_16 == k_16 &&
          v.out2 == k_out2 && v.act2 == k_act2) {
        fn = v.fn;
        break;
      }
    }
  }
  const int n_threads = kThreads;
  {
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(fn), kSmemMaxDynamic);
    if (rc) return rc;
  }
  if (p.w_mcast) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(n_threads));
    cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(p.w_mcast);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // co-resident clusters (a GPC with an odd number of free SMs cannot host a last pair): size the persistent grid to them
    {
      static std::mutex mu2;
      static std::map<std::pair<const void*, int>, int> max_clusters;   // key: (fn, device * 16 + cluster size)
      std::lock_guard<std::mutex> lock(mu2);
      const std::pair<const void*, int> ckey(reinterpret_cast<const void*>(fn), current_device() * 16 + p.w_mcast);
      auto itc = max_clusters.find(ckey);
      int n_cl = 0;
      if (itc == max_clusters.end()) {
        cudaLaunchConfig_t q = cfg;
        q.dynamicSmemBytes = kSmemMaxDynamic;
        CTTA_CUDA(cudaOccupancyMaxActiveClusters(&n_cl, fn, &q));
        max_clusters[ckey] = n_cl;
        if (getenv("CTTA_DEBUG") != nullptr)
          fprintf(stderr, "ctta_gemm: %d co-resident clusters of %d CTAs (persistent grid %d, %d SMs)\n", n_cl, p.w_mcast, grid,
                  sm_count());
      } else {
        n_cl = itc->second;
      }
      if (n_cl >= 1 && p.w_mcast * n_cl < grid) cfg.gridDim = dim3(static_cast<unsigned>(p.w_mcast * n_cl));
    }
    CTTA_CUDA(cudaLaunchKernelEx(&cfg, fn, tmap_a, tmap_b, tmap_out, tmap_out2, tmap_res, p));
    ::ctta::count_launch();
    return 0;
  }
  fn<<<grid, n_threads, smem_bytes, stream>>>(tmap_a, tmap_b, tmap_out, tmap_out2, tmap_res, p);
  CTTA_LAUNCH_CHECK();
  return 0;
}
