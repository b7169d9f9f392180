// Fused HiFi-GAN ResBlock pair for the narrow stages (C = 32 / 64):
//
//     x' = x + c2( lrelu( c1( lrelu(x) ) ) )        audioldm/hifigan/models.py:56-63 (one (c1, c2) pair of a ResBlock)
//
// in ONE kernel, on the LeakyReLU'ed 16-bit streams of vae.Generator.forward_btc: in = lx = lrelu(x) [B, T, C], out =
// lrelu(x') [B, T, C]; c1 is the dilated conv (k taps, dilation d), c2 the plain one (k taps).  The unfused path runs two
// implicit-GEMM launches and round-trips the hidden tensor through HBM (2+2 bytes per element for c1, 2+2+2 for c2);
// here the hidden tile never leaves shared memory: HBM traffic is lx in + lrelu(x') out = 2+2 bytes per element.
//
// Tile = 256 rows of time (two 128-row UMMA tiles) of the hidden tensor h, i.e. L = 256 - (k - 1) output rows:
//   X   TMA box of lx rows [t0 - p2 - p1, ...), p1 = d (k-1)/2, p2 = (k-1)/2, out-of-range rows zero filled (= the convs'
//       zero padding), K-major SW128 rows of 128 B (C = 32: the upper 64 B are TMA zero fill and never read)
//   c1  acc1[m] += X[rows m*128 + j*d ...] . W1[j]      tap j = a row-shifted UMMA view of the resident box
//   E1  h = lrelu(acc1 + b1) (0 outside [0, T): c2's zero padding) -> 16-bit K-major SW128 tile H in shared memory
//   c2  acc2[m] += H[rows m*128 + j ...] . W2[j]
//   E2  lrelu(acc2 + b2 + x) with x = un-lrelu(lx) re-read from global memory (L2 hits) -> 16-bit staging tile -> TMA store
//       of L rows
// All taps of W1 and W2 stay resident in shared memory for the whole (persistent) CTA.  Warp roles: warp 0 = TMA
// loader (X ring), warp 1 = single-thread tcgen05.mma issuer, warps 2..9 = E1, warps 10..17 = E2 (thread == row of one
// 128-row tile in both).  Tiles run through S "slots" (accumulators + H buffer), so that c1 of tile n + 1, E1 of tile
// n + 1 and E2 of tile n overlap.
//
// Shared-memory plan per shape (host side below): two slots wherever they fit, X ring 1-3 deep, and 0 / 1 / 2 own output
// staging buffers (with own buffers the hidden-tile slot is released by the issuer's tcgen05.commit after c2; without,
// the output is staged in the slot itself and released by the E2 warps once their TMA stores have read it).
//
// What bounds it (tools/trace_pair.py, profiles/r2_resblock_pair.txt): not HBM.  The MMA count is that of the two-launch
// path (~45-60 cycles per tcgen05.mma at N <= 64: every instruction re-reads a 128 x 16 A slice from shared memory); at
// C = 32 the E2 warp set is the longest stage (~2000 of ~2400 cycles per tile at k = 3), at C = 64 the residual loads
// (thread == row puts the lanes of a warp 128 B apart: 32 distinct lines per load instruction).  Per pair it runs 1.2-1.9 x
// faster than two launches, bit-identically.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "ctta_internal.h"
#include "ctta_ptx.cuh"

namespace ctta {

int make_tmap_ex(CUtensorMap* m, int dtype, CUtensorMapSwizzle swz, const void* base, int rank, const cuuint64_t* dims,
                 const cuuint64_t* strides_bytes, const cuuint32_t* box);

namespace rbp {

constexpr int kThreads = 576;          // 18 warps
constexpr int kEpiWarp0 = 2;           // warps 2..9: E1 (hidden tile)
constexpr int kEpi2Warp0 = 10;         // warps 10..17: E2 (output tile)
constexpr int kHRows = 256;            // hidden rows per tile (two UMMA M tiles)
constexpr int kHBufRows = 272;         // H buffer: the second M tile's taps read up to 10 rows past 256
constexpr int kMaxX = 4, kMaxSlots = 2;

struct Params {
  int C, taps, dil, p1, p2, L;         // L = output rows per tile
  int T, batch, tiles_per_seq, total_tiles;
  int rx, rxh;                         // rows of the X box (two TMA boxes of rxh rows)
  int n_x, n_slots;                    // X ring depth, accumulator / H slots
  int w_tap_bytes;                     // C * 128
  int off_w1, off_w2, off_x, x_bytes, off_h, h_bytes;
  int x_al32;                          // x is 32-byte aligned: residual rows by 256-bit loads
  int off_st, n_st;                    // own output staging buffers (0 = the output rows are staged in H(s))
  const float* b1;
  const float* b2;
  long long* trace;                    // RBP_TRACE builds only (tools/trace_pair.py): clock64 stamps of CTA 0, else null
  const void* x;                       // the input tensor itself: residual rows are re-read from global memory (an L2 hit)
  float slope, inv_slope;
  int is_bf16;
};

// Phase stamps of CTA 0 (20 slots per tile) for tools/trace_pair.py; compiled out of the product library.
#ifndef RBP_TRACE
#define RBP_TRACE 0
#endif
#if RBP_TRACE
#define RBP_STAMP(ev, n)                                                                   \
  do {                                                                                     \
    if (blockIdx.x == 0 && (n) < 48 && p.trace) p.trace[(n) * 20 + (ev)] = clock64();      \
  } while (0)
#else
#define RBP_STAMP(ev, n) \
  do {                   \
  } while (0)
#endif

struct Bars {
  uint64_t w_full, x_full[kMaxX], x_empty[kMaxX], acc1_full[kMaxSlots], acc1_free[kMaxSlots], h_full[kMaxSlots],
      h_free[kMaxSlots], acc2_full[kMaxSlots], acc2_free[kMaxSlots];
};
#define RB_BAR(field, idx) (bar_base + static_cast<uint32_t>(offsetof(Bars, field)) + 8u * static_cast<uint32_t>(idx))

template <int BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  if (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <int BF16>
__device__ __forceinline__ void unpack2(uint32_t v, float& a, float& b) {
  if (BF16) {
    a = __uint_as_float(v << 16);
    b = __uint_as_float(v & 0xFFFF0000u);
  } else {
    const __half2 h = *reinterpret_cast<const __half2*>(&v);
    const float2 f = __half22float2(h);
    a = f.x;
    b = f.y;
  }
}
// Packed fp32 pairs (FADD2 / FMUL2 on sm_100): IEEE round-to-nearest like the scalar forms, half the issue slots — the
// epilogue warps are issue / latency bound (two E1 + two E2 warps per scheduler).
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
// v > 0 ? v : s * v  ==  max(v, s v) for s in (0, 1],  min(v, s v) for s > 1 (the inverse map)
template <bool kInverse>
__device__ __forceinline__ void lrelu2(float& v0, float& v1, float s) {
  uint64_t a, b;
  float m0, m1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(v0), "f"(v1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(s));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(m0), "=f"(m1) : "l"(a));
  if (kInverse) {
    v0 = fminf(v0, m0);
    v1 = fminf(v1, m1);
  } else {
    v0 = fmaxf(v0, m0);
    v1 = fmaxf(v1, m1);
  }
}

template <int C, int BF16, int S>
__global__ void __launch_bounds__(kThreads, 1)
resblock_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                     const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out,
                     const __grid_constant__ CUtensorMap tmap_tail, const __grid_constant__ Params p) {
  constexpr int KK = C / 16;            // 16-element K slices per tap
  constexpr int NCH = C / 8;            // 16-byte chunks per row
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Bars bars_s;
  __shared__ __align__(16) float bias_s[2][64];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = pin_u32(smem_u32(&bars_s));
  const int n_my = (p.total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int NX = p.n_x;          // S (accumulator / H slots) is a template parameter: the epilogue's register sets are picked statically

  if (threadIdx.x == 0) {
    mbar_init(RB_BAR(w_full, 0), 1);
    for (int i = 0; i < kMaxX; ++i) {
      mbar_init(RB_BAR(x_full, i), 1);
      mbar_init(RB_BAR(x_empty, i), 1);        // tcgen05.commit of the issuer, once c1 has consumed the box
    }
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(RB_BAR(acc1_full, i), 1);
      mbar_init(RB_BAR(acc1_free, i), 8);
      mbar_init(RB_BAR(h_full, i), 8);
      mbar_init(RB_BAR(h_free, i), p.n_st == 0 ? 8 : 1);   // shared staging: every E2 warp, once its TMA store has read its rows; else the issuer's commit after c2
      mbar_init(RB_BAR(acc2_full, i), 1);
      mbar_init(RB_BAR(acc2_free, i), 8);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_out);
  }
  if (threadIdx.x < 128) bias_s[threadIdx.x >> 6][threadIdx.x & 63] = (threadIdx.x & 63) < C ? ((threadIdx.x >> 6) ? p.b2 : p.b1)[threadIdx.x & 63] : 0.f;
  const int tmem_cols = (S * 4 * C) <= 128 ? 128 : ((S * 4 * C) <= 256 ? 256 : 512);
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader: weights once, then the X ring
    if (elect_one_sync()) {
      const uint32_t wf = RB_BAR(w_full, 0);
      mbar_arrive_expect_tx(wf, 2u * p.taps * p.w_tap_bytes);
      for (int j = 0; j < p.taps; ++j) {
        tma_load_2d(base + p.off_w1 + j * p.w_tap_bytes, &tmap_w1, wf, j * 64, 0);
        tma_load_2d(base + p.off_w2 + j * p.w_tap_bytes, &tmap_w2, wf, j * 64, 0);
      }
      for (int n = 0; n < n_my; ++n) {
        const int tile = blockIdx.x + n * gridDim.x;
        const int b = tile / p.tiles_per_seq, t0 = (tile - b * p.tiles_per_seq) * p.L;
        const int dx = n % NX;
        mbar_wait(RB_BAR(x_empty, dx), static_cast<uint32_t>(((n / NX) & 1) ^ 1));
        const uint32_t full = RB_BAR(x_full, dx);
        RBP_STAMP(0, n);
        mbar_arrive_expect_tx(full, static_cast<uint32_t>(p.rx) * 128u);
        const uint32_t dst = base + p.off_x + dx * p.x_bytes;
        const int r0 = t0 - p.p2 - p.p1;          // negative / past-the-end rows are zero filled by the TMA unit
        tma_load_3d(dst, &tmap_x, full, 0, r0, b);
        tma_load_3d(dst + p.rxh * 128, &tmap_x, full, 0, r0 + p.rxh, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc(128, C, BF16);
      mbar_wait(RB_BAR(w_full, 0), 0);
      const uint32_t w1_lo = umma_desc_lo(base + p.off_w1), w2_lo = umma_desc_lo(base + p.off_w2);
      const uint32_t w_step = static_cast<uint32_t>(p.w_tap_bytes) >> 4;
      for (int n = 0; n < n_my + S - 1; ++n) {
        if (n < n_my) {
          // c1 of tile n: tap j = the X box shifted by j * d rows of 128 bytes (the swizzle phase follows the address bits)
          const int s = n % S, dx = n % NX;
          mbar_wait(RB_BAR(x_full, dx), static_cast<uint32_t>((n / NX) & 1));
          mbar_wait(RB_BAR(acc1_free, s), static_cast<uint32_t>(((n / S) & 1) ^ 1));
          tc_fence_after();
          RBP_STAMP(1, n);
          const uint32_t x_lo = umma_desc_lo(base + p.off_x + dx * p.x_bytes);
#pragma unroll 1
          for (int m = 0; m < 2; ++m) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(s * 4 * C + m * C);
#pragma unroll 1
            for (int j = 0; j < p.taps; ++j) {
              const uint32_t a_lo = x_lo + static_cast<uint32_t>(m * 128 + j * p.dil) * 8u;
              const uint32_t b_lo = w1_lo + static_cast<uint32_t>(j) * w_step;
#pragma unroll
              for (int kk = 0; kk < KK; ++kk)
                umma_f16(d_tmem, umma_desc_from_lo(a_lo + 2 * kk), umma_desc_from_lo(b_lo + 2 * kk), idesc, (j | kk) != 0 ? 1u : 0u);
            }
          }
          RBP_STAMP(2, n);
          umma_commit(RB_BAR(acc1_full, s));
          umma_commit(RB_BAR(x_empty, dx));      // the box goes back to the loader as soon as these MMAs have read it
        }
        const int i2 = n - (S - 1);
        if (i2 >= 0 && i2 < n_my) {
          // c2 of tile i2 on the hidden tile the epilogue warps wrote
          const int s = i2 % S;
          mbar_wait(RB_BAR(h_full, s), static_cast<uint32_t>((i2 / S) & 1));
          mbar_wait(RB_BAR(acc2_free, s), static_cast<uint32_t>(((i2 / S) & 1) ^ 1));
          tc_fence_after();
          RBP_STAMP(3, i2);
          const uint32_t h_lo = umma_desc_lo(base + p.off_h + s * p.h_bytes);
#pragma unroll 1
          for (int m = 0; m < 2; ++m) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(s * 4 * C + 2 * C + m * C);
#pragma unroll 1
            for (int j = 0; j < p.taps; ++j) {
              const uint32_t a_lo = h_lo + static_cast<uint32_t>(m * 128 + j) * 8u;
              const uint32_t b_lo = w2_lo + static_cast<uint32_t>(j) * w_step;
#pragma unroll
              for (int kk = 0; kk < KK; ++kk)
                umma_f16(d_tmem, umma_desc_from_lo(a_lo + 2 * kk), umma_desc_from_lo(b_lo + 2 * kk), idesc, (j | kk) != 0 ? 1u : 0u);
            }
          }
          RBP_STAMP(4, i2);
          umma_commit(RB_BAR(acc2_full, s));
          if (p.n_st != 0) umma_commit(RB_BAR(h_free, s));      // own staging buffers: H(s) is free once c2 has read it
        }
      }
    }
  } else if (warp < kEpi2Warp0) {
    // ------------------------------------------------------------------ E1 warps: thread == row of one 128-row hidden tile
    //   h = lrelu(c1 + b1), zero outside the sequence (c2's zero padding) -> K-major SW128 tile H(s)
    const int e = warp - kEpiWarp0;
    const int m = e >> 2;                           // which of the two 128-row tiles
    const int q = warp & 3;                         // TMEM lane quarter this warp may touch
    const int hr = m * 128 + q * 32 + lane;         // row of the 256-row hidden tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float slope = p.slope;
    for (int n = 0; n < n_my; ++n) {
      const int s = n % S;
      const int tile = blockIdx.x + n * gridDim.x;
      const int b = tile / p.tiles_per_seq, t0 = (tile - b * p.tiles_per_seq) * p.L;
      mbar_wait(RB_BAR(acc1_full, s), static_cast<uint32_t>((n / S) & 1));
      tc_fence_after();
      __syncwarp();          // tcgen05.ld is warp-collective (.sync.aligned): reconverge after the divergent wait loop
      if (warp == kEpiWarp0 && lane == 0) RBP_STAMP(5, n);
      const uint32_t a1 = tmem_base + lane_addr + static_cast<uint32_t>(s * 4 * C + m * C);
      const int th = t0 - p.p2 + hr;
      const bool valid = th >= 0 && th < p.T;
      const uint32_t h_row = base + p.off_h + s * p.h_bytes + static_cast<uint32_t>(hr) * 128u;
      const uint32_t xr = static_cast<uint32_t>(hr & 7);
      // 32 columns at a time (C = 64: two passes): 32 accumulator + 32 bias registers live, no spills; the bias loads are
      // issued with the tensor-memory load, so one shared-memory latency per pass instead of one per 8 columns
#pragma unroll
      for (int hf = 0; hf < C / 32; ++hf) {
        float bv[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = *reinterpret_cast<const float4*>(&bias_s[0][32 * hf + 4 * i]);
          bv[4 * i] = t4.x;
          bv[4 * i + 1] = t4.y;
          bv[4 * i + 2] = t4.z;
          bv[4 * i + 3] = t4.w;
        }
        uint32_t u[32];
        tmem_ld_x32(a1 + 32 * hf, u);
        tmem_ld_wait();
        if (hf == C / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(RB_BAR(acc1_free, s));
          if (warp == kEpiWarp0 && lane == 0) RBP_STAMP(6, n);
        }
        if (hf == 0) {
          mbar_wait(RB_BAR(h_free, s), static_cast<uint32_t>(((n / S) & 1) ^ 1));   // every E2 warp's store out of this buffer has read it
          if (warp == kEpiWarp0 && lane == 0) RBP_STAMP(7, n);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c0 = 8 * ch + 2 * i;
            float v0 = __uint_as_float(u[c0]), v1 = __uint_as_float(u[c0 + 1]);
            add2(v0, v1, bv[c0], bv[c0 + 1]);
            lrelu2<false>(v0, v1, slope);
            w[i] = valid ? pack2<BF16>(v0, v1) : 0u;
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h_row + ((static_cast<uint32_t>(4 * hf + ch) ^ xr) << 4)),
                       "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
        }
      }
      if (warp == kEpiWarp0 && lane == 0) RBP_STAMP(15, n);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(RB_BAR(h_full, s));
      if (warp == kEpiWarp0 && lane == 0) RBP_STAMP(8, n);
    }
  } else {
    // ------------------------------------------------------------------ E2 warps: thread == row of one 128-row output tile
    //   lrelu(c2 + b2 + x), x recovered from the lrelu'ed input -> 16-bit staging rows (in H(s), which c2 has consumed) ->
    //   per-warp TMA stores.  A separate warp set from E1, so that E1 of tile n + 1 and E2 of tile n run side by side.
    const int e = warp - kEpi2Warp0;
    const int m = e >> 2;
    const int q = warp & 3;
    const int hr = m * 128 + q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float slope = p.slope, inv_slope = p.inv_slope;
    int pending_slot = -1;   // slot whose staging rows this warp's last TMA store may still be reading
    for (int n = 0; n < n_my; ++n) {
      const int s = n % S;
      const int tile = blockIdx.x + n * gridDim.x;
      const int b = tile / p.tiles_per_seq, t0 = (tile - b * p.tiles_per_seq) * p.L;
      // residual rows (the lrelu'ed input at the output row's own time step) straight from global memory, issued before
      // anything is waited for: the X box holding the same rows was fetched a moment ago, so these are L2 hits whose
      // latency hides under the c2 wait.  (Reading them back from the X box with ld.shared was not reliable against the
      // TMA unit refilling the ring.)  Rows >= L belong to the next tile.
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(9, n);
      uint32_t res[C / 2];
      {
        const int tr = t0 + hr;
        if (hr < p.L && tr < p.T) {
          const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.x) +
                                                            (static_cast<size_t>(b) * p.T + tr) * (2u * C));
          if (p.x_al32) {
            // 256-bit loads (LDG.E.256): a row is 128 B away from the next lane's, so every load instruction touches 32
            // (C = 64) / 16 (C = 32) distinct lines and the L1 tag stage is the limit — half the instructions, half the lookups
#pragma unroll
            for (int ch = 0; ch < NCH; ch += 2)
              asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(res[4 * ch]), "=r"(res[4 * ch + 1]), "=r"(res[4 * ch + 2]), "=r"(res[4 * ch + 3]),
                             "=r"(res[4 * ch + 4]), "=r"(res[4 * ch + 5]), "=r"(res[4 * ch + 6]), "=r"(res[4 * ch + 7])
                           : "l"(src + ch));
          } else {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
              const uint4 v = __ldg(src + ch);
              res[4 * ch] = v.x;
              res[4 * ch + 1] = v.y;
              res[4 * ch + 2] = v.z;
              res[4 * ch + 3] = v.w;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < C / 2; ++i) res[i] = 0u;
        }
      }
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(17, n);
      // Shared staging (the output rows are staged in H(s), p.n_st == 0): hand back the rows of this warp's previous TMA
      // store — wait until it has READ them, then h_free — BEFORE waiting for c2: with one slot, c2 of this tile cannot
      // start until E1 has rewritten the buffer, which waits for exactly this arrive.
      if (p.n_st == 0 && pending_slot >= 0) {
        if (elect_one_sync()) {
          tma_store_wait_read<0>();
          mbar_arrive(RB_BAR(h_free, pending_slot));
        }
        pending_slot = -1;
        __syncwarp();
      }
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(10, n);
      mbar_wait(RB_BAR(acc2_full, s), static_cast<uint32_t>((n / S) & 1));
      tc_fence_after();
      __syncwarp();
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(11, n);
      const uint32_t a2 = tmem_base + lane_addr + static_cast<uint32_t>(s * 4 * C + 2 * C + m * C);
      // staging tile: dense rows of 2 C bytes, SW128 on the linear address (what the TMA store un-swizzles).  Own buffers
      // (p.n_st = 1 / 2, each warp re-uses only its own rows) where shared memory allows: E1 of tile n + S then waits for
      // c2 of tile n alone (h_free committed by the issuer), not for this warp set's TMA stores.
      const uint32_t st_base = p.n_st == 0 ? base + p.off_h + s * p.h_bytes
                                           : base + p.off_st + (p.n_st == 2 ? (n & 1) : 0) * (256 * 2 * C);
#pragma unroll
      for (int hf = 0; hf < C / 32; ++hf) {
        float bv[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = *reinterpret_cast<const float4*>(&bias_s[1][32 * hf + 4 * i]);
          bv[4 * i] = t4.x;
          bv[4 * i + 1] = t4.y;
          bv[4 * i + 2] = t4.z;
          bv[4 * i + 3] = t4.w;
        }
        uint32_t u[32];
        tmem_ld_x32(a2 + 32 * hf, u);
        tmem_ld_wait();
        if (hf == C / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(RB_BAR(acc2_free, s));
          if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(12, n);
        }
        if (hf == 0 && p.n_st != 0) {      // this warp's store out of the rows it is about to overwrite has read them
          if (elect_one_sync()) {
            if (p.n_st == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          __syncwarp();
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c0 = 8 * ch + 2 * i;
            float x0, x1;
            unpack2<BF16>(res[16 * hf + 4 * ch + i], x0, x1);
            lrelu2<true>(x0, x1, inv_slope);               // x from lrelu(x)
            float v0 = __uint_as_float(u[c0]), v1 = __uint_as_float(u[c0 + 1]);
            add2(v0, v1, bv[c0], bv[c0 + 1]);
            add2(v0, v1, x0, x1);
            lrelu2<false>(v0, v1, slope);
            w[i] = pack2<BF16>(v0, v1);
          }
          const uint32_t lin = static_cast<uint32_t>(hr) * (2u * C) + static_cast<uint32_t>(4 * hf + ch) * 16u;
          const uint32_t phys = lin ^ (((lin >> 7) & 7u) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_base + phys), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                       "r"(w[3]));
        }
      }
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(16, n);
      fence_proxy_async();
      __syncwarp();
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(13, n);
      // every warp stores its own 32 rows (no CTA-wide rendezvous): rows >= L belong to the next tile, so the warp that
      // straddles L uses the shorter box and the ones past it store nothing; rows past the end of the sequence are clipped
      if (elect_one_sync()) {     // the same thread every time (full mask): bulk async-groups are per thread (wait_read / wait_all)
        const int r0 = m * 128 + q * 32;
        const uint32_t src = st_base + static_cast<uint32_t>(r0) * (2u * C);
        if (r0 + 32 <= p.L) tma_store_3d(&tmap_out, src, 0, C == 64 ? t0 + r0 : (t0 + r0) / 2, b);
        else if (r0 < p.L) tma_store_3d(&tmap_tail, src, 0, C == 64 ? t0 + r0 : (t0 + r0) / 2, b);
        tma_store_commit();
      }
      pending_slot = s;
      __syncwarp();
      if (warp == kEpi2Warp0 && lane == 0) RBP_STAMP(14, n);
    }
    __syncwarp();
    if (elect_one_sync()) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace rbp
}  // namespace ctta

using namespace ctta;

extern "C" int ctta_resblock_pair_supported(int32_t c, int32_t taps, int32_t dilation, int32_t t) {
  if (c != 32 && c != 64) return 0;
  if (c == 32 && (t & 1)) return 0;     // the 64-byte output rows are stored pairwise as [t / 2, 64] (128-byte TMA rows)
  if (taps < 3 || taps > 11 || (taps & 1) == 0 || dilation < 1) return 0;
  if (dilation * (taps - 1) > 56) return 0;
  // resident weights of both convs + at least two X boxes + one H buffer must fit shared memory
  const int w = 2 * taps * c * 128;
  const int rx = (256 + dilation * (taps - 1) + 15) / 16 * 16;
  return (w + 2 * rx * 128 + rbp::kHBufRows * 128 + 4096 <= 227 * 1024) ? 1 : 0;
}

extern "C" int ctta_resblock_pair(const void* x, void* out, int32_t dtype, int32_t batch, int32_t t, int32_t c,
                                  const void* w1, const float* b1, const void* w2, const float* b2, int32_t taps,
                                  int32_t dilation, float slope, void* stream_v) {
  using namespace rbp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && out && w1 && w2 && b1 && b2, "resblock_pair: null argument");
  CTTA_REQUIRE(dtype == CTTA_F16 || dtype == CTTA_BF16, "resblock_pair: 16-bit streams only");
  CTTA_REQUIRE(batch > 0 && t > 0 && slope > 0.f, "resblock_pair: bad shape / slope");
  if (!ctta_resblock_pair_supported(c, taps, dilation, t))
    return set_error(CTTA_ERR_UNSUPPORTED, "resblock_pair: c=%d taps=%d dilation=%d not supported", c, taps, dilation);
  for (const void* ptr : {x, static_cast<const void*>(out), w1, w2})
    CTTA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "resblock_pair: operands must be 16-byte aligned");
  Params p{};
  p.C = c;
  p.taps = taps;
  p.dil = dilation;
  p.p2 = (taps - 1) / 2;
  p.p1 = p.p2 * dilation;
  p.L = kHRows - (taps - 1);
  p.T = t;
  p.batch = batch;
  p.tiles_per_seq = (t + p.L - 1) / p.L;
  p.total_tiles = p.tiles_per_seq * batch;
  p.rx = (kHRows + 2 * p.p1 + 15) / 16 * 16;
  p.rxh = p.rx / 2;
  p.w_tap_bytes = c * 128;
  p.b1 = b1;
  p.b2 = b2;
  p.x = x;
  p.x_al32 = (reinterpret_cast<uintptr_t>(x) & 31) == 0 ? 1 : 0;
  p.trace = nullptr;
#if RBP_TRACE
  if (const char* e = getenv("CTTA_RBP_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
#endif
  p.slope = slope;
  p.inv_slope = 1.f / slope;
  p.is_bf16 = dtype == CTTA_BF16;
  // shared-memory plan: W1 | W2 | X ring | H slots, every region 1024-byte aligned
  const int w_bytes = (taps * p.w_tap_bytes + 1023) / 1024 * 1024;
  p.x_bytes = (p.rx * 128 + 1023) / 1024 * 1024;
  p.h_bytes = (kHBufRows * 128 + 1023) / 1024 * 1024;
  const int budget = 227 * 1024 - 4096 - 2 * w_bytes;   // 1 KiB alignment slack + static barriers / biases
  // Shared-memory plan, in order of preference (CTTA_RBP_SLOTS=1 / CTTA_RBP_STAGING=0|1|2 force a variant for A/B runs):
  // two accumulator / H slots (c1 + E1 of tile n + 1 under c2 + E2 of tile n) with own output staging buffers (two, then
  // one), then with the output staged in H(s); X ring 3 -> 1 deep (one box is enough with two slots: the box of tile
  // n + 1 lands while the issuer runs c2 of tile n - 1); last, one slot.
  int slots = 2, nx = 3, n_st = 2;
  if (const char* e = getenv("CTTA_RBP_SLOTS")) slots = atoi(e) == 1 ? 1 : 2;
  int st_max = 2;
  if (const char* e = getenv("CTTA_RBP_STAGING")) st_max = atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e));
  const int st_bytes = 256 * 2 * c;
  int x_max = 3;
  if (const char* e = getenv("CTTA_RBP_X")) x_max = atoi(e) < 1 ? 1 : (atoi(e) > 3 ? 3 : atoi(e));
  bool ok = false;
  for (int pass = 0; pass < 2 && !ok; ++pass) {
    const int sl = pass == 0 ? slots : 1;
    // a two-deep X ring matters more than the staging buffers (measured), so: x >= 2 with staging 2, 1, 0, then x = 1
    for (int x_min = 2; x_min >= (sl == 2 ? 1 : 2) && !ok; --x_min)
      for (int st = st_max; st >= 0 && !ok; --st)
        for (int x = x_max; x >= x_min && !ok; --x)
          if (sl * p.h_bytes + x * p.x_bytes + st * st_bytes <= budget) {
            slots = sl;
            nx = x;
            n_st = st;
            ok = true;
          }
  }
  if (!ok) return set_error(CTTA_ERR_UNSUPPORTED, "resblock_pair: shared memory plan does not fit");
  p.n_slots = slots;
  p.n_x = nx;
  p.n_st = n_st;
  p.off_w1 = 0;
  p.off_w2 = w_bytes;
  p.off_x = 2 * w_bytes;
  p.off_h = p.off_x + nx * p.x_bytes;
  p.off_st = p.off_h + slots * p.h_bytes;
  const int smem_bytes = p.off_st + n_st * st_bytes + 1024;
  if (getenv("CTTA_DEBUG") != nullptr)
    fprintf(stderr, "ctta_resblock_pair: c=%d taps=%d dil=%d slots=%d x=%d staging=%d smem=%d\n", c, taps, dilation, slots, nx, n_st, smem_bytes);

  CUtensorMap tx, tw1, tw2, tout, ttail;
  {
    cuuint64_t dims[3] = {(cuuint64_t)c, (cuuint64_t)t, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * (cuuint64_t)t};
    cuuint32_t box[3] = {64, (cuuint32_t)p.rxh, 1};
    int rc = make_tmap_ex(&tx, dtype, CU_TENSOR_MAP_SWIZZLE_128B, x, 3, dims, strides, box);
    if (rc) return rc;
    // output boxes: 32 rows per epilogue warp, and the shorter box of the warp that straddles L = 256 - (taps - 1)
    const int tail_rows = 32 - (taps - 1);
    if (c == 64) {
      cuuint32_t obox[3] = {64, 32, 1}, tbox[3] = {64, (cuuint32_t)tail_rows, 1};
      rc = make_tmap_ex(&tout, dtype, CU_TENSOR_MAP_SWIZZLE_128B, out, 3, dims, strides, obox);
      if (rc) return rc;
      rc = make_tmap_ex(&ttail, dtype, CU_TENSOR_MAP_SWIZZLE_128B, out, 3, dims, strides, tbox);
    } else {
      // C = 32: rows of 64 B are stored pairwise, the tensor viewed as [t / 2, 64] (same memory; L and every t0 are even)
      cuuint64_t pd[3] = {64, (cuuint64_t)t / 2, (cuuint64_t)batch};
      cuuint64_t ps[2] = {128, (cuuint64_t)t * 64};
      cuuint32_t obox[3] = {64, 16, 1}, tbox[3] = {64, (cuuint32_t)tail_rows / 2, 1};
      rc = make_tmap_ex(&tout, dtype, CU_TENSOR_MAP_SWIZZLE_128B, out, 3, pd, ps, obox);
      if (rc) return rc;
      rc = make_tmap_ex(&ttail, dtype, CU_TENSOR_MAP_SWIZZLE_128B, out, 3, pd, ps, tbox);
    }
    if (rc) return rc;
    // packed weights [n = C, taps * 64] (ops.pack_conv1d): tap j is the {64 x C} box at column 64 j
    cuuint64_t wd[2] = {(cuuint64_t)taps * 64, (cuuint64_t)c};
    cuuint64_t ws[1] = {(cuuint64_t)taps * 64 * 2};
    cuuint32_t wb[2] = {64, (cuuint32_t)c};
    rc = make_tmap_ex(&tw1, dtype, CU_TENSOR_MAP_SWIZZLE_128B, w1, 2, wd, ws, wb);
    if (rc) return rc;
    rc = make_tmap_ex(&tw2, dtype, CU_TENSOR_MAP_SWIZZLE_128B, w2, 2, wd, ws, wb);
    if (rc) return rc;
  }
  typedef void (*Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);
  Fn fn;
  if (slots == 2)
    fn = c == 32 ? (p.is_bf16 ? resblock_pair_kernel<32, 1, 2> : resblock_pair_kernel<32, 0, 2>)
                 : (p.is_bf16 ? resblock_pair_kernel<64, 1, 2> : resblock_pair_kernel<64, 0, 2>);
  else
    fn = c == 32 ? (p.is_bf16 ? resblock_pair_kernel<32, 1, 1> : resblock_pair_kernel<32, 0, 1>)
                 : (p.is_bf16 ? resblock_pair_kernel<64, 1, 1> : resblock_pair_kernel<64, 0, 1>);
  int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(fn), 227 * 1024 - 2048);
  if (rc) return rc;
  int grid = sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  fn<<<grid, kThreads, smem_bytes, stream>>>(tx, tw1, tw2, tout, ttail, p);
  CTTA_LAUNCH_CHECK();
  return 0;
}

