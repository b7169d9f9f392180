// Fused flash-style self-attention on tcgen05 tensor cores for head_dim 64 (the UNet's 51-wide heads padded to 64).
//
// One CTA = one (batch, head) and 256 query rows = two independent 128-row tiles that share every K/V tile.
// 12 warps: warps 0..3 / 4..7 = softmax warp groups of query tile 0 / 1 (thread == query row, so the row max / row sum
// need no shuffles; setmaxnreg gives them 216 registers), warp 8 = TMA loader (Q once, K/V through a 3-stage ring),
// warp 9 = single-thread tcgen05.mma issuer.  The two softmax groups share the SM's MUFU pipes, so group 1 starts half
// an iteration late: one group's exponentials then overlap the other's TMEM loads / packing / stores.  Per key tile of 128:
//     S_t   = Q_t K^T            UMMA 128x128x64, fp32 in TMEM (S_t: 128 columns)
//     P_t   = exp2(c S_t - c m)  softmax warps: tcgen05.ld -> registers -> packed 16-bit P back into TMEM (tcgen05.st)
//     O_t  += P_t V              UMMA 128x64x128 with A = P_t from tensor memory; V is the MN-major B operand exactly
//                                as TMA lands it ([keys, d])
// The issuer runs QK^T of key tile j+1 as soon as the softmax warps are done with S(j), so the tensor core works under
// the packing / storing of P(j) and the softmax of the other query tile.
//
// At head dim 64 the softmax, not the tensor pipe, is the bound: 16 ex2/clk/SM = 1024 MUFU cycles per 128x128 tile against
// 736 tensor cycles, and each softmax warp's own serial chain on top (two warps per scheduler).  Three measures (r2):
//  * OPTIMISTIC softmax: a tile's probabilities are evaluated against the running max of the EARLIER tiles without
//    looking for the tile's own row max; the tile's row sum (needed anyway) is the overflow detector (sum <= 2^15 => every
//    P <= 2^15 fits the 16-bit P).  Only when a row fails it, or on the first tile, the scores are re-read from tensor
//    memory, the true max is taken, O_t / l are rescaled and the tile is re-evaluated.  Exact: softmax is shift
//    invariant, and the same stale max is used for the row sum.  Saves 64 FMNMX3 + a dependent reduction per tile.
//  * 3 of every 8 pairs of exponentials are evaluated on the FMA pipe instead of MUFU (Cody-Waite split + degree-3 minimax
//    polynomial, 7.5e-5 relative error, below the 4.9e-4 rounding of the 16-bit P).
//  * instruction diet: the phase timestamps and the 16-bit type are compile-time (a run-time dtype flag issued both
//    converts, each timestamp ~10 instructions of predicate arithmetic per tile), barrier addresses are pinned in
//    registers, hot waits are a 2-instruction spin.
// Measured on the 64 x 5 x 4096^2 site: 2.09 -> 1.64 ms (profiles/r2_attn_*.txt).  Rejected by measurement: a four-stream
// split-KV layout (16 softmax warps, P through shared memory): S 2x128 + O 4x64 fills tensor memory, P then has to go
// through shared memory and the kernel becomes shared-memory / issue bound (2.07 ms).
// Keys >= kv_len[b] get -inf, i.e. the reference's additive -10000 mask bias.
//
// Reference call site: F.scaled_dot_product_attention, diffusers/models/attention_processor.py:1127-1129.
#include <cuda.h>
#include <cstdlib>
#include "ctta_internal.h"
#include "ctta_ptx.cuh"

namespace ctta {

int make_tmap_ex(CUtensorMap* m, int dtype, CUtensorMapSwizzle swz, const void* base, int rank, const cuuint64_t* dims,
                 const cuuint64_t* strides_bytes, const cuuint32_t* box);

namespace atc {

constexpr int kTile = 128;                   // query rows per tile == keys per tile
constexpr int kD = 64;
constexpr int kTileBytes = kTile * kD * 2;   // 16 KiB: one [128 x 64] 16-bit operand tile (rows of 128 B, SW128)
constexpr int kStages = 3;
constexpr int kThreads = 384;
constexpr int kLoaderWarp = 8, kMmaWarp = 9;
constexpr int kSmemQ = 0;
constexpr int kSmemKV = kSmemQ + 2 * kTileBytes;             // K then V per stage
constexpr int kSmemP = kSmemKV + kStages * 2 * kTileBytes;   // per query tile: two [128 x 64] K-major halves
constexpr int kSmemTotal = kSmemP;                           // 128 KiB (P lives in tensor memory)
constexpr int kTmemCols = 512;
constexpr int kColS = 0, kColO = 256, kColP = 384;   // S_t at 128 t, O_t at 256 + 64 t, P_t (16-bit pairs) at 384 + 64 t
constexpr int kDefaultPoly = 3;               // 3 of 8 pairs (37.5 %) of the exponentials on the FMA pipe (measured best of 0 / 2 / 3 / 4)
constexpr float kSumLimit = 32768.f;         // optimistic softmax: a tile whose probabilities sum to <= 2^15 cannot overflow the 16-bit P

struct Params {
  int lq, lk, heads;
  const int* kv_len;
  float scale_log2;
  unsigned short* o;
  long long o_bs, o_ls;
  int is_bf16;
  long long* dbg;   // optional phase timestamps of CTA 0 / softmax warp 0 (tools/attn_phases.py), else NULL
};
// compiled in only for the DBG instantiation: each stamp costs ~10 instructions of predicate arithmetic per key tile
#define ATC_STAMP(slot)                                                            \
  do {                                                                             \
    if (DBG && blockIdx.x + blockIdx.y + blockIdx.z == 0 && warp == 0 && lane == 0 && j >= 8 && j < 16) \
      p.dbg[(j - 8) * 8 + (slot)] = clock64();                                     \
  } while (0)

// Hot-loop wait: a plain try_wait spin (2 instructions per round); the wall-clock bound that turns a protocol bug into a
// trap instead of a hang is only consulted every 4096 failed rounds.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {   // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > 20000000000ULL) __trap();
    }
  }
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major B operand (V as [keys, d] rows of 128 B, SW128): 8-key groups are 1024 B apart (stride byte offset);
// the leading byte offset (between 64-wide d blocks) is unused for N = 64.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// D[tmem] (+)= A[tmem] * B[smem]: the 16-bit A operand (P) is read from tensor memory, row = lane, two K elements per
// 32-bit column (8 columns per 16-wide K slice)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int BF16>
__device__ __forceinline__ uint32_t pack_p(float a, float b) {   // compile-time type: a runtime flag issues both converts
  uint32_t r;
  if (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// ---- packed fp32 pairs (FFMA2 / FADD2)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for two values on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x) (magic-number add), f in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (7.5e-5 max relative error), 2^n by adding n to the exponent field.
// x is clamped to >= -125 so that the exponent add never leaves the normal range (-inf / masked keys -> 2^-125 ~ 0).
__device__ __forceinline__ void exp2_poly2(float& x0, float& x1) {
  const float kMagic = 12582912.f;   // 1.5 * 2^23
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  const uint64_t x = pk2(x0, x1);
  const uint64_t t = add2(x, pk2(kMagic, kMagic));
  const uint64_t n = add2(t, pk2(-kMagic, -kMagic));
  const uint64_t f = fma2(n, pk2(-1.f, -1.f), x);
  uint64_t r = fma2(f, pk2(0.05517132207751274f, 0.05517132207751274f), pk2(0.24261054396629333f, 0.24261054396629333f));
  r = fma2(r, f, pk2(0.6932609677314758f, 0.6932609677314758f));
  r = fma2(r, f, pk2(0.9999281167984009f, 0.9999281167984009f));
  float t0, t1, r0, r1;
  upk2(t, t0, t1);
  upk2(r, r0, r1);
  x0 = __int_as_float(__float_as_int(r0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(r1) + (__float_as_int(t1) << 23));
}

// POLY: of every 8 pairs of exponentials, POLY pairs are evaluated on the FMA pipe instead of MUFU (0 .. 4 -> 0 .. 50 %)
template <int POLY, int BF16, int DBG>

__global__ void __launch_bounds__(kThreads, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q_full;
  __shared__ __align__(8) uint64_t bar_kv_full[kStages], bar_kv_empty[kStages];
  __shared__ __align__(8) uint64_t bar_s_full[2], bar_s_free[2], bar_p_full[2], bar_pv_done[2];
  __shared__ __align__(8) uint64_t bar_stagger;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * 2 * kTile;

  int kvl = p.kv_len ? p.kv_len[b] : p.lk;
  kvl = max(1, min(kvl, p.lk));
  const int n_tiles = (kvl + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_q_full), 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bar_kv_full[s]), 1);
      mbar_init(smem_u32(&bar_kv_empty[s]), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_s_full[t]), 1);
      mbar_init(smem_u32(&bar_s_free[t]), 4);   // one arrive per softmax warp
      mbar_init(smem_u32(&bar_p_full[t]), 4);
      mbar_init(smem_u32(&bar_pv_done[t]), 1);
    }
    mbar_init(smem_u32(&bar_stagger), 4);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == kLoaderWarp) {
    // ------------------------------------------------------------------ TMA loader
    if (elect_one_sync()) {
      const uint32_t qf = smem_u32(&bar_q_full);
      mbar_arrive_expect_tx(qf, 2 * kTileBytes);
      tma_load_4d(base + kSmemQ, &tmap_q, qf, 0, head, q0, b);
      tma_load_4d(base + kSmemQ + kTileBytes, &tmap_q, qf, 0, head, q0 + kTile, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(smem_u32(&bar_kv_empty[stage]), phase ^ 1u);
        const uint32_t full = smem_u32(&bar_kv_full[stage]);
        mbar_arrive_expect_tx(full, 2 * kTileBytes);
        const uint32_t dst = base + kSmemKV + stage * 2 * kTileBytes;
        tma_load_4d(dst, &tmap_k, full, 0, head, j * kTile, b);
        tma_load_4d(dst + kTileBytes, &tmap_v, full, 0, head, j * kTile, b);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc_qk = umma_idesc(kTile, kTile, BF16);                 // 128 x 128, A and B K-major
      const uint32_t idesc_pv = umma_idesc(kTile, kD, BF16) | (1u << 16);       // 128 x 64, B (= V) MN-major
      mbar_wait(smem_u32(&bar_q_full), 0);
      mbar_wait(smem_u32(&bar_kv_full[0]), 0);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t a = base + kSmemQ + t * kTileBytes, bk = base + kSmemKV;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + kColS + t * kTile, umma_desc_sw128(a + kk * 32), umma_desc_sw128(bk + kk * 32), idesc_qk,
                   kk != 0 ? 1u : 0u);
        umma_commit(smem_u32(&bar_s_full[t]));
      }
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t par = static_cast<uint32_t>(j & 1);
        const int stage = j % kStages;
        if (j + 1 < n_tiles) {
          const int nstage = (j + 1) % kStages;
          mbar_wait(smem_u32(&bar_kv_full[nstage]), static_cast<uint32_t>(((j + 1) / kStages) & 1));
          const uint32_t bk = base + kSmemKV + nstage * 2 * kTileBytes;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(smem_u32(&bar_s_free[t]), par);   // softmax has S_t(j) in registers
            tc_fence_after();
            const uint32_t a = base + kSmemQ + t * kTileBytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_base + kColS + t * kTile, umma_desc_sw128(a + kk * 32), umma_desc_sw128(bk + kk * 32),
                       idesc_qk, kk != 0 ? 1u : 0u);
            umma_commit(smem_u32(&bar_s_full[t]));
          }
        }
        const uint32_t bv = base + kSmemKV + stage * 2 * kTileBytes + kTileBytes;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(smem_u32(&bar_p_full[t]), par);     // P_t(j) is in shared memory, O_t rescaled if needed
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < 8; ++s)
            umma_f16_ts(tmem_base + kColO + t * kD, tmem_base + kColP + t * 64 + s * 8, umma_desc_sw128_mn(bv + s * 2048),
                        idesc_pv, (j | s) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&bar_pv_done[t]));
        }
        umma_commit(smem_u32(&bar_kv_empty[stage]));
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------------------------------------------ softmax warps (thread == query row)
    const int t = warp >> 2;                        // query tile of this warp group
    const int q = warp & 3;                         // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;                  // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + kColS + t * kTile;
    const uint32_t o_addr = tmem_base + lane_addr + kColO + t * kD;
    const uint32_t p_addr = tmem_base + lane_addr + kColP + t * 64;
    const float c = p.scale_log2;
    // barrier addresses pinned in registers (ptxas otherwise re-derives the shared-window address at each use)
    const uint32_t a_s_full = pin_u32(smem_u32(&bar_s_full[t])), a_s_free = pin_u32(smem_u32(&bar_s_free[t]));
    const uint32_t a_p_full = pin_u32(smem_u32(&bar_p_full[t])), a_pv_done = pin_u32(smem_u32(&bar_pv_done[t]));
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t par = static_cast<uint32_t>(j & 1);
      ATC_STAMP(0);
      mbar_wait_lean(a_s_full, par);
      tc_fence_after();
      ATC_STAMP(1);
      uint32_t u[128];
      float* s = reinterpret_cast<float*>(u);
      const int valid = kvl - j * kTile;            // keys of this tile below kv_len
      bool waited_pv = false;
      float tile_sum = 0.f;
      // OPTIMISTIC softmax: the probabilities of a tile are first evaluated against the running max m_used of the earlier
      // tiles WITHOUT looking for the tile's own row max (64 FMNMX3 + a dependent reduction per tile).  The row sum of
      // the tile, which is needed anyway, is the overflow detector: sum <= 2^15 implies every P <= 2^15, exactly
      // representable in the 16-bit P (and a NaN / inf sum fails the test).  Only when some row of the warp fails it — a
      // key beat the stale max by more than ~2^15, or this is the first tile — the scores are read from tensor memory
      // again (S_t is released to the next Q K^T only after the check), the true row max is taken, O_t and the row sum are
      // rescaled and the tile is re-evaluated.  Any m works for softmax as long as nothing overflows, so this is exact.
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld_x32(s_addr + ch * 32, u + ch * 32);
        tmem_ld_wait();
        if (valid < kTile) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) s[i] = -INFINITY;
        }
        if (attempt == 1 || j == 0) {
          // true row maximum with three-input max (FMNMX3), eight independent chains
          float mx[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) mx[k] = fmaxf(s[2 * k], s[2 * k + 1]);
#pragma unroll
          for (int i = 16; i < 128; i += 16) {
#pragma unroll
            for (int k = 0; k < 8; ++k) mx[k] = max3(mx[k], s[i + 2 * k], s[i + 2 * k + 1]);
          }
          const float m_new = fmaxf(max3(max3(mx[0], mx[1], mx[2]), max3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])), m_used);
          if (j > 0) {
            // rescale O_t (and the row sum) to the new max; needs PV(j-1) to have landed in TMEM
            mbar_wait_lean(a_pv_done, par ^ 1u);
            waited_pv = true;
            tc_fence_after();
            const float alpha = fast_exp2((m_used - m_new) * c);
            l_sum *= alpha;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[32];
              tmem_ld_x32(o_addr + hh * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x32(o_addr + hh * 32, o);
            }
            tmem_st_wait();
          }
          m_used = m_new;
        }
        if (attempt == 0 && j == 0 && t == 1) mbar_wait(smem_u32(&bar_stagger), 0);   // start half an iteration behind group 0
        ATC_STAMP(3);
        const float nmc = -m_used * c;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        // exponent arguments and row sums on packed fp32 pairs (FFMA2 / FADD2); of every 8 pairs, POLY go to the FMA pipe
#pragma unroll
        for (int i = 0; i < 128; i += 16) {
#pragma unroll
          for (int pr = 0; pr < 8; ++pr) fma2_ss(s[i + 2 * pr], s[i + 2 * pr + 1], c, nmc);
#pragma unroll
          for (int pr = 0; pr < 8; ++pr) {
            if (pr >= 8 - POLY) {
              exp2_poly2(s[i + 2 * pr], s[i + 2 * pr + 1]);
            } else {
              s[i + 2 * pr] = fast_exp2(s[i + 2 * pr]);
              s[i + 2 * pr + 1] = fast_exp2(s[i + 2 * pr + 1]);
            }
          }
#pragma unroll
          for (int pr = 0; pr < 8; pr += 2) {
            add2_acc(l0, l1, s[i + 2 * pr], s[i + 2 * pr + 1]);
            add2_acc(l2, l3, s[i + 2 * pr + 2], s[i + 2 * pr + 3]);
          }
        }
        tile_sum = (l0 + l1) + (l2 + l3);
        const bool bad = !(tile_sum <= kSumLimit);
        if (attempt == 1 || j == 0 || !__any_sync(0xffffffffu, bad)) break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_s_free);         // S_t may be overwritten by the next Q K^T
      l_sum += tile_sum;
      ATC_STAMP(4);
      if (j == 0 && t == 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_stagger));
      }
      if (j > 0 && !waited_pv) mbar_wait_lean(a_pv_done, par ^ 1u);   // PV(j-1) no longer reads P_t
      ATC_STAMP(5);
      {
        // P_t -> tensor memory as packed 16-bit pairs (the A operand of the P.V UMMA): row = lane, 64 columns
        uint32_t w[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = pack_p<BF16>(s[64 * hh + 2 * i], s[64 * hh + 2 * i + 1]);
          tmem_st_x32(p_addr + hh * 32, w);
        }
        tmem_st_wait();
      }
      tc_fence_before();     // orders the P store and the O rescale (tcgen05.st) before the arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(a_p_full);
      ATC_STAMP(6);
    }
    // ---- epilogue: O_t / l -> 16-bit, this thread's row
    mbar_wait(smem_u32(&bar_pv_done[t]), static_cast<uint32_t>((n_tiles - 1) & 1));
    tc_fence_after();
    uint32_t o[64];
    tmem_ld_x32(o_addr, o);
    tmem_ld_x32(o_addr + 32, o + 32);
    tmem_ld_wait();
    const float inv = 1.f / l_sum;
    const int grow = q0 + t * kTile + row;
    if (grow < p.lq) {
      unsigned short* dst = p.o + static_cast<long long>(b) * p.o_bs + static_cast<long long>(grow) * p.o_ls + head * kD;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 w;
        w.x = pack_p<BF16>(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
        w.y = pack_p<BF16>(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
        w.z = pack_p<BF16>(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
        w.w = pack_p<BF16>(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + 8 * g) = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace atc
long long* g_attn_dbg = nullptr;
namespace atc {

static int make_qkv_tmap(CUtensorMap* m, int dtype, const void* ptr, int heads, int len, int batch, long long ls,
                         long long bs) {
  cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)heads, (cuuint64_t)len, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)kD * 2, (cuuint64_t)ls * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[4] = {(cuuint32_t)kD, 1, (cuuint32_t)kTile, 1};
  return make_tmap_ex(m, dtype, CU_TENSOR_MAP_SWIZZLE_128B, ptr, 4, dims, strides, box);
}

}  // namespace atc

// Launches the tcgen05 kernel; returns 1 if the problem is not eligible (caller falls back to the mma.sync kernel
// for short key sequences), 0 on success, < 0 on error.
int attention_tc_launch(const void* q, const void* k, const void* v, void* o, int dtype, int batch, int heads, int lq,
                        int lk, long long q_bs, long long q_ls, long long k_bs, long long k_ls, long long v_bs,
                        long long v_ls, long long o_bs, long long o_ls, const int* kv_len, float scale,
                        cudaStream_t stream) {
  using namespace atc;
  if (lq < kTile || lk < kTile || getenv("CTTA_ATTN_LEGACY") != nullptr) return 1;
  if (o_ls % 8 != 0 || o_bs % 8 != 0 || (reinterpret_cast<uintptr_t>(o) & 15) != 0) return 1;
  CUtensorMap tq, tk, tv;
  int rc = make_qkv_tmap(&tq, dtype, q, heads, lq, batch, q_ls, q_bs);
  if (rc) return rc;
  rc = make_qkv_tmap(&tk, dtype, k, heads, lk, batch, k_ls, k_bs);
  if (rc) return rc;
  rc = make_qkv_tmap(&tv, dtype, v, heads, lk, batch, v_ls, v_bs);
  if (rc) return rc;
  Params p{};
  p.lq = lq;
  p.lk = lk;
  p.heads = heads;
  p.kv_len = kv_len;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<unsigned short*>(o);
  p.o_bs = o_bs;
  p.o_ls = o_ls;
  p.is_bf16 = dtype == CTTA_BF16;
  p.dbg = nullptr;
  if (getenv("CTTA_ATTN_DEBUG") != nullptr) {
    static long long* dbg_buf = nullptr;   // debug only: managed memory the tool reads back through ctta_attention_debug
    if (dbg_buf == nullptr) cudaMallocManaged(&dbg_buf, 64 * sizeof(long long));
    p.dbg = dbg_buf;
    g_attn_dbg = dbg_buf;
  }
  // share of the exponentials evaluated on the FMA pipe, in eighths: CTTA_ATTN_POLY = 0 .. 4 -> 0 .. 50 % (A/B switch)
  typedef void (*Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);
  static const Fn fns[2][5] = {{flash_attn_tc_kernel<0, 0, 0>, flash_attn_tc_kernel<1, 0, 0>, flash_attn_tc_kernel<2, 0, 0>,
                                flash_attn_tc_kernel<3, 0, 0>, flash_attn_tc_kernel<4, 0, 0>},
                               {flash_attn_tc_kernel<0, 1, 0>, flash_attn_tc_kernel<1, 1, 0>, flash_attn_tc_kernel<2, 1, 0>,
                                flash_attn_tc_kernel<3, 1, 0>, flash_attn_tc_kernel<4, 1, 0>}};
  int poly = kDefaultPoly;
  if (const char* e = getenv("CTTA_ATTN_POLY")) poly = atoi(e);
  if (poly < 0 || poly > 4) poly = kDefaultPoly;
  Fn fn = fns[p.is_bf16 ? 1 : 0][poly];
  if (p.dbg != nullptr) fn = p.is_bf16 ? flash_attn_tc_kernel<kDefaultPoly, 1, 1> : flash_attn_tc_kernel<kDefaultPoly, 0, 1>;
  rc = ensure_dynamic_smem(reinterpret_cast<const void*>(fn), kSmemTotal + 1024);
  if (rc) return rc;
  dim3 grid((lq + 2 * kTile - 1) / (2 * kTile), heads, batch);
  fn<<<grid, kThreads, kSmemTotal + 1024, stream>>>(tq, tk, tv, p);
  CTTA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctta

// debug helper for tools/attn_phases.py: copies the 64 phase timestamps of the last CTTA_ATTN_DEBUG launch
extern "C" int ctta_attention_debug(long long* host_out) {
  if (ctta::g_attn_dbg == nullptr) return CTTA_ERR_INVALID;
  cudaDeviceSynchronize();
  for (int i = 0; i < 64; ++i) host_out[i] = ctta::g_attn_dbg[i];
  return 0;
}
