// Implicit-GEMM convolution / linear kernel for sm_100a.
//
//   * one persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer + TMEM owner,
//     warps 2..5 = epilogue (TMEM -> registers -> global, one accumulator row per thread)
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
#include "ctta_internal.h"
#include "ctta_ptx.cuh"

namespace ctta {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // 64 x 16-bit = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr int kSmemMaxDynamic = 232448 - 1024;     // 227 KiB minus the static barriers
constexpr int kSmemBudget = kSmemMaxDynamic - 1024;  // minus alignment slack

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
};

struct TileCoord {
  int n0;          // first output channel of the tile
  int c1, c2, c3;  // A box origin (mode dependent)
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile) {
  TileCoord tc;
  const int m_tile = tile / p.n_tiles_n;
  tc.n0 = (tile - m_tile * p.n_tiles_n) * p.block_n;
  if (p.a_mode == CTTA_A_ROWS) {
    tc.c1 = m_tile * kBlockM;
    tc.c2 = 0;
    tc.c3 = 0;
  } else if (p.a_mode == CTTA_A_CONV1D) {
    const int b = m_tile / p.tiles_per_img;
    tc.c1 = (m_tile - b * p.tiles_per_img) * kBlockM;
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
        for (int i = 0; i < 8; ++i) v[i] += f[i];
      }
    }
    const long long ooff = orow * p.out_ld + col;
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
    if (p.out_scale != 1.f) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= p.out_scale;
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
      x += p.res_dtype == CTTA_F32 ? reinterpret_cast<const float*>(p.residual)[off]
                                   : ld16(p.residual, p.res_dtype == CTTA_BF16, off);
    }
    const long long ooff = orow * p.out_ld + c;
    if (p.accumulate) {
      x += p.out_dtype == CTTA_F32 ? reinterpret_cast<const float*>(p.out)[ooff]
                                   : ld16(p.out, p.out_dtype == CTTA_BF16, ooff);
    }
    x *= p.out_scale;
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

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tmem_full[s]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[s]), 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int total_tiles = p.n_tiles_m * p.n_tiles_n;
  const int k_iters = p.ntaps * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.stage_bytes);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        int kb = 0;
        for (int j = 0; j < p.ntaps; ++j) {
          const int d0 = p.tap_d0[j], d1 = p.tap_d1[j];
          for (int kc = 0; kc < p.k_chunks; ++kc, ++kb) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            mbar_arrive_expect_tx(full, tx_bytes);
            const uint32_t a_dst = tiles_base + static_cast<uint32_t>(stage * p.stage_bytes);
            const uint32_t b_dst = a_dst + kATileBytes;
            if (p.a_mode == CTTA_A_ROWS) {
              tma_load_2d(a_dst, &tmap_a, full, kc * kBlockK, tc.c1);
            } else if (p.a_mode == CTTA_A_CONV1D) {
              tma_load_3d(a_dst, &tmap_a, full, kc * kBlockK, tc.c1 + d0, tc.c2);
            } else {
              tma_load_4d(a_dst, &tmap_a, full, kc * kBlockK, tc.c1 + d0, tc.c2 + d1, tc.c3);
            }
            tma_load_2d(b_dst, &tmap_b, full, kb * kBlockK, tc.n0);
            if (++stage == p.n_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(kBlockM, p.block_n, p.is_bf16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bar_tmem_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
        for (int kb = 0; kb < k_iters; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = tiles_base + static_cast<uint32_t>(stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            umma_f16(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), idesc,
                     (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));  // frees the smem stage once these MMAs retire
          if (++stage == p.n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(smem_u32(&bar_tmem_full[acc]));  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int lr = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
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

      mbar_wait(smem_u32(&bar_tmem_full[acc]), acc_phase);
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
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[acc]));
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

  tc_fence_before();
  __syncthreads();
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

static int make_tmap(CUtensorMap* m, int is_bf16, const void* base, int rank, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(CTTA_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                  const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(CTTA_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                     static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                     rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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
  if (d->accumulate) CTTA_REQUIRE(d->out != nullptr, "ctta_gemm: accumulate needs out");
  if (d->rowadd) CTTA_REQUIRE(d->rowadd_rows >= 1, "ctta_gemm: rowadd_rows must be >= 1");

  GemmKParams p{};
  p.a_mode = d->a_mode;
  p.is_bf16 = d->ab_dtype == CTTA_BF16;
  p.N = d->n;
  const int n_tiles_n = (d->n + 255) / 256;
  int block_n = ((d->n + n_tiles_n - 1) / n_tiles_n + 15) / 16 * 16;
  if (block_n < 16) block_n = 16;
  p.block_n = block_n;
  p.n_tiles_n = (d->n + block_n - 1) / block_n;
  int acc_stride = 32;
  while (acc_stride < block_n) acc_stride *= 2;
  p.acc_stride = acc_stride;
  p.tmem_cols = 2 * acc_stride;
  p.k_chunks = (d->c + kBlockK - 1) / kBlockK;
  p.ntaps = d->ntaps;
  p.stage_bytes = kATileBytes + block_n * kBlockK * 2;
  int n_stages = kSmemBudget / p.stage_bytes;
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  p.n_stages = n_stages;
  p.H = d->h;
  p.W = d->w;
  p.n_img = d->n_img;
  p.rows_per_img = d->rows_per_img;
  for (int j = 0; j < d->ntaps; ++j) {
    p.tap_d0[j] = d->tap_d0[j];
    p.tap_d1[j] = d->tap_d1[j];
  }

  const int c_pad = p.k_chunks * kBlockK;
  CUtensorMap tmap_a, tmap_b;
  const int esz = 2;
  if (d->a_mode == CTTA_A_ROWS) {
    p.n_tiles_m = (d->rows_per_img + kBlockM - 1) / kBlockM;
    cuuint64_t dims[2] = {(cuuint64_t)d->c, (cuuint64_t)d->rows_per_img};
    cuuint64_t strides[1] = {(cuuint64_t)d->a_ld * esz};
    cuuint32_t box[2] = {kBlockK, kBlockM};
    int rc = make_tmap(&tmap_a, p.is_bf16, d->a, 2, dims, strides, box);
    if (rc) return rc;
  } else if (d->a_mode == CTTA_A_CONV1D) {
    p.tiles_per_img = (d->rows_per_img + kBlockM - 1) / kBlockM;
    p.n_tiles_m = p.tiles_per_img * d->n_img;
    cuuint64_t dims[3] = {(cuuint64_t)d->c, (cuuint64_t)d->w, (cuuint64_t)d->n_img};
    cuuint64_t strides[2] = {(cuuint64_t)d->a_ld * esz, (cuuint64_t)d->a_ld * esz * (cuuint64_t)d->w};
    cuuint32_t box[3] = {kBlockK, kBlockM, 1};
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
  {
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
  p.out_off = d->out_off;

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

  const int smem_bytes = p.n_stages * p.stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CTTA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMaxDynamic));
    attr_set = true;
  }
  const int total_tiles = p.n_tiles_m * p.n_tiles_n;
  int grid = sm_count();
  if (grid > total_tiles) grid = total_tiles;
  gemm_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(tmap_a, tmap_b, p);
  CTTA_LAUNCH_CHECK();
  return 0;
}
