// GroupNorm (moments + apply/SiLU/upsample/concat -> 16-bit operand) and LayerNorm for channels-last tensors.
// Bandwidth-bound: 16-byte vector loads/stores, warp-shuffle + shared-memory reductions, grid sized to fill 148 SMs.
// Reference call sites: see ctta_groupnorm_stats / ctta_groupnorm_apply / ctta_layernorm in include/ctta.h.
#include "ctta_internal.h"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ctta {

__device__ __forceinline__ float4 load4(const void* base, int dtype, long long elem_off) {
  if (dtype == CTTA_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
  const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(base) + elem_off);
  float4 r;
  if (dtype == CTTA_BF16) {
    r.x = __uint_as_float(u.x << 16);
    r.y = __uint_as_float(u.x & 0xFFFF0000u);
    r.z = __uint_as_float(u.y << 16);
    r.w = __uint_as_float(u.y & 0xFFFF0000u);
  } else {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    r = make_float4(a.x, a.y, b.x, b.y);
  }
  return r;
}
__device__ __forceinline__ unsigned short cvt16n(float v, int dtype) {
  if (dtype == CTTA_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __half_as_ushort(__float2half_rn(v));
}
__device__ __forceinline__ void store4(void* base, int dtype, long long elem_off, float4 v) {
  if (dtype == CTTA_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off) = v;
  } else {
    uint2 u;
    u.x = static_cast<uint32_t>(cvt16n(v.x, dtype)) | (static_cast<uint32_t>(cvt16n(v.y, dtype)) << 16);
    u.y = static_cast<uint32_t>(cvt16n(v.z, dtype)) | (static_cast<uint32_t>(cvt16n(v.w, dtype)) << 16);
    *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base) + elem_off) = u;
  }
}

// ---------------------------------------------------------------------------------------------- GN moments
// grid (slabs, n_img); block (32, 8): x = 4-channel vector column (coalesced), y = row within the slab.
// Each thread owns fixed vector columns, so its partial sums belong to one group per column.
constexpr int kGnMaxGroups = 64;
constexpr int kGnMaxChannels = 2048;   // widest normalised tensor: torch.cat([x, skip]) of the UNet's first up block

__global__ void __launch_bounds__(256) gn_moments_kernel(const void* __restrict__ x, int x_dtype, int c, int ld,
                                                         const void* __restrict__ x2, int c2, int ld2, int hw,
                                                         int rows_per_slab, int groups, float* __restrict__ stats) {
  __shared__ float s_sum[kGnMaxGroups], s_sq[kGnMaxGroups];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < groups) {
    s_sum[tid] = 0.f;
    s_sq[tid] = 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y;
  const int ctot = c + c2;
  const int cpg = ctot / groups;
  const int nvec = ctot >> 2;
  const int r0 = blockIdx.x * rows_per_slab;
  const int r1 = min(hw, r0 + rows_per_slab);
  for (int v = threadIdx.x; v < nvec; v += 32) {
    const int ch = v << 2;
    const bool second = ch >= c;
    const void* src = second ? x2 : x;
    const int sld = second ? ld2 : ld;
    const int sch = second ? ch - c : ch;
    float s = 0.f, q = 0.f;
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 f = load4(src, x_dtype, (static_cast<long long>(n) * hw + r) * sld + sch);
      s += (f.x + f.y) + (f.z + f.w);
      q += (f.x * f.x + f.y * f.y) + (f.z * f.z + f.w * f.w);
    }
    const int g = ch / cpg;
    atomicAdd(&s_sum[g], s);
    atomicAdd(&s_sq[g], q);
  }
  __syncthreads();
  if (tid < groups) {
    atomicAdd(&stats[(n * groups + tid) * 2 + 0], s_sum[tid]);
    atomicAdd(&stats[(n * groups + tid) * 2 + 1], s_sq[tid]);
  }
}

// ---------------------------------------------------------------------------------------------- GN apply
// One item = (pixel, 8-channel vector): 32 B (fp32) or 16 B (16-bit) in, 16 B out.  Each thread keeps kGnItems items in
// flight (all loads issued before any use; 2 items for fp32 input, 4 for 16-bit input = 64 B per thread) so a
// resident SM has >= 100 KB of reads outstanding; 32-bit index math.
// grid (blocks, n_img).

struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 load8(const void* base, int dtype, long long elem_off) {
  Vec8 r;
  if (dtype == CTTA_F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(base) + elem_off));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == CTTA_BF16) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      } else {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        r.v[2 * i] = t.x;
        r.v[2 * i + 1] = t.y;
      }
    }
  }
  return r;
}
__device__ __forceinline__ Vec8 unpack8(const uint4& u, int dtype) {
  Vec8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dtype == CTTA_BF16) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      r.v[2 * i] = t.x;
      r.v[2 * i + 1] = t.y;
    }
  }
  return r;
}
__device__ __forceinline__ uint4 pack8(const Vec8& f, int dtype) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dtype == CTTA_BF16) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f.v[2 * i], f.v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(f.v[2 * i + 1]), "f"(f.v[2 * i]));
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void store8(void* base, int dtype, long long elem_off, const Vec8& f) {
  if (dtype == CTTA_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    o[0] = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
    o[1] = make_float4(f.v[4], f.v[5], f.v[6], f.v[7]);
  } else {
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(base) + elem_off) = pack8(f, dtype);
  }
}

template <int kGnItems, bool kIn16>
__global__ void __launch_bounds__(256, kIn16 ? 3 : 5) gn_apply_kernel(const void* __restrict__ x, int x_dtype, int c, int ld,
                                                       const void* __restrict__ x2, int c2, int ld2, int h, int w,
                                                       int groups, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int act, int up, void* __restrict__ y, int y_dtype,
                                                       int y_ld, void* __restrict__ raw, int raw_ld, int nshift) {
  // per-channel affine of this image: y = x * s_a[c] + s_b[c] with s_a = rstd[g] gamma[c], s_b = beta[c] - mean[g] s_a
  // (built once per block: one FMA per element in the loop, no per-item group lookup / gamma / beta loads)
  __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
  __shared__ __align__(16) float s_a[kGnMaxChannels], s_b[kGnMaxChannels];
  const int n = blockIdx.y;
  const int ctot = c + c2;
  const int hw = h * w;
  const int cpg = stats ? ctot / groups : ctot;
  if (stats && threadIdx.x < groups) {
    const float cnt = static_cast<float>(hw) * cpg;
    const float s = stats[(n * groups + threadIdx.x) * 2 + 0];
    const float q = stats[(n * groups + threadIdx.x) * 2 + 1];
    const float mean = s / cnt;
    const float var = fmaxf(q / cnt - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (stats) {
    for (int ch = threadIdx.x; ch < ctot; ch += blockDim.x) {
      const int g = ch / cpg;
      const float a = s_rstd[g] * gamma[ch];
      s_a[ch] = a;
      s_b[ch] = beta[ch] - s_mean[g] * a;
    }
    __syncthreads();
  }
  const unsigned nvec = static_cast<unsigned>(ctot) >> 3;
  const unsigned total = static_cast<unsigned>(hw) * nvec;  // < 2^31: checked on the host
  const unsigned step = gridDim.x * blockDim.x;
  const long long img_row0 = static_cast<long long>(n) * hw;
  for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < total; base += step * kGnItems) {
    // 16-bit inputs stay packed (one uint4 per item) until they are processed, so 8 items = 128 B per thread can be
    // in flight at 4 registers each
    Vec8 f[kIn16 ? 1 : kGnItems];
    uint4 r16[kIn16 ? kGnItems : 1];
#pragma unroll
    for (int k = 0; k < kGnItems; ++k) {
      const unsigned idx = base + k * step;
      const unsigned pix_k = nshift >= 0 ? idx >> nshift : idx / nvec;   // nvec is a power of two except for some concats
      const unsigned ch_k = (idx - pix_k * nvec) << 3;
      if (idx < total) {
        const bool second = ch_k >= static_cast<unsigned>(c);
        const long long off = (img_row0 + pix_k) * (second ? ld2 : ld) + (second ? ch_k - c : ch_k);
        if (kIn16)
          r16[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(second ? x2 : x) + off));
        else
          f[k] = load8(second ? x2 : x, x_dtype, off);
      }
    }
#pragma unroll
    for (int k = 0; k < kGnItems; ++k) {
      const unsigned idx = base + k * step;   // (recomputed rather than kept: registers are what limits occupancy)
      if (idx >= total) continue;
      const unsigned pix_k = nshift >= 0 ? idx >> nshift : idx / nvec;
      const unsigned ch_k = (idx - pix_k * nvec) << 3;
      Vec8& fk = f[kIn16 ? 0 : k];
      if (kIn16) fk = unpack8(r16[k], x_dtype);
      if (raw) store8(raw, y_dtype, (img_row0 + pix_k) * raw_ld + ch_k, fk);
      if (stats) {
        const float4 a0 = *reinterpret_cast<const float4*>(s_a + ch_k), a1 = *reinterpret_cast<const float4*>(s_a + ch_k + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_b + ch_k), b1 = *reinterpret_cast<const float4*>(s_b + ch_k + 4);
        fk.v[0] = fmaf(fk.v[0], a0.x, b0.x);
        fk.v[1] = fmaf(fk.v[1], a0.y, b0.y);
        fk.v[2] = fmaf(fk.v[2], a0.z, b0.z);
        fk.v[3] = fmaf(fk.v[3], a0.w, b0.w);
        fk.v[4] = fmaf(fk.v[4], a1.x, b1.x);
        fk.v[5] = fmaf(fk.v[5], a1.y, b1.y);
        fk.v[6] = fmaf(fk.v[6], a1.z, b1.z);
        fk.v[7] = fmaf(fk.v[7], a1.w, b1.w);
      }
      if (act == CTTA_ACT_SILU) {
        if (kIn16) {
          // 16-bit input: 4 B/element of traffic makes two MUFU ops per element (ex2 + rcp) the bound at full HBM
          // speed, so use x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op; tanh.approx is good to 2^-11
          // relative, the rounding of the 16-bit value written next
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float hx = 0.5f * fk.v[i];
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hx));
            fk.v[i] = fmaf(hx, th, hx);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) fk.v[i] = __fdividef(fk.v[i], 1.f + __expf(-fk.v[i]));
        }
      }
      if (!up) {
        store8(y, y_dtype, (img_row0 + pix_k) * y_ld + ch_k, fk);
      } else {
        const int ph = pix_k / w, pw = pix_k - ph * w;
        const int w2 = 2 * w;
        const long long o = (static_cast<long long>(n) * 4 * hw + static_cast<long long>(2 * ph) * w2 + 2 * pw);
        store8(y, y_dtype, o * y_ld + ch_k, fk);
        store8(y, y_dtype, (o + 1) * y_ld + ch_k, fk);
        store8(y, y_dtype, (o + w2) * y_ld + ch_k, fk);
        store8(y, y_dtype, (o + w2 + 1) * y_ld + ch_k, fk);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per group of R rows, rows held in registers (NV float4 per lane per row, R * NV = 8 so that every lane has
// 128 B of loads in flight whatever the row length); two-pass variance; 16-bit output with zeroed pad columns.
template <int NV, int R>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int m, int d, int ld,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        void* __restrict__ y, int y_dtype, int y_ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * R;
  if (row0 >= m) return;
  const int nvec = ld >> 2;
  float4 f[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = min(row0 + r, m - 1);
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ld);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      f[r][i] = v < nvec ? __ldg(src + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float4 ga[NV], be[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    ga[i] = v < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    be[i] = v < nvec ? __ldg(reinterpret_cast<const float4*>(beta) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s[R], q[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = 4 * (lane + i * 32);
      if (c0 + 0 >= d) f[r][i].x = 0.f;
      if (c0 + 1 >= d) f[r][i].y = 0.f;
      if (c0 + 2 >= d) f[r][i].z = 0.f;
      if (c0 + 3 >= d) f[r][i].w = 0.f;
      s[r] += (f[r][i].x + f[r][i].y) + (f[r][i].z + f[r][i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
  const float inv_d = 1.f / d;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s[r] *= inv_d;   // mean
    q[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = 4 * (lane + i * 32);
      const float a = c0 + 0 < d ? f[r][i].x - s[r] : 0.f;
      const float b = c0 + 1 < d ? f[r][i].y - s[r] : 0.f;
      const float c = c0 + 2 < d ? f[r][i].z - s[r] : 0.f;
      const float e = c0 + 3 < d ? f[r][i].w - s[r] : 0.f;
      q[r] += (a * a + b * b) + (c * c + e * e);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= m) break;
    const float rstd = rsqrtf(q[r] * inv_d + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const int c0 = 4 * v;
        float4 o4;
        o4.x = c0 + 0 < d ? (f[r][i].x - s[r]) * rstd * ga[i].x + be[i].x : 0.f;
        o4.y = c0 + 1 < d ? (f[r][i].y - s[r]) * rstd * ga[i].y + be[i].y : 0.f;
        o4.z = c0 + 2 < d ? (f[r][i].z - s[r]) * rstd * ga[i].z + be[i].z : 0.f;
        o4.w = c0 + 3 < d ? (f[r][i].w - s[r]) * rstd * ga[i].w + be[i].w : 0.f;
        store4(y, y_dtype, static_cast<long long>(row0 + r) * y_ld + c0, o4);
      }
    }
  }
}

}  // namespace ctta

using namespace ctta;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int ctta_groupnorm_stats(const void* x, int32_t x_dtype, int32_t c, int32_t ld, const void* x2, int32_t c2,
                                    int32_t ld2, int32_t n_img, int32_t hw, int32_t groups, float* stats,
                                    void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && stats && n_img > 0 && hw > 0, "groupnorm_stats: null / empty input");
  CTTA_REQUIRE(groups > 0 && groups <= kGnMaxGroups && (c + c2) % groups == 0, "groupnorm_stats: bad groups");
  CTTA_REQUIRE(c % 4 == 0 && c2 % 4 == 0 && ((c + c2) / groups) % 4 == 0 && ld % 4 == 0 && ld2 % 4 == 0,
               "groupnorm_stats: channels per group and strides must be multiples of 4");
  CTTA_REQUIRE(al16(x) && (!x2 || al16(x2)), "groupnorm_stats: inputs must be 16-byte aligned");
  CTTA_REQUIRE(c2 == 0 || x2, "groupnorm_stats: c2 > 0 without x2");
  CTTA_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * n_img * groups, stream));
  // enough slabs to fill the machine (~4 CTAs per SM over all images), at least 8 rows each
  int slabs = (sm_count() * 4 + n_img - 1) / n_img;
  int rows_per_slab = (hw + slabs - 1) / slabs;
  if (rows_per_slab < 8) rows_per_slab = 8;
  slabs = (hw + rows_per_slab - 1) / rows_per_slab;
  dim3 grid(slabs, n_img), block(32, 8);
  gn_moments_kernel<<<grid, block, 0, stream>>>(x, x_dtype, c, ld, x2, c2, ld2, hw, rows_per_slab, groups, stats);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_groupnorm_apply(const void* x, int32_t x_dtype, int32_t c, int32_t ld, const void* x2, int32_t c2,
                                    int32_t ld2, int32_t n_img, int32_t h, int32_t w, int32_t groups,
                                    const float* stats, const float* gamma, const float* beta, float eps,
                                    int32_t act, int32_t upsample2x, void* y, int32_t y_dtype, int32_t y_ld, void* raw_out,
                                    int32_t raw_ld, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && n_img > 0 && h > 0 && w > 0, "groupnorm_apply: null / empty input");
  CTTA_REQUIRE(c % 8 == 0 && c2 % 8 == 0 && ld % 8 == 0 && ld2 % 8 == 0 && y_ld % 8 == 0 && raw_ld % 8 == 0,
               "groupnorm_apply: channels and strides must be multiples of 8");
  CTTA_REQUIRE(static_cast<long long>(h) * w * ((c + c2) / 8) < (1LL << 31) - (1 << 24),
               "groupnorm_apply: image too large for 32-bit indexing");
  CTTA_REQUIRE(al16(x) && (!x2 || al16(x2)) && al16(y) && (!raw_out || al16(raw_out)),
               "groupnorm_apply: tensors must be 16-byte aligned");
  if (stats) {
    CTTA_REQUIRE(c + c2 <= kGnMaxChannels, "groupnorm_apply: at most %d channels", kGnMaxChannels);
    CTTA_REQUIRE(gamma && beta && groups > 0 && groups <= kGnMaxGroups && (c + c2) % groups == 0 &&
                     ((c + c2) / groups) % 4 == 0 && al16(gamma) && al16(beta),
                 "groupnorm_apply: bad group configuration");
  }
  CTTA_REQUIRE(!(raw_out && upsample2x), "groupnorm_apply: raw_out and upsample are exclusive");
  CTTA_REQUIRE(act == CTTA_ACT_NONE || act == CTTA_ACT_SILU, "groupnorm_apply: act must be NONE or SILU");
  const long long total = static_cast<long long>(h) * w * ((c + c2) / 8);
  const int items = x_dtype == CTTA_F32 ? 2 : 8;
  const int nvec_h = (c + c2) / 8;
  int nshift = -1;
  if ((nvec_h & (nvec_h - 1)) == 0) {
    nshift = 0;
    while ((1 << nshift) < nvec_h) ++nshift;
  }
  long long blocks = (total + 256 * items - 1) / (256 * items);
  const long long cap = (static_cast<long long>(sm_count()) * 32 + n_img - 1) / n_img;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), n_img);
  if (items == 2)
    gn_apply_kernel<2, false><<<grid, 256, 0, stream>>>(x, x_dtype, c, ld, x2, c2, ld2, h, w, groups, stats, gamma, beta,
                                                 eps, act, upsample2x, y, y_dtype, y_ld, raw_out, raw_ld, nshift);
  else
    gn_apply_kernel<8, true><<<grid, 256, 0, stream>>>(x, x_dtype, c, ld, x2, c2, ld2, h, w, groups, stats, gamma, beta,
                                                 eps, act, upsample2x, y, y_dtype, y_ld, raw_out, raw_ld, nshift);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_layernorm(const float* x, int32_t m, int32_t d, int32_t ld, const float* gamma, const float* beta,
                              float eps, void* y, int32_t y_dtype, int32_t y_ld, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && gamma && beta && m > 0, "layernorm: null / empty input");
  CTTA_REQUIRE(d > 0 && d <= ld && ld % 4 == 0 && ld <= 1024 && y_ld % 4 == 0 && y_ld >= ld,
               "layernorm: need d <= ld <= 1024, ld %% 4 == 0 (d=%d ld=%d)", d, ld);
  CTTA_REQUIRE(al16(x) && al16(y), "layernorm: tensors must be 16-byte aligned");
  CTTA_REQUIRE(al16(gamma) && al16(beta), "layernorm: gamma / beta must be 16-byte aligned");
  const int warps_per_block = 8;
  const int nv = (ld / 4 + 31) / 32;   // float4 per lane per row
  if (nv <= 2) {
    const int blocks = ((m + 3) / 4 + warps_per_block - 1) / warps_per_block;
    layernorm_kernel<2, 4><<<blocks, 256, 0, stream>>>(x, m, d, ld, gamma, beta, eps, y, y_dtype, y_ld);
  } else if (nv <= 4) {
    const int blocks = ((m + 1) / 2 + warps_per_block - 1) / warps_per_block;
    layernorm_kernel<4, 2><<<blocks, 256, 0, stream>>>(x, m, d, ld, gamma, beta, eps, y, y_dtype, y_ld);
  } else {
    const int blocks = (m + warps_per_block - 1) / warps_per_block;
    layernorm_kernel<8, 1><<<blocks, 256, 0, stream>>>(x, m, d, ld, gamma, beta, eps, y, y_dtype, y_ld);
  }
  CTTA_LAUNCH_CHECK();
  return 0;
}
