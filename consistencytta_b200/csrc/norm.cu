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
// One thread per (pixel, 4-channel vector). grid (blocks, n_img).
__global__ void __launch_bounds__(256) gn_apply_kernel(const void* __restrict__ x, int x_dtype, int c, int ld,
                                                       const void* __restrict__ x2, int c2, int ld2, int h, int w,
                                                       int groups, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int act, int up, void* __restrict__ y, int y_dtype,
                                                       int y_ld, void* __restrict__ raw, int raw_ld) {
  __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
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
  const int nvec = ctot >> 2;
  const long long total = static_cast<long long>(hw) * nvec;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % nvec);
    const int pix = static_cast<int>(idx / nvec);
    const int ch = v << 2;
    const bool second = ch >= c;
    const void* src = second ? x2 : x;
    const int sld = second ? ld2 : ld;
    const int sch = second ? ch - c : ch;
    float4 f = load4(src, x_dtype, (static_cast<long long>(n) * hw + pix) * sld + sch);
    if (raw) store4(raw, y_dtype, (static_cast<long long>(n) * hw + pix) * raw_ld + ch, f);
    if (stats) {
      const int g = ch / cpg;
      const float m = s_mean[g], rs = s_rstd[g];
      const float4 ga = *reinterpret_cast<const float4*>(gamma + ch);
      const float4 be = *reinterpret_cast<const float4*>(beta + ch);
      f.x = (f.x - m) * rs * ga.x + be.x;
      f.y = (f.y - m) * rs * ga.y + be.y;
      f.z = (f.z - m) * rs * ga.z + be.z;
      f.w = (f.w - m) * rs * ga.w + be.w;
    }
    if (act == CTTA_ACT_SILU) {
      f.x = f.x / (1.f + __expf(-f.x));
      f.y = f.y / (1.f + __expf(-f.y));
      f.z = f.z / (1.f + __expf(-f.z));
      f.w = f.w / (1.f + __expf(-f.w));
    }
    if (!up) {
      store4(y, y_dtype, (static_cast<long long>(n) * hw + pix) * y_ld + ch, f);
    } else {
      const int ph = pix / w, pw = pix - ph * w;
      const int w2 = 2 * w;
      const long long o = (static_cast<long long>(n) * 4 * hw + static_cast<long long>(2 * ph) * w2 + 2 * pw);
      store4(y, y_dtype, o * y_ld + ch, f);
      store4(y, y_dtype, (o + 1) * y_ld + ch, f);
      store4(y, y_dtype, (o + w2) * y_ld + ch, f);
      store4(y, y_dtype, (o + w2 + 1) * y_ld + ch, f);
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row, row held in registers (ld <= 1024 fp32 -> <= 8 float4 per lane), two-pass variance.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int m, int d, int ld,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        void* __restrict__ y, int y_dtype, int y_ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= m) return;
  const int nvec = ld >> 2;
  const float* row = x + static_cast<long long>(warp) * ld;
  float4 f[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      f[i] = *reinterpret_cast<const float4*>(row + 4 * v);
      const int c0 = 4 * v;
      if (c0 + 0 >= d) f[i].x = 0.f;
      if (c0 + 1 >= d) f[i].y = 0.f;
      if (c0 + 2 >= d) f[i].z = 0.f;
      if (c0 + 3 >= d) f[i].w = 0.f;
      s += (f[i].x + f[i].y) + (f[i].z + f[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const int c0 = 4 * v;
      const float a = c0 + 0 < d ? f[i].x - mean : 0.f;
      const float b = c0 + 1 < d ? f[i].y - mean : 0.f;
      const float c = c0 + 2 < d ? f[i].z - mean : 0.f;
      const float e = c0 + 3 < d ? f[i].w - mean : 0.f;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / d + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const int c0 = 4 * v;
      float4 o4;
      o4.x = c0 + 0 < d ? (f[i].x - mean) * rstd * gamma[c0 + 0] + beta[c0 + 0] : 0.f;
      o4.y = c0 + 1 < d ? (f[i].y - mean) * rstd * gamma[c0 + 1] + beta[c0 + 1] : 0.f;
      o4.z = c0 + 2 < d ? (f[i].z - mean) * rstd * gamma[c0 + 2] + beta[c0 + 2] : 0.f;
      o4.w = c0 + 3 < d ? (f[i].w - mean) * rstd * gamma[c0 + 3] + beta[c0 + 3] : 0.f;
      store4(y, y_dtype, static_cast<long long>(warp) * y_ld + c0, o4);
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
  CTTA_REQUIRE(c % 4 == 0 && c2 % 4 == 0 && ld % 4 == 0 && ld2 % 4 == 0 && y_ld % 4 == 0 && raw_ld % 4 == 0,
               "groupnorm_apply: channels and strides must be multiples of 4");
  CTTA_REQUIRE(al16(x) && (!x2 || al16(x2)) && al16(y) && (!raw_out || al16(raw_out)),
               "groupnorm_apply: tensors must be 16-byte aligned");
  if (stats) {
    CTTA_REQUIRE(gamma && beta && groups > 0 && groups <= kGnMaxGroups && (c + c2) % groups == 0 &&
                     ((c + c2) / groups) % 4 == 0 && al16(gamma) && al16(beta),
                 "groupnorm_apply: bad group configuration");
  }
  CTTA_REQUIRE(!(raw_out && upsample2x), "groupnorm_apply: raw_out and upsample are exclusive");
  CTTA_REQUIRE(act == CTTA_ACT_NONE || act == CTTA_ACT_SILU, "groupnorm_apply: act must be NONE or SILU");
  const long long total = static_cast<long long>(h) * w * ((c + c2) / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (static_cast<long long>(sm_count()) * 16 + n_img - 1) / n_img;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), n_img);
  gn_apply_kernel<<<grid, 256, 0, stream>>>(x, x_dtype, c, ld, x2, c2, ld2, h, w, groups, stats, gamma, beta,
                                            eps, act, upsample2x, y, y_dtype, y_ld, raw_out, raw_ld);
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
  const int warps_per_block = 8;
  const int blocks = (m + warps_per_block - 1) / warps_per_block;
  layernorm_kernel<<<blocks, 256, 0, stream>>>(x, m, d, ld, gamma, beta, eps, y, y_dtype, y_ld);
  CTTA_LAUNCH_CHECK();
  return 0;
}
