// Small bandwidth / latency bound kernels around the tensor-core path: stride-2 im2col gather, layout conversion
// at the module boundary, timestep/guidance features, small-M fp32 linear, waveform post-processing, CFG mix.
// Reference call sites: see include/ctta.h.
#include "ctta_internal.h"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ctta {

// ------------------------------------------------------------------------------------------- im2col (stride 2)
// out[(n, oh, ow), (i*3 + j)*c + ch] = x[n, 2*oh + i - 1, 2*ow + j - 1, ch]  (zero outside); 8 channels / thread.
__global__ void __launch_bounds__(256) im2col_s2_kernel(const uint4* __restrict__ x, int n_img, int h, int w, int c8,
                                                        uint4* __restrict__ out) {
  const int oh_n = h >> 1, ow_n = w >> 1;
  const long long total = static_cast<long long>(n_img) * oh_n * ow_n * 9 * c8;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % c8);
    long long r = idx / c8;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int ow = static_cast<int>(r % ow_n);
    r /= ow_n;
    const int oh = static_cast<int>(r % oh_n);
    const int n = static_cast<int>(r / oh_n);
    const int ih = 2 * oh + tap / 3 - 1, iw = 2 * ow + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < h && iw >= 0 && iw < w) val = x[((static_cast<long long>(n) * h + ih) * w + iw) * c8 + v];
    out[idx] = val;
  }
}

// ------------------------------------------------------------------------------------------- layout conversion
__device__ __forceinline__ unsigned short cvt16e(float v, int dtype) {
  if (dtype == CTTA_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __half_as_ushort(__float2half_rn(v));
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int n_img, int c, int hw, void* __restrict__ y,
                                    int y_dtype, int y_ld, float scale, const float* __restrict__ scale_dev) {
  const long long total = static_cast<long long>(n_img) * hw * c;
  if (scale_dev != nullptr) scale *= __ldg(scale_dev);
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % c);
    const long long r = idx / c;  // n * hw + pix
    const int pix = static_cast<int>(r % hw);
    const int n = static_cast<int>(r / hw);
    const float v = x[(static_cast<long long>(n) * c + ch) * hw + pix] * scale;
    if (y_dtype == CTTA_F32) reinterpret_cast<float*>(y)[r * y_ld + ch] = v;
    else reinterpret_cast<unsigned short*>(y)[r * y_ld + ch] = cvt16e(v, y_dtype);
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int n_img, int c, int hw, int ld,
                                    float* __restrict__ y) {
  const long long total = static_cast<long long>(n_img) * hw * c;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pix = static_cast<int>(idx % hw);
    const long long r = idx / hw;  // n * c + ch
    const int ch = static_cast<int>(r % c);
    const int n = static_cast<int>(r / c);
    y[idx] = x[(static_cast<long long>(n) * hw + pix) * ld + ch];
  }
}

// ------------------------------------------------------------------------------------------- embeddings
// t_feat[b, i] = cos(t f_i), t_feat[b, 128 + i] = sin(t f_i), f_i = exp(-ln(1e4) i / 128)   (flip_sin_to_cos)
// g_feat[b, i] = cos(2 pi w W_i), g_feat[b, E + i] = sin(2 pi w W_i)                          (E = 512)
// evaluated in double: the reference evaluates the guidance features in float64 (unet_2d_condition_guided.py:703-708).
__global__ void time_features_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                     const float* __restrict__ gw, int batch, int e, float* __restrict__ t_feat,
                                     float* __restrict__ g_feat) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    const float f = expf(-logf(10000.f) * static_cast<float>(i) / 128.f);
    const float arg = t[b] * f;  // fp32 like the reference (embeddings.py:49-60)
    t_feat[b * 256 + i] = cosf(arg);
    t_feat[b * 256 + 128 + i] = sinf(arg);
  }
  for (int i = threadIdx.x; i < e; i += blockDim.x) {
    const double arg = static_cast<double>(w[b]) * static_cast<double>(gw[i]) * 2.0 * 3.14159265358979323846;
    g_feat[b * 2 * e + i] = static_cast<float>(cos(arg));
    g_feat[b * 2 * e + e + i] = static_cast<float>(sin(arg));
  }
}

// y[m, n] = act_out( sum_k act_in(x[m, k]) W[n, k] + b[n] ) (+ y).  One warp per output column, 8 rows per pass.
__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, int m, int k,
                                                           const float* __restrict__ wgt,
                                                           const float* __restrict__ bias, int n, int act_in,
                                                           int act_out, int accumulate, float* __restrict__ y) {
  const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (col >= n) return;
  const float* wr = wgt + static_cast<long long>(col) * k;
  for (int m0 = 0; m0 < m; m0 += 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int kk = lane * 4; kk < k; kk += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(wr + kk);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (m0 + r < m) {
          float4 xv = *reinterpret_cast<const float4*>(x + static_cast<long long>(m0 + r) * k + kk);
          if (act_in == CTTA_ACT_SILU) {
            xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w);
          }
          acc[r] += (xv.x * wv.x + xv.y * wv.y) + (xv.z * wv.z + xv.w * wv.w);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (m0 + r < m) {
          float v = acc[r] + (bias ? bias[col] : 0.f);
          if (act_out == CTTA_ACT_SILU) v = silu_f(v);
          float* o = y + static_cast<long long>(m0 + r) * n + col;
          *o = accumulate ? *o + v : v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- waveform post-proc
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__global__ void minmax_init_kernel(float* mm) {
  mm[0] = __int_as_float(0x7f800000);   // +inf (min)
  mm[1] = __int_as_float(0xff800000);   // -inf (max)
}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, long long numel,
                                                     float* __restrict__ mm) {
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[8], s_hi[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_lo[warp] = lo;
    s_hi[warp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      lo = fminf(lo, s_lo[i]);
      hi = fmaxf(hi, s_hi[i]);
    }
    atomic_min_f(&mm[0], lo);
    atomic_max_f(&mm[1], hi);
  }
}
// (wav - (max + min) / 2) * 32768 -> int16 with numpy astype semantics: truncate toward zero, wrap modulo 2^16.
__global__ void __launch_bounds__(256) to_int16_kernel(const float* __restrict__ x, long long numel,
                                                       const float* __restrict__ mm, short* __restrict__ out) {
  const float centre = (mm[1] + mm[0]) / 2.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = (x[i] - centre) * 32768.f;
    const int iv = __float2int_rz(v);
    out[i] = static_cast<short>(static_cast<unsigned short>(iv & 0xFFFF));
  }
}

__global__ void __launch_bounds__(256) cfg_mix_kernel(const float* __restrict__ x, long long half_numel, float s,
                                                      float* __restrict__ y) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < half_numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    y[i] = (1.f - s) * x[i] + s * x[half_numel + i];
  }
}

__global__ void __launch_bounds__(256) lrelu_cast_kernel(const float4* __restrict__ x, long long n4, float slope,
                                                         uint2* __restrict__ y, int y_dtype) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 f = x[i];
    f.x = f.x > 0.f ? f.x : f.x * slope;
    f.y = f.y > 0.f ? f.y : f.y * slope;
    f.z = f.z > 0.f ? f.z : f.z * slope;
    f.w = f.w > 0.f ? f.w : f.w * slope;
    uint2 u;
    u.x = static_cast<uint32_t>(cvt16e(f.x, y_dtype)) | (static_cast<uint32_t>(cvt16e(f.y, y_dtype)) << 16);
    u.y = static_cast<uint32_t>(cvt16e(f.z, y_dtype)) | (static_cast<uint32_t>(cvt16e(f.w, y_dtype)) << 16);
    y[i] = u;
  }
}

static int grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace ctta

using namespace ctta;

extern "C" int ctta_im2col_s2(const void* x, int32_t n_img, int32_t h, int32_t w, int32_t c, void* a, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && a && n_img > 0, "im2col_s2: null input");
  CTTA_REQUIRE(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "im2col_s2: need even h, w and c %% 8 == 0");
  const long long total = static_cast<long long>(n_img) * (h / 2) * (w / 2) * 9 * (c / 8);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), n_img, h, w, c / 8,
                                                             reinterpret_cast<uint4*>(a));
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_nchw_to_nhwc(const float* x, int32_t n_img, int32_t c, int32_t hw, void* y, int32_t y_dtype,
                                 int32_t y_ld, float scale, const float* scale_dev, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && n_img > 0 && c > 0 && hw > 0 && y_ld >= c, "nchw_to_nhwc: bad arguments");
  const long long total = static_cast<long long>(n_img) * hw * c;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, n_img, c, hw, y, y_dtype, y_ld, scale, scale_dev);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_nhwc_to_nchw(const float* x, int32_t n_img, int32_t c, int32_t hw, int32_t ld, float* y,
                                 void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && n_img > 0 && c > 0 && hw > 0 && ld >= c, "nhwc_to_nchw: bad arguments");
  const long long total = static_cast<long long>(n_img) * hw * c;
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, n_img, c, hw, ld, y);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_time_features(const float* t, const float* w, const float* gw, int32_t batch, float* t_feat,
                                  float* g_feat, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(t && w && gw && t_feat && g_feat && batch > 0, "time_features: bad arguments");
  time_features_kernel<<<batch, 128, 0, stream>>>(t, w, gw, batch, 512, t_feat, g_feat);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_small_linear(const float* x, int32_t m, int32_t k, const float* wgt, const float* bias, int32_t n,
                                 int32_t act_in, int32_t act_out, int32_t accumulate, float* y, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && wgt && y && m > 0 && n > 0 && k > 0 && k % 4 == 0, "small_linear: bad arguments (k %% 4)");
  small_linear_kernel<<<(n + 7) / 8, 256, 0, stream>>>(x, m, k, wgt, bias, n, act_in, act_out, accumulate, y);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_wave_minmax(const float* wav, int64_t numel, float* minmax, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(wav && minmax && numel > 0, "wave_minmax: bad arguments");
  minmax_init_kernel<<<1, 1, 0, stream>>>(minmax);
  CTTA_LAUNCH_CHECK();
  minmax_kernel<<<grid_for(numel, 256 * 8), 256, 0, stream>>>(wav, numel, minmax);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_wave_to_int16(const float* wav, int64_t numel, const float* minmax, int16_t* out, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(wav && minmax && out && numel > 0, "wave_to_int16: bad arguments");
  to_int16_kernel<<<grid_for(numel, 256 * 4), 256, 0, stream>>>(wav, numel, minmax, out);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_lrelu_cast(const float* x, int64_t numel, float slope, void* y, int32_t y_dtype, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && numel > 0 && numel % 4 == 0, "lrelu_cast: numel must be a positive multiple of 4");
  CTTA_REQUIRE(y_dtype == CTTA_F16 || y_dtype == CTTA_BF16, "lrelu_cast: 16-bit output only");
  CTTA_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
               "lrelu_cast: misaligned tensors");
  lrelu_cast_kernel<<<grid_for(numel / 4, 256 * 4), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), numel / 4, slope,
                                                                    reinterpret_cast<uint2*>(y), y_dtype);
  CTTA_LAUNCH_CHECK();
  return 0;
}

namespace ctta {
struct MrfPtrs {
  const uint4* x[4];
};
__device__ __forceinline__ void mrf_unpack8(const uint4& u, int is_bf16, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (is_bf16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
// 8 elements per thread per input (16-byte loads), all inputs in flight before the first use
__global__ void __launch_bounds__(256) mrf_combine_kernel(MrfPtrs in, int n_in, long long nvec, int is_bf16, float in_inv,
                                                          float out_scale, float out_slope, uint4* __restrict__ y) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < n_in) u[k] = __ldg(in.x[k] + i);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < n_in) {
        float f[8];
        mrf_unpack8(u[k], is_bf16, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j] < 0.f ? f[j] * in_inv : f[j];
      }
    }
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = acc[2 * j] * out_scale, b = acc[2 * j + 1] * out_scale;
      a = a < 0.f ? a * out_slope : a;
      b = b < 0.f ? b * out_slope : b;
      if (is_bf16) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
      } else {
        const __half2 h = __floats2half2_rn(a, b);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    y[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
}  // namespace ctta

extern "C" int ctta_mrf_combine(const void* const* x, int32_t n_in, int64_t numel, int32_t dtype, float in_slope,
                                float out_scale, float out_slope, void* y, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && n_in >= 1 && n_in <= 4 && numel > 0 && numel % 8 == 0, "mrf_combine: need 1..4 inputs, numel %% 8 == 0");
  CTTA_REQUIRE(dtype == CTTA_F16 || dtype == CTTA_BF16, "mrf_combine: 16-bit tensors only");
  CTTA_REQUIRE(in_slope > 0.f, "mrf_combine: in_slope must be positive");
  ctta::MrfPtrs in{};
  for (int k = 0; k < n_in; ++k) {
    CTTA_REQUIRE(x[k] && (reinterpret_cast<uintptr_t>(x[k]) & 15) == 0, "mrf_combine: inputs must be 16-byte aligned");
    in.x[k] = reinterpret_cast<const uint4*>(x[k]);
  }
  CTTA_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "mrf_combine: output must be 16-byte aligned");
  ctta::mrf_combine_kernel<<<ctta::grid_for(numel / 8, 256 * 2), 256, 0, stream>>>(
      in, n_in, numel / 8, dtype == CTTA_BF16, 1.f / in_slope, out_scale, out_slope, reinterpret_cast<uint4*>(y));
  CTTA_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------- single-channel conv tail
// A convolution with ONE output channel is a bandwidth problem, not a GEMM: as an implicit GEMM it would issue
// taps * C / 16 tensor-core instructions per 128 pixels for one useful accumulator column.  It is restated as
//   z[p, j] = sum_c x[p, c] * w[j, c]      (pointwise GEMM, N = taps: C / 16 instructions per 128 pixels, x read once)
//   y[p]    = act(bias + sum_j z[p + shift_j, j])                                                   (this kernel)
// which computes the same products and only reorders the fp32 summation.
struct TapSumParams {
  int ntaps;
  short d0[CTTA_MAX_TAPS], d1[CTTA_MAX_TAPS];
};
__global__ void __launch_bounds__(256) tap_sum_kernel(const float* __restrict__ z, int z_ld, int n_img, int h, int w,
                                                      const __grid_constant__ TapSumParams tp, const float* __restrict__ bias,
                                                      int act, float* __restrict__ out, void* __restrict__ out16, int out16_dtype) {
  const long long hw = static_cast<long long>(h) * w;
  const long long total = hw * n_img;
  const float b0 = bias != nullptr ? __ldg(bias) : 0.f;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int rem = static_cast<int>(p % hw);
    const int ph = rem / w, pw = rem - ph * w;
    float acc = b0;
    for (int j = 0; j < tp.ntaps; ++j) {
      const int hh = ph + tp.d1[j], ww = pw + tp.d0[j];
      if (hh >= 0 && hh < h && ww >= 0 && ww < w)
        acc += __ldg(z + (p + static_cast<long long>(tp.d1[j]) * w + tp.d0[j]) * z_ld + j);
    }
    if (act == CTTA_ACT_TANH) acc = tanhf(acc);
    else if (act == CTTA_ACT_SILU) acc = acc / (1.f + __expf(-acc));
    if (out) out[p] = acc;
    if (out16) reinterpret_cast<unsigned short*>(out16)[p] = cvt16e(acc, out16_dtype);
  }
}

extern "C" int ctta_tap_sum(const float* z, int32_t z_ld, int32_t n_img, int32_t h, int32_t w, int32_t ntaps,
                            const int16_t* tap_d0, const int16_t* tap_d1, const float* bias, int32_t act, float* out,
                            void* out16, int32_t out16_dtype, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(z && tap_d0 && tap_d1 && (out || out16), "tap_sum: null argument");
  CTTA_REQUIRE(ntaps >= 1 && ntaps <= CTTA_MAX_TAPS && z_ld >= ntaps && n_img >= 1 && h >= 1 && w >= 1,
               "tap_sum: bad geometry (ntaps=%d z_ld=%d n_img=%d h=%d w=%d)", ntaps, z_ld, n_img, h, w);
  CTTA_REQUIRE(act == CTTA_ACT_NONE || act == CTTA_ACT_TANH || act == CTTA_ACT_SILU, "tap_sum: unsupported activation");
  TapSumParams tp{};
  tp.ntaps = ntaps;
  for (int j = 0; j < ntaps; ++j) {
    tp.d0[j] = tap_d0[j];
    tp.d1[j] = tap_d1[j];
  }
  const long long total = static_cast<long long>(n_img) * h * w;
  tap_sum_kernel<<<grid_for(total, 256), 256, 0, stream>>>(z, z_ld, n_img, h, w, tp, bias, act, out, out16, out16_dtype);
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_cfg_mix(const float* x, int64_t half_numel, float s, float* y, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && half_numel > 0, "cfg_mix: bad arguments");
  cfg_mix_kernel<<<grid_for(half_numel, 256 * 4), 256, 0, stream>>>(x, half_numel, s, y);
  CTTA_LAUNCH_CHECK();
  return 0;
}
