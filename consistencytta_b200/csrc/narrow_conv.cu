// Single-output-channel convolutions (VAE conv_out 128 -> 1, HiFi-GAN conv_post 32 -> 1): bandwidth-bound reductions
// over the channel axis, not GEMMs — a 128 x 16 tcgen05 tile would spend its 60-cycle issue floor per 16-wide K slice on
// one useful output column.  Here the C / 8 lanes that share a pixel each own 8 channels (one 16-byte load per tap,
// fully coalesced across the warp), multiply by fp32 weights held in shared memory and combine with shuffles.
// Reference call sites: audioldm/variational_autoencoder/modules.py:680 (conv_out), audioldm/hifigan/models.py:114-115
// (conv_post + tanh).  Dispatched from ctta_gemm for n == 1 multi-tap problems.
#include "ctta_internal.h"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ctta {

struct NarrowParams {
  const unsigned short* a;
  long long a_ld;
  int c, n_img, h, w, ntaps, is_bf16;
  short d0[CTTA_MAX_TAPS], d1[CTTA_MAX_TAPS];
  const unsigned short* wgt;
  int c_pad;
  const float* bias;
  int act;
  float out_scale;
  float* out;
  long long out_ld;
  unsigned short* out2;
  long long out2_ld;
  int out2_bf16;
};

template <int LPP>  // lanes per pixel = c / 8
__global__ void __launch_bounds__(256) narrow_conv_kernel(const __grid_constant__ NarrowParams p) {
  extern __shared__ float w_s[];   // [ntaps][c]
  for (int i = threadIdx.x; i < p.ntaps * p.c; i += blockDim.x) {
    const int j = i / p.c, ch = i - j * p.c;
    const unsigned short raw = p.wgt[j * p.c_pad + ch];
    w_s[i] = p.is_bf16 ? __uint_as_float(static_cast<uint32_t>(raw) << 16) : __half2float(__ushort_as_half(raw));
  }
  __syncthreads();
  constexpr int PPW = 32 / LPP;   // pixels per warp per iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPP;
  const long long hw = static_cast<long long>(p.h) * p.w;
  const long long total = hw * p.n_img;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long g = ((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5); g * PPW < total; g += warps) {
    const long long pix = g * PPW + lane / LPP;
    const bool valid = pix < total;
    const long long img = valid ? pix / hw : 0;
    const int rem = valid ? static_cast<int>(pix - img * hw) : 0;
    const int ph = rem / p.w, pw = rem - ph * p.w;
    float acc = 0.f;
    for (int j = 0; j < p.ntaps; ++j) {
      const int hh = ph + p.d1[j], ww = pw + p.d0[j];
      if (valid && hh >= 0 && hh < p.h && ww >= 0 && ww < p.w) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.a + ((img * p.h + hh) * p.w + ww) * p.a_ld + sub * 8));
        const float* wj = w_s + j * p.c + sub * 8;
        const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float x0, x1;
          if (p.is_bf16) {
            x0 = __uint_as_float(wd[i] << 16);
            x1 = __uint_as_float(wd[i] & 0xFFFF0000u);
          } else {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&wd[i]));
            x0 = t.x;
            x1 = t.y;
          }
          acc = fmaf(x0, wj[2 * i], acc);
          acc = fmaf(x1, wj[2 * i + 1], acc);
        }
      }
    }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (valid && sub == 0) {
      float v = acc + (p.bias != nullptr ? __ldg(p.bias) : 0.f);
      if (p.act == CTTA_ACT_TANH) v = tanhf(v);
      else if (p.act == CTTA_ACT_SILU) v = v / (1.f + __expf(-v));
      v *= p.out_scale;
      if (p.out) p.out[pix * p.out_ld] = v;
      if (p.out2)
        p.out2[pix * p.out2_ld] = p.out2_bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v))
                                               : __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
    }
  }
}

// returns 1 when the problem is not a single-output-channel multi-tap convolution this kernel covers
int narrow_conv_launch(const ctta_gemm_desc* d, cudaStream_t stream) {
  if (d->n != 1 || d->ntaps < 2 || d->a_mode == CTTA_A_ROWS || d->residual || d->rowadd || d->accumulate || d->stats ||
      d->out_stride != 1 || d->out_off != 0 || (d->out && d->out_dtype != CTTA_F32) || d->wgt_img_stride != 0 ||
      (d->act != CTTA_ACT_NONE && d->act != CTTA_ACT_TANH && d->act != CTTA_ACT_SILU) ||
      (d->out2 && d->act2 != CTTA_ACT_NONE) || d->rows_per_img != d->h * d->w || d->out_rows_per_img != d->rows_per_img)
    return 1;
  const int lpp = d->c / 8;
  if (d->c % 8 != 0 || (lpp != 4 && lpp != 8 && lpp != 16 && lpp != 32)) return 1;
  NarrowParams p{};
  p.a = reinterpret_cast<const unsigned short*>(d->a);
  p.a_ld = d->a_ld;
  p.c = d->c;
  p.n_img = d->n_img;
  p.h = d->h;
  p.w = d->w;
  p.ntaps = d->ntaps;
  p.is_bf16 = d->ab_dtype == CTTA_BF16;
  for (int j = 0; j < d->ntaps; ++j) {
    p.d0[j] = d->tap_d0[j];
    p.d1[j] = d->tap_d1[j];
  }
  p.wgt = reinterpret_cast<const unsigned short*>(d->wgt);
  p.c_pad = (d->c + 63) / 64 * 64;
  p.bias = d->bias;
  p.act = d->act;
  p.out_scale = d->out_scale;
  p.out = reinterpret_cast<float*>(d->out);
  p.out_ld = d->out_ld;
  p.out2 = reinterpret_cast<unsigned short*>(d->out2);
  p.out2_ld = d->out2_ld;
  p.out2_bf16 = p.is_bf16;
  const long long total = static_cast<long long>(d->h) * d->w * d->n_img;
  const int ppw = 32 / lpp;
  long long blocks = (total / ppw + 7) / 8 + 1;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  const int smem = d->ntaps * d->c * static_cast<int>(sizeof(float));
  const int grid = static_cast<int>(blocks);
  switch (lpp) {
    case 4: narrow_conv_kernel<4><<<grid, 256, smem, stream>>>(p); break;
    case 8: narrow_conv_kernel<8><<<grid, 256, smem, stream>>>(p); break;
    case 16: narrow_conv_kernel<16><<<grid, 256, smem, stream>>>(p); break;
    default: narrow_conv_kernel<32><<<grid, 256, smem, stream>>>(p); break;
  }
  CTTA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctta
