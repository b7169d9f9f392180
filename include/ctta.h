/*
 * libctta — C-ABI of the B200 (sm_100a) kernels behind the ConsistencyTTA single-step generation path
 * (guided UNet forward -> AudioLDM VAE decode -> HiFi-GAN vocoder).
 *
 * The reference (Bai-YT/ConsistencyTTA) has no FFI layer: its boundary is the Python nn.Module API
 * (SURVEY.md 8b).  The Python modules in consistencytta_b200/ keep that API and call these entry points
 * through ctypes with raw device pointers; every function replaces one family of ATen call sites of the
 * reference, cited per function as  <reference file>:<line>.
 *
 * Conventions: every function returns 0 on success or a negative ctta_status; ctta_last_error() gives the
 * message of the last failure on the calling thread.  All pointers are DEVICE pointers unless named host_*.
 * No function allocates, synchronises or touches a stream other than the `stream` argument (a cudaStream_t
 * passed as void*).  Activations are channels-last: images [N, H, W, C], sequences [B, T, C], tokens [M, C].
 */
#ifndef CTTA_H_
#define CTTA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CTTA_OK = 0,
  CTTA_ERR_INVALID = -1,   /* bad argument (shape, alignment, enum) */
  CTTA_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  CTTA_ERR_UNSUPPORTED = -3
} ctta_status;

typedef enum { CTTA_F32 = 0, CTTA_F16 = 1, CTTA_BF16 = 2 } ctta_dtype;

/* epilogue activations */
typedef enum {
  CTTA_ACT_NONE = 0,
  CTTA_ACT_SILU = 1,
  CTTA_ACT_GEGLU = 2, /* columns interleaved (value, gate): out[:, j] = v[2j] * gelu_erf(v[2j+1]); out has N/2 cols */
  CTTA_ACT_TANH = 3,
  CTTA_ACT_LRELU = 4
} ctta_act;

/* how the A operand (activations) is addressed */
typedef enum {
  CTTA_A_ROWS = 0,   /* A is a row-major matrix [M, K] (Linear, 1x1 conv, pre-gathered im2col)            */
  CTTA_A_CONV1D = 1, /* A is [B, T, C]; tap j reads rows t + tap_d0[j] (zero outside [0, T))              */
  CTTA_A_CONV2D = 2  /* A is [N, H, W, C]; tap j reads pixel (h + tap_d1[j], w + tap_d0[j]), zero padded  */
} ctta_a_mode;

#define CTTA_MAX_TAPS 16

const char* ctta_last_error(void);
int ctta_version(void);
/* number of kernels launched by this library since load (all threads); used for bench.py's gpu_launches */
long long ctta_launch_count(void);

/*
 * Implicit-GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators, persistent, warp-specialised):
 *     acc[m, n] = sum_{tap j} sum_{c < C} A(m shifted by tap j, c) * W[n, j * c_pad + c]
 *     v   = act(acc + bias[n] + rowadd[img(m), n])
 *     v   = (v + residual[orow(m), n]) * out_scale
 *     out[orow(m), n] = v (+ out[orow(m), n] if accumulate) ;   out2[orow(m), n] = f16/bf16( act2(v) )   (optional)
 *     stats[img(m), group(n)] += (v, v^2)                                                              (optional)
 * Replaces, in the reference: F.conv2d 3x3/1x1 (diffusers/models/resnet.py:570,590,593,157;
 * unet_2d_condition_guided.py:863,940; audioldm/variational_autoencoder/modules.py:56,159,167,173,207-209,228,
 * 658,680; autoencoder.py:99), F.linear (attention_processor.py:1110-1135; transformer_2d.py:267,295;
 * attention.py:378,431), F.conv1d and F.conv_transpose1d (audioldm/hifigan/models.py:59,61,102,105,114),
 * plus the elementwise ops fused as epilogue (resnet.py:573-576,595; attention.py:305-332,432;
 * hifigan/models.py:58-62,104,108-115).
 */
typedef struct {
  /* ---- A operand (16-bit, channels-last) */
  const void* a;
  int32_t a_mode;     /* ctta_a_mode */
  int32_t ab_dtype;   /* CTTA_F16 or CTTA_BF16 (A and W) */
  int32_t c;          /* channels (K per tap) actually present in A */
  int32_t a_ld;       /* elements between consecutive rows / pixels / time steps of A (>= c, multiple of 8) */
  int32_t n_img;      /* ROWS: 1;  CONV1D: B;  CONV2D: N */
  int32_t h, w;       /* CONV2D: H, W;  CONV1D: h = 1, w = T (input length);  ROWS: h = 1, w = M */
  int32_t rows_per_img; /* logical GEMM rows per image: ROWS: M; CONV1D: number of output positions computed
                           (before the output mapping below); CONV2D: H * W */
  /* ---- taps */
  int32_t ntaps;
  int16_t tap_d0[CTTA_MAX_TAPS]; /* shift along w / t */
  int16_t tap_d1[CTTA_MAX_TAPS]; /* shift along h */
  /* ---- W operand: [n, ntaps * c_pad] row-major 16-bit, c_pad = c rounded up to 64, zero padded */
  const void* wgt;
  int32_t n;          /* output channels (GEMM N, before GEGLU halving) */
  /* ---- epilogue */
  const float* bias;      /* [n] or NULL */
  const float* rowadd;    /* [n_img_rowadd, rowadd_ld] fp32 or NULL; indexed by img(m) = m / rowadd_rows */
  int32_t rowadd_ld;
  int32_t rowadd_rows;    /* logical rows sharing one rowadd row (e.g. H*W for a time embedding) */
  int32_t act;            /* ctta_act */
  float act_slope;        /* LRELU slope */
  const void* residual;   /* [*, res_ld] or NULL */
  int32_t res_dtype;      /* CTTA_F32 / CTTA_F16 / CTTA_BF16 */
  int32_t res_ld;
  int32_t accumulate;     /* out += v (TMA reduce-add into global memory); excludes out2 */
  float out_scale;
  void* out;              /* may be NULL when only out2 is wanted */
  int32_t out_dtype;
  int32_t out_ld;
  void* out2;             /* optional 16-bit second output (dtype = ab_dtype) */
  int32_t out2_ld;
  int32_t act2;           /* CTTA_ACT_NONE / SILU / LRELU applied to v for out2 */
  float act2_slope;
  /* ---- output row mapping: logical row r of image i goes to row  i * out_rows_per_img + r * out_stride + out_off,
   *      and is dropped unless 0 <= r * out_stride + out_off < out_rows_per_img (transposed-conv phases). */
  int32_t out_rows_per_img;
  int32_t out_stride;
  int32_t out_off;
  /* ---- optional fused GroupNorm moments of the result v (the value written to out): stats[img, g, 0] += sum v,
   *      stats[img, g, 1] += sum v^2 over the n / stats_groups channels of group g and all rows of image img
   *      (the statistics pass of the FOLLOWING F.group_norm: resnet.py:555,581, modules.py:38-41).  Zeroed here.
   *      Needs the TMA epilogue (n >= 32, aligned pitches), out_stride == 1, channels per group a power of two >= 4;
   *      ROWS mode: image = row / stats_rows_per_img (a multiple of 32).  Returns CTTA_ERR_UNSUPPORTED otherwise. */
  float* stats;
  int32_t stats_groups;
  int32_t stats_rows_per_img;
  /* ---- 16-bit residual given as its LeakyReLU'ed copy: negative residual values are multiplied by res_neg_scale
   *      (= 1 / slope) before the add, 0 means 1.  HiFi-GAN's ResBlock x + c2(...) (hifigan/models.py:56-63) reads the
   *      same 16-bit tensor the previous conv consumed as its operand instead of a separate fp32 residual stream. */
  float res_neg_scale;
  /* ---- batched GEMM: != 0 gives image i its own weight matrix at wgt + i * wgt_img_stride elements (CONV1D mode, one
   *      tap, c % 64 == 0): attention scores q.k^T and P.V of the VAE AttnBlock (modules.py:216-225) for all samples. */
  int64_t wgt_img_stride;
  /* ---- nearest-2x-upsample + conv fused as four phase convolutions on the LOW-resolution input (resnet.py:146-157,
   *      modules.py:53-57): out_up_phase = 1 + 2 * ph + pw writes logical pixel (h, w) of image i to output pixel
   *      (2h + ph, 2w + pw) of an [n_img, 2h, 2w, out_ld] tensor (CONV2D mode, fp32 out, TMA epilogue); 0 = off.
   *      The 3x3 kernel collapses to 2x2 taps per phase (ops.pack_upsample2x_conv2d): 4/9 of the FLOPs and no
   *      materialised upsampled operand.  stats_keep != 0 accumulates into `stats` without zeroing it first. */
  int32_t out_up_phase;
  int32_t stats_keep;
} ctta_gemm_desc;

int ctta_gemm(const ctta_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Bandwidth-bound kernels (vectorised, coalesced, warp-shuffle reductions).
 * ---------------------------------------------------------------------------------------------------------- */

/* GroupNorm raw moments over channels-last input x[n_img, hw, c] (row stride ld): (sum x, sum x^2) per
 * (image, group).  Optional second source x2 (c2 channels, row stride ld2) is the channel-concatenated skip tensor
 * (torch.cat at diffusers/models/unet_2d_blocks.py:2032,2134).  stats = float[n_img, groups, 2] (zeroed here).
 * Reference: F.group_norm at resnet.py:555,581; transformer_2d.py:259; unet_2d_condition_guided.py:938;
 * audioldm/variational_autoencoder/modules.py:38-41. */
int ctta_groupnorm_stats(const void* x, int32_t x_dtype, int32_t c, int32_t ld, const void* x2, int32_t c2,
                         int32_t ld2, int32_t n_img, int32_t hw, int32_t groups, float* stats, void* stream);

/* y = act((x - mean) * rsqrt(var + eps) * gamma + beta), mean/var from the raw moments, written 16-bit, channels-last, optionally nearest-upsampled 2x
 * (F.interpolate at resnet.py:146, modules.py:54) — the operand of the following conv/linear.
 * With stats == NULL the normalisation is skipped (plain cast / concat / upsample).  act: NONE or SILU.
 * Optional raw 16-bit copy of the (concatenated) input into raw_out (operand of the 1x1 shortcut conv). */
int ctta_groupnorm_apply(const void* x, int32_t x_dtype, int32_t c, int32_t ld, const void* x2, int32_t c2,
                         int32_t ld2, int32_t n_img, int32_t h, int32_t w, int32_t groups, const float* stats,
                         const float* gamma, const float* beta, float eps, int32_t act, int32_t upsample2x,
                         void* y, int32_t y_dtype, int32_t y_ld, void* raw_out, int32_t raw_ld, void* stream);

/* LayerNorm over the first d of ld columns of fp32 x[m, ld]; writes 16-bit y[m, y_ld] (pad columns zeroed).
 * Reference: F.layer_norm at diffusers/models/attention.py:293,309,322. */
int ctta_layernorm(const float* x, int32_t m, int32_t d, int32_t ld, const float* gamma, const float* beta,
                   float eps, void* y, int32_t y_dtype, int32_t y_ld, void* stream);

/* Fused flash-style attention, head_dim padded to 64: q[b, lq, heads, 64] etc. given as element strides.
 * kv_len[b] (may be NULL) is the number of valid keys per batch row — the reference's additive -10000 bias on
 * padded text tokens (unet_2d_condition_guided.py:793-795) underflows to exactly 0 weight in fp32.
 * Reference: F.scaled_dot_product_attention at attention_processor.py:1127-1129; bmm+softmax at
 * audioldm/variational_autoencoder/modules.py:216-225 (heads = 1, d = 512 handled as 8 x 64 chunks). */
int ctta_attention(const void* q, const void* k, const void* v, void* o, int32_t dtype, int32_t batch,
                   int32_t heads, int32_t lq, int32_t lk, int32_t head_dim, int64_t q_bs, int64_t q_ls,
                   int64_t k_bs, int64_t k_ls, int64_t v_bs, int64_t v_ls, int64_t o_bs, int64_t o_ls,
                   const int32_t* kv_len, float scale, void* stream);

/* Debug only: with CTTA_ATTN_DEBUG set in the environment the tcgen05 attention kernel records clock64() at the phase
 * boundaries of one softmax warp (CTA 0, key tiles 8..15) into managed memory; this copies the 64 stamps to the host
 * (synchronises the device).  tools/attn_phases.py prints them.  Not used by the product path. */
int ctta_attention_debug(long long* host_out);

/* y[r, :] = softmax(scale * x[r, :]) for fp32 scores x[rows, ld] -> 16-bit probabilities (VAE AttnBlock,
 * audioldm/variational_autoencoder/modules.py:217-218; scores and P.V run through ctta_gemm). */
int ctta_softmax_rows(const float* x, int32_t rows, int32_t cols, int64_t ld, float scale, void* y, int32_t y_dtype,
                      int64_t y_ld, void* stream);

/* Gather for the three stride-2 3x3 convs (resnet.py:206): x[n, h, w, c] 16-bit -> a[n*(h/2)*(w/2), 9*c]. */
int ctta_im2col_s2(const void* x, int32_t n_img, int32_t h, int32_t w, int32_t c, void* a, void* stream);

/* NCHW fp32 <-> channels-last conversions at the module boundary. dst 16-bit or fp32.  y = x * scale * (*scale_dev):
 * `scale_dev` (device float, may be NULL = 1) carries the scheduler prologue `z_N = noise * init_noise_sigma` +
 * `scale_model_input` (easy_inference/consistencytta.py:160,173; scheduling_heun_discrete.py:151-172) as DATA, so one
 * captured CUDA graph serves every sigma of the multi-step sampler. */
int ctta_nchw_to_nhwc(const float* x, int32_t n_img, int32_t c, int32_t hw, void* y, int32_t y_dtype, int32_t y_ld,
                      float scale, const float* scale_dev, void* stream);
int ctta_nhwc_to_nchw(const float* x, int32_t n_img, int32_t c, int32_t hw, int32_t ld, float* y, void* stream);

/* Timestep + guidance embedding front end (embeddings.py:25-65,222-249; unet_2d_condition_guided.py:803-816):
 * t_feat[b, 256] = [cos(t f_i), sin(t f_i)], g_feat[b, 1024] = [cos(2 pi w W), sin(2 pi w W)]. */
int ctta_time_features(const float* t, const float* w, const float* gw, int32_t batch, float* t_feat, float* g_feat,
                       void* stream);
/* Small-M fp32 linear: y[m, n] = act_in(x)[m, k] @ W[n, k]^T + b (+ y if accumulate). act_in / act_out: NONE or SILU.
 * For the embedding MLPs and the 22 time_emb_proj GEMVs batched as one call (embeddings.py:193-198; resnet.py:573). */
int ctta_small_linear(const float* x, int32_t m, int32_t k, const float* wgt, const float* bias, int32_t n,
                      int32_t act_in, int32_t act_out, int32_t accumulate, float* y, void* stream);

/* Waveform post-processing (audioldm/hifigan/utilities.py:84-86): batch-global min/max, centre, * 32768,
 * truncate toward zero to int16 (numpy astype semantics incl. wrap-around). minmax = float[2] scratch. */
int ctta_wave_minmax(const float* wav, int64_t numel, float* minmax, void* stream);
int ctta_wave_to_int16(const float* wav, int64_t numel, const float* minmax, int16_t* out, void* stream);

/* y = 16-bit(leaky_relu(x, slope)) elementwise: operand of the next up-sampling stage after the MRF sum
 * (audioldm/hifigan/models.py:104,112-113). */
int ctta_lrelu_cast(const float* x, int64_t numel, float slope, void* y, int32_t y_dtype, void* stream);

/* One (c1, c2) pair of a HiFi-GAN ResBlock fused in a single kernel (audioldm/hifigan/models.py:56-63):
 *     x' = x + c2( lrelu( c1( lrelu(x) ) ) ),  c1 = Conv1d(C, C, taps, dilation, "same"), c2 = Conv1d(C, C, taps, 1, "same")
 * on the LeakyReLU'ed 16-bit streams of the vocoder: x_lrelu = lrelu(x, slope) [batch, t, c] in, out = lrelu(x', slope)
 * [batch, t, c]; w1 / w2 are K-major packed weights [c, taps * 64] (taps of 64 zero-padded input channels each), b1 / b2
 * fp32 [c].  The hidden tensor lives only in shared memory.  ctta_resblock_pair_supported() tells whether a
 * (c, taps, dilation) fits (c in {32, 64}, odd taps <= 11, even t for c = 32, resident weights + tiles within 227 KiB); otherwise the two
 * convolutions run through ctta_gemm.  Returns CTTA_ERR_UNSUPPORTED for an unsupported combination.  x_lrelu is read
 * twice (halo'd TMA boxes for c1, and the residual rows through the read-only data path) and must not overlap out;
 * the result is bit-identical to the two ctta_gemm launches. */
int ctta_resblock_pair_supported(int32_t c, int32_t taps, int32_t dilation, int32_t t);
int ctta_resblock_pair(const void* x_lrelu, void* out, int32_t dtype, int32_t batch, int32_t t, int32_t c, const void* w1,
                       const float* b1, const void* w2, const float* b2, int32_t taps, int32_t dilation, float slope,
                       void* stream);

/* HiFi-GAN multi-receptive-field sum (audioldm/hifigan/models.py:106-112 followed by the LeakyReLU at :104 / :113):
 * the n_in (<= 4) ResBlock outputs arrive as their LeakyReLU'ed 16-bit copies (negative values scaled by in_slope);
 * y = 16-bit( lrelu( out_scale * sum_i lrelu^-1(x_i), out_slope ) ), the operand of the next up-sampling stage. */
int ctta_mrf_combine(const void* const* x, int32_t n_in, int64_t numel, int32_t dtype, float in_slope, float out_scale,
                     float out_slope, void* y, void* stream);

/* Tail of a single-output-channel convolution (AudioLDM VAE conv_out 128 -> 1, 3x3:
 * audioldm/variational_autoencoder/modules.py:680; HiFi-GAN conv_post 32 -> 1, k = 7, + tanh:
 * audioldm/hifigan/models.py:114-115).  The convolution is restated as a pointwise GEMM z[p, j] = x[p, :] . w[j, :]
 * (ctta_gemm, CTTA_A_ROWS, n = taps) followed by y[p] = act(bias + sum_j z[p + (d1_j, d0_j), j]) with zero padding
 * outside the h x w image.  z: fp32 [n_img * h * w, z_ld]; out fp32 and / or out16 (16-bit, out16_dtype) [n_img*h*w]. */
int ctta_tap_sum(const float* z, int32_t z_ld, int32_t n_img, int32_t h, int32_t w, int32_t ntaps,
                 const int16_t* tap_d0, const int16_t* tap_d1, const float* bias, int32_t act, float* out, void* out16,
                 int32_t out16_dtype, void* stream);

/* (1 - s) * uncond + s * cond on the two batch halves (models/audio_consistency_model.py:453-456). */
int ctta_cfg_mix(const float* x, int64_t half_numel, float s, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTTA_H_ */
