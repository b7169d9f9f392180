// Fused flash-style attention for head_dim 64 (the UNet's 51-wide heads zero-padded to 64) + a row softmax used
// by the VAE's single-head 512-wide attention (scores and P.V run on the tcgen05 GEMM).
//
// CTA = 4 warps x 16 query rows (Br = 64), key/value tiles of 64 rows double-buffered with cp.async, XOR-swizzled
// shared memory, ldmatrix fragments, online softmax in fp32 with exp2f.  Keys >= kv_len[b] are masked, which is
// what the reference's additive -10000 bias on padded text tokens evaluates to in fp32.
// Reference call sites: F.scaled_dot_product_attention (diffusers/models/attention_processor.py:1127-1129);
// softmax at audioldm/variational_autoencoder/modules.py:218.
#include "ctta_internal.h"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ctta {

constexpr int kBr = 64, kBc = 64, kD = 64;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <bool BF>
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  if (BF) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <bool BF>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// byte offset of (row, 16-byte chunk) inside a [rows][64 x 16-bit] tile with the XOR swizzle
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// cooperative load of a 64 x 64 16-bit tile (rows >= valid_rows are zero-filled)
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const unsigned short* g, long long row_stride,
                                          int valid_rows, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, chunk = idx & 7;
    const bool ok = row < valid_rows;
    const unsigned short* src = g + (ok ? static_cast<long long>(row) * row_stride + chunk * 8 : 0);
    cp_async16(smem_tile + sw_off(row, chunk), src, ok ? 16 : 0);
  }
}

template <bool BF>
__global__ void __launch_bounds__(128) flash_attn_kernel(const unsigned short* __restrict__ q,
                                                         const unsigned short* __restrict__ k,
                                                         const unsigned short* __restrict__ v,
                                                         unsigned short* __restrict__ o, int lq, int lk, long long q_bs,
                                                         long long q_ls, long long k_bs, long long k_ls, long long v_bs,
                                                         long long v_ls, long long o_bs, long long o_ls,
                                                         const int* __restrict__ kv_len, float scale_log2) {
  __shared__ __align__(128) unsigned char smem[(kBr + 4 * kBc) * kD * 2];  // Q, K0, K1, V0, V1 = 40 KiB
  const uint32_t sQ = s_u32(smem);
  const uint32_t sK = sQ + kBr * kD * 2;
  const uint32_t sV = sK + 2 * kBc * kD * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kBr;

  int kvl = kv_len ? kv_len[b] : lk;
  kvl = max(1, min(kvl, lk));
  const int n_tiles = (kvl + kBc - 1) / kBc;

  const unsigned short* qg = q + b * q_bs + static_cast<long long>(q0) * q_ls + head * kD;
  const unsigned short* kg = k + b * k_bs + head * kD;
  const unsigned short* vg = v + b * v_bs + head * kD;

  load_tile(sQ, qg, q_ls, lq - q0, tid);
  load_tile(sK, kg, k_ls, kvl, tid);
  load_tile(sV, vg, v_ls, kvl, tid);
  cp_async_commit();

  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      const int k1 = (t + 1) * kBc;
      load_tile(sK + (buf ^ 1) * kBc * kD * 2, kg + static_cast<long long>(k1) * k_ls, k_ls, kvl - k1, tid);
      load_tile(sV + (buf ^ 1) * kBc * kD * 2, vg + static_cast<long long>(k1) * v_ls, v_ls, kvl - k1, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int m = lane >> 3;
        const int row = warp * 16 + (m & 1) * 8 + (lane & 7);
        const int chunk = ks * 2 + (m >> 1);
        ldsm_x4(sQ + sw_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t sKb = sK + buf * kBc * kD * 2, sVb = sV + buf * kBc * kD * 2;

    // ---- S = Q K^T
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        const int m = lane >> 3;
        const int row = nb * 8 + (m >> 1) * 8 + (lane & 7);
        const int chunk = ks * 2 + (m & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sKb + sw_off(row, chunk), b0, b1, b2, b3);
        mma16816<BF>(s[nb], qf[ks], b0, b1);
        mma16816<BF>(s[nb + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask the tail of the last tile
    const int k0 = t * kBc;
    if (k0 + kBc > kvl) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = k0 + nb * 8 + (lane & 3) * 2;
        if (col >= kvl) s[nb][0] = s[nb][2] = -INFINITY;
        if (col + 1 >= kvl) s[nb][1] = s[nb][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g = lane/4 and g + 8)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(m_run[r], mx[r]);  // finite: every tile has at least one unmasked key
      corr[r] = exp2f((m_run[r] - mnew[r]) * scale_log2);
      m_run[r] = mnew[r];
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = exp2f((s[nb][0] - mnew[0]) * scale_log2);
      s[nb][1] = exp2f((s[nb][1] - mnew[0]) * scale_log2);
      s[nb][2] = exp2f((s[nb][2] - mnew[1]) * scale_log2);
      s[nb][3] = exp2f((s[nb][3] - mnew[1]) * scale_log2);
      l_run[0] += s[nb][0] + s[nb][1];
      l_run[1] += s[nb][2] + s[nb][3];
      o_acc[nb][0] *= corr[0];
      o_acc[nb][1] *= corr[0];
      o_acc[nb][2] *= corr[1];
      o_acc[nb][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2<BF>(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack2<BF>(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack2<BF>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack2<BF>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        const int m = lane >> 3;
        const int row = kk * 16 + (m & 1) * 8 + (lane & 7);
        const int chunk = nb + (m >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sVb + sw_off(row, chunk), b0, b1, b2, b3);
        mma16816<BF>(o_acc[nb], pa, b0, b1);
        mma16816<BF>(o_acc[nb + 1], pa, b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // ---- finalise: divide by the row sums, write 16-bit
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    l_run[r] = 1.f / l_run[r];
  }
  const int g = lane >> 2, tq = lane & 3;
  const int row0 = q0 + warp * 16 + g;
  unsigned short* og = o + b * o_bs + head * kD;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int col = nb * 8 + tq * 2;
    if (row0 < lq)
      *reinterpret_cast<uint32_t*>(og + static_cast<long long>(row0) * o_ls + col) =
          pack2<BF>(o_acc[nb][0] * l_run[0], o_acc[nb][1] * l_run[0]);
    if (row0 + 8 < lq)
      *reinterpret_cast<uint32_t*>(og + static_cast<long long>(row0 + 8) * o_ls + col) =
          pack2<BF>(o_acc[nb][2] * l_run[1], o_acc[nb][3] * l_run[1]);
  }
}

// ---------------------------------------------------------------------------------------------- row softmax
// y[r, :] = softmax(scale * x[r, :]) for fp32 x; one CTA per row, 16-bit output.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, int cols, long long ld,
                                                           float scale_log2, unsigned short* __restrict__ y,
                                                           long long y_ld, int is_bf16) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* row = x + blockIdx.x * ld;
  unsigned short* out = y + blockIdx.x * y_ld;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float mx = -INFINITY;
  for (int c = tid * 4; c < cols; c += 1024) {
    const float4 f = *reinterpret_cast<const float4*>(row + c);
    mx = fmaxf(mx, fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int c = tid * 4; c < cols; c += 1024) {
    const float4 f = *reinterpret_cast<const float4*>(row + c);
    sum += exp2f((f.x - mx) * scale_log2) + exp2f((f.y - mx) * scale_log2) + exp2f((f.z - mx) * scale_log2) +
           exp2f((f.w - mx) * scale_log2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bcast = 1.f / s;
  }
  __syncthreads();
  const float inv = bcast;
  for (int c = tid * 4; c < cols; c += 1024) {
    const float4 f = *reinterpret_cast<const float4*>(row + c);
    const float p0 = exp2f((f.x - mx) * scale_log2) * inv, p1 = exp2f((f.y - mx) * scale_log2) * inv;
    const float p2 = exp2f((f.z - mx) * scale_log2) * inv, p3 = exp2f((f.w - mx) * scale_log2) * inv;
    uint2 u;
    if (is_bf16) {
      u.x = pack2<true>(p0, p1);
      u.y = pack2<true>(p2, p3);
    } else {
      u.x = pack2<false>(p0, p1);
      u.y = pack2<false>(p2, p3);
    }
    *reinterpret_cast<uint2*>(out + c) = u;
  }
}

// attention_tc.cu: tcgen05 kernel for long key sequences; returns 1 when the problem is not eligible
int attention_tc_launch(const void* q, const void* k, const void* v, void* o, int dtype, int batch, int heads, int lq,
                        int lk, long long q_bs, long long q_ls, long long k_bs, long long k_ls, long long v_bs,
                        long long v_ls, long long o_bs, long long o_ls, const int* kv_len, float scale,
                        cudaStream_t stream);

}  // namespace ctta

using namespace ctta;

extern "C" int ctta_attention(const void* q, const void* k, const void* v, void* o, int32_t dtype, int32_t batch,
                              int32_t heads, int32_t lq, int32_t lk, int32_t head_dim, int64_t q_bs, int64_t q_ls,
                              int64_t k_bs, int64_t k_ls, int64_t v_bs, int64_t v_ls, int64_t o_bs, int64_t o_ls,
                              const int32_t* kv_len, float scale, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(q && k && v && o, "attention: null tensor");
  CTTA_REQUIRE(head_dim == kD, "attention: head_dim must be 64 (pad 51 -> 64), got %d", head_dim);
  CTTA_REQUIRE(dtype == CTTA_F16 || dtype == CTTA_BF16, "attention: 16-bit tensors only");
  CTTA_REQUIRE(batch > 0 && heads > 0 && lq > 0 && lk > 0, "attention: empty problem");
  CTTA_REQUIRE(q_ls % 8 == 0 && k_ls % 8 == 0 && v_ls % 8 == 0 && o_ls % 2 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 &&
                   v_bs % 8 == 0 && o_bs % 2 == 0,
               "attention: strides must keep 16-byte alignment");
  CTTA_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(o) & 3) == 0,
               "attention: tensors must be 16-byte aligned");
  CTTA_REQUIRE(batch <= 65535 && heads <= 65535, "attention: batch / heads exceed grid limits");
  {
    // self-attention sized problems run on the tcgen05 kernel; short key sequences (text cross-attention, the
    // 64-token mid block) stay on the mma.sync kernel below, where a 128-key tile would be mostly padding
    const int rc = attention_tc_launch(q, k, v, o, dtype, batch, heads, lq, lk, q_bs, q_ls, k_bs, k_ls, v_bs, v_ls, o_bs,
                                       o_ls, kv_len, scale, stream);
    if (rc <= 0) return rc;
  }
  dim3 grid((lq + kBr - 1) / kBr, heads, batch);
  const float scale_log2 = scale * 1.4426950408889634f;
  if (dtype == CTTA_BF16) {
    flash_attn_kernel<true><<<grid, 128, 0, stream>>>(
        reinterpret_cast<const unsigned short*>(q), reinterpret_cast<const unsigned short*>(k),
        reinterpret_cast<const unsigned short*>(v), reinterpret_cast<unsigned short*>(o), lq, lk, q_bs, q_ls, k_bs,
        k_ls, v_bs, v_ls, o_bs, o_ls, kv_len, scale_log2);
  } else {
    flash_attn_kernel<false><<<grid, 128, 0, stream>>>(
        reinterpret_cast<const unsigned short*>(q), reinterpret_cast<const unsigned short*>(k),
        reinterpret_cast<const unsigned short*>(v), reinterpret_cast<unsigned short*>(o), lq, lk, q_bs, q_ls, k_bs,
        k_ls, v_bs, v_ls, o_bs, o_ls, kv_len, scale_log2);
  }
  CTTA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ctta_softmax_rows(const float* x, int32_t rows, int32_t cols, int64_t ld, float scale, void* y,
                                 int32_t y_dtype, int64_t y_ld, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  CTTA_REQUIRE(x && y && rows > 0 && cols > 0, "softmax_rows: bad arguments");
  CTTA_REQUIRE(cols % 4 == 0 && ld % 4 == 0 && y_ld % 4 == 0, "softmax_rows: cols and strides must be multiples of 4");
  CTTA_REQUIRE(y_dtype == CTTA_F16 || y_dtype == CTTA_BF16, "softmax_rows: 16-bit output only");
  softmax_rows_kernel<<<rows, 256, 0, stream>>>(x, cols, ld, scale * 1.4426950408889634f,
                                                reinterpret_cast<unsigned short*>(y), y_ld, y_dtype == CTTA_BF16);
  CTTA_LAUNCH_CHECK();
  return 0;
}
