// Gated DeltaNet (Qwen3-Next linear attention) prefill kernels — fp32 CUDA-core version 1.
//
// Restates python/krasis/linear_attention.py:695-844 (_forward_chunked) + :593-693 (_chunked_inner):
//   gdn_prep_kernel        un-interleave qkvz/ba (:337-391), causal depthwise conv(4)+SiLU (:722-741),
//                          l2norm(q), l2norm(k), q*dk^-1/2 (:114-117,164,765-770), beta=sigmoid(b),
//                          g=-exp(A_log)*softplus(a+dt_bias) (:752-755)  — BF16 rounding points as torch produces them
//   gdn_conv_state_kernel  conv state = last 4 pre-conv inputs (:725-728)
//   gdn_chunk_prepare      per (head, chunk of 64): gcum, decay, A=-(K.beta K^T).decay strictly lower,
//                          (I-A)^-1 applied to [V.beta | K.beta.e^gcum] by forward substitution (:634-646), and the
//                          masked intra-chunk matrix (Q K^T).decay (:43-47)
//   gdn_chunk_scan         sequential pass over chunks per (head, 32-wide slice of dv): the _chunk_step recurrence (:43-60)
//   gdn_post_kernel        gated RMSNorm (:987-1004)
// The in/out projections are dense_gemm_kernel (tcgen05).  All delta-rule math is fp32 like the reference.
#include <cstdlib>

#include "moe_common.cuh"
#include "prof.cuh"
#include "mma_sync.cuh"
#include "ptx.cuh"

namespace kb2 {

constexpr int kGC = 64;   // chunk size (linear_attention.py:702)

struct GdnDims {
  int H, nk, nv, dk, dv, K;   // K = conv kernel size (4)
  float eps, scale;
  int ba_ld;                  // row stride of the ba projection output (2*nv rounded up to 16)
};

__device__ __forceinline__ float bf16r(float x) { return bf16_round_rn(x); }

// column of mixed_qkvz (per key-head group [q dk | k dk | v r*dv | z r*dv]) holding conv channel c
__device__ __forceinline__ int qkvz_col(const GdnDims& d, int c) {
  const int r = d.nv / d.nk, G = 2 * d.dk + 2 * r * d.dv, kd = d.nk * d.dk;
  if (c < kd) return (c / d.dk) * G + (c % d.dk);
  if (c < 2 * kd) {
    const int cc = c - kd;
    return (cc / d.dk) * G + d.dk + (cc % d.dk);
  }
  const int cc = c - 2 * kd, vh = cc / d.dv, kh = vh / r;
  return kh * G + 2 * d.dk + (vh % r) * d.dv + (cc % d.dv);
}

// one CTA per token
__global__ void __launch_bounds__(256) gdn_prep_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qkvz,
                                                       const __nv_bfloat16* __restrict__ ba,
                                                       const __nv_bfloat16* __restrict__ conv_w,      // [C][K]
                                                       const __nv_bfloat16* __restrict__ conv_state,  // [C][K]
                                                       const float* __restrict__ A_log, const float* __restrict__ dt_bias,
                                                       __nv_bfloat16* __restrict__ qn, __nv_bfloat16* __restrict__ kn,
                                                       __nv_bfloat16* __restrict__ vc, float* __restrict__ beta,
                                                       float* __restrict__ g, int M) {
  extern __shared__ float s_conv[];   // [C]
  const int t = blockIdx.x;
  const int kd = d.nk * d.dk, vd = d.nv * d.dv, C = 2 * kd + vd;
  const int ld = 2 * kd + 2 * vd;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int col = qkvz_col(d, c);
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < d.K; ++j) {
      const int tt = t - (d.K - 1) + j;          // input time index; negative -> conv state
      const float x = tt >= 0 ? __bfloat162float(qkvz[(long long)tt * ld + col])
                              : __bfloat162float(conv_state[c * d.K + (d.K + tt)]);
      acc = fmaf(__bfloat162float(conv_w[c * d.K + j]), x, acc);
    }
    const float y = bf16r(acc);                   // conv output in BF16
    s_conv[c] = bf16r(y / (1.0f + expf(-y)));     // SiLU, BF16
  }
  __syncthreads();
  // v: straight copy
  for (int c = threadIdx.x; c < vd; c += blockDim.x) vc[(long long)t * vd + c] = __float2bfloat16_rn(s_conv[2 * kd + c]);
  // q, k: l2norm per head with torch's BF16 elementwise rounding (x*x -> bf16, sum fp32 -> bf16, +eps, rsqrt, x*inv)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int hh = warp; hh < 2 * d.nk; hh += nwarps) {
    const float* src = s_conv + hh * d.dk;        // q heads then k heads are contiguous in the conv channel order
    float s = 0.f;
    for (int i = lane; i < d.dk; i += 32) s += bf16r(src[i] * src[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = bf16r(rsqrtf(bf16r(bf16r(s) + 1e-6f)));
    const bool is_q = hh < d.nk;
    __nv_bfloat16* dst = (is_q ? qn : kn) + (long long)t * kd + (is_q ? hh : hh - d.nk) * d.dk;
    for (int i = lane; i < d.dk; i += 32) {
      float v = bf16r(src[i] * inv);
      if (is_q) v = bf16r(v * d.scale);
      dst[i] = __float2bfloat16_rn(v);
    }
  }
  // gates
  const int r = d.nv / d.nk;
  for (int h = threadIdx.x; h < d.nv; h += blockDim.x) {
    const int kh = h / r, j = h % r;
    const float b = __bfloat162float(ba[(long long)t * d.ba_ld + kh * 2 * r + j]);
    const float a = __bfloat162float(ba[(long long)t * d.ba_ld + kh * 2 * r + r + j]);
    beta[(long long)t * d.nv + h] = bf16r(1.0f / (1.0f + expf(-b)));          // sigmoid on a BF16 tensor
    const float x = a + dt_bias[h];
    const float sp = x > 20.f ? x : log1pf(expf(x));                           // F.softplus (threshold 20)
    g[(long long)t * d.nv + h] = -expf(A_log[h]) * sp;
  }
}

// Tiled variant for dk == dv == 32*CPL, conv width KW: one warp per (32-token tile, head) slides down the tokens with the
// last KW-1 inputs in registers, so every qkvz element is read once (the per-token kernel above reads it KW times) and the
// l2norm needs no shared memory.  Same arithmetic, same rounding points, same summation order as gdn_prep_kernel.
template <int CPL, int KW>
__global__ void __launch_bounds__(256) gdn_prep_tiled_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qkvz,
                                                             const __nv_bfloat16* __restrict__ ba,
                                                             const __nv_bfloat16* __restrict__ conv_w,
                                                             const __nv_bfloat16* __restrict__ conv_state,
                                                             const float* __restrict__ A_log,
                                                             const float* __restrict__ dt_bias,
                                                             __nv_bfloat16* __restrict__ qn, __nv_bfloat16* __restrict__ kn,
                                                             __nv_bfloat16* __restrict__ vc, float* __restrict__ beta,
                                                             float* __restrict__ g, int M, int n_tiles) {
  constexpr int TT = 32;
  const int lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n_units = 2 * d.nk + d.nv + 1;
  const int tile = (int)(wg / n_units), u = (int)(wg % n_units);
  if (tile >= n_tiles) return;
  const int t_begin = tile * TT, t_end = min(M, t_begin + TT);
  const int kd = d.nk * d.dk, vd = d.nv * d.dv, ld = 2 * kd + 2 * vd, r = d.nv / d.nk;
  if (u == 2 * d.nk + d.nv) {             // gates for this token tile
    for (int t = t_begin; t < t_end; ++t)
      for (int h = lane; h < d.nv; h += 32) {
        const int kh = h / r, j = h % r;
        const float b = __bfloat162float(ba[(long long)t * d.ba_ld + kh * 2 * r + j]);
        const float a = __bfloat162float(ba[(long long)t * d.ba_ld + kh * 2 * r + r + j]);
        beta[(long long)t * d.nv + h] = bf16r(1.0f / (1.0f + expf(-b)));
        const float x = a + dt_bias[h];
        const float sp = x > 20.f ? x : log1pf(expf(x));
        g[(long long)t * d.nv + h] = -expf(A_log[h]) * sp;
      }
    return;
  }
  const bool is_q = u < d.nk, is_k = !is_q && u < 2 * d.nk;
  const int c_base = is_q ? u * d.dk : (is_k ? kd + (u - d.nk) * d.dk : 2 * kd + (u - 2 * d.nk) * d.dv);
  int col[CPL];
  float w[CPL][KW], hist[CPL][KW - 1];
#pragma unroll
  for (int e = 0; e < CPL; ++e) {
    const int c = c_base + lane + 32 * e;
    col[e] = qkvz_col(d, c);
#pragma unroll
    for (int j = 0; j < KW; ++j) w[e][j] = __bfloat162float(conv_w[c * KW + j]);
#pragma unroll
    for (int j = 0; j < KW - 1; ++j) {
      const int tt = t_begin - (KW - 1) + j;
      hist[e][j] = tt >= 0 ? __bfloat162float(qkvz[(long long)tt * ld + col[e]])
                           : __bfloat162float(conv_state[c * KW + (KW + tt)]);
    }
  }
  for (int t = t_begin; t < t_end; ++t) {
    float sv[CPL];
#pragma unroll
    for (int e = 0; e < CPL; ++e) {
      const float x = __bfloat162float(qkvz[(long long)t * ld + col[e]]);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < KW - 1; ++j) acc = fmaf(w[e][j], hist[e][j], acc);
      acc = fmaf(w[e][KW - 1], x, acc);
#pragma unroll
      for (int j = 0; j < KW - 2; ++j) hist[e][j] = hist[e][j + 1];
      hist[e][KW - 2] = x;
      const float y = bf16r(acc);
      sv[e] = bf16r(y / (1.0f + expf(-y)));
    }
    if (!is_q && !is_k) {
#pragma unroll
      for (int e = 0; e < CPL; ++e)
        vc[(long long)t * vd + (u - 2 * d.nk) * d.dv + lane + 32 * e] = __float2bfloat16_rn(sv[e]);
      continue;
    }
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < CPL; ++e) ss += bf16r(sv[e] * sv[e]);
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = bf16r(rsqrtf(bf16r(bf16r(ss) + 1e-6f)));
    __nv_bfloat16* dst = (is_q ? qn : kn) + (long long)t * kd + (is_q ? u : u - d.nk) * d.dk;
#pragma unroll
    for (int e = 0; e < CPL; ++e) {
      float v = bf16r(sv[e] * inv);
      if (is_q) v = bf16r(v * d.scale);
      dst[lane + 32 * e] = __float2bfloat16_rn(v);
    }
  }
}


// Vectorised variant (dk == dv == 128, conv width 4): one warp per (token tile, head); every lane owns FOUR CONSECUTIVE
// channels of the head, so a token row is one 8-byte load / store per lane (the kernel above issues four 2-byte ones) and
// the conv window lives in registers.  SiLU uses ex2-based __expf / __fdividef: the f32 result differs from expf by <= 2 ulp,
// which moves a BF16 rounding only when the value sits within 2^-22 of a rounding boundary (p ~ 6e-5 per element).
// Rounding points are those of gdn_prep_kernel; the l2norm partial sums are grouped per lane differently (f32 addition order).
template <int TT>
__global__ void __launch_bounds__(256) gdn_prep_vec_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qkvz,
                                                           const __nv_bfloat16* __restrict__ ba,
                                                           const __nv_bfloat16* __restrict__ conv_w,
                                                           const __nv_bfloat16* __restrict__ conv_state,
                                                           const float* __restrict__ A_log, const float* __restrict__ dt_bias,
                                                           __nv_bfloat16* __restrict__ qn, __nv_bfloat16* __restrict__ kn,
                                                           __nv_bfloat16* __restrict__ vc, float* __restrict__ beta,
                                                           float* __restrict__ g, int M, int n_tiles) {
  constexpr int KW = 4;
  const int lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n_units = 2 * d.nk + d.nv + 1;
  const int tile = (int)(wg / n_units), u = (int)(wg % n_units);
  if (tile >= n_tiles) return;
  const int t_begin = tile * TT, t_end = min(M, t_begin + TT);
  const int kd = d.nk * d.dk, vd = d.nv * d.dv, ld = 2 * kd + 2 * vd, r = d.nv / d.nk;
  if (u == 2 * d.nk + d.nv) {             // gates for this token tile
    for (int t = t_begin; t < t_end; ++t)
      for (int h = lane; h < d.nv; h += 32) {
        const int kh = h / r, j = h % r;
        const float b = __bfloat162float(ba[(long long)t * d.ba_ld + kh * 2 * r + j]);
        const float a = __bfloat162float(ba[(long long)t * d.ba_ld + kh * 2 * r + r + j]);
        beta[(long long)t * d.nv + h] = bf16r(1.0f / (1.0f + expf(-b)));
        const float x = a + dt_bias[h];
        const float sp = x > 20.f ? x : log1pf(expf(x));
        g[(long long)t * d.nv + h] = -expf(A_log[h]) * sp;
      }
    return;
  }
  const bool is_q = u < d.nk, is_k = !is_q && u < 2 * d.nk;
  const int c0 = (is_q ? u * d.dk : (is_k ? kd + (u - d.nk) * d.dk : 2 * kd + (u - 2 * d.nk) * d.dv)) + lane * 4;   // conv channel
  const int col0 = qkvz_col(d, c0);                                 // the 4 channels are consecutive columns of qkvz
  float w[4][KW], hist[4][KW - 1];
  {
    const uint4 w01 = *reinterpret_cast<const uint4*>(conv_w + (long long)c0 * KW);          // 16 bf16 = 4 channels x 4 taps
    const uint4 w23 = *reinterpret_cast<const uint4*>(conv_w + (long long)c0 * KW + 8);
    const uint32_t ww[8] = {w01.x, w01.y, w01.z, w01.w, w23.x, w23.y, w23.z, w23.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      w[e][0] = __uint_as_float(ww[2 * e] << 16); w[e][1] = __uint_as_float(ww[2 * e] & 0xFFFF0000u);
      w[e][2] = __uint_as_float(ww[2 * e + 1] << 16); w[e][3] = __uint_as_float(ww[2 * e + 1] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int j = 0; j < KW - 1; ++j) {
    const int tt = t_begin - (KW - 1) + j;
    if (tt >= 0) {
      const uint2 x = *reinterpret_cast<const uint2*>(qkvz + (long long)tt * ld + col0);
      hist[0][j] = __uint_as_float(x.x << 16); hist[1][j] = __uint_as_float(x.x & 0xFFFF0000u);
      hist[2][j] = __uint_as_float(x.y << 16); hist[3][j] = __uint_as_float(x.y & 0xFFFF0000u);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) hist[e][j] = __bfloat162float(conv_state[(long long)(c0 + e) * KW + (KW + tt)]);
    }
  }
  __nv_bfloat16* dst = is_q ? qn + (long long)u * d.dk : (is_k ? kn + (long long)(u - d.nk) * d.dk : vc + (long long)(u - 2 * d.nk) * d.dv);
  const int dst_ld = (is_q || is_k) ? kd : vd;
  // rows are consumed strictly in order (the convolution history lives in registers), so the only memory-level parallelism a warp
  // has is how far ahead it loads: four rows (1 KB per warp) in flight instead of one took this kernel off the latency bound
  constexpr int kAhead = 4;
  uint2 ring[kAhead];
#pragma unroll
  for (int j = 0; j < kAhead; ++j)
    ring[j] = t_begin + j < t_end ? *reinterpret_cast<const uint2*>(qkvz + (long long)(t_begin + j) * ld + col0) : make_uint2(0u, 0u);
  for (int t0 = t_begin; t0 < t_end; t0 += kAhead) {
#pragma unroll
   for (int jr = 0; jr < kAhead; ++jr) {
    const int t = t0 + jr;
    if (t >= t_end) break;                                   // warp-uniform
    const uint2 cur = ring[jr];
    if (t + kAhead < t_end) ring[jr] = *reinterpret_cast<const uint2*>(qkvz + (long long)(t + kAhead) * ld + col0);
    const float x[4] = {__uint_as_float(cur.x << 16), __uint_as_float(cur.x & 0xFFFF0000u), __uint_as_float(cur.y << 16),
                        __uint_as_float(cur.y & 0xFFFF0000u)};
    float sv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < KW - 1; ++j) acc = fmaf(w[e][j], hist[e][j], acc);
      acc = fmaf(w[e][KW - 1], x[e], acc);
#pragma unroll
      for (int j = 0; j < KW - 2; ++j) hist[e][j] = hist[e][j + 1];
      hist[e][KW - 2] = x[e];
      const float y = bf16r(acc);
      sv[e] = bf16r(__fdividef(y, 1.0f + __expf(-y)));
    }
    if (is_q || is_k) {
      float ss = (bf16r(sv[0] * sv[0]) + bf16r(sv[1] * sv[1])) + (bf16r(sv[2] * sv[2]) + bf16r(sv[3] * sv[3]));
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = bf16r(rsqrtf(bf16r(bf16r(ss) + 1e-6f)));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sv[e] = bf16r(sv[e] * inv);
        if (is_q) sv[e] = bf16r(sv[e] * d.scale);
      }
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(sv[0], sv[1]), hi = __floats2bfloat162_rn(sv[2], sv[3]);
    *reinterpret_cast<uint2*>(dst + (long long)t * dst_ld + lane * 4) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
   }
  }
}

// new conv state = last K pre-conv inputs (older entries come from the previous state when M < K)
__global__ void gdn_conv_state_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qkvz,
                                      __nv_bfloat16* __restrict__ conv_state, int M) {
  const int kd = d.nk * d.dk, vd = d.nv * d.dv, C = 2 * kd + vd, ld = 2 * kd + 2 * vd;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int col = qkvz_col(d, c);
  __nv_bfloat16 old[8], nw[8];
  for (int j = 0; j < d.K; ++j) old[j] = conv_state[c * d.K + j];
  for (int j = 0; j < d.K; ++j) {
    const int tt = M - d.K + j;
    nw[j] = tt >= 0 ? qkvz[(long long)tt * ld + col] : old[d.K + tt];
  }
  for (int j = 0; j < d.K; ++j) conv_state[c * d.K + j] = nw[j];
}


__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src_gmem, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
               "r"(valid ? 16 : 0)
               : "memory");
}
// ------------------------------------------------------------------------------------------------
// chunk prepare: grid (n_chunks, nv), 256 threads, 2 CTAs / SM (~100 KB smem each).  dk, dv <= 128, multiples of 16.
// outputs per (head, chunk): vcorr [64][dv], kcd [64][dk], intra [64][64], gcum [64]   (all fp32)
//   smem: sk bf16 [64][dk+8] | sA f32 [64][68] | sB f32 [64][dv+dk+8] (its head doubles as the q tile until Q K^T is done)
// ------------------------------------------------------------------------------------------------
constexpr int kLdA = kGC + 4;   // [row][k] fp32 operands: row stride = 4 mod 32 words -> conflict-free A fragments

// TC = true (dk == dv == 128): outputs in the operand layouts of gdn_scan_tc_kernel (gdn_tc.cu):
//   vcorr  [hc][dv/32][64][36] fp32 (one padded 9 KB slice per scan CTA)
//   kcd    [hc][hi c0 | hi c1 | lo c0 | lo c1] BF16 hi/lo pair, each an 8 KB K-major 128B-swizzled UMMA image (64 rows x 64 k)
//   intra  [hc][hi | lo] same image format (64 x 64)
__device__ __forceinline__ uint32_t prep_sw128_off(int row, int k) {       // element (row, k) of a 64-row K-major SW128 image pair
  const int kk = k & 63;
  return (uint32_t)((k >> 6) * 8192 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1)));
}
__device__ __forceinline__ void prep_split_bf16(float x, unsigned short& hi, unsigned short& lo) {
  const uint32_t h = bf16_bits_rn(x);                                     // packed converter, not the quarter-rate F2F (ptx.cuh)
  const uint32_t l = bf16_bits_rn(x - __uint_as_float(h << 16));
  hi = (unsigned short)h;
  lo = (unsigned short)l;
}

template <bool TC>
__global__ void __launch_bounds__(256, 2) gdn_chunk_prepare_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qn,
                                                                   const __nv_bfloat16* __restrict__ kn,
                                                                   const __nv_bfloat16* __restrict__ vc,
                                                                   const float* __restrict__ beta,
                                                                   const float* __restrict__ g, int M, int n_chunks,
                                                                   float* __restrict__ vcorr, float* __restrict__ kcd,
                                                                   float* __restrict__ intra,
                                                                   float* __restrict__ gcum_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int ch = blockIdx.x, h = blockIdx.y, r = d.nv / d.nk, kh = h / r;
  const int dk = d.dk, dv = d.dv, kd = d.nk * dk, vd = d.nv * dv;
  const int ldkb = dk + 8, ldb = dv + dk + 8;          // bf16 row stride (halves); fp32 RHS row stride (8 mod 32 words)
  __nv_bfloat16* sk = reinterpret_cast<__nv_bfloat16*>(smraw);                   // [64][ldkb]
  float* sA = reinterpret_cast<float*>(smraw + (size_t)kGC * ldkb * 2);          // [64][kLdA]
  float* sB = sA + kGC * kLdA;                                                   // [64][ldb]  right-hand sides -> solution
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(sB);                      // [64][ldkb] (dead before sB is written)
  float* sg = sB + kGC * ldb;                                                    // [64] gcum
  float* sb = sg + kGC;                                                          // [64] beta
  const int t0 = ch * kGC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const long long hc = (long long)h * n_chunks + ch;
  {
    const int cpr = dk / 8;                            // 16-byte pieces per row
    for (int idx = tid; idx < kGC * cpr; idx += 256) {
      const int i = idx / cpr, c = (idx % cpr) * 8, t = t0 + i;
      const bool ok = t < M;
      const long long off = (long long)(ok ? t : 0) * kd + kh * dk + c;
      cp_async16_zfill(sq + i * ldkb + c, qn + off, ok);
      cp_async16_zfill(sk + i * ldkb + c, kn + off, ok);
    }
    cp_async_commit();
  }
  if (tid < kGC) {
    const int t = t0 + tid;
    sb[tid] = t < M ? beta[(long long)t * d.nv + h] : 0.f;
    sg[tid] = t < M ? g[(long long)t * d.nv + h] : 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (tid == 0) {                         // cumulative sum in token order (matches torch.cumsum), in registers
    float s = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < kGC / 4; ++i4) {
      float4 v = *reinterpret_cast<float4*>(sg + i4 * 4);
      v.x = s + v.x; v.y = v.x + v.y; v.z = v.y + v.z; v.w = v.z + v.w;
      s = v.w;
      *reinterpret_cast<float4*>(sg + i4 * 4) = v;
    }
  }
  // K K^T (warps 0-3) and Q K^T (warps 4-7): q, k hold BF16 values -> exact in TF32, fp32 accumulate
  const int mt = warp & 3;
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  warp_mma_tiles<8, false, false>(acc, (warp < 4 ? sk : sq) + mt * 16 * ldkb, ldkb, 1, sk, 1, ldkb, 0, dk);   // B(k=c, n=j) = sk[j][c]
  __syncthreads();                        // gcum ready; q tile dead -> sB may be written
  {
    float* o_intra = intra + hc * kGC * kGC;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int i = mt * 16 + gq + e2 * 8, j = nt * 8 + 2 * tq;
        const float gi = sg[i], d0 = expf(gi - sg[j]), d1 = expf(gi - sg[j + 1]);
        const float a0 = acc[nt][e2 * 2], a1 = acc[nt][e2 * 2 + 1];
        if (warp < 4) {
          const float bi = sb[i];
          *reinterpret_cast<float2*>(sA + i * kLdA + j) =
              make_float2(j < i ? -(a0 * bi) * d0 : 0.f, j + 1 < i ? -(a1 * bi) * d1 : 0.f);
        } else if (TC) {
          unsigned short h0, l0, h1, l1;
          prep_split_bf16(j <= i ? a0 * d0 : 0.f, h0, l0);
          prep_split_bf16(j + 1 <= i ? a1 * d1 : 0.f, h1, l1);
          unsigned char* img = reinterpret_cast<unsigned char*>(intra) + hc * 16384;
          const uint32_t off = prep_sw128_off(i, j);
          *reinterpret_cast<uint32_t*>(img + off) = (uint32_t)h0 | ((uint32_t)h1 << 16);
          *reinterpret_cast<uint32_t*>(img + 8192 + off) = (uint32_t)l0 | ((uint32_t)l1 << 16);
        } else {
          *reinterpret_cast<float2*>(o_intra + i * kGC + j) = make_float2(j <= i ? a0 * d0 : 0.f, j + 1 <= i ? a1 * d1 : 0.f);
        }
      }
    }
  }
  // right-hand sides: [ v*beta | k*beta*exp(gcum) ], 8 columns per thread step
  {
    const int cpr = (dv + dk) / 8;
    for (int idx = tid; idx < kGC * cpr; idx += 256) {
      const int i = idx / cpr, c = (idx % cpr) * 8, t = t0 + i;
      const float bi = sb[i];
      uint4 raw = make_uint4(0, 0, 0, 0);
      float f;
      if (c < dv) {
        if (t < M) raw = *reinterpret_cast<const uint4*>(vc + (long long)t * vd + h * dv + c);
        f = bi;
      } else {
        raw = *reinterpret_cast<const uint4*>(sk + i * ldkb + (c - dv));
        f = bi * expf(sg[i]);
      }
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[2 * e] = __uint_as_float(w[e] << 16) * f;
        o[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u) * f;
      }
      *reinterpret_cast<float4*>(sB + i * ldb + c) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(sB + i * ldb + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  __syncthreads();
  // blocked forward substitution X = (I - A)^-1 B: 16-row blocks; off-diagonal blocks on mma (3xTF32), the 16x16
  // diagonal block in registers (one thread per column; A[i][j] is a warp-wide broadcast)
  const int ncols = dv + dk;
  for (int b = 0; b < kGC / 16; ++b) {
    if (b > 0) {
      for (int nt0 = warp * 4; nt0 * 8 < ncols; nt0 += 32) {        // 8 warps x 4 n-tiles = 256 columns per pass
        float ac[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r0 = b * 16 + gq, c0 = (nt0 + nt) * 8 + 2 * tq;
          const bool ok = c0 < ncols;
          const float2 lo = ok ? *reinterpret_cast<const float2*>(sB + r0 * ldb + c0) : make_float2(0.f, 0.f);
          const float2 hi = ok ? *reinterpret_cast<const float2*>(sB + (r0 + 8) * ldb + c0) : make_float2(0.f, 0.f);
          ac[nt][0] = lo.x; ac[nt][1] = lo.y; ac[nt][2] = hi.x; ac[nt][3] = hi.y;
        }
        warp_mma_tiles<4, true, true>(ac, sA + b * 16 * kLdA, kLdA, 1, sB + nt0 * 8, ldb, 1, 0, b * 16);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r0 = b * 16 + gq, c0 = (nt0 + nt) * 8 + 2 * tq;
          if (c0 < ncols) {
            *reinterpret_cast<float2*>(sB + r0 * ldb + c0) = make_float2(ac[nt][0], ac[nt][1]);
            *reinterpret_cast<float2*>(sB + (r0 + 8) * ldb + c0) = make_float2(ac[nt][2], ac[nt][3]);
          }
        }
      }
      __syncthreads();
    }
    for (int c = tid; c < ncols; c += 256) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = sB[(b * 16 + i) * ldb + c];
#pragma unroll
      for (int i = 1; i < 16; ++i) {
        const float* arow = sA + (b * 16 + i) * kLdA + b * 16;
        float a = x[i];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (j4 * 4 < i) {
            const float4 av = *reinterpret_cast<const float4*>(arow + j4 * 4);
            if (j4 * 4 + 0 < i) a = fmaf(av.x, x[j4 * 4 + 0], a);
            if (j4 * 4 + 1 < i) a = fmaf(av.y, x[j4 * 4 + 1], a);
            if (j4 * 4 + 2 < i) a = fmaf(av.z, x[j4 * 4 + 2], a);
            if (j4 * 4 + 3 < i) a = fmaf(av.w, x[j4 * 4 + 3], a);
          }
        }
        x[i] = a;
      }
#pragma unroll
      for (int i = 1; i < 16; ++i) sB[(b * 16 + i) * ldb + c] = x[i];
    }
    __syncthreads();
  }
  {
    const int cpr = ncols / 4;
    for (int idx = tid; idx < kGC * cpr; idx += 256) {
      const int i = idx / cpr, c = (idx % cpr) * 4;
      const float4 v = *reinterpret_cast<const float4*>(sB + i * ldb + c);
      if (TC) {
        if (c < dv) {
          *reinterpret_cast<float4*>(vcorr + ((hc * (dv / 32) + (c >> 5)) * kGC + i) * 36 + (c & 31)) = v;
        } else {
          unsigned short h[4], l[4];
          prep_split_bf16(v.x, h[0], l[0]); prep_split_bf16(v.y, h[1], l[1]);
          prep_split_bf16(v.z, h[2], l[2]); prep_split_bf16(v.w, h[3], l[3]);
          unsigned char* img = reinterpret_cast<unsigned char*>(kcd) + hc * 32768;
          const uint32_t off = prep_sw128_off(i, c - dv);
          *reinterpret_cast<uint2*>(img + off) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
          *reinterpret_cast<uint2*>(img + 16384 + off) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
        }
        continue;
      }
      if (c < dv) *reinterpret_cast<float4*>(vcorr + hc * kGC * dv + i * dv + c) = v;
      else *reinterpret_cast<float4*>(kcd + hc * kGC * dk + i * dk + (c - dv)) = v;
    }
  }
  if (tid < kGC) gcum_out[hc * kGC + tid] = sg[tid];
}

// ------------------------------------------------------------------------------------------------
// chunk scan: grid (nv, dv/kSV), 256 threads, one CTA per SM; state slice S[dk][kSV] in smem (fp32), updated in place.
// The per-chunk operands (q, k bf16; k_cumdecay, intra, value_corrected slice, gcum fp32) are double-buffered with
// cp.async: chunk c+1 streams in while chunk c runs its three mma phases.
// ------------------------------------------------------------------------------------------------
constexpr int kSVMax = 32;         // widest dv slice per CTA
constexpr int kLdS = kSVMax + 8;   // [k][n] operands: row stride = 8 mod 32 -> conflict-free B fragments

struct ScanStage {                 // byte offsets inside one stage
  int q, k, kc, I, V, g, bytes;
};
__host__ __device__ inline ScanStage scan_stage(int dk) {
  ScanStage s;
  s.q = 0;
  s.k = s.q + kGC * (dk + 8) * 2;
  s.kc = s.k + kGC * (dk + 8) * 2;
  s.I = s.kc + kGC * (dk + 4) * 4;
  s.V = s.I + kGC * kLdA * 4;
  s.g = s.V + kGC * kLdS * 4;
  s.bytes = s.g + kGC * 4;
  return s;
}

// kSV = dv slice width per CTA (32, 16 or 8): narrower slices keep every SM busy when a rank holds few heads
template <int kSV>
__global__ void __launch_bounds__(256, 1) gdn_chunk_scan_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qn,
                                                                const __nv_bfloat16* __restrict__ kn,
                                                                const float* __restrict__ vcorr,
                                                                const float* __restrict__ kcd,
                                                                const float* __restrict__ intra,
                                                                const float* __restrict__ gcum, int M, int n_chunks,
                                                                float* __restrict__ state,       // [nv][dk][dv] in/out
                                                                __nv_bfloat16* __restrict__ core_out) {  // [M][nv][dv] (core_attn_out.to(bf16))
  extern __shared__ __align__(16) unsigned char smraw[];
  const int h = blockIdx.x, sl = blockIdx.y, r = d.nv / d.nk, kh = h / r;
  const int dk = d.dk, dv = d.dv, kd = d.nk * dk, vd = d.nv * dv;
  const int ldk = dk + 4, ldkb = dk + 8;
  const ScanStage L = scan_stage(dk);
  float* S = reinterpret_cast<float*>(smraw);      // [dk][kLdS]
  float* sO = S + dk * kLdS;                       // [64][kLdS] decayed v_new
  unsigned char* stage0 = reinterpret_cast<unsigned char*>(sO + kGC * kLdS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  auto prefetch = [&](int ch, int buf) {
    unsigned char* st = stage0 + (size_t)buf * L.bytes;
    __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(st + L.q);
    __nv_bfloat16* sk = reinterpret_cast<__nv_bfloat16*>(st + L.k);
    float* skc = reinterpret_cast<float*>(st + L.kc);
    float* sI = reinterpret_cast<float*>(st + L.I);
    float* sV = reinterpret_cast<float*>(st + L.V);
    float* sg = reinterpret_cast<float*>(st + L.g);
    const int t0 = ch * kGC;
    const long long hc = (long long)h * n_chunks + ch;
    const int cq = dk / 8;
    for (int idx = tid; idx < kGC * cq; idx += 256) {
      const int i = idx / cq, c = (idx % cq) * 8, tt = t0 + i;
      const bool ok = tt < M;
      const long long off = (long long)(ok ? tt : 0) * kd + kh * dk + c;
      cp_async16_zfill(sq + i * ldkb + c, qn + off, ok);
      cp_async16_zfill(sk + i * ldkb + c, kn + off, ok);
    }
    const int ck = dk / 4;
    for (int idx = tid; idx < kGC * ck; idx += 256) {
      const int i = idx / ck, c = (idx % ck) * 4;
      cp_async16(skc + i * ldk + c, kcd + hc * kGC * dk + i * dk + c);
    }
    for (int idx = tid; idx < kGC * (kGC / 4); idx += 256) {
      const int i = idx / (kGC / 4), c = (idx % (kGC / 4)) * 4;
      cp_async16(sI + i * kLdA + c, intra + hc * kGC * kGC + i * kGC + c);
    }
    for (int idx = tid; idx < kGC * (kSV / 4); idx += 256) {
      const int i = idx / (kSV / 4), c = (idx % (kSV / 4)) * 4;
      cp_async16(sV + i * kLdS + c, vcorr + hc * kGC * dv + i * dv + sl * kSV + c);
    }
    if (tid < kGC / 4) cp_async16(sg + tid * 4, gcum + hc * kGC + tid * 4);
  };

  prefetch(0, 0);
  cp_async_commit();
  for (int idx = tid; idx < dk * kSV; idx += 256) {
    const int k = idx / kSV, c = idx % kSV;
    S[k * kLdS + c] = state[((long long)h * dk + k) * dv + sl * kSV + c];
  }
  constexpr int NT = kSV / 8, NTW = NT >= 2 ? NT / 2 : 1, NH = NT / NTW;   // n-tiles per warp in phases 1-2, column halves
  const int mt = warp & 3, nh = warp >> 2, n0 = nh * NTW * 8;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int t0 = ch * kGC, buf = ch & 1;
    unsigned char* st = stage0 + (size_t)buf * L.bytes;
    const __nv_bfloat16* sq = reinterpret_cast<const __nv_bfloat16*>(st + L.q);
    const __nv_bfloat16* sk = reinterpret_cast<const __nv_bfloat16*>(st + L.k);
    const float* skc = reinterpret_cast<const float*>(st + L.kc);
    const float* sI = reinterpret_cast<const float*>(st + L.I);
    float* sV = reinterpret_cast<float*>(st + L.V);
    const float* sg = reinterpret_cast<const float*>(st + L.g);
    cp_async_wait<0>();                            // this chunk's stage has landed (this thread's copies) ...
    __syncthreads();                               // ... everybody's; and step ch-1 is done with the other stage and S
    if (ch + 1 < n_chunks) {
      prefetch(ch + 1, buf ^ 1);
      cp_async_commit();
    }
    // (1) warp (mt, nh) owns rows [16mt, 16mt+16) x columns [n0, n0 + 8*NTW): VP = kcd @ S (fp32 x fp32 -> 3xTF32) and
    //     IT = q @ S (q exact in TF32 -> 2 mma), sharing the S fragments; then v_new = vcorr - VP and its decayed copy
    float it[NTW][4];
    if (nh < NH) {
      float vp[NTW][4];
#pragma unroll
      for (int a = 0; a < NTW; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) vp[a][b] = it[a][b] = 0.f;
      const float* A1 = skc + mt * 16 * ldk;
      const __nv_bfloat16* A2 = sq + mt * 16 * ldkb;
      const float* Bs = S + n0;
      for (int k0 = 0; k0 < dk; k0 += 8) {
        uint32_t ah[4], al[4], aq[4], dummy;
        split_tf32(A1[g * ldk + k0 + t], ah[0], al[0]);
        split_tf32(A1[(g + 8) * ldk + k0 + t], ah[1], al[1]);
        split_tf32(A1[g * ldk + k0 + t + 4], ah[2], al[2]);
        split_tf32(A1[(g + 8) * ldk + k0 + t + 4], ah[3], al[3]);
        ld_tf32<false>(A2 + g * ldkb + k0 + t, aq[0], dummy);
        ld_tf32<false>(A2 + (g + 8) * ldkb + k0 + t, aq[1], dummy);
        ld_tf32<false>(A2 + g * ldkb + k0 + t + 4, aq[2], dummy);
        ld_tf32<false>(A2 + (g + 8) * ldkb + k0 + t + 4, aq[3], dummy);
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
          uint32_t bh[2], bl[2];
          split_tf32(Bs[(k0 + t) * kLdS + nt * 8 + g], bh[0], bl[0]);
          split_tf32(Bs[(k0 + t + 4) * kLdS + nt * 8 + g], bh[1], bl[1]);
          mma_tf32(vp[nt], al, bh);
          mma_tf32(vp[nt], ah, bl);
          mma_tf32(vp[nt], ah, bh);
          mma_tf32(it[nt], aq, bl);
          mma_tf32(it[nt], aq, bh);
        }
      }
      const float gl = sg[kGC - 1];
      const int r0 = mt * 16 + g;
      const float d0 = expf(gl - sg[r0]), d1 = expf(gl - sg[r0 + 8]);
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int c0 = n0 + nt * 8 + 2 * t;
        float2 a = *reinterpret_cast<float2*>(sV + r0 * kLdS + c0), b = *reinterpret_cast<float2*>(sV + (r0 + 8) * kLdS + c0);
        a.x -= vp[nt][0]; a.y -= vp[nt][1]; b.x -= vp[nt][2]; b.y -= vp[nt][3];
        *reinterpret_cast<float2*>(sV + r0 * kLdS + c0) = a;
        *reinterpret_cast<float2*>(sV + (r0 + 8) * kLdS + c0) = b;
        *reinterpret_cast<float2*>(sO + r0 * kLdS + c0) = make_float2(d0 * a.x, d0 * a.y);
        *reinterpret_cast<float2*>(sO + (r0 + 8) * kLdS + c0) = make_float2(d1 * b.x, d1 * b.y);
      }
    }
    __syncthreads();
    // (2) out = exp(gcum_i) * IT + intra @ v_new (lower-triangular: k < 16(mt+1))
    if (nh < NH) {
      const float e0 = expf(sg[mt * 16 + g]), e1 = expf(sg[mt * 16 + g + 8]);
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        it[nt][0] *= e0; it[nt][1] *= e0; it[nt][2] *= e1; it[nt][3] *= e1;
      }
      warp_mma_tiles<NTW, true, true>(it, sI + mt * 16 * kLdA, kLdA, 1, sV + n0, kLdS, 1, 0, (mt + 1) * 16);
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int r0 = mt * 16 + g, c0 = sl * kSV + n0 + nt * 8 + 2 * t;
        const int ta = t0 + r0, tb = t0 + r0 + 8;
        if (ta < M) *reinterpret_cast<__nv_bfloat162*>(core_out + (long long)ta * vd + h * dv + c0) = __floats2bfloat162_rn(it[nt][0], it[nt][1]);
        if (tb < M) *reinterpret_cast<__nv_bfloat162*>(core_out + (long long)tb * vd + h * dv + c0) = __floats2bfloat162_rn(it[nt][2], it[nt][3]);
      }
    }
    // (3) S = S * exp(g_last) + K^T @ (decayed v_new): warp w owns state rows [16w, 16w+16) (dk = 128 -> 8 warps).
    //     Independent of (2): no barrier in between.
    {
      const float egl = expf(sg[kGC - 1]);
      for (int mt3 = warp; mt3 < dk / 16; mt3 += 8) {
        float sc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int r0 = mt3 * 16 + g, c0 = nt * 8 + 2 * t;
          const float2 a = *reinterpret_cast<const float2*>(S + r0 * kLdS + c0), b = *reinterpret_cast<const float2*>(S + (r0 + 8) * kLdS + c0);
          sc[nt][0] = a.x * egl; sc[nt][1] = a.y * egl; sc[nt][2] = b.x * egl; sc[nt][3] = b.y * egl;
        }
        // A = K^T: A(row = kk, k = i) = sk[i*ldkb + kk]
        warp_mma_tiles<NT, false, true>(sc, sk + mt3 * 16, 1, ldkb, sO, kLdS, 1, 0, kGC);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int r0 = mt3 * 16 + g, c0 = nt * 8 + 2 * t;
          *reinterpret_cast<float2*>(S + r0 * kLdS + c0) = make_float2(sc[nt][0], sc[nt][1]);
          *reinterpret_cast<float2*>(S + (r0 + 8) * kLdS + c0) = make_float2(sc[nt][2], sc[nt][3]);
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  for (int idx = tid; idx < dk * kSV; idx += 256) {
    const int k = idx / kSV, c = idx % kSV;
    state[((long long)h * dk + k) * dv + sl * kSV + c] = S[k * kLdS + c];
  }
}

// gated RMSNorm: one warp per (token, head) — or, for dv == 128, per (token, HP heads) with all loads issued up front
template <int HP>
__global__ void __launch_bounds__(256) gdn_post128_kernel(GdnDims d, const __nv_bfloat16* __restrict__ core,
                                                          const __nv_bfloat16* __restrict__ qkvz,
                                                          const float* __restrict__ norm_w, int M,
                                                          __nv_bfloat16* __restrict__ out) {   // [M][nv*128]
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, groups = d.nv / HP;
  if (wid >= (long long)M * groups) return;
  const int t = (int)(wid / groups), h0 = (int)(wid % groups) * HP, r = d.nv / d.nk;
  const int kd = d.nk * d.dk, vd = d.nv * 128, ld = 2 * kd + 2 * vd, G = 2 * d.dk + 2 * r * 128;
  const int c = lane * 4;
  uint2 xr[HP], zr[HP];
#pragma unroll
  for (int i = 0; i < HP; ++i) {
    const int h = h0 + i;
    xr[i] = *reinterpret_cast<const uint2*>(core + (long long)t * vd + h * 128 + c);
    zr[i] = *reinterpret_cast<const uint2*>(qkvz + (long long)t * ld + (h / r) * G + 2 * d.dk + r * 128 + (h % r) * 128 + c);
  }
  const float4 nw = *reinterpret_cast<const float4*>(norm_w + c);
  const float wv[4] = {nw.x, nw.y, nw.z, nw.w};
#pragma unroll
  for (int i = 0; i < HP; ++i) {
    const float x[4] = {__uint_as_float(xr[i].x << 16), __uint_as_float(xr[i].x & 0xFFFF0000u), __uint_as_float(xr[i].y << 16),
                        __uint_as_float(xr[i].y & 0xFFFF0000u)};
    const float z[4] = {__uint_as_float(zr[i].x << 16), __uint_as_float(zr[i].x & 0xFFFF0000u), __uint_as_float(zr[i].y << 16),
                        __uint_as_float(zr[i].y & 0xFFFF0000u)};
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(ss / 128.f + d.eps);
    float o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xn = bf16r(wv[j] * (x[j] * inv));
      const float sz = bf16r(__fdividef(z[j], 1.0f + __expf(-z[j])));
      o4[j] = xn * sz;
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(o4[0], o4[1]), hi = __floats2bfloat162_rn(o4[2], o4[3]);
    *reinterpret_cast<uint2*>(out + (long long)t * vd + (h0 + i) * 128 + c) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

__global__ void __launch_bounds__(256) gdn_post_kernel(GdnDims d, const __nv_bfloat16* __restrict__ core,
                                                       const __nv_bfloat16* __restrict__ qkvz,
                                                       const float* __restrict__ norm_w, int M,
                                                       __nv_bfloat16* __restrict__ out) {   // [M][nv*dv]
  const int wid = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (wid >= M * d.nv) return;
  const int t = wid / d.nv, h = wid % d.nv, r = d.nv / d.nk;
  const int kd = d.nk * d.dk, vd = d.nv * d.dv, ld = 2 * kd + 2 * vd, G = 2 * d.dk + 2 * r * d.dv;
  const int zcol = (h / r) * G + 2 * d.dk + r * d.dv + (h % r) * d.dv;
  float ss = 0.f;
  for (int c = lane; c < d.dv; c += 32) {
    const float x = __bfloat162float(core[(long long)t * vd + h * d.dv + c]);          // core_attn_out.to(bf16), done by the scan
    ss += x * x;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = rsqrtf(ss / d.dv + d.eps);
  for (int c = lane; c < d.dv; c += 32) {
    const float x = __bfloat162float(core[(long long)t * vd + h * d.dv + c]);
    const float xn = bf16r(norm_w[c] * (x * inv));
    const float z = __bfloat162float(qkvz[(long long)t * ld + zcol + c]);
    const float sz = bf16r(z / (1.0f + expf(-z)));
    out[(long long)t * vd + h * d.dv + c] = __float2bfloat16_rn(xn * sz);
  }
}

// ------------------------------------------------------------------------------------------------
size_t gdn_prepare_smem(const GdnDims& d) {
  return (size_t)kGC * (d.dk + 8) * 2 + sizeof(float) * (kGC * kLdA + kGC * (d.dv + d.dk + 8) + 2 * kGC);
}
size_t gdn_scan_smem(const GdnDims& d) {   // same for every slice width
  return sizeof(float) * (d.dk * kLdS + kGC * kLdS) + 2 * (size_t)scan_stage(d.dk).bytes;
}

cudaError_t launch_gdn_scan_tc(const void* qn, const void* kn, const void* kcd_img, const void* intra_img, const float* vcorr,
                               const float* gcum, float* state, void* core_out, int M, int n_chunks, int nk, int nv,
                               cudaStream_t s);

// KB2_GDN_LEGACY=1 keeps the mma.sync scan for dk == dv == 128 too (A/B comparison in tests; never set in production)
static void launch_gdn_post(const GdnDims& d, const void* core, const void* qkvz, const float* norm_w, int M, void* out,
                            cudaStream_t s) {
  if (d.dv == 128 && d.nv % 4 == 0) {
    const long long nw = (long long)M * (d.nv / 4);
    gdn_post128_kernel<4><<<(unsigned)((nw * 32 + 255) / 256), 256, 0, s>>>(d, (const __nv_bfloat16*)core, (const __nv_bfloat16*)qkvz,
                                                                           norm_w, M, (__nv_bfloat16*)out);
    return;
  }
  const long long nw = (long long)M * d.nv;
  gdn_post_kernel<<<(unsigned)((nw * 32 + 255) / 256), 256, 0, s>>>(d, (const __nv_bfloat16*)core, (const __nv_bfloat16*)qkvz, norm_w, M,
                                                                  (__nv_bfloat16*)out);
}

static bool gdn_legacy_env() {
  const char* e = getenv("KB2_GDN_LEGACY");            // read per call so one test process can run both paths
  return e && e[0] == '1';
}
static bool gdn_use_tc(const GdnDims& d) { return !gdn_legacy_env() && d.dk == 128 && d.dv == 128; }

cudaError_t launch_gdn_prepare_tc(const void* qn, const void* kn, const void* vc, const float* beta, const float* g,
                                  void* kcd_img, void* intra_img, float* vcorr, float* gcum, int M, int n_chunks, int nk, int nv,
                                  int num_sms, cudaStream_t s);
// Which chunk-prepare feeds the tcgen05 scan (dk == dv == 128).  QCN layer, 8192 tokens: the mma.sync prepare of round 1 takes 332 us,
// the tcgen05 prepare (gdn_tc.cu) 148 us after this round's work (460 us in its first version: profiles/r02c ... r02u).  The
// mma.sync kernel stays for other head sizes and as the A/B reference of tests/test_gpu_scale_parity.py:
// KB2_GDN_PREPARE_MMA_SYNC=1 picks it, =0 the tcgen05 kernel; unset -> kDefaultPrepareTc.
constexpr bool kDefaultPrepareTc = true;
static bool gdn_prepare_mma_sync_env() {
  const char* e = getenv("KB2_GDN_PREPARE_MMA_SYNC");
  return e ? e[0] == '1' : !kDefaultPrepareTc;
}
static int gdn_num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

cudaError_t launch_gdn_core(const GdnDims& d, const void* qkvz, const void* ba, const void* conv_w, void* conv_state,
                            const float* A_log, const float* dt_bias, const float* norm_w, float* rec_state,
                            void* qn, void* kn, void* vc, float* beta, float* g, float* vcorr, float* kcd, float* intra,
                            float* gcum, float* core, void* normed_out, int M, cudaStream_t s) {
  if (d.dv % kSVMax || d.dk > 128 || d.dk % 16 || d.dv > 128 || d.K > 8) return cudaErrorInvalidValue;
  const int C = 2 * d.nk * d.dk + d.nv * d.dv;
  const int n_chunks = (M + kGC - 1) / kGC;
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    cudaFuncSetAttribute(gdn_chunk_prepare_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(gdn_chunk_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(gdn_chunk_scan_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gdn_chunk_scan_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gdn_chunk_scan_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gdn_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    once.mark(dev);
  }
  { KernelSpan ks(K_GDN_PREP, s);
  if (d.dk == 128 && d.dv == 128 && d.K == 4 && !gdn_legacy_env()) {
    constexpr int TT = 64;
    const int n_tiles = (M + TT - 1) / TT;
    const long long warps = (long long)n_tiles * (2 * d.nk + d.nv + 1);
    gdn_prep_vec_kernel<TT><<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(
        d, (const __nv_bfloat16*)qkvz, (const __nv_bfloat16*)ba, (const __nv_bfloat16*)conv_w,
        (const __nv_bfloat16*)conv_state, A_log, dt_bias, (__nv_bfloat16*)qn, (__nv_bfloat16*)kn, (__nv_bfloat16*)vc, beta, g,
        M, n_tiles);
  } else if (d.dk == 128 && d.dv == 128 && d.K == 4) {
    const int n_tiles = (M + 31) / 32;
    const long long warps = (long long)n_tiles * (2 * d.nk + d.nv + 1);
    gdn_prep_tiled_kernel<4, 4><<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(
        d, (const __nv_bfloat16*)qkvz, (const __nv_bfloat16*)ba, (const __nv_bfloat16*)conv_w,
        (const __nv_bfloat16*)conv_state, A_log, dt_bias, (__nv_bfloat16*)qn, (__nv_bfloat16*)kn, (__nv_bfloat16*)vc, beta, g,
        M, n_tiles);
  } else {
    gdn_prep_kernel<<<M, 256, sizeof(float) * C, s>>>(d, (const __nv_bfloat16*)qkvz, (const __nv_bfloat16*)ba,
                                                      (const __nv_bfloat16*)conv_w, (const __nv_bfloat16*)conv_state, A_log,
                                                      dt_bias, (__nv_bfloat16*)qn, (__nv_bfloat16*)kn, (__nv_bfloat16*)vc,
                                                      beta, g, M);
  } }
  { KernelSpan ks(K_GDN_CONV_STATE, s);
  gdn_conv_state_kernel<<<(C + 255) / 256, 256, 0, s>>>(d, (const __nv_bfloat16*)qkvz, (__nv_bfloat16*)conv_state, M); }
  const long long nw = (long long)M * d.nv;
  if (gdn_use_tc(d)) {
    { KernelSpan ks(K_GDN_PREPARE, s);
    if (gdn_prepare_mma_sync_env()) {
      gdn_chunk_prepare_kernel<true><<<dim3(n_chunks, d.nv), 256, gdn_prepare_smem(d), s>>>(
          d, (const __nv_bfloat16*)qn, (const __nv_bfloat16*)kn, (const __nv_bfloat16*)vc, beta, g, M, n_chunks, vcorr, kcd,
          intra, gcum);
    } else {
      static int sms = gdn_num_sms();
      cudaError_t e = launch_gdn_prepare_tc(qn, kn, vc, beta, g, kcd, intra, vcorr, gcum, M, n_chunks, d.nk, d.nv, sms, s);
      if (e != cudaSuccess) return e;
    } }
    { KernelSpan ks(K_GDN_SCAN, s);
    cudaError_t e = launch_gdn_scan_tc(qn, kn, kcd, intra, vcorr, gcum, rec_state, core, M, n_chunks, d.nk, d.nv, s);
    if (e != cudaSuccess) return e; }
    KernelSpan ks(K_GDN_POST, s);
    launch_gdn_post(d, core, qkvz, norm_w, M, normed_out, s);
    return cudaGetLastError();
  }
  { KernelSpan ks(K_GDN_PREPARE, s);
  gdn_chunk_prepare_kernel<false><<<dim3(n_chunks, d.nv), 256, gdn_prepare_smem(d), s>>>(
      d, (const __nv_bfloat16*)qn, (const __nv_bfloat16*)kn, (const __nv_bfloat16*)vc, beta, g, M, n_chunks, vcorr, kcd,
      intra, gcum); }
  // slice width: keep >= ~128 CTAs in flight (one per SM) when a rank holds only a few heads
  const int sv = d.nv * (d.dv / 32) >= 96 ? 32 : (d.nv * (d.dv / 16) >= 96 ? 16 : 8);
#define KB2_SCAN(SV)                                                                                                  \
  gdn_chunk_scan_kernel<SV><<<dim3(d.nv, d.dv / SV), 256, gdn_scan_smem(d), s>>>(                                      \
      d, (const __nv_bfloat16*)qn, (const __nv_bfloat16*)kn, vcorr, kcd, intra, gcum, M, n_chunks, rec_state,       \
      (__nv_bfloat16*)core)
  { KernelSpan ks(K_GDN_SCAN, s);
  if (sv == 32) KB2_SCAN(32); else if (sv == 16) KB2_SCAN(16); else KB2_SCAN(8); }
#undef KB2_SCAN
  KernelSpan ks(K_GDN_POST, s);
  launch_gdn_post(d, core, qkvz, norm_w, M, normed_out, s);
  return cudaGetLastError();
}

}  // namespace kb2
