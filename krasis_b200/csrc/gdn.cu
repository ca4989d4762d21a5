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
#include "moe_common.cuh"

namespace kb2 {

constexpr int kGC = 64;   // chunk size (linear_attention.py:702)

struct GdnDims {
  int H, nk, nv, dk, dv, K;   // K = conv kernel size (4)
  float eps, scale;
  int ba_ld;                  // row stride of the ba projection output (2*nv rounded up to 16)
};

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

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

// ------------------------------------------------------------------------------------------------
// chunk prepare: grid (n_chunks, nv), 256 threads.  dk, dv <= 128, dk + dv == 256 is NOT required.
// outputs per (head, chunk): vcorr [64][dv], kcd [64][dk], intra [64][64], gcum [64]   (all fp32)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gdn_chunk_prepare_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qn,
                                                                const __nv_bfloat16* __restrict__ kn,
                                                                const __nv_bfloat16* __restrict__ vc,
                                                                const float* __restrict__ beta,
                                                                const float* __restrict__ g, int M, int n_chunks,
                                                                float* __restrict__ vcorr, float* __restrict__ kcd,
                                                                float* __restrict__ intra, float* __restrict__ gcum_out) {
  extern __shared__ float sm[];
  const int ch = blockIdx.x, h = blockIdx.y, r = d.nv / d.nk, kh = h / r;
  const int dk = d.dk, dv = d.dv, kd = d.nk * dk, vd = d.nv * dv;
  float* sq = sm;                       // [64][dk+1]
  float* sk = sq + kGC * (dk + 1);      // [64][dk+1]
  float* sA = sk + kGC * (dk + 1);      // [64][65]
  float* sB = sA + kGC * 65;            // [64][dv + dk]   right-hand sides / solution
  float* sg = sB + kGC * (dv + dk);     // [64] gcum
  float* sb = sg + kGC;                 // [64] beta
  const int t0 = ch * kGC;
  const int tid = threadIdx.x;
  for (int i = tid; i < kGC; i += 256) {
    const int t = t0 + i;
    sb[i] = t < M ? beta[(long long)t * d.nv + h] : 0.f;
    sg[i] = t < M ? g[(long long)t * d.nv + h] : 0.f;
  }
  for (int idx = tid; idx < kGC * dk; idx += 256) {
    const int i = idx / dk, c = idx % dk, t = t0 + i;
    sq[i * (dk + 1) + c] = t < M ? __bfloat162float(qn[(long long)t * kd + kh * dk + c]) : 0.f;
    sk[i * (dk + 1) + c] = t < M ? __bfloat162float(kn[(long long)t * kd + kh * dk + c]) : 0.f;
  }
  __syncthreads();
  if (tid == 0) {                         // cumulative sum in token order (matches torch.cumsum)
    float s = 0.f;
    for (int i = 0; i < kGC; ++i) {
      s += sg[i];
      sg[i] = s;
    }
  }
  __syncthreads();
  // right-hand sides: [ v*beta | k*beta*exp(gcum) ]
  for (int idx = tid; idx < kGC * (dv + dk); idx += 256) {
    const int i = idx / (dv + dk), c = idx % (dv + dk), t = t0 + i;
    float val;
    if (c < dv) val = t < M ? __bfloat162float(vc[(long long)t * vd + h * dv + c]) * sb[i] : 0.f;
    else val = sk[i * (dk + 1) + (c - dv)] * sb[i] * expf(sg[i]);
    sB[i * (dv + dk) + c] = val;
  }
  // A[i][j] = -(k_i.beta_i . k_j) * exp(gcum_i - gcum_j), j < i ; intra[i][j] = (q_i . k_j) * exp(gcum_i - gcum_j), j <= i
  float* o_intra = intra + ((long long)h * n_chunks + ch) * kGC * kGC;
  for (int idx = tid; idx < kGC * kGC; idx += 256) {
    const int i = idx / kGC, j = idx % kGC;
    float a = 0.f, qk = 0.f;
    if (j <= i) {
      float kk = 0.f;
      for (int c = 0; c < dk; ++c) {
        const float kj = sk[j * (dk + 1) + c];
        kk = fmaf(sk[i * (dk + 1) + c], kj, kk);
        qk = fmaf(sq[i * (dk + 1) + c], kj, qk);
      }
      const float dec = expf(sg[i] - sg[j]);
      qk *= dec;
      a = j < i ? -(kk * sb[i]) * dec : 0.f;
    }
    sA[i * 65 + j] = a;
    o_intra[idx] = qk;
  }
  __syncthreads();
  // forward substitution in place: X[i] = B[i] + sum_{j<i} A[i][j] X[j]; one thread owns one column of sB
  // (bank = column % 32 -> conflict-free; A[i][j] is a warp-wide broadcast)
  for (int c = tid; c < dv + dk; c += 256) {
    const int ldb = dv + dk;
    for (int i = 1; i < kGC; ++i) {
      float acc = sB[i * ldb + c];
      const float* arow = sA + i * 65;
#pragma unroll 4
      for (int j = 0; j < i; ++j) acc = fmaf(arow[j], sB[j * ldb + c], acc);
      sB[i * ldb + c] = acc;
    }
    float* dst = c < dv ? vcorr + ((long long)h * n_chunks + ch) * kGC * dv + c
                        : kcd + ((long long)h * n_chunks + ch) * kGC * dk + (c - dv);
    const int ldd = c < dv ? dv : dk;
    for (int i = 0; i < kGC; ++i) dst[i * ldd] = sB[i * ldb + c];
  }
  for (int i = tid; i < kGC; i += 256) gcum_out[((long long)h * n_chunks + ch) * kGC + i] = sg[i];
}

// ------------------------------------------------------------------------------------------------
// chunk scan: grid (nv, dv/32), 256 threads; state slice S[dk][32] in smem (fp32), updated in place.
// ------------------------------------------------------------------------------------------------
constexpr int kSV = 32;   // dv slice width

__global__ void __launch_bounds__(256) gdn_chunk_scan_kernel(GdnDims d, const __nv_bfloat16* __restrict__ qn,
                                                             const __nv_bfloat16* __restrict__ kn,
                                                             const float* __restrict__ vcorr,
                                                             const float* __restrict__ kcd,
                                                             const float* __restrict__ intra,
                                                             const float* __restrict__ gcum, int M, int n_chunks,
                                                             float* __restrict__ state,       // [nv][dk][dv] in/out
                                                             float* __restrict__ core_out) {  // [M][nv][dv]
  extern __shared__ float sm[];
  const int h = blockIdx.x, sl = blockIdx.y, r = d.nv / d.nk, kh = h / r;
  const int dk = d.dk, dv = d.dv, kd = d.nk * dk, vd = d.nv * dv;
  float* S = sm;                          // [dk][kSV]
  float* sq = S + dk * kSV;               // [64][dk+1]
  float* sk = sq + kGC * (dk + 1);        // [64][dk+1]
  float* skc = sk + kGC * (dk + 1);       // [64][dk+1]  k_cumdecay
  float* sI = skc + kGC * (dk + 1);       // [64][65]
  float* sV = sI + kGC * 65;              // [64][kSV]   value_corrected slice -> v_new
  float* sO = sV + kGC * kSV;             // [64][kSV]
  float* sg = sO + kGC * kSV;             // [64]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < dk * kSV; idx += 256) {
    const int k = idx / kSV, c = idx % kSV;
    S[idx] = state[((long long)h * dk + k) * dv + sl * kSV + c];
  }
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int t0 = ch * kGC;
    const long long hc = (long long)h * n_chunks + ch;
    __syncthreads();
    for (int idx = tid; idx < kGC * dk; idx += 256) {
      const int i = idx / dk, c = idx % dk, t = t0 + i;
      sq[i * (dk + 1) + c] = t < M ? __bfloat162float(qn[(long long)t * kd + kh * dk + c]) : 0.f;
      sk[i * (dk + 1) + c] = t < M ? __bfloat162float(kn[(long long)t * kd + kh * dk + c]) : 0.f;
      skc[i * (dk + 1) + c] = kcd[hc * kGC * dk + idx];
    }
    for (int idx = tid; idx < kGC * kGC; idx += 256) sI[(idx / kGC) * 65 + (idx % kGC)] = intra[hc * kGC * kGC + idx];
    for (int idx = tid; idx < kGC * kSV; idx += 256) {
      const int i = idx / kSV, c = idx % kSV;
      sV[idx] = vcorr[hc * kGC * dv + i * dv + sl * kSV + c];
    }
    for (int i = tid; i < kGC; i += 256) sg[i] = gcum[hc * kGC + i];
    __syncthreads();
    // (1) v_new = vcorr - kcd @ S ;  inter = exp(gcum_i) * (q_i @ S)      [64 x 32], 8 outputs per thread
    {
      const int c = tid % kSV, i0 = tid / kSV;                 // i0 in 0..7, rows i0, i0+8, ...
      float vp[8], it[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) vp[u] = it[u] = 0.f;
      for (int k = 0; k < dk; ++k) {
        const float s = S[k * kSV + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + 8 * u;
          vp[u] = fmaf(skc[i * (dk + 1) + k], s, vp[u]);
          it[u] = fmaf(sq[i * (dk + 1) + k], s, it[u]);
        }
      }
      __syncthreads();                                           // everyone done reading S/sV inputs? (sV read below)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + 8 * u;
        sV[i * kSV + c] = sV[i * kSV + c] - vp[u];              // v_new (own element only)
        sO[i * kSV + c] = it[u] * expf(sg[i]);
      }
    }
    __syncthreads();
    // (2) out = inter + intra @ v_new
    {
      const int c = tid % kSV, i0 = tid / kSV;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + 8 * u;
        float acc = sO[i * kSV + c];
        for (int j = 0; j <= i; ++j) acc = fmaf(sI[i * 65 + j], sV[j * kSV + c], acc);
        const int t = t0 + i;
        if (t < M) core_out[(long long)t * vd + h * dv + sl * kSV + c] = acc;
      }
    }
    // (3) S = S * exp(g_last) + sum_i k_i^T * (exp(g_last - gcum_i) * v_new_i)     [dk x 32], dk/8 rows per thread
    __syncthreads();                                             // (2) is done with sO
    {
      const float gl = sg[kGC - 1];
      for (int idx = tid; idx < kGC * kSV; idx += 256) sO[idx] = expf(gl - sg[idx / kSV]) * sV[idx];   // decayed v_new
      __syncthreads();
      const float egl = expf(gl);
      const int c = tid % kSV, k0 = tid / kSV;                  // k rows k0, k0+8, ...
      for (int kk = k0; kk < dk; kk += 8) {
        float acc = S[kk * kSV + c] * egl;
#pragma unroll 8
        for (int i = 0; i < kGC; ++i) acc = fmaf(sk[i * (dk + 1) + kk], sO[i * kSV + c], acc);
        S[kk * kSV + c] = acc;
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < dk * kSV; idx += 256) {
    const int k = idx / kSV, c = idx % kSV;
    state[((long long)h * dk + k) * dv + sl * kSV + c] = S[idx];
  }
}

// gated RMSNorm: one warp per (token, head)
__global__ void __launch_bounds__(256) gdn_post_kernel(GdnDims d, const float* __restrict__ core,
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
    const float x = bf16r(core[(long long)t * vd + h * d.dv + c]);          // core_attn_out.to(bf16)
    ss += x * x;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = rsqrtf(ss / d.dv + d.eps);
  for (int c = lane; c < d.dv; c += 32) {
    const float x = bf16r(core[(long long)t * vd + h * d.dv + c]);
    const float xn = bf16r(norm_w[c] * (x * inv));
    const float z = __bfloat162float(qkvz[(long long)t * ld + zcol + c]);
    const float sz = bf16r(z / (1.0f + expf(-z)));
    out[(long long)t * vd + h * d.dv + c] = __float2bfloat16_rn(xn * sz);
  }
}

// ------------------------------------------------------------------------------------------------
size_t gdn_prepare_smem(const GdnDims& d) {
  return sizeof(float) * (2 * kGC * (d.dk + 1) + kGC * 65 + kGC * (d.dv + d.dk) + 2 * kGC);
}
size_t gdn_scan_smem(const GdnDims& d) {
  return sizeof(float) * (d.dk * kSV + 3 * kGC * (d.dk + 1) + kGC * 65 + 2 * kGC * kSV + kGC);
}

cudaError_t launch_gdn_core(const GdnDims& d, const void* qkvz, const void* ba, const void* conv_w, void* conv_state,
                            const float* A_log, const float* dt_bias, const float* norm_w, float* rec_state,
                            void* qn, void* kn, void* vc, float* beta, float* g, float* vcorr, float* kcd, float* intra,
                            float* gcum, float* core, void* normed_out, int M, cudaStream_t s) {
  if (d.dv % kSV || d.dk > 128 || d.dv > 128 || d.K > 8) return cudaErrorInvalidValue;
  const int C = 2 * d.nk * d.dk + d.nv * d.dv;
  const int n_chunks = (M + kGC - 1) / kGC;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(gdn_chunk_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gdn_chunk_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gdn_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    configured = true;
  }
  gdn_prep_kernel<<<M, 256, sizeof(float) * C, s>>>(d, (const __nv_bfloat16*)qkvz, (const __nv_bfloat16*)ba,
                                                    (const __nv_bfloat16*)conv_w, (const __nv_bfloat16*)conv_state, A_log,
                                                    dt_bias, (__nv_bfloat16*)qn, (__nv_bfloat16*)kn, (__nv_bfloat16*)vc,
                                                    beta, g, M);
  gdn_conv_state_kernel<<<(C + 255) / 256, 256, 0, s>>>(d, (const __nv_bfloat16*)qkvz, (__nv_bfloat16*)conv_state, M);
  gdn_chunk_prepare_kernel<<<dim3(n_chunks, d.nv), 256, gdn_prepare_smem(d), s>>>(
      d, (const __nv_bfloat16*)qn, (const __nv_bfloat16*)kn, (const __nv_bfloat16*)vc, beta, g, M, n_chunks, vcorr, kcd,
      intra, gcum);
  gdn_chunk_scan_kernel<<<dim3(d.nv, d.dv / kSV), 256, gdn_scan_smem(d), s>>>(
      d, (const __nv_bfloat16*)qn, (const __nv_bfloat16*)kn, vcorr, kcd, intra, gcum, M, n_chunks, rec_state, core);
  const long long nw = (long long)M * d.nv;
  gdn_post_kernel<<<(unsigned)((nw * 32 + 255) / 256), 256, 0, s>>>(d, core, (const __nv_bfloat16*)qkvz, norm_w, M,
                                                                  (__nv_bfloat16*)normed_out);
  return cudaGetLastError();
}

}  // namespace kb2
