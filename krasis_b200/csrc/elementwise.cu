// Row-wise HBM-bound kernels around the GEMMs: RMSNorm, fused residual-add + RMSNorm, per-row INT8 quantisation,
// SiLU*mul, sigmoid-gated scaling.
//   rmsnorm / fused_add_rmsnorm   flashinfer.norm.{rmsnorm,fused_add_rmsnorm} as used at python/krasis/layer.py:163-183,
//                                 283-308 (x = input + residual in fp32; residual <- bf16(x); out <- bf16(x * rsqrt(mean x^2 + eps) * w))
//   quant_rows_int8               python/krasis/weight_loader.py:25-43 (weights) and :66-70 (activations):
//                                 scale = clamp(amax, 1e-10) / 127 (fp32); q = clamp(round_half_even(x / scale), -128, 127)
//   silu_and_mul                  flashinfer.activation.silu_and_mul (layer.py:512): bf16(silu(x[:N]) * x[N:]) in fp32
//   sigmoid_gate_mul              layer.py:518-522: out *= sigmoid(F.linear(hidden, w[1,H])) with BF16 rounding at every tensor op
#include "moe_common.cuh"
#include "prof.cuh"

namespace kb2 {

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  __syncthreads();
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
  __syncthreads();
  return t;
}

// clamp(round_half_even(x / scale)) of the reference's `(x / scale).round().clamp(-128, 127)` (weight_loader.py:25-43, 46-99), exactly,
// without an IEEE division per element: y = x * rn(1 / scale) is within ~1e-5 of the correctly rounded quotient for |x / scale| <= 128,
// so rint(y) can only differ from rint(x / scale) when y sits within that distance of a half-integer — those (rare) elements take the
// true division.
__device__ __forceinline__ int quant_code_rhe(float x, float scale, float rscale) {
  const float y = x * rscale;
  float r = rintf(y);
  if (fabsf(fabsf(y - r) - 0.5f) < 1e-3f) r = rintf(x / scale);
  return (int)fminf(fmaxf(r, -128.f), 127.f);
}

// one WARP per row, row kept in registers (H == 256 * VPL): no shared memory, no block barriers
// kQuant: additionally emits the row-quantised INT8 copy of the OUTPUT row (the activation quantisation of int8_linear,
// weight_loader.py:46-99: scale = max(|y|, 1e-10) / 127, q = clamp(round_half_even(y / scale))) — the row is already in registers,
// so the shared expert's first W8A8 GEMM needs no separate quantisation pass over the normed activations.
template <bool kAdd, int VPL, bool kQuant = false>
__global__ void __launch_bounds__(256) rmsnorm_warp_kernel(__nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ residual,
                                                           const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                           int M, float eps, int8_t* __restrict__ q_out = nullptr,
                                                           float* __restrict__ q_scale = nullptr) {
  constexpr int H = 256 * VPL;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float f[VPL][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int v = lane + 32 * j;
    const uint4 a = *reinterpret_cast<const uint4*>(x + row * H + v * 8);
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[j][i] = __bfloat162float(pa[i]);
    if constexpr (kAdd) {
      const uint4 r = *reinterpret_cast<const uint4*>(residual + row * H + v * 8);
      const __nv_bfloat16* pr = reinterpret_cast<const __nv_bfloat16*>(&r);
      uint4 o;
      __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[j][i] += __bfloat162float(pr[i]);
        po[i] = __float2bfloat16_rn(f[j][i]);
      }
      *reinterpret_cast<uint4*>(residual + row * H + v * 8) = o;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += f[j][i] * f[j][i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = rsqrtf(ss / H + eps);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int v = lane + 32 * j;
    const float4 w0 = *reinterpret_cast<const float4*>(w + v * 8), w1 = *reinterpret_cast<const float4*>(w + v * 8 + 4);
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint4 o;
    __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      po[i] = __float2bfloat16_rn(f[j][i] * inv * ww[i]);
      if constexpr (kQuant) f[j][i] = __bfloat162float(po[i]);
    }
    *reinterpret_cast<uint4*>(out + row * H + v * 8) = o;
  }
  if constexpr (kQuant) {
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(f[j][i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float scale = fmaxf(mx, 1e-10f) / 127.0f;
    const float rscale = 1.0f / scale;
    if (lane == 0) q_scale[row] = scale;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      uint32_t o[2] = {0u, 0u};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int qi = quant_code_rhe(f[j][i], scale, rscale);
        o[i >> 2] |= (uint32_t)(qi & 0xFF) << (8 * (i & 3));
      }
      *reinterpret_cast<uint2*>(q_out + row * H + (lane + 32 * j) * 8) = make_uint2(o[0], o[1]);
    }
  }
}

// one CTA per row; H % 8 == 0; each thread owns 8-element vectors
template <bool kAdd>
__global__ void __launch_bounds__(256) rmsnorm_kernel(__nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ residual,
                                                      const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                      int H, float eps) {
  __shared__ float red[8];
  extern __shared__ float xs[];    // [H] fp32 copy of the (summed) row
  const long long row = blockIdx.x;
  float ss = 0.f;
  for (int v = threadIdx.x; v < H / 8; v += blockDim.x) {
    uint4 a = *reinterpret_cast<const uint4*>(x + row * H + v * 8);
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __bfloat162float(pa[i]);
    if constexpr (kAdd) {
      uint4 r = *reinterpret_cast<const uint4*>(residual + row * H + v * 8);
      const __nv_bfloat16* pr = reinterpret_cast<const __nv_bfloat16*>(&r);
      uint4 o;
      __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] += __bfloat162float(pr[i]);
        po[i] = __float2bfloat16_rn(f[i]);
      }
      *reinterpret_cast<uint4*>(residual + row * H + v * 8) = o;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xs[v * 8 + i] = f[i];
      ss += f[i] * f[i];
    }
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / H + eps);
  for (int v = threadIdx.x; v < H / 8; v += blockDim.x) {
    uint4 o;
    __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
    for (int i = 0; i < 8; ++i) po[i] = __float2bfloat16_rn(xs[v * 8 + i] * inv * w[v * 8 + i]);
    *reinterpret_cast<uint4*>(out + row * H + v * 8) = o;
  }
}

// one CTA per row of x [rows][K] bf16 -> q int8, scale (f32 and/or bf16)
__global__ void __launch_bounds__(256) quant_rows_int8_kernel(const __nv_bfloat16* __restrict__ x, int8_t* __restrict__ q,
                                                              float* __restrict__ scale_f32,
                                                              __nv_bfloat16* __restrict__ scale_bf16, int rows, int K) {
  // one warp per row, 16-byte loads, 8-byte stores (K % 8 == 0); same arithmetic as the scalar form: max is order-free,
  // x / scale is an IEEE division, rintf = round half to even
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float mx = 0.f;
  for (int v = lane; v < K / 8; v += 32) {
    const uint4 a = *reinterpret_cast<const uint4*>(x + row * K + v * 8);
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
#pragma unroll
    for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(__bfloat162float(pa[i])));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float scale = fmaxf(mx, 1e-10f) / 127.0f;
  const float rscale = 1.0f / scale;
  if (lane == 0) {
    if (scale_f32) scale_f32[row] = scale;
    if (scale_bf16) scale_bf16[row] = __float2bfloat16_rn(scale);
  }
  for (int v = lane; v < K / 8; v += 32) {
    const uint4 a = *reinterpret_cast<const uint4*>(x + row * K + v * 8);
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
    uint32_t o[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int qi = quant_code_rhe(__bfloat162float(pa[i]), scale, rscale);   // torch.round: half to even
      o[i >> 2] |= (uint32_t)(qi & 0xFF) << (8 * (i & 3));
    }
    *reinterpret_cast<uint2*>(q + row * K + v * 8) = make_uint2(o[0], o[1]);
  }
}

__global__ void __launch_bounds__(256) silu_and_mul_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           long long rows, int N) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * (N / 8)) return;
  const long long r = idx / (N / 8);
  const int v = (int)(idx % (N / 8));
  const uint4 g4 = *reinterpret_cast<const uint4*>(x + r * 2 * N + v * 8);
  const uint4 u4 = *reinterpret_cast<const uint4*>(x + r * 2 * N + N + v * 8);
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(&g4);
  const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(&u4);
  uint4 o;
  __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float gv = __bfloat162float(g[i]);
    po[i] = __float2bfloat16_rn(gv / (1.0f + __expf(-gv)) * __bfloat162float(u[i]));
  }
  *reinterpret_cast<uint4*>(out + r * N + v * 8) = o;
}

// act = bf16(silu(gate) * up) exactly as silu_and_mul_kernel, quantised per row exactly as quant_rows_int8_kernel, without writing the
// BF16 activation: one warp per row, N == 256 * VPL, row kept in registers
template <int VPL>
__global__ void __launch_bounds__(256) silu_mul_quant_kernel(const __nv_bfloat16* __restrict__ x, int8_t* __restrict__ q,
                                                             float* __restrict__ scale_f32, int rows) {
  constexpr int N = 256 * VPL;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float f[VPL][8];
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int v = lane + 32 * j;
    const uint4 g4 = *reinterpret_cast<const uint4*>(x + row * 2 * N + v * 8);
    const uint4 u4 = *reinterpret_cast<const uint4*>(x + row * 2 * N + N + v * 8);
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(&g4);
    const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(&u4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gv = __bfloat162float(g[i]);
      f[j][i] = __bfloat162float(__float2bfloat16_rn(gv / (1.0f + __expf(-gv)) * __bfloat162float(u[i])));
      mx = fmaxf(mx, fabsf(f[j][i]));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float scale = fmaxf(mx, 1e-10f) / 127.0f;
  const float rscale = 1.0f / scale;
  if (lane == 0) scale_f32[row] = scale;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    uint32_t o[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int qi = quant_code_rhe(f[j][i], scale, rscale);
      o[i >> 2] |= (uint32_t)(qi & 0xFF) << (8 * (i & 3));
    }
    *reinterpret_cast<uint2*>(q + row * N + (lane + 32 * j) * 8) = make_uint2(o[0], o[1]);
  }
}

// y[m] *= bf16(sigmoid(bf16(dot(h[m], w)))) ; one warp per row, 16-byte vectors (H % 8 == 0, N % 8 == 0)
__global__ void __launch_bounds__(256) sigmoid_gate_mul_kernel(const __nv_bfloat16* __restrict__ h,
                                                               const __nv_bfloat16* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ y, int M, int H, int N) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float acc = 0.f;
  for (int v = lane; v < H / 8; v += 32) {
    const uint4 a = *reinterpret_cast<const uint4*>(h + row * H + v * 8), b = *reinterpret_cast<const uint4*>(w + v * 8);
    const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
    const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(__bfloat162float(pa[i]), __bfloat162float(pb[i]), acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float lin = __bfloat162float(__float2bfloat16_rn(acc));
  const float gv = __bfloat162float(__float2bfloat16_rn(1.0f / (1.0f + expf(-lin))));
  for (int v = lane; v < N / 8; v += 32) {
    uint4 a = *reinterpret_cast<const uint4*>(y + row * N + v * 8);
    __nv_bfloat16* pa = reinterpret_cast<__nv_bfloat16*>(&a);
#pragma unroll
    for (int i = 0; i < 8; ++i) pa[i] = __float2bfloat16_rn(gv * __bfloat162float(pa[i]));
    *reinterpret_cast<uint4*>(y + row * N + v * 8) = a;
  }
}

__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 x = reinterpret_cast<const uint4*>(a)[i], y = reinterpret_cast<const uint4*>(b)[i];
  const __nv_bfloat162* px = reinterpret_cast<const __nv_bfloat162*>(&x);
  const __nv_bfloat162* py = reinterpret_cast<const __nv_bfloat162*>(&y);
  uint4 o;
  __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) po[j] = __hadd2(px[j], py[j]);
  reinterpret_cast<uint4*>(out)[i] = o;
}
cudaError_t launch_add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t s) {
  KernelSpan ks(K_ADD, s);
  if (n % 8 || n <= 0) return cudaErrorInvalidValue;
  add_bf16_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                 (__nv_bfloat16*)out, n / 8);
  return cudaGetLastError();
}

cudaError_t launch_rmsnorm(void* x, void* residual, const float* w, void* out, int M, int H, float eps, cudaStream_t s) {
  KernelSpan ks(K_RMSNORM, s);
  if (H % 8 || M <= 0) return cudaErrorInvalidValue;
  if (H == 2048) {
    if (residual)
      rmsnorm_warp_kernel<true, 8><<<(M + 7) / 8, 256, 0, s>>>((__nv_bfloat16*)x, (__nv_bfloat16*)residual, w, (__nv_bfloat16*)out, M, eps);
    else
      rmsnorm_warp_kernel<false, 8><<<(M + 7) / 8, 256, 0, s>>>((__nv_bfloat16*)x, nullptr, w, (__nv_bfloat16*)out, M, eps);
    return cudaGetLastError();
  }
  if (residual)
    rmsnorm_kernel<true><<<M, 256, H * sizeof(float), s>>>((__nv_bfloat16*)x, (__nv_bfloat16*)residual, w, (__nv_bfloat16*)out, H, eps);
  else
    rmsnorm_kernel<false><<<M, 256, H * sizeof(float), s>>>((__nv_bfloat16*)x, nullptr, w, (__nv_bfloat16*)out, H, eps);
  return cudaGetLastError();
}
// rmsnorm (+ residual add) that also emits the INT8 row quantisation of its output; false = geometry not covered by the fused kernel
bool rmsnorm_q8_supported(int H) { return H == 2048; }
cudaError_t launch_rmsnorm_q8(void* x, void* residual, const float* w, void* out, void* q, float* q_scale, int M, int H, float eps,
                              cudaStream_t s) {
  KernelSpan ks(K_RMSNORM, s);
  if (!rmsnorm_q8_supported(H) || M <= 0) return cudaErrorInvalidValue;
  if (residual)
    rmsnorm_warp_kernel<true, 8, true><<<(M + 7) / 8, 256, 0, s>>>((__nv_bfloat16*)x, (__nv_bfloat16*)residual, w, (__nv_bfloat16*)out, M,
                                                                   eps, (int8_t*)q, q_scale);
  else
    rmsnorm_warp_kernel<false, 8, true><<<(M + 7) / 8, 256, 0, s>>>((__nv_bfloat16*)x, nullptr, w, (__nv_bfloat16*)out, M, eps, (int8_t*)q,
                                                                    q_scale);
  return cudaGetLastError();
}
bool silu_mul_quant_supported(int N) { return N == 256 || N == 512 || N == 1024 || N == 2048; }
cudaError_t launch_silu_mul_quant(const void* x, void* q, float* scale_f32, int rows, int N, cudaStream_t s) {
  KernelSpan ks(K_SILU_MUL, s);
  if (rows <= 0) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  if (N == 256) silu_mul_quant_kernel<1><<<grid, 256, 0, s>>>(xb, (int8_t*)q, scale_f32, rows);
  else if (N == 512) silu_mul_quant_kernel<2><<<grid, 256, 0, s>>>(xb, (int8_t*)q, scale_f32, rows);
  else if (N == 1024) silu_mul_quant_kernel<4><<<grid, 256, 0, s>>>(xb, (int8_t*)q, scale_f32, rows);
  else if (N == 2048) silu_mul_quant_kernel<8><<<grid, 256, 0, s>>>(xb, (int8_t*)q, scale_f32, rows);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_quant_rows_int8(const void* x, void* q, float* scale_f32, void* scale_bf16, int rows, int K, cudaStream_t s) {
  KernelSpan ks(K_QUANT_ROWS, s);
  if (rows <= 0 || K % 8) return cudaErrorInvalidValue;
  quant_rows_int8_kernel<<<(rows + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)x, (int8_t*)q, scale_f32, (__nv_bfloat16*)scale_bf16,
                                                        rows, K);
  return cudaGetLastError();
}
cudaError_t launch_silu_and_mul(const void* x, void* out, int rows, int N, cudaStream_t s) {
  KernelSpan ks(K_SILU_MUL, s);
  if (N % 8 || rows <= 0) return cudaErrorInvalidValue;
  const long long total = (long long)rows * (N / 8);
  silu_and_mul_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, N);
  return cudaGetLastError();
}
cudaError_t launch_sigmoid_gate_mul(const void* h, const void* w, void* y, int M, int H, int N, cudaStream_t s) {
  KernelSpan ks(K_SIGMOID_GATE, s);
  if (M <= 0 || H % 8 || N % 8) return cudaErrorInvalidValue;
  sigmoid_gate_mul_kernel<<<(M + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)h, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, M, H, N);
  return cudaGetLastError();
}

}  // namespace kb2
