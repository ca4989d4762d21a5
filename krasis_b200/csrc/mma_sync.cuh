// Warp-level mma.sync m16n8k8 TF32 tiles with fp32-grade accuracy (3xTF32), shared by gdn.cu (round-1 Gated-DeltaNet kernels) and
// gdn_tc.cu (the blocked triangular inverse of the tcgen05 chunk-prepare kernel).
#pragma once
#include <cstdint>

#include <cuda_bf16.h>

namespace kb2 {

// ------------------------------------------------------------------------------------------------
// Warp-level 16x8 output tiles on mma.sync m16n8k8 TF32 with fp32-grade accuracy: an fp32 operand is split into
// hi = tf32(x), lo = tf32(x - hi) and the product is a_hi*b_hi + a_hi*b_lo + a_lo*b_hi ("3xTF32"); operands that hold
// BF16 values (q, k) are exact in TF32 and need no split.  The delta-rule recurrences stay at fp32 accuracy like the
// reference (python/krasis/linear_attention.py:776-779 casts everything to float32) while the small per-chunk matmuls
// leave the CUDA cores.  (These 64x64x128 blocks are below the tcgen05 M=64 tile and strictly sequential across chunks.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// operand element -> TF32 hi (and lo when SPLIT); BF16 bits shifted up ARE a TF32 value (no cvt, never split)
// fp32 -> (hi, lo) with hi = x rounded to TF32 by integer arithmetic (cvt.rna.tf32 costs ~4 instructions) and
// lo = x - hi (exact, <= 12 significant bits; the tensor core ignores the low 13 mantissa bits of a .tf32 register,
// so at most one bit of lo is dropped: |x - hi - lo| <= 2^-23 |x|)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
template <bool SPLIT>
__device__ __forceinline__ void ld_tf32(const float* p, uint32_t& hi, uint32_t& lo) {
  if (SPLIT) split_tf32(*p, hi, lo);
  else hi = to_tf32(*p);
}
template <bool SPLIT>
__device__ __forceinline__ void ld_tf32(const __nv_bfloat16* p, uint32_t& hi, uint32_t& lo) {
  static_assert(!SPLIT, "BF16 operands are exact in TF32");
  hi = (uint32_t)(*reinterpret_cast<const unsigned short*>(p)) << 16;
}
// c[NT][4] (16 rows x NT*8 cols) += A[16 x K] * B[K x NT*8];  A(r,k) = A[r*sar + k*sak], B(k,n) = B[k*sbk + n*sbn]
template <int NT, bool SPLIT_A, bool SPLIT_B, typename TA, typename TB>
__device__ __forceinline__ void warp_mma_tiles(float (&c)[NT][4], const TA* __restrict__ A, int sar, int sak,
                                               const TB* __restrict__ B, int sbk, int sbn, int k_begin, int k_end) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int k0 = k_begin; k0 < k_end; k0 += 8) {
    uint32_t ah[4], al[4];
    ld_tf32<SPLIT_A>(A + g * sar + (k0 + t) * sak, ah[0], al[0]);
    ld_tf32<SPLIT_A>(A + (g + 8) * sar + (k0 + t) * sak, ah[1], al[1]);
    ld_tf32<SPLIT_A>(A + g * sar + (k0 + t + 4) * sak, ah[2], al[2]);
    ld_tf32<SPLIT_A>(A + (g + 8) * sar + (k0 + t + 4) * sak, ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bh[2], bl[2];
      ld_tf32<SPLIT_B>(B + (k0 + t) * sbk + (nt * 8 + g) * sbn, bh[0], bl[0]);
      ld_tf32<SPLIT_B>(B + (k0 + t + 4) * sbk + (nt * 8 + g) * sbn, bh[1], bl[1]);
      if (SPLIT_A) mma_tf32(c[nt], al, bh);
      if (SPLIT_B) mma_tf32(c[nt], ah, bl);
      mma_tf32(c[nt], ah, bh);
    }
  }
}

}  // namespace kb2
