// Process-wide kernel accounting: every launch site of the library opens a KernelSpan.  It always counts the launch
// (kb2_total_launches, what bench.py reports as gpu_launches) and, while kb2_kernel_profile_enable(1) is in effect, brackets
// the launch with CUDA events on the launching stream so that bench.py can print a per-kernel time table measured live
// (the reference's KRASIS_LAYER_TIMING / EP breakdown, python/krasis/model.py:2839-2860, at kernel granularity).
#pragma once
#include <cuda_runtime.h>

namespace kb2 {

enum KernelId : int {
  K_ROUTER_GEMM = 0, K_ROUTER_TOPK, K_BINNING, K_GEMM1_GATE_UP, K_GEMM2_DOWN, K_COMBINE, K_DENSE_BF16, K_DENSE_I8,
  K_GDN_PREP, K_GDN_CONV_STATE, K_GDN_PREPARE, K_GDN_SCAN, K_GDN_POST, K_GQA_PREP, K_KV_GATHER, K_FMHA, K_MLA_PREP,
  K_RMSNORM, K_QUANT_ROWS, K_SILU_MUL, K_SIGMOID_GATE, K_ADD, K_QUANTIZE_GROUP, K_RETILE, K_NUM
};

void kernel_span_begin(int id, cudaStream_t s, int n_launches, int* slot);
void kernel_span_end(int id, cudaStream_t s, int slot);

struct KernelSpan {
  int id, slot = -1;
  cudaStream_t s;
  KernelSpan(int id_, cudaStream_t s_, int n_launches = 1) : id(id_), s(s_) { kernel_span_begin(id, s, n_launches, &slot); }
  ~KernelSpan() { if (slot >= 0) kernel_span_end(id, s, slot); }
  KernelSpan(const KernelSpan&) = delete;
  KernelSpan& operator=(const KernelSpan&) = delete;
};

}  // namespace kb2
