/*
 * krasis_b200 — C ABI of the B200-native MoE prefill path (drop-in boundary).
 *
 * Every entry point replaces one reference interface on the prefill hot path; the citation after
 * each declaration is the reference call site / PyO3 method it stands in for (paths relative to the
 * reference tree).  Conventions:
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - `*_dev` pointers are device memory owned by the caller, `*_host` pointers are host memory;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*) and performs no hidden
 *     device synchronisation, except the `_host` convenience calls which synchronise before returning;
 *   - return value 0 = ok; non-zero = error, text via kb2_last_error() (thread-local).
 *     KB2_ERR_STATE mirrors the reference's PyRuntimeError ("Model not loaded", "GPU weights not
 *     available"), KB2_ERR_VALUE mirrors PyValueError for size/shape mismatch
 *     (src/moe.rs:1543-1550,1790-1803,2285-2300), KB2_ERR_CUDA wraps a CUDA runtime error.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with KB2_ERR_CUDA.
 */
#ifndef KRASIS_B200_H
#define KRASIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KB2_API __attribute__((visibility("default")))
#else
#define KB2_API
#endif

#define KB2_OK 0
#define KB2_ERR_STATE 1
#define KB2_ERR_VALUE 2
#define KB2_ERR_CUDA 3

#define KB2_SCORE_SOFTMAX 0 /* softmax over all experts, then top-k   (python/krasis/layer.py:548-558) */
#define KB2_SCORE_SIGMOID 1 /* sigmoid (+ selection-only bias)         (python/krasis/layer.py:540-547) */
#define KB2_SCORE_TOPK_SOFTMAX 2 /* GPT-OSS: top-k on logits, softmax of the k (layer.py:536-539) */

#define KB2_FMT_INT4_G128 0 /* Krasis symmetric INT4, group 128 (src/weights/marlin.rs:145-207) */
#define KB2_FMT_INT8_G128 1 /* Krasis symmetric INT8, group 128 (src/weights/marlin.rs:65-114)  */
#define KB2_FMT_GGUF_Q8_0 2 /* native GGUF Q8_0 blocks (src/gguf.rs:574-593), losslessly re-tiled */
#define KB2_FMT_GGUF_Q4_K 3 /* native GGUF Q4_K super-blocks (src/gguf.rs:681-738), losslessly re-tiled */
/* The remaining expert types of real GGUF files (Q4_K_M pairs Q4_K gate/up with Q6_K down; 32-element fallbacks when K is not
 * a multiple of 256).  Decoded at load time — losslessly: same values, same single f32 rounding as
 * src/gguf.rs — into int8 codes + one (a, b) f32 pair per 16 elements, w = fma(a, code, -b) inside the grouped GEMM. */
#define KB2_FMT_GGUF_Q6_K 4 /* src/gguf.rs:813-866 (scale index as in src/gguf_kernels.rs:594-635 / ggml) */
#define KB2_FMT_GGUF_Q5_K 5 /* src/gguf.rs:740-811 */
#define KB2_FMT_GGUF_Q5_0 6 /* src/gguf.rs:599-633 */
#define KB2_FMT_GGUF_Q4_0 7 /* src/gguf.rs:635-664 */

typedef struct kb2_engine kb2_engine;

/* Mirrors the constructor arguments of GpuPrefillManager (python/krasis/gpu_prefill.py:326-346) plus
 * the routing configuration KrasisEngine.set_routing_config takes (src/moe.rs:2959). */
typedef struct kb2_config {
  int32_t hidden_size;            /* H */
  int32_t moe_intermediate_size;  /* I */
  int32_t n_routed_experts;       /* E (global) */
  int32_t num_experts_per_tok;    /* top-k */
  int32_t num_moe_layers;
  int32_t weight_format;          /* KB2_FMT_* */
  int32_t rank;                   /* EP: this engine owns experts [rank*(E/R), (rank+1)*(E/R)), last rank takes */
  int32_t num_ranks;              /*     the remainder — python/krasis/gpu_prefill.py:353-359                   */
  int32_t scoring_func;           /* KB2_SCORE_* */
  int32_t norm_topk_prob;         /* 0/1 */
  float routed_scaling_factor;
  int32_t max_tokens;             /* largest M a forward call may pass (sizes the scratch buffers) */
  int32_t device;                 /* CUDA device ordinal; forward calls cudaSetDevice(device) like gpu_prefill.py:4401 */
  int32_t w2_weight_format;       /* KB2_FMT_* of the down projection, or -1 = same as weight_format (GGUF files mix types:
                                     Q4_K gate/up with Q8_0 / Q6_K down, src/weights/mod.rs:646-647) */
} kb2_config;

/* lifecycle — KrasisEngine(...) + GpuPrefillManager(...) (src/moe.rs:1482, gpu_prefill.py:326) */
KB2_API int kb2_create(const kb2_config* cfg, kb2_engine** out);
KB2_API void kb2_destroy(kb2_engine* e);
KB2_API const char* kb2_last_error(void);
KB2_API const char* kb2_version(void);

/* introspection — KrasisEngine.num_moe_layers/hidden_size/num_experts/top_k/group_size/intermediate_size/
 * gpu_num_bits (src/moe.rs:1874-1965) and GpuPrefillManager.expert_start/expert_end */
KB2_API int kb2_get_config(const kb2_engine* e, kb2_config* out);
KB2_API int kb2_expert_range(const kb2_engine* e, int32_t* expert_start, int32_t* expert_end);

/* Bytes of one layer's LOCAL experts in the B200 tile layout: which = 0 w13 packed, 1 w13 scales,
 * 2 w2 packed, 3 w2 scales.  (Counterpart of the per-expert sizes in src/weights/mod.rs:955-970.) */
KB2_API size_t kb2_tiled_bytes(const kb2_engine* e, int which);

/* Weight hand-off, host side — replaces KrasisEngine.get_expert_{w13_packed,w13_scales,w2_packed,
 * w2_scales} / write_experts_range_into_pinned (src/moe.rs:1972-2097,2431) followed by the H2D copy in
 * GpuPrefillManager (gpu_prefill.py:1059-1062).  Inputs are the reference QUANTISER's outputs for the
 * local experts (src/weights/marlin.rs:145-207, 65-114), row-major [E_local][N][K/8] u32 (INT4) or
 * [E_local][N][K] i8 (INT8) with scales [E_local][N][K/128] bf16; w13 = [gate rows ; up rows]
 * (src/weights/mod.rs:346-349).  The engine re-tiles them on the device and owns the result. */
KB2_API int kb2_load_experts_host(kb2_engine* e, int moe_layer_idx, const void* w13_q_host, const void* w13_s_host,
                          const void* w2_q_host, const void* w2_s_host);

/* The Krasis group quantiser itself, on the device and bit-exact: quantize_int4 / quantize_int8
 * (src/weights/marlin.rs:145-207, 65-114).  w [rows][K] bf16 -> q ([rows][K/8] u32 for 4 bits | [rows][K] i8 for 8 bits) and
 * scales [rows][K/128] raw bf16 — the reference quantiser's output layout.  K % 128 == 0 (marlin.rs:152). */
KB2_API int kb2_quantize_group_dev(const void* w_bf16_dev, int32_t num_bits, void* q_dev, void* scales_dev, int64_t rows,
                                   int32_t k_cols, int32_t device, void* stream);
/* kb2_load_experts_host with DEVICE pointers (e.g. straight out of kb2_quantize_group_dev): re-tiles and owns the result. */
KB2_API int kb2_load_experts_dev(kb2_engine* e, int moe_layer_idx, const void* w13_q_dev, const void* w13_s_dev,
                                 const void* w2_q_dev, const void* w2_s_dev, void* stream);

/* GGUF expert tensors kept as native blocks (new capability; in the reference GGUF feeds only the CPU experts,
 * src/weights/mod.rs:3251-3585, src/gguf_kernels.rs:690-756).  gate/up: [E_local][I][row_bytes(H)], down: [E_local][H][row_bytes(I)]
 * with the block types given by weight_format / w2_weight_format; rows are [N][K] with K contiguous in blocks
 * (src/gguf_kernels.rs:9).  Merged `ffn_*_exps` tensors are exactly this layout per layer (src/weights/mod.rs:3455-3487). */
KB2_API int kb2_load_experts_gguf_host(kb2_engine* e, int moe_layer_idx, const void* gate_host, const void* up_host,
                                       const void* down_host);

/* Same, but the caller already holds device buffers in the B200 tile layout (sizes = kb2_tiled_bytes);
 * the engine only records the pointers (caller keeps ownership).  Used to build synthetic full-size
 * models without a host copy — the analogue of the persistent-expert buffers gpu_prefill.py:4128-4137. */
KB2_API int kb2_attach_experts_tiled_dev(kb2_engine* e, int moe_layer_idx, const void* w13_q_dev, const void* w13_s_dev,
                                 const void* w2_q_dev, const void* w2_s_dev);

/* B200-layout expert cache (krasis_b200/tile_cache.py; the counterpart of the reference's Marlin cache file,
 * src/weights/mod.rs:857-893,2462-2476,4117-4144): the four tiled buffers of one layer's LOCAL experts, host side.
 * kb2_export_experts_tiled_host copies buffer `which` (sizes = kb2_tiled_bytes; a size mismatch is KB2_ERR_VALUE like
 * write_experts_*_into, src/moe.rs:2285-2300) to host memory and synchronises; kb2_load_experts_tiled_host uploads four such
 * buffers and owns the device copies — no quantisation or re-tiling happens on this path. */
KB2_API int kb2_export_experts_tiled_host(kb2_engine* e, int moe_layer_idx, int which, void* dst_host, size_t dst_bytes);
KB2_API int kb2_load_experts_tiled_host(kb2_engine* e, int moe_layer_idx, const void* w13_q_host, const void* w13_s_host,
                                        const void* w2_q_host, const void* w2_s_host);

/* Re-tile on the device without attaching (exposed so tests can check the layout transform). */
KB2_API int kb2_retile_dev(kb2_engine* e, int weight_format, const void* src_q_dev, const void* src_s_dev, void* dst_q_dev,
                   void* dst_s_dev, int n_experts, int n_rows, int k_cols, void* stream);

/* Router weights — KrasisEngine.set_routing_weights(layer, gate_bf16, bias_f32) (src/moe.rs:2989) and the
 * layer attributes gate_weight / gate_bias / e_score_correction_bias (python/krasis/layer.py:526-560).
 * gate: [E][H] bf16; gate_bias, e_score_correction_bias: [E] f32 or NULL. */
KB2_API int kb2_set_router_host(kb2_engine* e, int moe_layer_idx, const void* gate_bf16_host, const float* gate_bias_host,
                        const float* e_score_correction_bias_host);

/* TransformerLayer.compute_routing (python/krasis/layer.py:526-560): hidden [M][H] bf16 ->
 * topk_ids [M][k] int32 (descending score, ties -> lower expert index), topk_weights [M][k] f32. */
KB2_API int kb2_route(kb2_engine* e, int moe_layer_idx, const void* hidden_dev, int32_t num_tokens, int32_t* topk_ids_dev,
              float* topk_weights_dev, void* stream);

/* GpuPrefillManager.forward(moe_layer_idx, hidden_states, topk_ids, topk_weights, routed_only)
 * (python/krasis/gpu_prefill.py:4374-4484): x [M][H] bf16, ids [M][k] int32 (GLOBAL expert ids; ids outside
 * this engine's expert range, or negative, contribute zero — gpu_prefill.py:4140-4149, src/moe.rs:2722),
 * weights [M][k] f32 -> out [M][H] bf16.  routed_only=1 returns the raw routed sum (EP partial sums);
 * otherwise out = bf16(rsf * routed) (+ shared_dev [M][H] bf16 if not NULL), gpu_prefill.py:4467-4482. */
KB2_API int kb2_moe_forward(kb2_engine* e, int moe_layer_idx, const void* x_dev, const int32_t* topk_ids_dev,
                    const float* topk_weights_dev, void* out_dev, int32_t num_tokens, int32_t routed_only,
                    const void* shared_dev, void* stream);

/* ---- Expert-parallel building blocks (one process per GPU; the collective between them is the caller's NCCL
 * all-to-all).  They replace the reference's replicated-token EP loop (python/krasis/model.py:3086-3211) with
 * dispatch of ROUTED ROWS ONLY:
 *   source rank : kb2_route -> kb2_ep_bin_rows (rows grouped by GLOBAL expert id, hence by owner rank)
 *                 -> all-to-all(rows, weights, ids) ->
 *   owner rank  : kb2_moe_forward_rows (one routing entry per row; returns bf16(w * expert(row)) in input order)
 *                 -> all-to-all back ->
 *   source rank : kb2_ep_combine_rows (sum over the k entries of each token in k order, rsf, + shared)
 * kb2_ep_bin_rows: x [M][H], ids/w [M][k] -> x_sorted [M*k][H], w_sorted/ids_sorted [M*k], slot_of [M][k]
 * (position of each (token, j) in the sorted order, -1 never occurs here), counts [E] int32. */
KB2_API int kb2_ep_bin_rows(kb2_engine* e, const void* x_dev, const int32_t* topk_ids_dev, const float* topk_weights_dev,
                            int32_t num_tokens, void* x_sorted_dev, float* w_sorted_dev, int32_t* ids_sorted_dev,
                            int32_t* slot_of_dev, int32_t* counts_dev, void* stream);
/* rows [n][H] bf16, expert_ids [n] int32 (GLOBAL ids; rows whose expert is not local yield zeros), weights [n] f32
 * -> out_rows [n][H] bf16.  n <= max_tokens * top_k. */
KB2_API int kb2_moe_forward_rows(kb2_engine* e, int moe_layer_idx, const void* rows_dev, const int32_t* expert_ids_dev,
                                 const float* weights_dev, void* out_rows_dev, int32_t n_rows, void* stream);
KB2_API int kb2_ep_combine_rows(kb2_engine* e, const void* rows_sorted_dev, const int32_t* slot_of_dev, int32_t num_tokens,
                                int32_t routed_only, const void* shared_dev, void* out_dev, void* stream);

/* The tail of GpuPrefillManager.forward applied to an already reduced routed sum (EP: after the reduce-scatter of the
 * partial sums): out = bf16(rsf * routed) (+ shared_dev) — python/krasis/gpu_prefill.py:4467-4482.  out may alias routed. */
KB2_API int kb2_finish_routed(kb2_engine* e, const void* routed_dev, const void* shared_dev, void* out_dev, int32_t num_tokens,
                              void* stream);
/* Expert parallelism with the reduce-scatter fused into the combine kernel (replaces kb2_moe_forward(routed_only=1) +
 * kb2_comm_reduce_scatter_bf16 + kb2_finish_routed; reference semantics: python/krasis/model.py:3086-3211).
 * peer_recv_host[r] = rank r's receive buffer as seen from this rank (kb2_comm_peer_alloc), layout [M / num_ranks][num_ranks][H] bf16.
 * Every rank: kb2_moe_forward_scatter(all M tokens, local expert slice) -> kb2_comm_barrier -> kb2_finish_routed_slots(own buffer). */
KB2_API int kb2_moe_forward_scatter(kb2_engine* e, int moe_layer_idx, const void* hidden_dev, const int32_t* topk_ids_dev,
                                    const float* topk_weights_dev, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank,
                                    int32_t num_tokens, void* stream);
KB2_API int kb2_finish_routed_slots(kb2_engine* e, const void* slots_dev, int32_t num_ranks, const void* shared_dev, void* out_dev,
                                    int32_t rows, void* stream);

/* Host-buffer variant of route + forward for callers that hold activations in (pinned) host memory —
 * the shape of KrasisEngine.submit_forward/sync_forward (src/moe.rs:2722,2809: bytes in, bytes out).
 * If topk_ids_host is NULL the engine routes with its own router weights.  Copies H2D, computes,
 * copies D2H and synchronises `stream` before returning. */
KB2_API int kb2_moe_forward_host(kb2_engine* e, int moe_layer_idx, const void* x_host, const int32_t* topk_ids_host,
                         const float* topk_weights_host, void* out_host, int32_t num_tokens, int32_t routed_only,
                         void* stream);

/* Prefill through a stack of MoE layers with HOST buffers: H2D of x [M][H] bf16, then for every layer in
 * [first_layer, first_layer + n_layers): on-device routing (kb2_route) + kb2_moe_forward of the SAME x, then
 * D2H of the last layer's output and a stream synchronise.  This is the shape of the reference's per-layer
 * prefill loop (python/krasis/model.py:2944-2955 -> layer.py:562-723) restricted to the MoE blocks; it exists
 * so end-to-end throughput can be measured through the C ABI with host<->device copies inside the call. */
KB2_API int kb2_prefill_moe_stack_host(kb2_engine* e, const void* x_host, void* out_host, int32_t num_tokens,
                                       int32_t first_layer, int32_t n_layers, void* stream);

/* Introspection for tests / profiling: after a forward, copies the per-local-expert token counts
 * (int32 [E_local]) of the last call to host. */
KB2_API int kb2_last_expert_counts(kb2_engine* e, int32_t* counts_host, void* stream);

/* Number of kernels this library has launched since creation (bench.py reports it as gpu_launches). */
KB2_API int64_t kb2_launch_count(const kb2_engine* e);

/* ================================================================================================
 * Attention blocks that feed the MoE path.
 * ================================================================================================ */

/* torch.nn.functional.linear for BF16 weights (the reference keeps attention projections in BF16,
 * python/krasis/attention.py:526-529, config.py:209): out[M][N] = x[M][K] . w[N][K]^T (+ bias[N] f32),
 * fp32 accumulate, BF16 (out_f32=0) or FP32 output.  K % 64 == 0, N % 16 == 0. */
KB2_API int kb2_linear_bf16(const void* x_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t M,
                            int32_t N, int32_t K, int32_t out_f32, int32_t device, void* stream);

/* flashinfer.norm.rmsnorm (residual_dev == NULL) / fused_add_rmsnorm (residual_dev != NULL: residual += x, in place)
 * as used at python/krasis/layer.py:163-183,283-308.  x, residual, out: [M][H] bf16 (out may alias x); weight [H] f32. */
KB2_API int kb2_rmsnorm(void* x_dev, void* residual_dev, const float* weight_dev, void* out_dev, int32_t M, int32_t H,
                        float eps, int32_t device, void* stream);
/* Per-row symmetric INT8 quantisation: quantize_to_int8 for weights (python/krasis/weight_loader.py:25-43, bf16 scale
 * out) and the activation half of int8_linear (:66-70, f32 scale out). */
KB2_API int kb2_quantize_rows_int8(const void* x_dev, void* q_dev, float* scale_f32_dev, void* scale_bf16_dev, int32_t rows,
                                   int32_t K, int32_t device, void* stream);
/* int8_linear (python/krasis/weight_loader.py:46-99): W8A8, exact INT32 accumulation on tcgen05 kind::i8, then
 * bf16(float(acc) * (x_scale * w_scale)).  Scratch: xq [M][K] int8, xs [M] f32. */
KB2_API int kb2_int8_linear(const void* x_dev, const void* wq_dev, const void* w_scale_bf16_dev, void* out_dev, void* xq_scratch_dev,
                            float* xs_scratch_dev, int32_t M, int32_t N, int32_t K, int32_t device, void* stream);
/* flashinfer.activation.silu_and_mul (layer.py:512): x [rows][2N] -> out [rows][N] bf16. */
/* Fused forms of the shared expert's W8A8 chain (layer.py:508-524 + weight_loader.py:46-99), bit-identical to the unfused calls:
 *   kb2_rmsnorm_q8            kb2_rmsnorm that also emits the INT8 row quantisation (q [M][H] int8, q_scale [M] f32) of its output
 *   kb2_int8_linear_q8        kb2_int8_linear on an already quantised activation
 *   kb2_silu_mul_int8_linear  out = int8_linear(silu(gate) * up, W): the activation is quantised straight from the [M][2K] gate|up rows
 *                             (act_scratch [M][K] bf16 is only touched for K outside {256, 512, 1024, 2048}) */
KB2_API int kb2_rmsnorm_q8(void* x_dev, void* residual_dev, const float* weight_dev, void* out_dev, void* q_dev, float* q_scale_dev,
                           int32_t M, int32_t H, float eps, int32_t device, void* stream);
KB2_API int kb2_int8_linear_q8(const void* xq_dev, const float* xs_dev, const void* wq_dev, const void* w_scale_bf16_dev, void* out_dev,
                               int32_t M, int32_t N, int32_t K, int32_t device, void* stream);
KB2_API int kb2_silu_mul_int8_linear(const void* gate_up_dev, const void* wq_dev, const void* w_scale_bf16_dev, void* out_dev,
                                     void* act_scratch_dev, void* xq_scratch_dev, float* xs_scratch_dev, int32_t M, int32_t N, int32_t K,
                                     int32_t device, void* stream);
KB2_API int kb2_silu_and_mul(const void* x_dev, void* out_dev, int32_t rows, int32_t N, int32_t device, void* stream);
/* Qwen3-Next shared-expert gate (layer.py:518-522): y[m][:] *= sigmoid(hidden[m] . gate_w), y [M][N], gate_w [H] bf16. */
KB2_API int kb2_sigmoid_gate_mul(const void* hidden_dev, const void* gate_w_dev, void* y_dev, int32_t M, int32_t H, int32_t N,
                                 int32_t device, void* stream);

/* out = a + b on BF16 tensors (the `output + shared` add of python/krasis/layer.py:684-685 in the EP path). */
KB2_API int kb2_add_bf16(const void* a_dev, const void* b_dev, void* out_dev, int64_t n, int32_t device, void* stream);

/* Gated DeltaNet linear attention — python/krasis/linear_attention.py:GatedDeltaNetAttention (prefill path
 * `_forward_chunked`, :695-844).  One handle holds the weights and the per-layer conv / recurrent state the
 * reference keeps on the object (:180-205); kb2_gdn_forward == forward(hidden, is_decode=False) for M tokens and
 * carries the state to the next call; kb2_gdn_reset_state == reset_state() (:207-214). */
typedef struct kb2_gdn kb2_gdn;
typedef struct kb2_gdn_config {
  int32_t hidden_size;      /* cfg.hidden_size */
  int32_t num_k_heads;      /* cfg.linear_num_key_heads   (16) */
  int32_t num_v_heads;      /* cfg.linear_num_value_heads (32) */
  int32_t k_head_dim;       /* cfg.linear_key_head_dim    (128) */
  int32_t v_head_dim;       /* cfg.linear_value_head_dim  (128) */
  int32_t conv_kernel;      /* cfg.linear_conv_kernel_dim (4) */
  float rms_norm_eps;
  int32_t max_tokens;
  int32_t num_layers;
  int32_t device;
} kb2_gdn_config;
KB2_API int kb2_gdn_create(const kb2_gdn_config* cfg, kb2_gdn** out);
KB2_API void kb2_gdn_destroy(kb2_gdn* h);
/* All weights BF16, host pointers, shapes as in the HF checkpoint (linear_attention.py:131-139):
 * in_proj_qkvz [2*kd+2*vd][H], in_proj_ba [2*nv][H], conv1d.weight [2*kd+vd][1][K], A_log [nv], dt_bias [nv],
 * norm.weight [dv], out_proj [H][vd]. */
KB2_API int kb2_gdn_set_weights_host(kb2_gdn* h, int layer, const void* in_proj_qkvz, const void* in_proj_ba,
                                     const void* conv1d_weight, const void* A_log, const void* dt_bias,
                                     const void* norm_weight, const void* out_proj);
KB2_API int kb2_gdn_forward(kb2_gdn* h, int layer, const void* hidden_dev, void* out_dev, int32_t num_tokens, void* stream);
KB2_API int kb2_gdn_reset_state(kb2_gdn* h, int layer, void* stream);
/* conv state [C][K] bf16 and recurrent state [nv][dk][dv] f32 (either pointer may be NULL); synchronises. */
KB2_API int kb2_gdn_get_state_host(kb2_gdn* h, int layer, void* conv_state_bf16_host, float* recurrent_state_host);

/* GQA attention — python/krasis/attention.py:GQAAttention.forward (:496-687): q/k/v projections (BF16), gated-q split,
 * per-head RMSNorm, partial half-split RoPE, FP8-E4M3 paged KV append (page = 16 tokens, NHD, unscaled cast), causal
 * attention over the paged cache, sigmoid output gate, o_proj.  The KV cache tensors stay with the caller, exactly like
 * PagedKVCache.get_gqa_layer_caches(layer_offset) + SequenceKVState.kv_indices (python/krasis/kv_cache.py). */
typedef struct kb2_gqa kb2_gqa;
typedef struct kb2_gqa_config {
  int32_t hidden_size;
  int32_t num_heads;        /* cfg.num_attention_heads */
  int32_t num_kv_heads;     /* cfg.num_key_value_heads */
  int32_t head_dim;         /* cfg.gqa_head_dim: 128 or 256 */
  int32_t rotary_dim;       /* int(head_dim * partial_rotary_factor) (config.py:469-473) */
  int32_t gated;            /* q_proj emits [q | gate] per head (attention.py:398-406) */
  float rope_theta;
  float rms_norm_eps;
  int32_t page_size;        /* 16 (kv_cache.py:26) */
  int32_t max_tokens;
  int32_t num_layers;
  int32_t device;
} kb2_gqa_config;
KB2_API int kb2_gqa_create(const kb2_gqa_config* cfg, kb2_gqa** out);
KB2_API void kb2_gqa_destroy(kb2_gqa* h);
/* BF16 host weights: q_proj [nh*d*(1+gated)][H], k_proj/v_proj [nkv*d][H], o_proj [H][nh*d], q_norm/k_norm [d] or NULL
 * (already shifted by +1 for Qwen3-Next, like the reference does at load, weight_loader.py:259-265). */
KB2_API int kb2_gqa_set_weights_host(kb2_gqa* h, int layer, const void* q_proj, const void* k_proj, const void* v_proj,
                                     const void* o_proj, const void* q_norm, const void* k_norm);
/* hidden [M][H] bf16; positions [M] int32 = first_position .. first_position+M-1 (prefill appends contiguously);
 * k/v cache of THIS layer: [num_pages][16][nkv][d] FP8-E4M3; kv_indices: page ids of the sequence covering
 * kv_len_after = first_position + M tokens; out [M][H] bf16. */
KB2_API int kb2_gqa_forward(kb2_gqa* h, int layer, const void* hidden_dev, const int32_t* positions_dev, int32_t first_position,
                            void* k_cache_layer_dev, void* v_cache_layer_dev, const int32_t* kv_indices_dev,
                            int32_t kv_len_after, void* out_dev, int32_t num_tokens, void* stream);

/* ---- MLA prefill (python/krasis/attention.py:45-374, MLAAttention) ------------------------------------------------- */
typedef struct kb2_mla kb2_mla;
typedef struct kb2_mla_config {
  int32_t hidden_size;
  int32_t num_heads;        /* cfg.num_attention_heads */
  int32_t qk_nope_head_dim; /* 128 */
  int32_t qk_rope_head_dim; /* 64 */
  int32_t v_head_dim;       /* 128 */
  int32_t kv_lora_rank;     /* 512 */
  int32_t q_lora_rank;      /* 0 = direct q_proj (V2-Lite); > 0 = q_a_proj -> layernorm -> q_b_proj (attention.py:248-259) */
  float rms_norm_eps;
  float sm_scale;           /* 1/sqrt(nope+rope) * yarn mscale^2 (attention.py:79-88) */
  int32_t page_size;        /* 16 */
  int32_t max_tokens;       /* rows of one forward call */
  int32_t max_kv_len;       /* longest sequence (cached + new tokens) a call may attend over */
  int32_t num_layers;
  int32_t device;
} kb2_mla_config;
KB2_API int kb2_mla_create(const kb2_mla_config* cfg, kb2_mla** out);
KB2_API void kb2_mla_destroy(kb2_mla* h);
/* BF16 host weights (attention.py:90-107): q_proj [nh*(nope+rope)][H] when q_lora_rank == 0 (q_a_* NULL), else q_a_proj
 * [q_lora][H], q_a_layernorm [q_lora], q_b_proj [nh*(nope+rope)][q_lora]; kv_a_proj_with_mqa [lora+rope][H];
 * kv_a_layernorm [lora]; w_kc [nh][nope][lora]; w_vc [nh][dv][lora]; o_proj [H][nh*dv].
 * rope_inv_freq: fp32 [rope/2] — the (YaRN-blended) inverse frequencies of attention.py:129-157. */
KB2_API int kb2_mla_set_weights_host(kb2_mla* h, int layer, const void* q_proj_or_q_b_proj, const void* q_a_proj,
                                     const void* q_a_layernorm, const void* kv_a_proj_with_mqa, const void* kv_a_layernorm,
                                     const void* w_kc, const void* w_vc, const void* o_proj, const float* rope_inv_freq);
/* hidden [M][H] bf16; positions [M] int32 = first_position .. first_position+M-1; latent caches of THIS layer:
 * ckv [num_pages][16][lora], kpe [num_pages][16][rope] FP8-E4M3 (kv_cache.py:99-117); out [M][H] bf16. */
KB2_API int kb2_mla_forward(kb2_mla* h, int layer, const void* hidden_dev, const int32_t* positions_dev, int32_t first_position,
                            void* ckv_cache_layer_dev, void* kpe_cache_layer_dev, const int32_t* kv_indices_dev,
                            int32_t kv_len_after, void* out_dev, int32_t num_tokens, void* stream);

/* GEMM -> reduce-scatter fused (head-parallel attention under token sharding): after kb2_{gdn,gqa,mla}_set_output_scatter the
 * block's out_proj GEMM stores every output row straight into the receive buffer of the rank that owns the token
 * (peer_recv_host[r] from kb2_comm_peer_alloc, layout [M / num_ranks][num_ranks][hidden] bf16, slot = src_rank) instead of out_dev
 * (ignored).  num_ranks = 0 restores the local output.  Consumer: kb2_comm_barrier, then kb2_sum_slots_bf16 on the own buffer. */
KB2_API int kb2_gdn_set_output_scatter(kb2_gdn* h, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank);
KB2_API int kb2_gqa_set_output_scatter(kb2_gqa* h, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank);
KB2_API int kb2_mla_set_output_scatter(kb2_mla* h, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank);
KB2_API int kb2_sum_slots_bf16(const void* slots_dev, int32_t num_ranks, void* out_dev, int32_t rows, int32_t H, int32_t device, void* stream);


/* ================================================================================================
 * Multi-GPU: the expert-parallel communicator.  One process per GPU; replaces the reference's replicated-token EP loop
 * with a pinned-host bounce and a GPU0 reduction (python/krasis/model.py:3086-3211, KrasisEngine.reduce_sum_bf16
 * src/moe.rs:2505) by NCCL collectives over NVLink owned by THIS library (bound with dlopen at first use: no link-time
 * dependency).  The host moves only the 128-byte unique id between its processes.  Token-sharded schedule of one layer:
 *   norm (M/R rows) -> kb2_comm_all_gather -> head-parallel attention -> kb2_comm_reduce_scatter_bf16 -> add+norm, router,
 *   shared expert (M/R rows) -> kb2_comm_all_gather (rows, ids, weights) -> kb2_moe_forward(routed_only) on the local
 *   expert slice -> kb2_comm_reduce_scatter_bf16 -> rsf * routed + shared.
 * ================================================================================================ */
typedef struct kb2_comm kb2_comm;
KB2_API int kb2_comm_unique_id(void* out128);                                   /* rank 0: ncclGetUniqueId */
KB2_API int kb2_comm_init(const void* unique_id128, int32_t rank, int32_t num_ranks, int32_t device, kb2_comm** out);
KB2_API void kb2_comm_destroy(kb2_comm* c);
/* recv [R][bytes_per_rank] <- every rank's send [bytes_per_rank]; stream-ordered */
KB2_API int kb2_comm_all_gather(kb2_comm* c, const void* send_dev, void* recv_dev, size_t bytes_per_rank, void* stream);
/* recv [elems_per_rank] = sum over ranks of send[rank_slice]; send holds R * elems_per_rank BF16 */
KB2_API int kb2_comm_reduce_scatter_bf16(kb2_comm* c, const void* send_dev, void* recv_dev, size_t elems_per_rank, void* stream);
KB2_API int kb2_comm_all_reduce_bf16(kb2_comm* c, const void* send_dev, void* recv_dev, size_t elems, void* stream);
KB2_API int kb2_comm_broadcast(kb2_comm* c, void* buf_dev, size_t bytes, int32_t root, void* stream);
/* Peer memory + barrier for collectives fused into compute kernels (one process per GPU, one node, NVLink / NVSwitch):
 *   kb2_comm_peer_alloc   collective: cudaMalloc(bytes) on every rank, CUDA IPC handles exchanged by all-gather;
 *                         ptrs_out_host[r] = pointer through which THIS rank loads / stores rank r's buffer (r == rank: local)
 *   kb2_comm_barrier      4-byte all-reduce on `stream`: a rank leaves only after every peer's stream has reached it
 *   kb2_moe_forward_scatter + kb2_finish_routed_slots (below, engine section) use them for the expert-parallel reduce-scatter:
 *   the combine kernel of every rank stores its partial rows straight into the token owner's receive buffer. */
KB2_API int kb2_comm_peer_alloc(kb2_comm* c, size_t bytes, void** ptrs_out_host);
KB2_API int kb2_comm_peer_free(kb2_comm* c, void** ptrs_host);
KB2_API int kb2_comm_barrier(kb2_comm* c, void* stream);
/* BF16 sum over ranks delivered to `root` only (recv_dev may be NULL elsewhere): one chunk of a pipelined reduce-scatter whose
 * chunk j is rank j's token shard (krasis_b200/model.py: attention over chunk j+1 runs while chunk j's partial outputs are reduced). */
KB2_API int kb2_comm_reduce_bf16(kb2_comm* c, const void* send_dev, void* recv_dev, size_t elems, int32_t root, void* stream);

/* Per-kernel device timing with CUDA events recorded on the launching stream (the reference's
 * KRASIS_LAYER_TIMING / EP breakdown, python/krasis/model.py:2839-2860).  kb2_profile_collect synchronises the
 * device and returns, per kernel class, the summed milliseconds and the number of launches since enable. */
#define KB2_PROF_ROUTER_LOGITS 0
#define KB2_PROF_ROUTER_TOPK 1
#define KB2_PROF_BINNING 2
#define KB2_PROF_GEMM1 3
#define KB2_PROF_GEMM2 4
#define KB2_PROF_COMBINE 5
#define KB2_PROF_NUM 6
KB2_API int kb2_profile_enable(kb2_engine* e, int on);
KB2_API int kb2_profile_collect(kb2_engine* e, double* total_ms, int64_t* n_spans);

/* Process-wide kernel accounting: every kernel this library launches is counted (kb2_total_launches) and, while
 * kb2_kernel_profile_enable(1) is in effect, timed with CUDA events on its launching stream.  kb2_kernel_profile_collect
 * synchronises the device and fills total_ms[k] / n_spans[k] for k < kb2_kernel_profile_num(), then clears the record. */
KB2_API int64_t kb2_total_launches(void);
KB2_API int kb2_kernel_profile_num(void);
KB2_API const char* kb2_kernel_profile_name(int id);
KB2_API int kb2_kernel_profile_enable(int on);
KB2_API int kb2_kernel_profile_collect(double* total_ms, int64_t* n_spans);

#ifdef __cplusplus
}
#endif
#endif /* KRASIS_B200_H */
