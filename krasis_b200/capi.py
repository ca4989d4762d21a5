"""ctypes binding of include/krasis_b200.h — the same stub a maintainer of the reference would add
(INTEGRATION.md).  No fallback: if the CUDA library is missing or fails to load, importing the compute
entry points raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libkrasis_b200.so")

KB2_OK, KB2_ERR_STATE, KB2_ERR_VALUE, KB2_ERR_CUDA = 0, 1, 2, 3
SCORE_SOFTMAX, SCORE_SIGMOID, SCORE_TOPK_SOFTMAX = 0, 1, 2
FMT_INT4_G128, FMT_INT8_G128, FMT_GGUF_Q8_0, FMT_GGUF_Q4_K = 0, 1, 2, 3
FMT_GGUF_Q6_K, FMT_GGUF_Q5_K, FMT_GGUF_Q5_0, FMT_GGUF_Q4_0 = 4, 5, 6, 7
GGUF_FORMATS = {"Q8_0": FMT_GGUF_Q8_0, "Q4_K": FMT_GGUF_Q4_K, "Q6_K": FMT_GGUF_Q6_K, "Q5_K": FMT_GGUF_Q5_K, "Q5_0": FMT_GGUF_Q5_0,
                "Q4_0": FMT_GGUF_Q4_0}
GGUF_ROW_BYTES = {FMT_GGUF_Q8_0: lambda k: k // 32 * 34, FMT_GGUF_Q4_K: lambda k: k // 256 * 144, FMT_GGUF_Q6_K: lambda k: k // 256 * 210,
                  FMT_GGUF_Q5_K: lambda k: k // 256 * 176, FMT_GGUF_Q5_0: lambda k: k // 32 * 22, FMT_GGUF_Q4_0: lambda k: k // 32 * 18}


class Config(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int32), ("moe_intermediate_size", C.c_int32), ("n_routed_experts", C.c_int32),
        ("num_experts_per_tok", C.c_int32), ("num_moe_layers", C.c_int32), ("weight_format", C.c_int32),
        ("rank", C.c_int32), ("num_ranks", C.c_int32), ("scoring_func", C.c_int32),
        ("norm_topk_prob", C.c_int32), ("routed_scaling_factor", C.c_float), ("max_tokens", C.c_int32),
        ("device", C.c_int32), ("w2_weight_format", C.c_int32),
    ]


class GdnConfig(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_k_heads", C.c_int32), ("num_v_heads", C.c_int32),
                ("k_head_dim", C.c_int32), ("v_head_dim", C.c_int32), ("conv_kernel", C.c_int32),
                ("rms_norm_eps", C.c_float), ("max_tokens", C.c_int32), ("num_layers", C.c_int32),
                ("device", C.c_int32)]


class GqaConfig(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
                ("head_dim", C.c_int32), ("rotary_dim", C.c_int32), ("gated", C.c_int32), ("rope_theta", C.c_float),
                ("rms_norm_eps", C.c_float), ("page_size", C.c_int32), ("max_tokens", C.c_int32),
                ("num_layers", C.c_int32), ("device", C.c_int32)]


class MlaConfig(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_heads", C.c_int32), ("qk_nope_head_dim", C.c_int32),
                ("qk_rope_head_dim", C.c_int32), ("v_head_dim", C.c_int32), ("kv_lora_rank", C.c_int32),
                ("q_lora_rank", C.c_int32), ("rms_norm_eps", C.c_float), ("sm_scale", C.c_float),
                ("page_size", C.c_int32), ("max_tokens", C.c_int32), ("max_kv_len", C.c_int32),
                ("num_layers", C.c_int32), ("device", C.c_int32)]


# name -> (restype, argtypes); must list every symbol include/krasis_b200.h declares
SIGNATURES = {
    "kb2_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "kb2_destroy": (None, [C.c_void_p]),
    "kb2_last_error": (C.c_char_p, []),
    "kb2_version": (C.c_char_p, []),
    "kb2_get_config": (C.c_int, [C.c_void_p, C.POINTER(Config)]),
    "kb2_expert_range": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "kb2_tiled_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "kb2_load_experts_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4),
    "kb2_quantize_group_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                         C.c_void_p]),
    "kb2_load_experts_dev": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5),
    "kb2_load_experts_gguf_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 3),
    "kb2_export_experts_tiled_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "kb2_load_experts_tiled_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4),
    "kb2_attach_experts_tiled_dev": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4),
    "kb2_retile_dev": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p]),
    "kb2_set_router_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kb2_route": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kb2_moe_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "kb2_ep_bin_rows": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 6),
    "kb2_moe_forward_rows": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p]),
    "kb2_ep_combine_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "kb2_finish_routed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "kb2_moe_forward_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "kb2_prefill_moe_stack_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_void_p]),
    "kb2_last_expert_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "kb2_launch_count": (C.c_int64, [C.c_void_p]),
    "kb2_linear_bf16": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.c_void_p]),
    "kb2_rmsnorm": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "kb2_quantize_rows_int8": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p]),
    "kb2_int8_linear": (C.c_int, [C.c_void_p] * 6 + [C.c_int32] * 4 + [C.c_void_p]),
    "kb2_silu_and_mul": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kb2_rmsnorm_q8": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "kb2_int8_linear_q8": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
    "kb2_silu_mul_int8_linear": (C.c_int, [C.c_void_p] * 7 + [C.c_int32] * 4 + [C.c_void_p]),
    "kb2_sigmoid_gate_mul": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p]),
    "kb2_add_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "kb2_gdn_create": (C.c_int, [C.POINTER(GdnConfig), C.POINTER(C.c_void_p)]),
    "kb2_gdn_destroy": (None, [C.c_void_p]),
    "kb2_gdn_set_weights_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7),
    "kb2_gdn_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "kb2_gdn_reset_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "kb2_gdn_get_state_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "kb2_gqa_create": (C.c_int, [C.POINTER(GqaConfig), C.POINTER(C.c_void_p)]),
    "kb2_gqa_destroy": (None, [C.c_void_p]),
    "kb2_gqa_set_weights_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 6),
    "kb2_gqa_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "kb2_mla_create": (C.c_int, [C.POINTER(MlaConfig), C.POINTER(C.c_void_p)]),
    "kb2_mla_destroy": (None, [C.c_void_p]),
    "kb2_mla_set_weights_host": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 9),
    "kb2_mla_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "kb2_comm_unique_id": (C.c_int, [C.c_void_p]),
    "kb2_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "kb2_comm_destroy": (None, [C.c_void_p]),
    "kb2_comm_all_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kb2_comm_reduce_scatter_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kb2_comm_all_reduce_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kb2_comm_broadcast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "kb2_gdn_set_output_scatter": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "kb2_gqa_set_output_scatter": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "kb2_mla_set_output_scatter": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "kb2_sum_slots_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kb2_comm_peer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "kb2_comm_peer_free": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "kb2_comm_barrier": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kb2_moe_forward_scatter": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p]),
    "kb2_finish_routed_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "kb2_comm_reduce_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "kb2_total_launches": (C.c_int64, []),
    "kb2_kernel_profile_num": (C.c_int, []),
    "kb2_kernel_profile_name": (C.c_char_p, [C.c_int]),
    "kb2_kernel_profile_enable": (C.c_int, [C.c_int]),
    "kb2_kernel_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "kb2_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "kb2_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}
PROF_NAMES = ["router_logits", "router_topk", "binning", "gemm1_gate_up_silu", "gemm2_down", "combine"]

_lib = None


def load():
    """Load the C-ABI library (works without a GPU; compute calls then fail with KB2_ERR_CUDA)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m krasis_b200.build` "
                "(krasis_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def kernel_profile(on: bool):
    check(load().kb2_kernel_profile_enable(int(bool(on))))


def kernel_profile_collect():
    """{kernel class: (total_ms, launches)} of every kernel timed since kernel_profile(True); synchronises the device."""
    lib = load()
    n = lib.kb2_kernel_profile_num()
    ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
    check(lib.kb2_kernel_profile_collect(ms, cnt))
    return {lib.kb2_kernel_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}


def total_launches() -> int:
    return int(load().kb2_total_launches())


class Kb2Error(RuntimeError):
    pass


def check(rc: int):
    """Map C error codes onto the exception types the reference's PyO3 layer raises
    (PyRuntimeError for state, PyValueError for shape — src/moe.rs:1543-1550,2285-2300)."""
    if rc == KB2_OK:
        return
    msg = load().kb2_last_error().decode()
    if rc == KB2_ERR_VALUE:
        raise ValueError(msg)
    raise Kb2Error(msg)
