"""Whole-model prefill oracle (torch / numpy on CPU; test infrastructure only): the layer loop of
python/krasis/model.py:2719-2955 + python/krasis/layer.py:242-460 assembled from the per-block oracles of this package.

`W` holds the EFFECTIVE weights (norm weights after the +1 convention), BF16 torch tensors:
  W["embed"], W["final_norm"], W["lm_head"], W["layers"][i] = dict(
      input_norm, post_norm, attn (dict for oracle.attention.*), dense=(gate_up [2I,H], down [H,I]) | None,
      gate [E,H], corr_bias | None, experts=(w13 [E,2I,H], w2 [E,H,I]) BF16 (quantised here with the Krasis group quantiser,
      src/weights/marlin.rs:145-207), shared=(gate_up, down, sigmoid_gate | None) | None)
shared_mode: "int8_gated" / "int8" -> layer.py:508-524 (W8A8); "int4_manager" -> one-expert INT4 MoE with weight 1 added after
the rsf scaling (gpu_prefill.py:4471-4480,4738-4801); "none".
"""
import numpy as np
import torch

from . import attention as A, dense as D, moe as omoe, quant, router
from .bf16 import f32_to_bf16_bits


def _bits(t):
    return t.detach().to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


def quantize_experts(w13, w2, bits=4):
    """BF16 [E, 2I, H], [E, H, I] -> oracle.moe.Int4Layer through the reference quantiser's restatement."""
    qf = quant.quantize_int4 if bits == 4 else quant.quantize_int8
    a = [qf(_bits(w13[e])) for e in range(w13.shape[0])]
    b = [qf(_bits(w2[e])) for e in range(w2.shape[0])]
    return omoe.Int4Layer(np.stack([x[0] for x in a]), np.stack([x[1] for x in a]), np.stack([x[0] for x in b]),
                          np.stack([x[1] for x in b]), bits)


def dense_mlp(h, gate_up, down):
    """layer.py:497-506 with INT8 weights (gate_proj / up_proj quantised per row -> one stacked matrix is the same math)."""
    act = D.silu_and_mul(D.int8_linear(h, *D.quantize_to_int8(gate_up)))
    return D.int8_linear(act, *D.quantize_to_int8(down))


def forward(cfg, W, tokens, positions, shared_mode, expert_bits=4):
    """cfg: krasis_b200.model.HybridMoEConfig (only read).  Returns logits [M, V] float32."""
    hidden = W["embed"][tokens]
    residual = None
    eps = cfg.rms_norm_eps
    k = cfg.num_experts_per_tok
    for i, lw in enumerate(W["layers"]):
        if residual is None:
            residual, hidden = hidden, D.rmsnorm(hidden, lw["input_norm"], eps)
        else:
            hidden, residual = D.fused_add_rmsnorm(hidden, residual, lw["input_norm"], eps)
        lt = cfg.layer_type(i)
        if lt == "linear_attention":
            attn, _, _ = A.gdn_layer_prefill(hidden, lw["attn"], dict(nk=cfg.linear_num_key_heads, nv=cfg.linear_num_value_heads,
                                                                     dk=cfg.linear_key_head_dim, dv=cfg.linear_value_head_dim, eps=eps))
        elif lt == "mla":
            attn, _, _ = A.mla_layer_prefill(hidden, lw["attn"], dict(nh=cfg.num_attention_heads, nope=cfg.qk_nope_head_dim,
                                                                     rope=cfg.qk_rope_head_dim, dv=cfg.v_head_dim, lora=cfg.kv_lora_rank,
                                                                     theta=cfg.rope_theta, eps=eps, rope_scaling=cfg.rope_scaling,
                                                                     q_lora=cfg.q_lora_rank or 0), positions)
        else:
            attn, _, _ = A.gqa_layer_prefill(hidden, lw["attn"], dict(nh=cfg.num_attention_heads, nkv=cfg.num_key_value_heads,
                                                                     d=cfg.gqa_head_dim, rotary_dim=cfg.rotary_dim,
                                                                     theta=cfg.rope_theta, eps=eps), positions)
        h, residual = D.fused_add_rmsnorm(attn, residual, lw["post_norm"], eps)
        if lw.get("dense") is not None:
            hidden = dense_mlp(h, *lw["dense"])
            continue
        layer = quantize_experts(*lw["experts"], bits=expert_bits)
        hn = h.float().numpy()
        cb = lw.get("corr_bias")
        ids, wts = router.compute_routing(hn, lw["gate"].float().numpy(), k, scoring_func=cfg.scoring_func,
                                          norm_topk_prob=cfg.norm_topk_prob,
                                          e_score_correction_bias=cb.float().numpy() if cb is not None else None)
        routed = omoe.moe_forward_gpu_path(layer, hn, ids, wts)
        shared = None
        if shared_mode in ("int8", "int8_gated"):
            gu, dn, gw = lw["shared"]
            shared = D.shared_expert_forward(h, D.quantize_to_int8(gu), D.quantize_to_int8(dn), gw if shared_mode == "int8_gated" else None)
            shared = shared.float().numpy()
        elif shared_mode == "int4_manager":
            gu, dn, _ = lw["shared"]
            sl = quantize_experts(gu.unsqueeze(0), dn.unsqueeze(0), bits=expert_bits)
            M = hn.shape[0]
            shared = omoe.moe_forward_gpu_path(sl, hn, np.zeros((M, 1), np.int32), np.ones((M, 1), np.float32))
        hidden = torch.from_numpy(omoe.finish_gpu_path(routed, cfg.routed_scaling_factor, shared)).to(torch.bfloat16)
    hidden, _ = D.fused_add_rmsnorm(hidden, residual, W["final_norm"], eps)
    return D.int8_linear(hidden, *D.quantize_to_int8(W["lm_head"])).float()
