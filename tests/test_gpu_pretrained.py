"""Real-format checkpoints through KrasisModel.from_pretrained: the test WRITES a tiny HF-layout model
directory (config.json + model.safetensors with the tensor names, dtypes and storage conventions of the real checkpoints:
Qwen3-Next norm weights stored as w - 1, A_log / dt_bias in F32, per-expert or stacked expert tensors, kv_b_proj, tied or
separate lm_head), loads it like the reference's KrasisModel.load() (python/krasis/model.py:505-611) and compares
forward(return_all_logits=True) with the oracle stack on the same tensors.
Bar (tests/test_prefill_vs_decode.py:124-131): cos >= 0.999 on the last token, greedy match within BF16 ties."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model as OM  # noqa: E402
from tests.test_loader_cpu import _write_safetensors  # noqa: E402

BF = torch.bfloat16


def _bf(t):
    return ("BF16", t.to(BF).contiguous().view(torch.int16).numpy().view(np.uint16))


def _f32(t):
    return ("F32", t.float().contiguous().numpy())


def _rn(g, *shape, std=0.05):
    return (torch.randn(*shape, generator=g) * std).to(BF)


def _norm(g, n):
    return (1 + 0.1 * torch.randn(n, generator=g)).to(BF)


def _check(model, cfg, W, M, shared_mode, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, cfg.vocab_size, (M,), generator=g)
    pos = torch.arange(M)
    got = model.forward(tok.cuda(), pos.cuda(), model.new_sequence(), return_all_logits=True).cpu()
    last = model.forward(tok.cuda(), pos.cuda(), model.new_sequence()).cpu()
    assert torch.equal(got[-1:], last)
    want = OM.forward(cfg, W, tok, pos, shared_mode)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
    assert cos[-1].item() >= 0.999, cos[-1].item()
    assert cos.median().item() >= 0.999 and (cos >= 0.995).float().mean().item() > 0.9, (cos.median().item(), cos.min().item())
    top_ok = want.gather(1, got.argmax(1)[:, None]).squeeze(1) >= want.max(1).values - 4 * 2 ** -8 * want.abs().max(1).values
    assert top_ok.float().mean().item() > 0.9


def build_v2lite_checkpoint():
    from krasis_b200.model import V2_LITE_ROPE
    g = torch.Generator().manual_seed(1)
    H, V, nh, E, k, I, Id, nl = 256, 512, 4, 8, 2, 128, 320, 3
    nope, rope, dv, lora = 128, 64, 128, 512
    hf = dict(model_type="deepseek_v2", hidden_size=H, num_hidden_layers=nl, vocab_size=V, rms_norm_eps=1e-6,
              num_attention_heads=nh, num_key_value_heads=nh, kv_lora_rank=lora, q_lora_rank=None, qk_nope_head_dim=nope,
              qk_rope_head_dim=rope, v_head_dim=dv, rope_theta=10000.0, rope_scaling=V2_LITE_ROPE, n_routed_experts=E,
              num_experts_per_tok=k, moe_intermediate_size=I, intermediate_size=Id, first_k_dense_replace=1, n_shared_experts=2,
              routed_scaling_factor=1.0, scoring_func="softmax", norm_topk_prob=False, tie_word_embeddings=False)
    t, layers = {}, []
    emb, fnorm, lm = _rn(g, V, H, std=0.5), _norm(g, H), _rn(g, V, H)
    t["model.embed_tokens.weight"], t["model.norm.weight"], t["lm_head.weight"] = _bf(emb), _bf(fnorm), _bf(lm)
    for i in range(nl):
        p = f"model.layers.{i}"
        lw = dict(input_norm=_norm(g, H), post_norm=_norm(g, H))
        t[f"{p}.input_layernorm.weight"], t[f"{p}.post_attention_layernorm.weight"] = _bf(lw["input_norm"]), _bf(lw["post_norm"])
        kv_b = _rn(g, nh * (nope + dv), lora, std=0.04)
        a = dict(q_proj=_rn(g, nh * (nope + rope), H), kv_a_proj_with_mqa=_rn(g, lora + rope, H), kv_a_layernorm=_norm(g, lora),
                 o_proj=_rn(g, H, nh * dv))
        for kk, v in a.items():
            t[f"{p}.self_attn.{kk}.weight"] = _bf(v)
        t[f"{p}.self_attn.kv_b_proj.weight"] = _bf(kv_b)
        kv = kv_b.reshape(nh, nope + dv, lora)
        a["w_kc"], a["w_vc"] = kv[:, :nope].contiguous(), kv[:, nope:].contiguous()
        lw["attn"] = a
        if i < 1:
            gp, up, dn = _rn(g, Id, H), _rn(g, Id, H), _rn(g, H, Id)
            t[f"{p}.mlp.gate_proj.weight"], t[f"{p}.mlp.up_proj.weight"], t[f"{p}.mlp.down_proj.weight"] = _bf(gp), _bf(up), _bf(dn)
            lw["dense"] = (torch.cat([gp, up]), dn)
        else:
            lw["gate"] = _rn(g, E, H, std=1.0)
            t[f"{p}.mlp.gate.weight"] = _bf(lw["gate"])
            w13, w2 = _rn(g, E, 2 * I, H), _rn(g, E, H, I)
            for e in range(E):
                t[f"{p}.mlp.experts.{e}.gate_proj.weight"] = _bf(w13[e, :I])
                t[f"{p}.mlp.experts.{e}.up_proj.weight"] = _bf(w13[e, I:])
                t[f"{p}.mlp.experts.{e}.down_proj.weight"] = _bf(w2[e])
            lw["experts"] = (w13, w2)
            sg, su, sd = _rn(g, 2 * I, H), _rn(g, 2 * I, H), _rn(g, H, 2 * I)
            t[f"{p}.mlp.shared_experts.gate_proj.weight"], t[f"{p}.mlp.shared_experts.up_proj.weight"] = _bf(sg), _bf(su)
            t[f"{p}.mlp.shared_experts.down_proj.weight"] = _bf(sd)
            lw["shared"] = (torch.cat([sg, su]), sd, None)
        layers.append(lw)
    return hf, t, dict(embed=emb, final_norm=fnorm, lm_head=lm, layers=layers)


def test_deepseek_v2_lite_layout_checkpoint(tmp_path):
    """MLA + first dense layer (intermediate size not a multiple of 128) + 2 ungated shared experts run as the manager's
    INT4 expert + softmax routing without renormalisation + per-expert tensors + separate lm_head."""
    from krasis_b200.model import KrasisModel
    hf, t, W = build_v2lite_checkpoint()
    I, nl = hf["moe_intermediate_size"], hf["num_hidden_layers"]
    json.dump(hf, open(tmp_path / "config.json", "w"))
    _write_safetensors(tmp_path / "model.safetensors", t)
    M = 150
    model = KrasisModel.from_pretrained(str(tmp_path), max_tokens=M)
    cfg = model.cfg
    assert cfg.is_mla and cfg.first_k_dense_replace == 1 and cfg.shared_width == 2 * I and not cfg.shared_expert_gate
    assert model.shared_mode == "int4_manager" and model.layer_types == ["mla"] * nl
    assert model.layers[0].dense is not None and model.layers[1].moe_idx == 0
    _check(model, cfg, W, M, "int4_manager", seed=3)


def build_qwen3_next_checkpoint():
    g = torch.Generator().manual_seed(2)
    H, V, E, k, I, nl = 256, 512, 8, 2, 128, 4
    nh, nkv, d, nk, nv, dk, dv, K = 4, 2, 128, 2, 4, 128, 128, 4
    hf = dict(model_type="qwen3_next", hidden_size=H, num_hidden_layers=nl, vocab_size=V, rms_norm_eps=1e-6, head_dim=d,
              num_attention_heads=nh, num_key_value_heads=nkv, partial_rotary_factor=0.5, rope_theta=10000.0,
              full_attention_interval=4, num_experts=E, num_experts_per_tok=k, moe_intermediate_size=I,
              shared_expert_intermediate_size=I, norm_topk_prob=True, linear_num_key_heads=nk, linear_num_value_heads=nv,
              linear_key_head_dim=dk, linear_value_head_dim=dv, linear_conv_kernel_dim=K, tie_word_embeddings=True)
    t, layers = {}, []
    one = lambda w: (w.float() - 1.0).to(BF)            # stored form of a Qwen3-Next norm weight; +1 in BF16 restores w exactly
    emb, fnorm = _rn(g, V, H, std=0.5), (1 + torch.randint(-8, 9, (H,), generator=g) / 64.0).to(BF)
    t["model.embed_tokens.weight"], t["model.norm.weight"] = _bf(emb), _bf(one(fnorm))
    qn = lambda n: (1 + torch.randint(-8, 9, (n,), generator=g) / 64.0).to(BF)   # values whose -1 / +1 round trip is exact in BF16
    kd, vd = nk * dk, nv * dv
    for i in range(nl):
        p = f"model.layers.{i}"
        lw = dict(input_norm=qn(H), post_norm=qn(H))
        t[f"{p}.input_layernorm.weight"], t[f"{p}.post_attention_layernorm.weight"] = _bf(one(lw["input_norm"])), _bf(one(lw["post_norm"]))
        if (i + 1) % 4:
            a = dict(in_proj_qkvz=_rn(g, 2 * kd + 2 * vd, H, std=0.15), in_proj_ba=_rn(g, 2 * nv, H, std=0.15),
                     out_proj=_rn(g, H, vd), conv1d_weight=_rn(g, 2 * kd + vd, 1, K, std=0.5),
                     A_log=_rn(g, nv, std=0.5), dt_bias=_rn(g, nv, std=0.5), norm_weight=_norm(g, dv))
            for kk, name in (("in_proj_qkvz", "in_proj_qkvz.weight"), ("in_proj_ba", "in_proj_ba.weight"), ("out_proj", "out_proj.weight"),
                             ("conv1d_weight", "conv1d.weight"), ("norm_weight", "norm.weight")):
                t[f"{p}.linear_attn.{name}"] = _bf(a[kk])
            t[f"{p}.linear_attn.A_log"], t[f"{p}.linear_attn.dt_bias"] = _f32(a["A_log"]), _f32(a["dt_bias"])   # F32 storage
        else:
            a = dict(q_proj=_rn(g, nh * d * 2, H), k_proj=_rn(g, nkv * d, H), v_proj=_rn(g, nkv * d, H), o_proj=_rn(g, H, nh * d),
                     q_norm=qn(d), k_norm=qn(d))
            for kk in ("q_proj", "k_proj", "v_proj", "o_proj"):
                t[f"{p}.self_attn.{kk}.weight"] = _bf(a[kk])
            t[f"{p}.self_attn.q_norm.weight"], t[f"{p}.self_attn.k_norm.weight"] = _bf(one(a["q_norm"])), _bf(one(a["k_norm"]))
        lw["attn"] = a
        lw["gate"] = _rn(g, E, H, std=1.0)
        t[f"{p}.mlp.gate.weight"] = _bf(lw["gate"])
        w13, w2 = _rn(g, E, 2 * I, H), _rn(g, E, H, I)
        for e in range(E):
            t[f"{p}.mlp.experts.{e}.gate_proj.weight"] = _bf(w13[e, :I])
            t[f"{p}.mlp.experts.{e}.up_proj.weight"] = _bf(w13[e, I:])
            t[f"{p}.mlp.experts.{e}.down_proj.weight"] = _bf(w2[e])
        lw["experts"] = (w13, w2)
        sg, su, sd, sgate = _rn(g, I, H), _rn(g, I, H), _rn(g, H, I), _rn(g, 1, H)
        t[f"{p}.mlp.shared_expert.gate_proj.weight"], t[f"{p}.mlp.shared_expert.up_proj.weight"] = _bf(sg), _bf(su)
        t[f"{p}.mlp.shared_expert.down_proj.weight"], t[f"{p}.mlp.shared_expert_gate.weight"] = _bf(sd), _bf(sgate)
        lw["shared"] = (torch.cat([sg, su]), sd, sgate)
        layers.append(lw)
    return hf, t, dict(embed=emb, final_norm=fnorm, lm_head=emb, layers=layers)


def test_qwen3_next_layout_checkpoint(tmp_path):
    """Hybrid Gated DeltaNet (dk = dv = 128: the tcgen05 scan) + gated GQA, norm weights stored as w - 1, A_log / dt_bias
    stored in F32, gated INT8 shared expert, renormalised top-k, tied embeddings."""
    from krasis_b200.model import KrasisModel
    hf, t, W = build_qwen3_next_checkpoint()
    json.dump(hf, open(tmp_path / "config.json", "w"))
    _write_safetensors(tmp_path / "model.safetensors", t)
    M = 150
    model = KrasisModel.from_pretrained(str(tmp_path), max_tokens=M)
    cfg = model.cfg
    assert cfg.norm_bias_one and cfg.gated_attention and cfg.shared_expert_gate and model.shared_mode == "int8_gated"
    assert model.layer_types == ["linear_attention"] * 3 + ["full_attention"]
    _check(model, cfg, W, M, "int8_gated", seed=4)
