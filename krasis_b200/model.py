"""Lean prefill re-host of the reference's model loop: hybrid linear-attention + GQA MoE models (Qwen3-Coder-Next,
Qwen3.5-35B-A3B), pure-GQA MoE models (Qwen3-235B) and MLA MoE models with leading dense layers (DeepSeek-V2-Lite).

Mirrors, for M > 1 (prefill):
  KrasisModel.forward                 python/krasis/model.py:2167-2207  (token_ids, positions, seq_states, return_all_logits)
  forward_prefill_layer_grouped       python/krasis/model.py:2719-2955  — without layer groups, expert streaming or
                                      token chunking: on a 180 GB B200 every expert is resident and 8K tokens fit one pass
  TransformerLayer.forward            python/krasis/layer.py:242-460    (pre-norm, attention, fused_add_rmsnorm, dense MLP | MoE)
  dense MLP / shared expert           python/krasis/layer.py:497-524    (INT8 W8A8, QuantConfig defaults config.py:208-211)
  final norm + lm_head                python/krasis/model.py:3380-3399
  ModelConfig.from_model_path         python/krasis/config.py:216-420   -> HybridMoEConfig.from_hf_config
  KrasisModel.load                    python/krasis/model.py:505-611    -> KrasisModel.from_pretrained
Layer pattern: full attention iff (i + 1) % full_attention_interval == 0 (python/krasis/config.py:336-342).

Shared experts follow the reference's three cases: gated (Qwen3-Next) -> the layer's INT8 W8A8 expert with
the sigmoid gate; ungated on ONE GPU (DeepSeek-V2-Lite) -> the manager's fused INT4 expert, added after the rsf scaling
(gpu_prefill.py:4471-4480,4738-4801); ungated under EP -> the layer's INT8 expert (model.py:3076-3081,3234-3240).

Multi-GPU (one process per GPU, `krasis_b200.parallel.Communicator` = NCCL behind the C ABI): the token rows of the residual
stream are SHARDED over the ranks; norms, router, shared expert, dense MLP and lm_head run on M/R rows.  Attention is
head-parallel on all M rows (all-gather in, reduce-scatter of the partial o_proj outputs out); experts are sliced by rank
exactly like the reference (gpu_prefill.py:353-359): all-gather of the routed rows / ids / weights, local expert slice with
routed_only, reduce-scatter of the partial sums.  This replaces python/krasis/model.py:3086-3211.

All arithmetic is in libkrasis_b200 kernels; torch holds buffers and does the embedding row gather.
"""
import json
import math
import os
from dataclasses import dataclass, field, fields
from types import SimpleNamespace
from typing import List, Optional

import torch

from . import capi
from .attention import (GatedDeltaNetAttention, GQAAttention, MLAAttention, MLAPagedKVCache, PagedKVCache,
                        SequenceKVState)
from .engine import KrasisEngine
from . import layers as L


@dataclass
class HybridMoEConfig:
    """The subset of python/krasis/config.py:ModelConfig the prefill path reads."""
    hidden_size: int = 2048
    num_hidden_layers: int = 48
    full_attention_interval: int = 4            # 0 => every layer is full attention (GQA, or MLA when kv_lora_rank is set)
    vocab_size: int = 151936
    rms_norm_eps: float = 1e-6
    norm_bias_one: bool = False                 # Qwen3-Next / Qwen3.5: stored norm weights get +1 at load (config.py:348)
    tie_word_embeddings: bool = False
    # MoE
    n_routed_experts: int = 512
    num_experts_per_tok: int = 10
    moe_intermediate_size: int = 512
    first_k_dense_replace: int = 0              # leading dense layers (DeepSeek)
    intermediate_size: int = 0                  # their MLP width
    n_shared_experts: int = 1
    shared_expert_intermediate_size: int = 512  # 0 with n_shared_experts > 0 => n_shared_experts * moe_intermediate_size
    shared_expert_gate: bool = True
    norm_topk_prob: bool = True
    routed_scaling_factor: float = 1.0
    scoring_func: str = "softmax"
    expert_bits: int = 4
    # GQA
    num_attention_heads: int = 16
    num_key_value_heads: int = 2
    gqa_head_dim: int = 256
    partial_rotary_factor: float = 0.25
    rope_theta: float = 10000000.0
    gated_attention: bool = True
    # MLA (kv_lora_rank None => GQA)
    kv_lora_rank: Optional[int] = None
    q_lora_rank: Optional[int] = None
    qk_nope_head_dim: int = 128
    qk_rope_head_dim: int = 64
    v_head_dim: int = 128
    rope_scaling: Optional[dict] = None
    # Gated DeltaNet
    linear_num_key_heads: int = 16
    linear_num_value_heads: int = 32
    linear_key_head_dim: int = 128
    linear_value_head_dim: int = 128
    linear_conv_kernel_dim: int = 4
    synthetic_router_std: float = 0.02          # std of the random router weights KrasisModel generates
    layers_prefix: str = "model"
    name: str = "custom"

    @property
    def rotary_dim(self) -> int:                 # config.py:469-473
        return int(self.gqa_head_dim * self.partial_rotary_factor)

    @property
    def is_mla(self) -> bool:
        return self.kv_lora_rank is not None

    @property
    def num_moe_layers(self) -> int:
        return self.num_hidden_layers - self.first_k_dense_replace

    @property
    def shared_width(self) -> int:
        if self.n_shared_experts <= 0:
            return 0
        return self.shared_expert_intermediate_size or self.n_shared_experts * self.moe_intermediate_size

    def is_full_attention_layer(self, i: int) -> bool:
        return self.full_attention_interval <= 0 or (i + 1) % self.full_attention_interval == 0

    def layer_type(self, i: int) -> str:
        if not self.is_full_attention_layer(i):
            return "linear_attention"
        return "mla" if self.is_mla else "full_attention"

    @classmethod
    def from_hf_config(cls, raw: dict, layers_prefix: str = "model", has_shared_gate: Optional[bool] = None) -> "HybridMoEConfig":
        """config.json -> config, with the reference's field fallbacks (python/krasis/config.py:300-420)."""
        cfg = raw.get("text_config", raw)
        arch = cfg.get("model_type", "")
        is_mla = "kv_lora_rank" in cfg
        if "first_k_dense_replace" in cfg:
            first_k = cfg["first_k_dense_replace"]
        elif "decoder_sparse_step" in cfg:
            first_k = 0 if cfg["decoder_sparse_step"] <= 1 else cfg["decoder_sparse_step"]
        else:
            first_k = 0
        n_shared = cfg.get("n_shared_experts", 0) or 0
        shared_inter = cfg.get("shared_expert_intermediate_size", 0) or 0
        if n_shared == 0 and shared_inter > 0:
            n_shared = 1
        rp = cfg.get("rope_parameters", {}) or {}
        nh = cfg["num_attention_heads"]
        head_dim = cfg.get("head_dim") or cfg["hidden_size"] // nh
        gated = has_shared_gate if has_shared_gate is not None else arch in ("qwen3_next", "qwen3_5_moe_text")
        return cls(hidden_size=cfg["hidden_size"], num_hidden_layers=cfg["num_hidden_layers"],
                   full_attention_interval=cfg.get("full_attention_interval", 0), vocab_size=cfg["vocab_size"],
                   rms_norm_eps=cfg.get("rms_norm_eps", 1e-6), norm_bias_one=arch in ("qwen3_next", "qwen3_5_moe_text"),
                   tie_word_embeddings=bool(cfg.get("tie_word_embeddings", raw.get("tie_word_embeddings", False))),
                   n_routed_experts=cfg.get("n_routed_experts", cfg.get("num_experts", cfg.get("num_local_experts", 0))),
                   num_experts_per_tok=cfg.get("num_experts_per_tok", cfg.get("experts_per_token", 0)),
                   moe_intermediate_size=cfg.get("moe_intermediate_size", cfg.get("intermediate_size", 0)),
                   first_k_dense_replace=first_k, intermediate_size=cfg.get("intermediate_size", 0),
                   n_shared_experts=n_shared, shared_expert_intermediate_size=shared_inter, shared_expert_gate=gated and n_shared > 0,
                   norm_topk_prob=cfg.get("norm_topk_prob", arch == "qwen3_5_moe_text"),
                   routed_scaling_factor=cfg.get("routed_scaling_factor", 1.0), scoring_func=cfg.get("scoring_func", "softmax"),
                   num_attention_heads=nh, num_key_value_heads=cfg.get("num_key_value_heads", nh), gqa_head_dim=head_dim,
                   partial_rotary_factor=cfg.get("partial_rotary_factor", rp.get("partial_rotary_factor", 1.0)),
                   rope_theta=cfg.get("rope_theta", rp.get("rope_theta", 10000.0)),
                   gated_attention=arch in ("qwen3_next", "qwen3_5_moe_text"),
                   kv_lora_rank=cfg.get("kv_lora_rank") if is_mla else None, q_lora_rank=cfg.get("q_lora_rank") if is_mla else None,
                   qk_nope_head_dim=cfg.get("qk_nope_head_dim", 128), qk_rope_head_dim=cfg.get("qk_rope_head_dim", 64),
                   v_head_dim=cfg.get("v_head_dim", 128), rope_scaling=cfg.get("rope_scaling") or None,
                   linear_num_key_heads=cfg.get("linear_num_key_heads", 16), linear_num_value_heads=cfg.get("linear_num_value_heads", 32),
                   linear_key_head_dim=cfg.get("linear_key_head_dim", 128), linear_value_head_dim=cfg.get("linear_value_head_dim", 128),
                   linear_conv_kernel_dim=cfg.get("linear_conv_kernel_dim", 4), layers_prefix=layers_prefix, name=arch or "custom")


QWEN3_CODER_NEXT = HybridMoEConfig(name="qwen3-coder-next")
# BASELINE config C3: 40 layers (30 linear + 10 gated GQA), 256 experts top-8; I = 512 per the HF config
QWEN35_35B_A3B = HybridMoEConfig(name="qwen3.5-35b-a3b", num_hidden_layers=40, n_routed_experts=256, num_experts_per_tok=8,
                                 vocab_size=248320)
# BASELINE config C2: MLA 16 heads, first layer dense (10944), 64 experts top-6, two ungated shared experts
V2_LITE_ROPE = {"beta_fast": 32, "beta_slow": 1, "factor": 40, "mscale": 0.707, "mscale_all_dim": 0.707,
                "original_max_position_embeddings": 4096, "type": "yarn"}
DEEPSEEK_V2_LITE = HybridMoEConfig(name="deepseek-v2-lite", num_hidden_layers=27, full_attention_interval=0, vocab_size=102400,
                                   n_routed_experts=64, num_experts_per_tok=6, moe_intermediate_size=1408,
                                   first_k_dense_replace=1, intermediate_size=10944, n_shared_experts=2,
                                   shared_expert_intermediate_size=0, shared_expert_gate=False, norm_topk_prob=False,
                                   num_attention_heads=16, num_key_value_heads=16, gqa_head_dim=192, partial_rotary_factor=1.0,
                                   rope_theta=10000.0, gated_attention=False, kv_lora_rank=512, q_lora_rank=None,
                                   rope_scaling=V2_LITE_ROPE)
# BASELINE config C5: 94 GQA layers 64/4/128, 128 experts top-8, no shared expert
QWEN3_235B = HybridMoEConfig(name="qwen3-235b-a22b", hidden_size=4096, num_hidden_layers=94, full_attention_interval=0,
                             n_routed_experts=128, num_experts_per_tok=8, moe_intermediate_size=1536, n_shared_experts=0,
                             shared_expert_intermediate_size=0, shared_expert_gate=False, num_attention_heads=64,
                             num_key_value_heads=4, gqa_head_dim=128, partial_rotary_factor=1.0, rope_theta=1000000.0,
                             gated_attention=False)
PRESETS = {"qcn": QWEN3_CODER_NEXT, "qwen35": QWEN35_35B_A3B, "v2lite": DEEPSEEK_V2_LITE, "q235b": QWEN3_235B}


# --------------------------------------------------------------------------------------------- head-parallel weight slices

def shard_gdn_weights(w: dict, cfg: HybridMoEConfig, rank: int, num_ranks: int) -> dict:
    """Head-parallel slice of a Gated-DeltaNet layer: key-head groups [rank*nk/R, (rank+1)*nk/R).
    in_proj_qkvz / in_proj_ba rows are already grouped per key head (linear_attention.py:337-391)."""
    nk, nv, dk, dv = cfg.linear_num_key_heads, cfg.linear_num_value_heads, cfg.linear_key_head_dim, cfg.linear_value_head_dim
    r = nv // nk
    k0, k1 = rank * nk // num_ranks, (rank + 1) * nk // num_ranks
    G = 2 * dk + 2 * r * dv
    kd, vd = nk * dk, nv * dv
    cw = w["conv1d_weight"]
    conv = torch.cat([cw[k0 * dk:k1 * dk], cw[kd + k0 * dk:kd + k1 * dk], cw[2 * kd + k0 * r * dv:2 * kd + k1 * r * dv]], dim=0)
    return dict(in_proj_qkvz=w["in_proj_qkvz"][k0 * G:k1 * G].contiguous(), in_proj_ba=w["in_proj_ba"][k0 * 2 * r:k1 * 2 * r].contiguous(),
                conv1d_weight=conv.contiguous(), A_log=w["A_log"][k0 * r:k1 * r].contiguous(), dt_bias=w["dt_bias"][k0 * r:k1 * r].contiguous(),
                norm_weight=w["norm_weight"], out_proj=w["out_proj"][:, k0 * r * dv:k1 * r * dv].contiguous())


def shard_gqa_weights(w: dict, cfg: HybridMoEConfig, rank: int, num_ranks: int):
    """Head-parallel slice of a GQA layer: query heads split evenly; a KV head is replicated on the ranks that
    hold query heads of its group.  Returns (weights, num_heads_local, num_kv_heads_local)."""
    nh, nkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.gqa_head_dim
    h0, h1 = rank * nh // num_ranks, (rank + 1) * nh // num_ranks
    grp = nh // nkv
    kv0, kv1 = h0 // grp, (h1 - 1) // grp + 1
    qw = d * (2 if cfg.gated_attention else 1)
    out = dict(q_proj=w["q_proj"][h0 * qw:h1 * qw].contiguous(), k_proj=w["k_proj"][kv0 * d:kv1 * d].contiguous(),
               v_proj=w["v_proj"][kv0 * d:kv1 * d].contiguous(), o_proj=w["o_proj"][:, h0 * d:h1 * d].contiguous(),
               q_norm=w.get("q_norm"), k_norm=w.get("k_norm"))
    return out, h1 - h0, kv1 - kv0


def shard_mla_weights(w: dict, cfg: HybridMoEConfig, rank: int, num_ranks: int):
    """Head-parallel slice of an MLA layer: per-head rows of q_proj / q_b_proj, per-head w_kc / w_vc, o_proj columns; the
    shared latent projection (kv_a_proj_with_mqa, its layernorm, q_a_*) is replicated.  Returns (weights, heads_local)."""
    nh, qd, dv = cfg.num_attention_heads, cfg.qk_nope_head_dim + cfg.qk_rope_head_dim, cfg.v_head_dim
    h0, h1 = rank * nh // num_ranks, (rank + 1) * nh // num_ranks
    out = dict(w)
    qk = "q_b_proj" if cfg.q_lora_rank else "q_proj"
    out[qk] = w[qk][h0 * qd:h1 * qd].contiguous()
    out["w_kc"], out["w_vc"] = w["w_kc"][h0:h1].contiguous(), w["w_vc"][h0:h1].contiguous()
    out["o_proj"] = w["o_proj"][:, h0 * dv:h1 * dv].contiguous()
    return out, h1 - h0


# --------------------------------------------------------------------------------------------- weight sources

class SyntheticWeights:
    """Random weights of the real shapes (N(0, 0.02^2) BF16; projections that write into the residual stream scaled by
    1/sqrt(2 L), GPT-2 / Megatron init, so the stream stays token-specific through dozens of random layers and the router
    sees a realistic, not collapsed, load); INT4 experts as random tile bytes with zero-mean nibbles generated on the device."""

    def __init__(self, cfg: HybridMoEConfig, device, seed: int = 0):
        self.cfg, self.dev = cfg, device
        self.g = torch.Generator(device=device).manual_seed(1000 + seed)          # identical on every rank
        self.ge = torch.Generator(device=device).manual_seed(5000 + seed)         # global expert tensors; ranks keep a slice
        self.res_scale = 1.0 / math.sqrt(2.0 * cfg.num_hidden_layers)

    def rnd(self, *shape, std=0.02):
        return (torch.randn(*shape, device=self.dev, generator=self.g) * std).to(torch.bfloat16)

    def norm(self, n):
        return (1 + 0.05 * torch.randn(n, device=self.dev, generator=self.g)).to(torch.bfloat16)

    def globals_(self):
        c = self.cfg
        return dict(embedding=self.rnd(c.vocab_size, c.hidden_size), final_norm=self.norm(c.hidden_size),
                    lm_head=self.rnd(c.vocab_size, c.hidden_size))

    def attention(self, i: int, lt: str) -> dict:
        c, H, rs = self.cfg, self.cfg.hidden_size, self.res_scale
        if lt == "linear_attention":
            kd, vd = c.linear_num_key_heads * c.linear_key_head_dim, c.linear_num_value_heads * c.linear_value_head_dim
            return dict(in_proj_qkvz=self.rnd(2 * kd + 2 * vd, H), in_proj_ba=self.rnd(2 * c.linear_num_value_heads, H),
                        out_proj=self.rnd(H, vd, std=0.02 * rs), conv1d_weight=self.rnd(2 * kd + vd, 1, c.linear_conv_kernel_dim, std=0.3),
                        A_log=self.rnd(c.linear_num_value_heads, std=0.5), dt_bias=self.rnd(c.linear_num_value_heads, std=0.5),
                        norm_weight=self.norm(c.linear_value_head_dim))
        if lt == "mla":
            nh, nope, rope, dv, lora = c.num_attention_heads, c.qk_nope_head_dim, c.qk_rope_head_dim, c.v_head_dim, c.kv_lora_rank
            w = dict(kv_a_proj_with_mqa=self.rnd(lora + rope, H), kv_a_layernorm=self.norm(lora),
                     w_kc=self.rnd(nh, nope, lora, std=0.04), w_vc=self.rnd(nh, dv, lora, std=0.04),
                     o_proj=self.rnd(H, nh * dv, std=0.02 * rs))
            if c.q_lora_rank:
                w.update(q_a_proj=self.rnd(c.q_lora_rank, H), q_a_layernorm=self.norm(c.q_lora_rank),
                         q_b_proj=self.rnd(nh * (nope + rope), c.q_lora_rank, std=0.04))
            else:
                w["q_proj"] = self.rnd(nh * (nope + rope), H)
            return w
        nh, nkv, d = c.num_attention_heads, c.num_key_value_heads, c.gqa_head_dim
        return dict(q_proj=self.rnd(nh * d * (2 if c.gated_attention else 1), H), k_proj=self.rnd(nkv * d, H), v_proj=self.rnd(nkv * d, H),
                    o_proj=self.rnd(H, nh * d, std=0.02 * rs), q_norm=self.norm(d), k_norm=self.norm(d))

    def layer_norms(self, i: int):
        return self.norm(self.cfg.hidden_size), self.norm(self.cfg.hidden_size)

    def dense_mlp(self, i: int):
        c = self.cfg
        return self.rnd(2 * c.intermediate_size, c.hidden_size), self.rnd(c.hidden_size, c.intermediate_size, std=0.02 * self.res_scale)

    def router(self, i: int):
        return self.rnd(self.cfg.n_routed_experts, self.cfg.hidden_size, std=self.cfg.synthetic_router_std), None

    def shared(self, i: int):
        c, Is = self.cfg, self.cfg.shared_width
        return (self.rnd(2 * Is, c.hidden_size), self.rnd(c.hidden_size, Is, std=0.02 * self.res_scale),
                self.rnd(1, c.hidden_size, std=0.05) if c.shared_expert_gate else None)

    def attach_experts(self, engine: KrasisEngine, moe_idx: int, num_ranks: int):
        """Random packed nibbles + BF16 group scales in the B200 tile layout (bandwidth-faithful); returns the tensors."""
        c, dev, bf = self.cfg, self.dev, torch.bfloat16
        ts = []
        e_loc, e_all = engine.expert_end - engine.expert_start, c.n_routed_experts
        for which in range(4):
            n = engine.tiled_bytes(which)              # bytes of this rank's experts; tiles are [expert][...]
            n_all, off = n // e_loc * e_all, n // e_loc * engine.expert_start
            if which in (0, 2):
                if c.expert_bits == 4:    # nibbles 1..15 = q in [-7, 7]: zero mean, the range the reference quantiser emits
                    full = (torch.randint(1, 16, (n_all,), dtype=torch.uint8, device=dev, generator=self.ge) * 16
                            + torch.randint(1, 16, (n_all,), dtype=torch.uint8, device=dev, generator=self.ge))
                else:
                    full = torch.randint(0, 256, (n_all,), dtype=torch.uint8, device=dev, generator=self.ge)
                ts.append(full[off:off + n].clone() if num_ranks > 1 else full)
            else:                           # group scales: |w| ~ 0.02; the down projection carries the residual-branch scale
                sc = 1.0 if which == 1 else self.res_scale
                full = ((torch.rand(n_all // 2, device=dev, generator=self.ge) * 0.004 + 0.002) * sc).to(bf)
                ts.append(full[off // 2:(off + n) // 2].clone() if num_ranks > 1 else full)
            del full
        engine.attach_tiled_layer(moe_idx, *ts)
        return ts


class SafetensorsWeights:
    """HF checkpoint directory -> the same per-layer dictionaries, with the reference's load-time conventions
    (python/krasis/weight_loader.py; krasis_b200/loader.py): norm +1 for Qwen3-Next, Qwen3.5 in_proj re-interleave,
    kv_b_proj split, F32/F16 -> BF16 conversion, experts BF16 -> device group quantiser -> B200 tiles."""

    def __init__(self, cfg: HybridMoEConfig, model_dir: str, device):
        from . import loader
        self.cfg, self.dev, self.ld = cfg, device, loader
        self.t = loader.open_model_safetensors(model_dir)
        self.p = cfg.layers_prefix

    def _d(self, t):
        return t.to(self.dev)

    def globals_(self):
        ld, t, p = self.ld, self.t, self.p
        emb = ld._st_bf16(t, f"{p}.embed_tokens.weight")
        fn = ld._st_bf16(t, f"{p}.norm.weight")
        if self.cfg.norm_bias_one:
            fn = ld.norm_plus_one(fn)
        lm = None
        for name in ("lm_head.weight", f"{p}.lm_head.weight", f"{p.split('.')[0]}.lm_head.weight"):      # weight_loader.py:172-190
            if name in t:
                lm = ld._st_bf16(t, name)
                break
        if lm is None:
            lm = emb                                                                                     # tied embeddings
        return dict(embedding=self._d(emb), final_norm=self._d(fn), lm_head=self._d(lm))

    def attention(self, i: int, lt: str) -> dict:
        c, ld = self.cfg, self.ld
        if lt == "linear_attention":
            w = ld.load_linear_attention_weights(self.t, self.p, i, c.linear_num_key_heads, c.linear_key_head_dim,
                                                 c.linear_num_value_heads, c.linear_value_head_dim)
        elif lt == "mla":
            w = ld.load_mla_weights(self.t, self.p, i, c.num_attention_heads, c.qk_nope_head_dim, c.v_head_dim, bool(c.q_lora_rank))
        else:
            w = ld.load_gqa_weights(self.t, self.p, i, c.norm_bias_one)
        return w

    def layer_norms(self, i: int):
        n = self.ld.load_layer_norms(self.t, self.p, i, self.cfg.norm_bias_one)
        return n["input_layernorm"], n["post_attention_layernorm"]

    def dense_mlp(self, i: int):
        ld, p = self.ld, f"{self.p}.layers.{i}.mlp"
        gu = torch.cat([ld._st_bf16(self.t, f"{p}.gate_proj.weight"), ld._st_bf16(self.t, f"{p}.up_proj.weight")], dim=0)
        return self._d(gu), self._d(ld._st_bf16(self.t, f"{p}.down_proj.weight"))

    def router(self, i: int):
        return self.ld.load_router(self.t, self.p, i)

    def shared(self, i: int):
        ld, t = self.ld, self.t
        p = f"{self.p}.layers.{i}.mlp.shared_experts"
        if f"{p}.gate_proj.weight" not in t:
            p = f"{self.p}.layers.{i}.mlp.shared_expert"                                                 # weight_loader.py:343-352
        gu = torch.cat([ld._st_bf16(t, f"{p}.gate_proj.weight"), ld._st_bf16(t, f"{p}.up_proj.weight")], dim=0)
        gate_name = f"{self.p}.layers.{i}.mlp.shared_expert_gate.weight"
        gate = ld._st_bf16(t, gate_name) if gate_name in t else None
        return self._d(gu), self._d(ld._st_bf16(t, f"{p}.down_proj.weight")), (self._d(gate) if gate is not None else None)

    def attach_experts(self, engine: KrasisEngine, moe_idx: int, num_ranks: int):
        import numpy as np
        prefix = self.ld.expert_prefix(self.t.keys())
        w13, w2 = self.ld.read_layer_experts_bf16(self.t, prefix, moe_idx + self.cfg.first_k_dense_replace,
                                                  engine.expert_start, engine.expert_end)
        engine.load_bf16_layer(moe_idx, torch.from_numpy(w13.view(np.int16)).view(torch.bfloat16),
                               torch.from_numpy(w2.view(np.int16)).view(torch.bfloat16))
        return None


class DenseMLP:
    """TransformerLayer._dense_mlp_forward (layer.py:497-506) with QuantConfig.dense_mlp = "int8" (config.py:211): gate_proj and
    up_proj are quantised per row, so stacking them into one [2I, H] W8A8 GEMM is the same arithmetic.  The down
    projection's K (= intermediate size, 10944 for DeepSeek-V2-Lite) is padded with zero columns to a multiple of 128."""

    def __init__(self, gate_up_bf16: torch.Tensor, down_bf16: torch.Tensor):
        self.gate_up = L.quantize_to_int8(gate_up_bf16.contiguous())
        q, s = L.quantize_to_int8(down_bf16.contiguous())
        self.I = down_bf16.shape[1]
        self.Ip = (self.I + 127) // 128 * 128
        if self.Ip != self.I:
            qp = torch.zeros((q.shape[0], self.Ip), dtype=torch.int8, device=q.device)
            qp[:, :self.I] = q
            q = qp
        self.down = (q, s)

    def forward(self, hidden: torch.Tensor) -> torch.Tensor:
        act = L.silu_and_mul(L.int8_linear(hidden, *self.gate_up))
        if self.Ip != self.I:
            ap = torch.zeros((act.shape[0], self.Ip), dtype=act.dtype, device=act.device)
            ap[:, :self.I] = act
            act = ap
        return L.int8_linear(act, *self.down)


class _ChunkKVState:
    """View of a SequenceKVState `offset` tokens ahead: what a layer sees while it processes a later token chunk of the same prefill
    call (the real state is advanced once per model forward, after the last layer)."""

    def __init__(self, state, offset: int):
        self._st, self._off = state, offset

    @property
    def seq_len(self):
        return self._st.seq_len + self._off

    def ensure_capacity(self, num_new_tokens: int):
        self._st.ensure_capacity(self._off + num_new_tokens)

    def kv_indices(self, device):
        return self._st.kv_indices(device)


class KrasisModel:
    """Prefill forward of a MoE transformer: synthetic weights of the real architecture (`KrasisModel(cfg, ...)`) or a real
    checkpoint (`KrasisModel.from_pretrained(model_dir, ...)`)."""

    def __init__(self, cfg: HybridMoEConfig, device: int = 0, max_tokens: int = 8192, rank: int = 0, num_ranks: int = 1,
                 seed: int = 0, comm=None, keep_weights: bool = False, weights=None, gguf_path: Optional[str] = None,
                 group=None):
        self.cfg, self.rank, self.num_ranks, self.comm = cfg, rank, num_ranks, comm
        if num_ranks > 1 and comm is None:
            raise ValueError("num_ranks > 1 needs a krasis_b200.parallel.Communicator (comm=...)")
        self._keep = keep_weights            # tests: keep torch copies of the weights for the oracle
        self.device = torch.device("cuda", device)
        self.max_tokens = max_tokens
        dev = self.device
        H, nl, R = cfg.hidden_size, cfg.num_hidden_layers, num_ranks
        src = weights if weights is not None else SyntheticWeights(cfg, dev, seed)
        self.weights_source = type(src).__name__
        gl = src.globals_()
        self.embedding = gl["embedding"]
        self.final_norm = gl["final_norm"].float()
        self.lm_head = L.quantize_to_int8(gl["lm_head"].contiguous())                    # QuantConfig.lm_head = "int8"
        self._lm_head_bf16 = gl["lm_head"] if keep_weights else None
        gg = {}
        if gguf_path is not None:                                                        # native GGUF expert blocks
            from . import loader
            self._gguf = loader.GgufFile(gguf_path)
            first = cfg.first_k_dense_replace
            _, _, _, t13, t2 = loader.gguf_expert_blocks(self._gguf, first, 0, 1, H, cfg.moe_intermediate_size)
            gg = dict(gguf_gate_up_type=t13, gguf_down_type=t2)
        self.engine = KrasisEngine(hidden_size=H, moe_intermediate_size=cfg.moe_intermediate_size,
                                   n_routed_experts=cfg.n_routed_experts, num_experts_per_tok=cfg.num_experts_per_tok,
                                   num_moe_layers=max(1, cfg.num_moe_layers), num_bits=cfg.expert_bits, rank=rank, num_ranks=num_ranks,
                                   scoring_func=cfg.scoring_func, norm_topk_prob=cfg.norm_topk_prob,
                                   routed_scaling_factor=cfg.routed_scaling_factor, max_tokens=max_tokens, device=device, **gg)
        self.layer_types = [cfg.layer_type(i) for i in range(nl)]
        self._kv_layer_offsets = []
        off = 0
        for t in self.layer_types:                                      # model.py:485-487: -1 for linear layers
            self._kv_layer_offsets.append(off if t != "linear_attention" else -1)
            off += t != "linear_attention"
        # shared-expert mode (see the module docstring)
        self.shared_mode = "none"
        if cfg.shared_width > 0:
            self.shared_mode = "int8_gated" if cfg.shared_expert_gate else ("int4_manager" if R == 1 else "int8")
        self._shared_engine = None
        if self.shared_mode == "int4_manager":
            self._shared_engine = KrasisEngine(hidden_size=H, moe_intermediate_size=cfg.shared_width, n_routed_experts=1,
                                               num_experts_per_tok=1, num_moe_layers=max(1, cfg.num_moe_layers),
                                               num_bits=cfg.expert_bits, max_tokens=max_tokens, device=device)
        if R > 1:
            if cfg.num_attention_heads % R or (any(t == "linear_attention" for t in self.layer_types) and cfg.linear_num_key_heads % R):
                raise ValueError("head-parallel attention needs the head counts to be divisible by the number of ranks")
            if not cfg.is_mla:
                grp = cfg.num_attention_heads // cfg.num_key_value_heads
                if (cfg.num_attention_heads // R) % grp and grp % (cfg.num_attention_heads // R):
                    raise ValueError("query heads per rank must align with the KV groups")
        acfg = SimpleNamespace(hidden_size=H, num_attention_heads=cfg.num_attention_heads,
                               num_key_value_heads=cfg.num_key_value_heads, gqa_head_dim=cfg.gqa_head_dim,
                               rotary_dim=cfg.rotary_dim, rope_theta=cfg.rope_theta, rms_norm_eps=cfg.rms_norm_eps,
                               linear_num_key_heads=cfg.linear_num_key_heads // R, linear_num_value_heads=cfg.linear_num_value_heads // R,
                               linear_key_head_dim=cfg.linear_key_head_dim, linear_value_head_dim=cfg.linear_value_head_dim,
                               linear_conv_kernel_dim=cfg.linear_conv_kernel_dim, qk_nope_head_dim=cfg.qk_nope_head_dim,
                               qk_rope_head_dim=cfg.qk_rope_head_dim, v_head_dim=cfg.v_head_dim, kv_lora_rank=cfg.kv_lora_rank,
                               q_lora_rank=cfg.q_lora_rank, rope_scaling=cfg.rope_scaling)
        shared_handles = {}
        self.kv_heads_local = cfg.num_key_value_heads
        self.layers = []
        for i, lt in enumerate(self.layer_types):
            lay = SimpleNamespace(layer_type=lt, moe_idx=None, dense=None, shared_expert=None)
            n_in, n_post = src.layer_norms(i)
            lay.input_norm, lay.post_attn_norm = n_in.to(dev).float(), n_post.to(dev).float()
            w = src.attention(i, lt)
            lay._w = w if keep_weights else None
            if lt == "linear_attention":
                wl = shard_gdn_weights(w, cfg, rank, R) if R > 1 else w
                lay.attention = GatedDeltaNetAttention(acfg, i, wl, dev, max_tokens=max_tokens, share_scratch_with=shared_handles.get(lt))
            elif lt == "mla":
                wl, nh_l = shard_mla_weights(w, cfg, rank, R) if R > 1 else (w, cfg.num_attention_heads)
                mcfg = SimpleNamespace(**{**acfg.__dict__, "num_attention_heads": nh_l})
                lay.attention = MLAAttention(mcfg, i, wl, dev, max_tokens=max_tokens, max_kv_len=max_tokens,
                                             share_scratch_with=shared_handles.get(lt))
            else:
                if R > 1:
                    wl, nh_l, nkv_l = shard_gqa_weights(w, cfg, rank, R)
                    self.kv_heads_local = nkv_l
                else:
                    wl, nh_l, nkv_l = w, cfg.num_attention_heads, cfg.num_key_value_heads
                gcfg = SimpleNamespace(**{**acfg.__dict__, "num_attention_heads": nh_l, "num_key_value_heads": nkv_l})
                lay.attention = GQAAttention(gcfg, i, wl, dev, max_tokens=max_tokens, share_scratch_with=shared_handles.get(lt))
            shared_handles.setdefault(lt, lay.attention)
            del w
            if i < cfg.first_k_dense_replace:                                # layer.py:497-506
                dw = src.dense_mlp(i)
                lay.dense = DenseMLP(*dw)
                lay._dense_w = dw if keep_weights else None
            else:
                m = i - cfg.first_k_dense_replace
                lay.moe_idx = m
                if gguf_path is not None:
                    from . import loader
                    gate, up, down, _, _ = loader.gguf_expert_blocks(self._gguf, i, self.engine.expert_start, self.engine.expert_end,
                                                                     H, cfg.moe_intermediate_size)
                    self.engine.load_gguf_layer(m, gate, up, down)
                    lay._experts = None
                else:
                    ts = src.attach_experts(self.engine, m, R)
                    lay._experts = ts if keep_weights else None
                gate, corr = src.router(i)
                self.engine.set_routing_weights(m, gate, e_score_correction_bias=corr)
                lay._gate = gate if keep_weights else None
                if self.shared_mode != "none":
                    sw = src.shared(i)
                    lay._shared_w = sw if keep_weights else None
                    if self.shared_mode == "int4_manager":
                        self._shared_engine.load_bf16_layer(m, sw[0].unsqueeze(0), sw[1].unsqueeze(0))
                    else:
                        lay.shared_expert = L.SharedExpert(sw[0], sw[1], sw[2] if self.shared_mode == "int8_gated" else None)
                    del sw
            self.layers.append(lay)
        self._ones = self._zero_ids = None
        self._side_stream = None
        # chunk-pipelined attention collectives under token sharding: OFF by default.  Measured on 2 GPUs (profiles/r02p_*): 70.0 ms per
        # step against 65.1 ms with one all-gather + one reduce-scatter per layer — the NCCL kernels that now run next to the attention
        # hold SMs that the persistent one-CTA-per-SM GEMM / scan kernels are sized for (Gated DeltaNet 24.8 -> 32.0 ms), which costs
        # more than the 2 x 1 ms of hidden communication.  KB2_PIPELINE_ATTENTION=1 enables it.
        self.pipeline_attention = os.environ.get("KB2_PIPELINE_ATTENTION", "0") == "1"
        # expert-parallel reduce-scatter fused into the combine kernel over peer memory (KB2_FUSED_EP=0: NCCL reduce-scatter)
        self.fused_ep = os.environ.get("KB2_FUSED_EP", "1") != "0"
        self._ep_recv, self._ep_recv_rows = None, 0
        self._graphs = {}
        # attention out_proj GEMM -> reduce-scatter fused the same way: OFF by default.  Measured on 2 GPUs (profiles/r02s_*): the
        # reduce-scatter itself drops from 3.1 to 2.2 ms per step, but the GEMM epilogue holds its TMEM accumulator while half of its
        # rows cross NVLink (Gated DeltaNet 24.6 -> 28.0 ms, GQA 7.0 -> 8.2 ms): 67.0 ms per step against 63.7.  The combine kernel
        # of the experts is HBM-bound elementwise work and hides the remote stores; a tensor-core epilogue does not.
        self.fused_attn_rs = os.environ.get("KB2_FUSED_ATTN_RS", "0") == "1"
        self._attn_recv, self._attn_recv_rows = None, 0

    # ------------------------------------------------------------------------------------------- real checkpoints
    @classmethod
    def from_pretrained(cls, model_dir: str, device: int = 0, max_tokens: int = 8192, rank: int = 0, num_ranks: int = 1,
                        comm=None, expert_bits: int = 4, gguf_path: Optional[str] = None, max_layers: Optional[int] = None,
                        keep_weights: bool = False):
        """KrasisModel(model_path, ...).load() of the reference (python/krasis/model.py:350-611) for the prefill path: HF
        config.json + safetensors (attention / norms / router / shared expert / dense MLP / lm_head, experts quantised on the
        device with the Krasis group quantiser) or, with gguf_path, native GGUF expert blocks for the routed experts."""
        from . import loader
        raw = json.load(open(os.path.join(model_dir, "config.json")))
        tensors = loader.open_model_safetensors(model_dir)
        prefix = next(k.split(".layers.")[0] for k in tensors if ".layers.0." in k)                       # config.py:_detect_layers_prefix
        has_gate = any(k.endswith("mlp.shared_expert_gate.weight") for k in tensors)
        cfg = HybridMoEConfig.from_hf_config(raw, layers_prefix=prefix, has_shared_gate=has_gate)
        cfg.expert_bits = expert_bits
        if max_layers is not None:
            cfg.num_hidden_layers = min(cfg.num_hidden_layers, max_layers)
        src = SafetensorsWeights(cfg, model_dir, torch.device("cuda", device))
        src.t = tensors
        return cls(cfg, device=device, max_tokens=max_tokens, rank=rank, num_ranks=num_ranks, comm=comm, weights=src,
                   gguf_path=gguf_path, keep_weights=keep_weights)

    # ------------------------------------------------------------------------------------------- sequences / KV cache
    def _make_kv_cache(self):
        n_full = sum(t != "linear_attention" for t in self.layer_types)
        # two sequences' worth of pages: the live request plus the sequence a captured CUDA graph keeps for its replays (forward_graphed)
        pages = 2 * ((self.max_tokens + 15) // 16 + 1)
        if self.cfg.is_mla:
            self.kv_cache = MLAPagedKVCache(max(n_full, 1), self.cfg.kv_lora_rank, self.cfg.qk_rope_head_dim, self.device, max_pages=pages)
        else:
            self.kv_cache = PagedKVCache(max(n_full, 1), self.kv_heads_local, self.cfg.gqa_head_dim, self.device, max_pages=pages)

    def new_sequence(self) -> List[SequenceKVState]:
        if not hasattr(self, "kv_cache"):
            self._make_kv_cache()
        for lay in self.layers:
            if lay.layer_type == "linear_attention":
                lay.attention.reset_state()
        for old in getattr(self, "_live_seqs", []):      # single request at a time (src/server.rs header): recycle pages
            old.free()
        self._live_seqs = [SequenceKVState(self.kv_cache)]
        return self._live_seqs

    # ------------------------------------------------------------------------------------------- forward
    def _moe(self, lay, h, M_loc, h_q8=None):
        """Router + shared expert + routed experts of one MoE layer on this rank's token rows h [M_loc, H]."""
        tm, m, R = self._timer, lay.moe_idx, self.num_ranks
        if R > 1:
            # the gather of this rank's rows runs on a side stream while the router and the shared expert (which need only the
            # local rows) run on the main stream; ids / weights follow as soon as the router has produced them
            main = torch.cuda.current_stream(self.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(self.device)
            side = self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                h_all = self.comm.all_gather_rows(h)
        with tm("router"):
            ids, w = self.engine.compute_routing(m, h)
        if R > 1:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ids_all = self.comm.all_gather_rows(ids)
                w_all = self.comm.all_gather_rows(w)
        shared = None
        with tm("shared_expert"):
            if lay.shared_expert is not None:
                shared = lay.shared_expert.forward(h, h_q8) if h_q8 is not None else lay.shared_expert.forward(h)
            elif self.shared_mode == "int4_manager":                        # gpu_prefill.py:4738-4801: one-expert MoE, weight 1
                if self._ones is None or self._ones.shape[0] < h.shape[0]:
                    self._zero_ids = torch.zeros((self.max_tokens, 1), dtype=torch.int32, device=self.device)
                    self._ones = torch.ones((self.max_tokens, 1), dtype=torch.float32, device=self.device)
                shared = self._shared_engine.moe_forward(m, h, self._zero_ids[:h.shape[0]], self._ones[:h.shape[0]], routed_only=True)
        if R == 1:
            with tm("routed_experts"):
                return self.engine.moe_forward(m, h, ids, w, shared=shared)
        with tm("ep_all_gather"):
            main.wait_stream(side)                                          # gathered rows / ids / weights are in place
            for t in (h_all, ids_all, w_all):
                t.record_stream(main)
        if self.fused_ep:
            # reduce-scatter fused into the combine kernel: every rank's partial rows go straight into the owner's receive buffer over
            # NVLink; two buffers alternate by layer so that a fast rank cannot overwrite slots a slow rank is still summing
            rows = h.shape[0]
            if self._ep_recv is None or self._ep_recv_rows != rows:
                nbytes = rows * R * self.cfg.hidden_size * 2
                self._ep_recv = [self.comm.peer_alloc(nbytes), self.comm.peer_alloc(nbytes)]
                self._ep_recv_rows = rows
            ptrs = self._ep_recv[m & 1]
            with tm("routed_experts"):
                self.engine.moe_forward_scatter(m, h_all, ids_all, w_all, ptrs, self.rank)
            with tm("ep_reduce_scatter"):
                self.comm.barrier()
            with tm("routed_experts"):
                return self.engine.finish_slots(ptrs[self.rank], R, rows, shared, h)
        with tm("routed_experts"):
            part = self.engine.moe_forward(m, h_all, ids_all, w_all, routed_only=True)       # local expert slice, all tokens
        with tm("ep_reduce_scatter"):
            routed = self.comm.reduce_scatter_rows(part)
        with tm("routed_experts"):
            return self.engine.finish(routed, shared)                                        # bf16(rsf * routed) + shared

    def _attention_pipelined(self, lay, i, hidden, positions, st, M):
        """Head-parallel attention with the all-gather and the reduce-scatter cut into one chunk per rank (chunk j = rank j's token
        shard) and pipelined against the attention itself: broadcast_j brings rank j's normed rows to everyone, every rank runs its
        heads over chunk j as one step of a chunked prefill (conv / recurrent state and the KV cache carry over exactly like two
        prefill calls of the reference), reduce_j sums the partial o_proj outputs on rank j while chunk j + 1 is being computed.
        Exposed communication: the first broadcast and the last reduce, 1/R of each collective."""
        R, rows, tm = self.num_ranks, M // self.num_ranks, self._timer
        main = torch.cuda.current_stream(self.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(self.device)
        side = self._side_stream
        full = torch.empty((M, hidden.shape[1]), dtype=hidden.dtype, device=hidden.device)
        full[self.rank * rows:(self.rank + 1) * rows].copy_(hidden)
        ev_in = [torch.cuda.Event() for _ in range(R)]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for j in range(R):
                self.comm.broadcast(full[j * rows:(j + 1) * rows], root=j)
                ev_in[j].record(side)
        full.record_stream(side)
        own = torch.empty((rows, hidden.shape[1]), dtype=hidden.dtype, device=hidden.device)
        own.record_stream(side)
        name = "gdn_attention" if lay.layer_type == "linear_attention" else ("mla_attention" if lay.layer_type == "mla" else "gqa_attention")
        for j in range(R):
            main.wait_event(ev_in[j])
            x = full[j * rows:(j + 1) * rows]
            with tm(name):
                if lay.layer_type == "linear_attention":
                    part = lay.attention.forward(x, is_decode=False)
                else:
                    part = lay.attention.forward(x, positions[j * rows:(j + 1) * rows], self.kv_cache, _ChunkKVState(st, j * rows),
                                                 self._kv_layer_offsets[i], num_new_tokens=rows)
            ev = torch.cuda.Event()
            ev.record(main)
            part.record_stream(side)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                self.comm.reduce_rows(part, root=j, out=own if j == self.rank else None)
        with tm("attention_reduce_scatter"):
            main.wait_stream(side)
        return own

    def forward(self, token_ids: torch.Tensor, positions: torch.Tensor, seq_states: List[SequenceKVState],
                return_all_logits: bool = False) -> torch.Tensor:
        """model.py:2167: token_ids [M] int64/int32 on the device, positions [M]; returns logits [1, V] (last token)
        or [M, V] (return_all_logits) in float32 — on every rank."""
        cfg, R = self.cfg, self.num_ranks
        M = token_ids.shape[0]
        if M > self.max_tokens:
            raise ValueError(f"{M} tokens > max_tokens={self.max_tokens}")
        if M % R:
            raise ValueError(f"token-sharded prefill needs the token count ({M}) to be a multiple of the rank count ({R})")
        tm = self._timer
        st = seq_states[0]
        lo, hi = self.rank * (M // R), (self.rank + 1) * (M // R)
        hidden = self.embedding[token_ids[lo:hi].long()]                    # row gather (model.py:2744), this rank's rows
        residual = None
        eps = cfg.rms_norm_eps
        for i, lay in enumerate(self.layers):
            with tm("norms"):
                if residual is None:                                       # layer.py:275-285
                    residual = hidden
                    hidden = L.rmsnorm(hidden, lay.input_norm, eps)
                else:
                    L.fused_add_rmsnorm(hidden, residual, lay.input_norm, eps)
            if R > 1 and self.pipeline_attention:
                attn = self._attention_pipelined(lay, i, hidden, positions, st, M)
            else:
                if R > 1:
                    with tm("attention_all_gather"):
                        hidden = self.comm.all_gather_rows(hidden)             # head-parallel attention sees every token
                fused_rs = R > 1 and self.fused_attn_rs
                if fused_rs:
                    # GEMM -> reduce-scatter fused: the out_proj epilogue stores its rows into the token owners' receive buffers
                    rows = M // R
                    if self._attn_recv is None or self._attn_recv_rows != rows:
                        nbytes = rows * R * cfg.hidden_size * 2
                        self._attn_recv = [self.comm.peer_alloc(nbytes), self.comm.peer_alloc(nbytes)]
                        self._attn_recv_rows = rows
                    ptrs = self._attn_recv[i & 1]
                    lay.attention.set_output_scatter(ptrs, self.rank)
                if lay.layer_type == "linear_attention":
                    with tm("gdn_attention"):
                        attn = lay.attention.forward(hidden, is_decode=False)
                else:
                    with tm("mla_attention" if lay.layer_type == "mla" else "gqa_attention"):
                        attn = lay.attention.forward(hidden, positions, self.kv_cache, st, self._kv_layer_offsets[i], num_new_tokens=M)
                if fused_rs:
                    lay.attention.set_output_scatter(None)
                    with tm("attention_reduce_scatter"):
                        self.comm.barrier()
                        attn = L.sum_slots(ptrs[self.rank], R, rows, cfg.hidden_size, self.device)
                elif R > 1:
                    with tm("attention_reduce_scatter"):
                        attn = self.comm.reduce_scatter_rows(attn)             # sum of the partial o_proj outputs, this rank's rows
            h_q8 = None
            with tm("norms"):
                if lay.dense is None and isinstance(lay.shared_expert, L.SharedExpert):
                    h_q8 = L.fused_add_rmsnorm_q8(attn, residual, lay.post_attn_norm, eps)   # + the shared expert's activation quantisation
                else:
                    L.fused_add_rmsnorm(attn, residual, lay.post_attn_norm, eps)   # layer.py:305-309
            if lay.dense is not None:
                with tm("dense_mlp"):
                    hidden = lay.dense.forward(attn)
            else:
                hidden = self._moe(lay, attn, hi - lo, h_q8)
        st.advance(M)
        with tm("final_norm_lm_head"):
            L.fused_add_rmsnorm(hidden, residual, self.final_norm, eps)    # model.py:3380-3386
            if return_all_logits:
                out = L.int8_linear(hidden, *self.lm_head)
                if R > 1:
                    out = self.comm.all_gather_rows(out)
                return out.float()
            out = torch.empty((1, cfg.vocab_size), dtype=torch.float32, device=self.device)
            if self.rank == R - 1:                                          # the last token lives on the last rank
                out.copy_(L.int8_linear(hidden[-1:].contiguous(), *self.lm_head).float())
            if R > 1:
                self.comm.broadcast(out, root=R - 1)
        return out

    # ---- fixed-shape prefill replayed from a CUDA graph
    def forward_graphed(self, token_ids: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """One prefill of a FRESH sequence of token_ids.shape[0] tokens, replayed from a CUDA graph (captured on first use per length;
        token ids and positions are copied into static buffers).  A whole-model step is ~110 kernel launches per layer: at one GPU the
        device is the bottleneck, but with token sharding over N GPUs every kernel shrinks and the Python / ctypes launch path becomes
        it.  Everything the step launches — our kernels through the C ABI, the NCCL collectives, the row gather of the embedding — is
        stream-ordered on torch's current stream, so the step captures as is.  Falls back to the eager path if capture fails."""
        M = token_ids.shape[0]
        ent = self._graphs.get(M)
        if ent is None:
            ent = self._capture_graph(token_ids, positions)
            self._graphs[M] = ent
        if ent is False:
            return self.forward(token_ids, positions, self.new_sequence())
        graph, st_tok, st_pos, out, seq = ent
        st_tok.copy_(token_ids, non_blocking=True)
        st_pos.copy_(positions, non_blocking=True)
        graph.replay()
        seq[0].seq_len = M                       # host-side bookkeeping of what the replay did on the device
        return out

    def _capture_graph(self, token_ids, positions):
        M = token_ids.shape[0]
        try:
            st_tok, st_pos = token_ids.clone(), positions.clone()
            if not hasattr(self, "kv_cache"):
                self._make_kv_cache()
            # the graph owns this sequence: its page table is baked into the captured launches, so it is never recycled by new_sequence()
            seq = [SequenceKVState(self.kv_cache)]
            seq[0].ensure_capacity(M)
            seq[0].kv_indices(self.device)       # page table on the device before capture (host -> device copies cannot be captured)

            def fresh_step():
                seq[0].seq_len = 0
                for lay in self.layers:
                    if lay.layer_type == "linear_attention":
                        lay.attention.reset_state()
                return self.forward(st_tok, st_pos, seq)

            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):        # warm-up off the default stream: lazy allocations (peer buffers, scratch) happen here
                for _ in range(2):
                    fresh_step()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fresh_step()
            return (graph, st_tok, st_pos, out, seq)
        except Exception as exc:                 # noqa: BLE001 — any capture problem means "use the eager path", never a wrong result
            import warnings
            warnings.warn(f"CUDA graph capture of the prefill step failed ({exc}); using eager launches")
            try:
                torch.cuda.synchronize(self.device)
            except Exception:                    # noqa: BLE001
                pass
            return False

    # ---- optional per-component device timing (the reference's KRASIS_LAYER_TIMING, model.py:2832,3355-3373)
    class _Span:
        def __init__(self, model, name):
            self.m, self.n = model, name

        def __enter__(self):
            if self.m.timing:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if self.m.timing:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.m._spans.append((self.n, self.e0, e1))

    timing = False

    def _timer(self, name):
        return KrasisModel._Span(self, name)

    def timing_start(self):
        self.timing, self._spans = True, []

    def timing_collect(self):
        torch.cuda.synchronize()
        out = {}
        for n, a, b in self._spans:
            out[n] = out.get(n, 0.0) + a.elapsed_time(b)
        self.timing, self._spans = False, []
        return out
