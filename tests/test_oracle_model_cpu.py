"""CPU checks of the whole-model oracle and of the checkpoint -> config path (no GPU): the tiny real-format checkpoints of
tests/test_gpu_pretrained.py parse into the expected configuration, the loader applies the load-time conventions, and the
oracle stack runs end to end on them."""
import json

import numpy as np
import torch

from krasis_b200 import loader as Ld
from krasis_b200.model import HybridMoEConfig
from oracle import model as OM
from tests.test_gpu_pretrained import build_qwen3_next_checkpoint, build_v2lite_checkpoint
from tests.test_loader_cpu import _write_safetensors


def test_v2lite_checkpoint_config_and_oracle_forward(tmp_path):
    hf, t, W = build_v2lite_checkpoint()
    cfg = HybridMoEConfig.from_hf_config(hf, has_shared_gate=False)
    assert cfg.is_mla and cfg.layer_type(0) == "mla" and cfg.first_k_dense_replace == 1 and cfg.num_moe_layers == 2
    assert cfg.shared_width == 2 * cfg.moe_intermediate_size and not cfg.norm_topk_prob and not cfg.norm_bias_one
    tok = torch.randint(0, cfg.vocab_size, (40,), generator=torch.Generator().manual_seed(0))
    out = OM.forward(cfg, W, tok, torch.arange(40), "int4_manager")
    assert out.shape == (40, cfg.vocab_size) and torch.isfinite(out).all() and out.std() > 0


def test_qwen3_next_checkpoint_conventions_and_oracle_forward(tmp_path):
    hf, t, W = build_qwen3_next_checkpoint()
    json.dump(hf, open(tmp_path / "config.json", "w"))
    _write_safetensors(tmp_path / "model.safetensors", t)
    cfg = HybridMoEConfig.from_hf_config(hf, has_shared_gate=True)
    assert cfg.norm_bias_one and cfg.gated_attention and cfg.norm_topk_prob and cfg.full_attention_interval == 4
    ts = Ld.open_model_safetensors(str(tmp_path))
    # the +1 convention restores the effective weights exactly; F32-stored A_log / dt_bias are converted, not reinterpreted
    n = Ld.load_layer_norms(ts, "model", 0, True)
    assert torch.equal(n["input_layernorm"], W["layers"][0]["input_norm"])
    la = Ld.load_linear_attention_weights(ts, "model", 0, 2, 128, 4, 128)
    assert torch.equal(la["A_log"], W["layers"][0]["attn"]["A_log"]) and la["A_log"].dtype == torch.bfloat16
    assert torch.equal(la["norm_weight"], W["layers"][0]["attn"]["norm_weight"])          # gated norm is NOT shifted
    ga = Ld.load_gqa_weights(ts, "model", 3, True)
    assert torch.equal(ga["q_norm"], W["layers"][3]["attn"]["q_norm"])
    tok = torch.randint(0, cfg.vocab_size, (40,), generator=torch.Generator().manual_seed(0))
    out = OM.forward(cfg, W, tok, torch.arange(40), "int8_gated")
    assert out.shape == (40, cfg.vocab_size) and torch.isfinite(out).all() and out.std() > 0
