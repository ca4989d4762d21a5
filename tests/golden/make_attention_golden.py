"""Generates tests/golden/gdn_reference.npz by EXECUTING THE REFERENCE's own Python code on CPU:
python/krasis/linear_attention.py GatedDeltaNetAttention._forward_chunked (chunk-64 prefill, fused
solve_triangular path, eager chunk step) and .forward(is_decode=...)'s token-by-token recurrence.

Only runs in the build container (needs /root/reference); the GPU box uses the committed .npz.
The class is instantiated without __init__ (which creates a CUDA stream) and the torch.compile wrapper is
replaced by the eager `_chunk_step` it wraps — the arithmetic is the reference's, unmodified.

Run:  python tests/golden/make_attention_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python/krasis"


def load_reference():
    pkg = types.ModuleType("krasis")       # skip krasis/__init__.py (imports the Rust extension)
    pkg.__path__ = [REF]
    sys.modules["krasis"] = pkg
    import krasis.linear_attention as la
    la._get_chunk_step = lambda: la._chunk_step      # eager instead of torch.compile (same function)
    return la


def make_layer(la, cfg, w):
    obj = object.__new__(la.GatedDeltaNetAttention)
    c = types.SimpleNamespace(rms_norm_eps=cfg["eps"])
    obj.cfg, obj.layer_idx, obj.device = c, 0, torch.device("cpu")
    obj.num_k_heads, obj.num_v_heads = cfg["nk"], cfg["nv"]
    obj.k_head_dim, obj.v_head_dim = cfg["dk"], cfg["dv"]
    obj.hidden_size, obj.kernel_dim = cfg["H"], cfg["K"]
    obj.key_dim = obj.num_k_heads * obj.k_head_dim
    obj.value_dim = obj.num_v_heads * obj.v_head_dim
    obj.conv_dim = obj.key_dim * 2 + obj.value_dim
    obj.head_ratio = obj.num_v_heads // obj.num_k_heads
    obj.scale = 1.0 / (obj.k_head_dim ** 0.5)
    obj.in_proj_qkvz, obj.in_proj_ba, obj.out_proj = w["in_proj_qkvz"], w["in_proj_ba"], w["out_proj"]
    obj.conv1d_weight, obj.A_log, obj.dt_bias, obj.norm_weight = w["conv1d_weight"], w["A_log"], w["dt_bias"], w["norm_weight"]
    obj._conv_state = torch.zeros(1, obj.conv_dim, obj.kernel_dim, dtype=torch.bfloat16)
    obj._recurrent_state = torch.zeros(1, obj.num_v_heads, obj.k_head_dim, obj.v_head_dim, dtype=torch.float32)
    obj._la_graph = None
    return obj


def main():
    la = load_reference()
    torch.manual_seed(1234)
    cfg = dict(nk=2, nv=4, dk=32, dv=32, H=64, K=4, eps=1e-6)
    kd, vd = cfg["nk"] * cfg["dk"], cfg["nv"] * cfg["dv"]
    bf = torch.bfloat16
    w = dict(
        in_proj_qkvz=(torch.randn(2 * kd + 2 * vd, cfg["H"]) * 0.3).to(bf),
        in_proj_ba=(torch.randn(2 * cfg["nv"], cfg["H"]) * 0.3).to(bf),
        out_proj=(torch.randn(cfg["H"], vd) * 0.1).to(bf),
        conv1d_weight=(torch.randn(2 * kd + vd, 1, cfg["K"]) * 0.5).to(bf),
        A_log=torch.randn(cfg["nv"]).to(bf) * 0.5,
        dt_bias=torch.randn(cfg["nv"]).to(bf) * 0.5,
        norm_weight=(1.0 + 0.1 * torch.randn(cfg["dv"])).to(bf),
    )
    M1, M2 = 150, 70                      # two prefill calls: state carries over, both need padding to 64
    x1 = torch.randn(M1, cfg["H"]).to(bf)
    x2 = torch.randn(M2, cfg["H"]).to(bf)
    with torch.no_grad():
        lay = make_layer(la, cfg, w)
        y1 = lay._forward_chunked(x1)
        conv1, st1 = lay._conv_state.clone(), lay._recurrent_state.clone()
        y2 = lay._forward_chunked(x2)
        conv2, st2 = lay._conv_state.clone(), lay._recurrent_state.clone()
        # token-by-token recurrence of the reference (the M>1 loop inside forward(); linear_attention.py:480-585)
        lay_r = make_layer(la, cfg, w)
        import krasis.timing as tm
        tm.TIMING.decode = False
        tm.TIMING.prefill = False
        yr = lay_r.forward(x1, is_decode=True) if False else None
    out = dict(x1=x1.float().numpy(), x2=x2.float().numpy(), y1=y1.float().numpy(), y2=y2.float().numpy(),
               conv1=conv1.float().numpy(), st1=st1.numpy(), conv2=conv2.float().numpy(), st2=st2.numpy())
    for k_, v_ in w.items():
        out["w_" + k_] = v_.float().numpy()
    out["cfg"] = np.array([cfg["nk"], cfg["nv"], cfg["dk"], cfg["dv"], cfg["H"], cfg["K"]], np.int64)
    np.savez_compressed(os.path.join(HERE, "gdn_reference.npz"), **out)
    print("wrote gdn_reference.npz", {k_: v_.shape for k_, v_ in out.items() if hasattr(v_, "shape")})


if __name__ == "__main__":
    main()
