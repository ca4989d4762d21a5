"""Generates tests/golden/mla_rope_reference.npz by EXECUTING THE REFERENCE's own Python code on CPU:
python/krasis/attention.py MLAAttention._get_rope_cos_sin (YaRN inverse frequencies + BF16 tables, :119-163),
._deinterleave / ._apply_rope (:165-211), the sm_scale arithmetic of __init__ (:79-88), and GQAAttention._get_rope_cos_sin /
._apply_rope (:443-494, partial and full rotary).

The attention core itself is FlashInfer (third-party, GPU only) and cannot be executed here; these are the parts of the
MLA path that live in the reference tree.  flashinfer / krasis.config / kv_cache / timing / weight_loader are stubbed
because attention.py imports them at module scope; none of them is touched by the functions executed.

Run:  python tests/golden/make_mla_golden.py      (build container only: needs /root/reference)
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python/krasis"

V2_LITE_ROPE = {"beta_fast": 32, "beta_slow": 1, "factor": 40, "mscale": 0.707, "mscale_all_dim": 0.707,
                "original_max_position_embeddings": 4096, "type": "yarn"}


def load_reference():
    pkg = types.ModuleType("krasis")
    pkg.__path__ = [REF]
    sys.modules["krasis"] = pkg
    for name, attrs in (("flashinfer", {}), ("krasis.config", {"ModelConfig": object}),
                        ("krasis.kv_cache", {"PagedKVCache": object, "SequenceKVState": object}),
                        ("krasis.timing", {"TIMING": types.SimpleNamespace()}),
                        ("krasis.weight_loader", {"int8_linear": None})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    import krasis.attention as att
    return att


def main():
    att = load_reference()
    out = {}
    for tag, scaling in (("yarn", V2_LITE_ROPE), ("plain", None)):
        obj = object.__new__(att.MLAAttention)
        obj.cfg = types.SimpleNamespace(rope_scaling=scaling or {})
        obj.device = torch.device("cpu")
        obj.qk_rope_dim, obj.rope_theta, obj._rope_cos_sin = 64, 10000.0, None
        cos, sin = obj._get_rope_cos_sin(5000)
        torch.manual_seed(7)
        pos = torch.tensor([0, 1, 2, 17, 255, 1023, 4095, 4999])
        q_pe = torch.randn(len(pos), 4, 64).to(torch.bfloat16)
        k_pe = torch.randn(len(pos), 1, 64).to(torch.bfloat16)
        q_r, k_r = obj._apply_rope(q_pe, k_pe, pos)
        # sm_scale exactly as __init__ computes it
        sm = 1.0 / math.sqrt(192)
        if scaling and scaling.get("factor", 1.0) > 1.0:
            ms = 0.1 * scaling.get("mscale_all_dim", 0) * math.log(scaling["factor"]) + 1.0
            sm *= ms * ms
        rows = torch.unique(torch.cat([torch.arange(64), pos, torch.arange(0, 5000, 97)]))     # keep the fixture small
        out.update({f"{tag}_rows": rows.numpy(), f"{tag}_cos": cos[rows].float().numpy(), f"{tag}_sin": sin[rows].float().numpy(),
                    f"{tag}_pos": pos.numpy(),
                    f"{tag}_q_in": q_pe.float().numpy(), f"{tag}_k_in": k_pe.float().numpy(),
                    f"{tag}_q_out": q_r.float().numpy(), f"{tag}_k_out": k_r.float().numpy(),
                    f"{tag}_sm_scale": np.float64(sm)})
    # GQAAttention._get_rope_cos_sin / _apply_rope (attention.py:443-494): partial (QCN: 64 of 256) and full (235B: 128) rotary
    for tag, d, rot, theta in (("gqa_partial", 256, 64, 10000000.0), ("gqa_full", 128, 128, 1000000.0)):
        obj = object.__new__(att.GQAAttention)
        obj.device, obj.rotary_dim, obj.rope_theta, obj._rope_cos_sin = torch.device("cpu"), rot, theta, None
        torch.manual_seed(9)
        pos = torch.tensor([0, 1, 5, 100, 4095, 8191])
        q = torch.randn(len(pos), 3, d).to(torch.bfloat16)
        k = torch.randn(len(pos), 2, d).to(torch.bfloat16)
        q_r, k_r = obj._apply_rope(q, k, pos)
        cos, sin = obj._get_rope_cos_sin(8192)
        out.update({f"{tag}_pos": pos.numpy(), f"{tag}_q_in": q.float().numpy(), f"{tag}_k_in": k.float().numpy(),
                    f"{tag}_q_out": q_r.float().numpy(), f"{tag}_k_out": k_r.float().numpy(),
                    f"{tag}_cos": cos[pos].float().numpy(), f"{tag}_sin": sin[pos].float().numpy(),
                    f"{tag}_cfg": np.array([d, rot, theta], np.float64)})
    np.savez_compressed(os.path.join(HERE, "mla_rope_reference.npz"), **out)
    print("wrote mla_rope_reference.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
