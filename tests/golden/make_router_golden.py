"""Generates tests/golden/router_reference.npz by EXECUTING THE REFERENCE's own Python code on CPU:
python/krasis/layer.py TransformerLayer.compute_routing (:526-560) — fp32 matmul, softmax / sigmoid / GPT-OSS scoring,
selection bias, top-k, renormalisation — for the four routing flavours of the reference's model families.

layer.py imports flashinfer and sibling modules at module scope; they are stubbed (none is touched by compute_routing).
Run:  python tests/golden/make_router_golden.py      (build container only: needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python/krasis"


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
    # layer.py imports more siblings lazily / at module scope: give every missing krasis.* module an empty stub
    import importlib.abc
    import importlib.machinery

    class _Stub(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path, target=None):
            if fullname.startswith("krasis.") and not os.path.exists(os.path.join(REF, fullname.split(".", 1)[1] + ".py")):
                return importlib.machinery.ModuleSpec(fullname, self)
            return None

        def create_module(self, spec):
            m = types.ModuleType(spec.name)
            m.__getattr__ = lambda name: object
            return m

        def exec_module(self, module):
            pass

    sys.meta_path.insert(0, _Stub())
    import krasis.layer as layer
    return layer


CASES = {   # name: (E, k, scoring, norm_topk_prob, swiglu_limit, gate_bias, correction_bias)
    "qwen_softmax_norm": (64, 8, "softmax", True, 0.0, False, False),
    "deepseek_softmax": (64, 6, "softmax", False, 0.0, False, False),
    "kimi_sigmoid_bias_norm": (128, 8, "sigmoid", True, 0.0, False, True),
    "gptoss_topk_softmax": (32, 4, "softmax", False, 7.0, True, False),
}


def main():
    layer = load_reference()
    out = {}
    for idx, (name, (E, k, scoring, norm, swiglu, has_gb, has_cb)) in enumerate(CASES.items()):
        torch.manual_seed(700 + idx)
        M, H = 96, 256
        obj = object.__new__(layer.TransformerLayer)
        obj.cfg = types.SimpleNamespace(num_experts_per_tok=k, scoring_func=scoring, norm_topk_prob=norm, swiglu_limit=swiglu)
        obj.gate_weight = (torch.randn(E, H) * 0.05).to(torch.bfloat16)
        obj.gate_bias = (torch.randn(E) * 0.1).float() if has_gb else None
        obj.e_score_correction_bias = (torch.randn(E) * 0.05).to(torch.bfloat16) if has_cb else None
        hidden = torch.randn(M, H).to(torch.bfloat16)
        ids, w = obj.compute_routing(hidden)
        bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)          # BF16 tensors stored as their bit patterns
        out.update({f"{name}_hidden": bits(hidden), f"{name}_gate": bits(obj.gate_weight),
                    f"{name}_ids": ids.numpy(), f"{name}_w": w.numpy()})
        if has_gb:
            out[f"{name}_gate_bias"] = obj.gate_bias.numpy()
        if has_cb:
            out[f"{name}_corr_bias"] = bits(obj.e_score_correction_bias)
    np.savez_compressed(os.path.join(HERE, "router_reference.npz"), **out)
    print("wrote router_reference.npz", sorted(out)[:6], "...")


if __name__ == "__main__":
    main()
