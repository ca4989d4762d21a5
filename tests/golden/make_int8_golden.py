"""Generates tests/golden/int8_reference.npz by EXECUTING THE REFERENCE's own Python code on CPU:
python/krasis/weight_loader.py quantize_to_int8 (:25-43) and int8_linear (:46-99, torch._int_mm).
`safetensors` and krasis.config are stubbed (imported at module scope, untouched by the two functions).
Run:  python tests/golden/make_int8_golden.py      (build container only: needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python/krasis"


def main():
    pkg = types.ModuleType("krasis")
    pkg.__path__ = [REF]
    sys.modules["krasis"] = pkg
    cfgm = types.ModuleType("krasis.config")
    cfgm.ModelConfig = cfgm.PPRankConfig = cfgm.QuantConfig = object
    sys.modules["krasis.config"] = cfgm
    import krasis.weight_loader as wl
    torch.manual_seed(42)
    M, K, N = 40, 512, 192
    x = torch.randn(M, K).to(torch.bfloat16)
    x[3] = 0                                             # all-zero row: scale clamps at 1e-10 / 127
    w = (torch.randn(N, K) * 0.03).to(torch.bfloat16)
    wq, ws = wl.quantize_to_int8(w)
    y = wl.int8_linear(x, wq, ws)
    bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)
    np.savez_compressed(os.path.join(HERE, "int8_reference.npz"), x=bits(x), w=bits(w), wq=wq.numpy(), ws=bits(ws), y=bits(y))
    print("wrote int8_reference.npz", tuple(y.shape))


if __name__ == "__main__":
    main()
