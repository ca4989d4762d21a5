"""Generates tests/golden/marlin_inverse_reference.npz by EXECUTING THE REFERENCE's own Python code on CPU:
python/krasis/triton_moe.py inverse_marlin_repack (:73-135) and inverse_scale_permute (:138-180) on random Marlin-order
words — the reference's own statement of the INT4 Marlin layout (weight permutation, tile transpose, nibble packing, scale
permutation), against which krasis_b200.marlin_cache and oracle.quant are pinned.
Run:  python tests/golden/make_marlin_golden.py      (build container only: needs /root/reference; imports triton)
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
    import krasis.triton_moe as tm
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (K, N, gs) in {"grouped": (256, 128, 128), "single_group": (128, 64, 128)}.items():
        wm = torch.randint(-2 ** 31, 2 ** 31 - 1, (2, K // 16, 2 * N), dtype=torch.int64, generator=g).to(torch.int32)
        sm = torch.randint(0, 2 ** 15, (2, K // gs, N), dtype=torch.int64, generator=g).to(torch.int16).view(torch.bfloat16)
        p = tm.inverse_marlin_repack(wm, K, N, 4)
        s = tm.inverse_scale_permute(sm, K, N, gs)
        out.update({f"{tag}_wm": wm.numpy().view(np.uint32), f"{tag}_sm": sm.view(torch.int16).numpy().view(np.uint16),
                    f"{tag}_packed": p.numpy().view(np.uint32), f"{tag}_scales": s.contiguous().view(torch.int16).numpy().view(np.uint16),
                    f"{tag}_cfg": np.array([K, N, gs])})
    np.savez_compressed(os.path.join(HERE, "marlin_inverse_reference.npz"), **out)
    print("wrote marlin_inverse_reference.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
