"""Writes tests/golden/reference_kats.json and tests/golden/oracle_vectors.npz.

reference_kats.json  — the known-answer vectors the reference's own Rust unit tests hold
                       for this path, transcribed as (input recipe, expected, tolerance)
                       with the Rust test that owns each one (file:line).  Nothing is
                       computed by the oracle here: expectations are the reference's.
oracle_vectors.npz   — frozen oracle outputs on seeded inputs (regression pins for the
                       oracle itself and small fixtures for the GPU parity tests, so the
                       GPU box does not need to recompute slow numpy loops).

Run from the repo root:  python tests/golden/make_golden.py
(/root/reference is NOT needed: the reference cannot be executed here — no Rust
toolchain — so there is nothing to import from it)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import quant, router, moe as omoe, gguf_blocks as G  # noqa: E402
from oracle.bf16 import f32_to_bf16_bits, bf16_bits_to_f32, round_bf16  # noqa: E402


def kats():
    return {
        "_about": "Known-answer tests transcribed from the reference's Rust unit tests.",
        "gguf_block_table": {  # src/gguf.rs:892-899 + :56-85
            "src": "src/gguf.rs:892-899",
            "entries": {"Q4_K": [256, 144], "Q5_K": [256, 176], "Q6_K": [256, 210], "F32": [1, 4],
                        "Q8_0": [32, 34], "Q4_0": [32, 18], "Q5_0": [32, 22]},
        },
        "get_scale_min_k4": {  # src/gguf.rs:902-908
            "src": "src/gguf.rs:902-908",
            "scales": [0x3F, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01, 0x80, 0, 0, 0, 0],
            "j": 0, "expect_sc": 0x3F, "expect_mn": 0x04,
        },
        "q4_0_all_0x88": {  # src/gguf_kernels.rs:770-778
            "src": "src/gguf_kernels.rs:770-778", "d": 1.0, "byte": 0x88, "x": 1.0,
            "expect": 0.0, "tol": 1e-6,
        },
        "q8_0_simple": {  # src/gguf_kernels.rs:782-792
            "src": "src/gguf_kernels.rs:782-792", "d": 0.1, "q": 10, "x": 1.0, "n": 32,
            "expect": 32.0, "tol": 0.5,
        },
        "int16_quant_roundtrip_g32": {  # src/gguf_kernels.rs:795-828
            "src": "src/gguf_kernels.rs:795-828", "k": 64,
            "recipe": "bf16_truncate((i - 32) * 0.1)", "max_abs_err": 0.01, "sums_exact": True,
        },
        "int4_roundtrip_synthetic": {  # src/weights/marlin.rs:851-884
            "src": "src/weights/marlin.rs:851-884", "rows": 4, "cols": 128, "group_size": 128,
            "recipe": "bf16_rne((i/len - 0.5) * 0.2)", "max_err_lt": 0.02,
        },
        "marlin_repack_roundtrip": {  # src/weights/marlin.rs:969-1009
            "src": "src/weights/marlin.rs:969-1009", "n": 64, "k": 128, "group_size": 128,
            "recipe": "bf16_rne((i/len - 0.5) * 0.2)", "max_diff": 0.0,
        },
        "perm_tables": {"src": "src/weights/marlin.rs:936-966", "weight_perm_len": 1024,
                        "scale_perm_len": 64, "scale_perm_single_len": 32},
        "moe_behaviour": {
            "src": "src/moe.rs:3892-3916",
            "all_ids_minus1_gives_zero": True, "identical_tokens_identical_rows_tol": 1e-6,
        },
        "synthetic_generators": {  # src/moe.rs:3316-3334, :3503
            "src": "src/moe.rs:3316-3334,3503",
            "gate_down": "bf16_rne(((i/len) - 0.5) * 0.1)", "up": "bf16_rne(((i/len) - 0.3) * 0.1)",
            "act_a": "bf16_rne((((3i+1)/H) - 0.5) * 0.2)", "act_b": "bf16_rne((((7i+13)/H) - 0.5) * 0.1)",
        },
    }


def oracle_vectors():
    rng = np.random.default_rng(20260923)
    out = {}
    # --- quantiser
    w = f32_to_bf16_bits(rng.normal(0, 0.02, (64, 256)).astype(np.float32))
    p4, s4 = quant.quantize_int4(w)
    q8, s8 = quant.quantize_int8(w)
    out.update(qw=w, q4_packed=p4, q4_scales=s4, q8_data=q8, q8_scales=s8)
    # --- router (E=64, k=6, H=128)
    hid = round_bf16(rng.normal(0, 1, (96, 128)).astype(np.float32))
    gate = round_bf16(rng.normal(0, 0.02, (64, 128)).astype(np.float32))
    ids, rw = router.compute_routing(hid, gate, 6)
    ids_n, rw_n = router.compute_routing(hid, gate, 6, norm_topk_prob=True)
    out.update(r_hidden=hid, r_gate=gate, r_ids=ids, r_w=rw, r_w_norm=rw_n)
    # --- small MoE layer in both numerics (E=8, H=256, I=128, k=2, M=24)
    layer = omoe.make_int_layer(rng, 8, 256, 128, bits=4)
    x = round_bf16(rng.normal(0, 1, (24, 256)).astype(np.float32))
    g2 = round_bf16(rng.normal(0, 0.05, (8, 256)).astype(np.float32))
    mids, mw = router.compute_routing(x, g2, 2, norm_topk_prob=True)
    y_gpu = omoe.moe_forward_gpu_path(layer, x, mids, mw)
    y_cpu = omoe.moe_forward_cpu_int(layer, x, mids, mw)
    out.update(m_w13_q=layer.w13_q, m_w13_s=layer.w13_s, m_w2_q=layer.w2_q, m_w2_s=layer.w2_s,
               m_x=x, m_ids=mids, m_w=mw, m_y_gpu=y_gpu, m_y_cpu=y_cpu)
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **oracle_vectors())
    print("wrote", os.listdir(HERE))
