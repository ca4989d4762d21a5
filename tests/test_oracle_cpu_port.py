"""The C/AVX2 port of the reference's CPU expert path (oracle/cpu_moe.c) against the numpy
restatement of the same functions, with the reference's own tolerances
(kernel/avx2.rs:2757,3050: integer-vs-FMA rel < 1 %; moe.rs:3423-3424: max|d| < 0.01)."""
import numpy as np

from oracle import cpu_ref, moe as omoe, router
from oracle.bf16 import round_bf16, f32_to_bf16_bits


def _setup(seed, E, H, I, k, M):
    rng = np.random.default_rng(seed)
    layer = omoe.make_int_layer(rng, E, H, I)
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids, w = router.route_from_logits(rng.normal(0, 1, (M, E)).astype(np.float32), k, norm_topk_prob=True)
    u = cpu_ref.to_unified(layer.w13_q, layer.w13_s) + cpu_ref.to_unified(layer.w2_q, layer.w2_s)
    return layer, x, ids, w, u


def test_c_port_matches_numpy_cpu_path():
    layer, x, ids, w, u = _setup(3, 8, 256, 128, 2, 16)
    ids[3, 1] = -1                                         # skipped expert (src/moe.rs:2722)
    want = omoe.moe_forward_cpu_int(layer, x, ids, w)
    got = cpu_ref.moe_forward_int4(*u, f32_to_bf16_bits(x), ids, w)
    assert np.abs(got - want).max() < 1e-3 * np.abs(want).max()      # ~20-bit fast exp in the AVX2 SiLU


def test_c_port_thread_count_invariant_bitwise():
    # kernel/avx2.rs:2629 "parallel == serial bit-exact"
    _, x, ids, w, u = _setup(4, 8, 512, 256, 4, 8)
    a = cpu_ref.moe_forward_int4(*u, f32_to_bf16_bits(x), ids, w, nthreads=1)
    b = cpu_ref.moe_forward_int4(*u, f32_to_bf16_bits(x), ids, w, nthreads=4)
    assert np.array_equal(a, b)


def test_c_port_behavioural_kats():
    _, x, ids, w, u = _setup(5, 8, 256, 128, 2, 6)
    x[:] = x[0]
    ids[:] = ids[0]
    w[:] = w[0]
    y = cpu_ref.moe_forward_int4(*u, f32_to_bf16_bits(x), ids, w)
    assert np.abs(y - y[0]).max() < 1e-6                              # moe.rs:3892-3900
    z = cpu_ref.moe_forward_int4(*u, f32_to_bf16_bits(x), np.full_like(ids, -1), w)
    assert not z.any()                                                # moe.rs:3903-3916
