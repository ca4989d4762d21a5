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


def test_c_port_gguf_path_matches_numpy_restatement():
    """moe_forward_gguf (src/moe.rs:990-1110) over Q4_K gate/up + Q8_0 down (the Q4_K_M-style pairing when the down K is not a
    multiple of 256) and over Q4_K throughout: the C/AVX2 port against the numpy restatement of gguf_kernels.rs:271-426
    (float64 accumulation there, fp32 lanes here: 1e-4 relative), plus the reference's own known-answer vectors
    (gguf_kernels.rs:782-792: Q8_0 d = 0.1, q = 10, x = 1 -> ~32) and thread-count invariance."""
    from oracle import gguf_blocks as G
    rng = np.random.default_rng(11)
    E, H, I, k, M = 6, 512, 256, 3, 9
    for t2 in (G.GGML_Q8_0, G.GGML_Q4_K):
        layer = omoe.make_gguf_layer(rng, E, H, I, G.GGML_Q4_K, t2)
        x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
        ids = np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)
        ids[2, 1] = -1
        w = rng.dirichlet(np.ones(k), M).astype(np.float32)
        got = cpu_ref.moe_forward_gguf(layer.gate, layer.up, layer.down, G.GGML_Q4_K, t2, H, I, f32_to_bf16_bits(x), ids, w)
        want = np.zeros((M, H), np.float32)
        for m in range(M):
            for j in range(k):
                if ids[m, j] >= 0:
                    want[m] += w[m, j] * omoe.expert_forward_cpu_gguf(layer, int(ids[m, j]), x[m])
        assert np.abs(got - want).max() < 1e-4 * np.abs(want).max()
        one = cpu_ref.moe_forward_gguf(layer.gate, layer.up, layer.down, G.GGML_Q4_K, t2, H, I, f32_to_bf16_bits(x), ids, w, nthreads=1)
        assert np.array_equal(got, one)
    # known-answer vector of the reference's own unit test (gguf_kernels.rs:782-792), through the full expert path shape:
    # a Q8_0 "down" whose every block is d = fp16(0.1), q = 10 maps a hidden vector h to 0.1 * 10 * sum(h)
    Hk, Ik = 64, 32
    blk = np.zeros((1, Hk, 34), np.uint8)
    blk[..., 0:2] = np.frombuffer(np.float16(0.1).tobytes(), np.uint8)
    blk[..., 2:] = 10
    gate = np.zeros((1, Ik, Hk // 32 * 34), np.uint8)          # zero gate/up -> hidden = silu(0) * 0 = 0 -> output 0
    out = cpu_ref.moe_forward_gguf(gate, gate, blk, G.GGML_Q8_0, G.GGML_Q8_0, Hk, Ik, f32_to_bf16_bits(np.ones((1, Hk), np.float32)),
                                   np.zeros((1, 1), np.int32), np.ones((1, 1), np.float32))
    assert not out.any()
