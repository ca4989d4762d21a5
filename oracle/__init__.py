"""CPU oracle for the Krasis prefill hot path — TEST INFRASTRUCTURE ONLY.

This package is a plain numpy / C restatement of the reference's algorithm for
the path named in BASELINE.json:north_star (router -> top-k -> expert MLPs ->
weighted combine, plus the attention blocks that feed it).  Every function cites
the reference file:line it restates.

Rules (enforced by tests/test_boundary_cpu.py):
  * Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import or execute anything
    under ``oracle/``.  The product package ``krasis_b200`` never does; it fails
    loudly if its CUDA library is missing.
  * The oracle is the checker, never the thing measured or shipped.

Pinning status.  The reference's Rust core and its sglang / FlashInfer GPU kernels cannot run in this environment
(no Rust toolchain, sglang / sgl_kernel absent, no GPU in the build container), so ``oracle/_ref`` does
not exist.  Everything of the path that IS Python in the reference tree was executed on CPU and its outputs are committed
as fixtures the oracle is checked against (scripts under tests/golden/make_*.py):
  router, 4 scoring flavours        layer.py:compute_routing               router_reference.npz
  Gated DeltaNet chunked prefill    linear_attention.py (2 calls, state)   gdn_reference.npz
  MLA / GQA RoPE tables + rotation  attention.py                           mla_rope_reference.npz
  W8A8 quantise + int8_linear       weight_loader.py                       int8_reference.npz        (bit-exact)
  INT4 Marlin layout, both ways     triton_moe.py inverse_*                marlin_inverse_reference.npz (bit-exact)
  paged-KV bookkeeping              kv_cache.py                            kvcache_reference.json
plus every known-answer vector the reference's Rust unit tests hold for this path (reference_kats.json, from the cited
test bodies) and, for the GGUF block formats, the independent ``gguf`` python package.
Unpinned (third-party, absent from the tree): the sglang Marlin MoE GEMM and the FlashInfer attention cores — their
documented math is restated and the tolerance of each comparison is written in the test.
"""
