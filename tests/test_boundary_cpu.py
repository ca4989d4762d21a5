"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/krasis_b200.h declares, argument validation mirrors the reference's PyValueError /
PyRuntimeError split, and the product never touches oracle/ or a CPU fallback."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "krasis_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kb2_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from krasis_b200 import capi
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/krasis_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == syms, "capi.SIGNATURES and the header disagree"


def test_config_struct_matches_header(lib):
    from krasis_b200 import capi
    assert C.sizeof(capi.Config) == 14 * 4
    assert lib.kb2_version().decode().startswith("krasis_b200")


def test_create_validates_shapes_like_the_reference(lib):
    from krasis_b200 import capi
    h = C.c_void_p()
    bad = capi.Config(2000, 512, 64, 6, 1, 0, 0, 1, 0, 0, 1.0, 16, 0, -1)    # H not a multiple of 256
    assert lib.kb2_create(C.byref(bad), C.byref(h)) == capi.KB2_ERR_VALUE
    assert b"hidden_size" in lib.kb2_last_error()
    bad = capi.Config(2048, 512, 64, 6, 1, 8, 0, 1, 0, 0, 1.0, 16, 0, -1)    # unknown weight format (0..7 are defined)
    assert lib.kb2_create(C.byref(bad), C.byref(h)) == capi.KB2_ERR_VALUE
    bad = capi.Config(2048, 512, 64, 6, 1, 0, 3, 2, 0, 0, 1.0, 16, 0, -1)    # rank >= num_ranks
    assert lib.kb2_create(C.byref(bad), C.byref(h)) == capi.KB2_ERR_VALUE


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from krasis_b200 import capi, KrasisEngine
    h = C.c_void_p()
    ok = capi.Config(2048, 512, 64, 6, 1, 0, 0, 1, 0, 0, 1.0, 16, 0, -1)
    assert lib.kb2_create(C.byref(ok), C.byref(h)) == capi.KB2_ERR_CUDA
    assert b"no CPU fallback" in lib.kb2_last_error()
    with pytest.raises(capi.Kb2Error):
        KrasisEngine(hidden_size=2048, moe_intermediate_size=512, n_routed_experts=64, num_experts_per_tok=6,
                     num_moe_layers=1)


def test_python_surface_validates_before_touching_cuda():
    from krasis_b200 import KrasisEngine
    with pytest.raises(ValueError):
        KrasisEngine(hidden_size=2048, moe_intermediate_size=512, n_routed_experts=64, num_experts_per_tok=6,
                     num_moe_layers=1, num_bits=3)
    with pytest.raises(ValueError):
        KrasisEngine(hidden_size=2048, moe_intermediate_size=512, n_routed_experts=64, num_experts_per_tok=6,
                     num_moe_layers=1, scoring_func="tanh")


def test_product_never_imports_oracle_or_falls_back():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "krasis_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"
                assert "/root/reference" not in src, f"{f} reads the reference tree"
    eng = open(os.path.join(pkg, "engine.py")).read()
    for banned in ("torch.matmul", "torch.topk", "torch.softmax", "F.linear", "torch.mm", "@ "):
        assert banned not in eng, f"engine.py contains {banned!r}: arithmetic belongs in the CUDA library"


def test_paged_kv_bookkeeping_matches_reference_execution():
    """PagedKVCache / SequenceKVState host logic vs a trace produced by executing the reference's kv_cache.py on CPU
    (tests/golden/make_kvcache_golden.py): sizing rule, pool shape/dtype, free-list order, page lists, kv_indices/indptr,
    last_page_len after every step."""
    import json
    import types
    import torch
    from krasis_b200.attention import PagedKVCache, SequenceKVState
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kvcache_reference.json")))
    cfg = types.SimpleNamespace(num_key_value_heads=2, gqa_head_dim=256)
    sized = PagedKVCache.from_config(cfg, 12, "cpu", max_mb=3)
    assert (sized.max_pages, sized.max_context_tokens) == (G["sized_pages"], G["sized_tokens"])
    cache = PagedKVCache.from_config(cfg, 1, "cpu", max_pages=24)
    assert list(cache.k_cache.shape) == G["k_cache_shape"] and str(cache.k_cache.dtype) == G["k_cache_dtype"]
    seqs = {}
    for op, want in zip(G["trace"], G["states"]):
        if op[0] == "new":
            seqs[op[1]] = SequenceKVState(cache, op[1])
        elif op[0] == "ensure":
            seqs[op[1]].ensure_capacity(op[2])
        elif op[0] == "advance":
            seqs[op[1]].advance(op[2])
        else:
            seqs[op[1]].free()
        assert cache.free_page_count == want["free"]
        for i, s in seqs.items():
            w = want[str(i)]
            assert s.pages == w["pages"] and s.seq_len == w["seq_len"] and s.last_page_len() == w["last_page_len"]
            assert s.kv_indices("cpu").tolist() == w["kv_indices"] and s.kv_indptr("cpu").tolist() == w["kv_indptr"]
    with pytest.raises(RuntimeError):
        cache.alloc_pages(10 ** 6)


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/krasis_b200.h must be consumable by a C99 compiler (cgo / Rust bindgen / ctypes users) and by C++."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "krasis_b200.h"\nint main(void){ kb2_config c; kb2_mla_config m; kb2_gqa_config g; (void)c; (void)m; (void)g; return 0; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", f"-I{inc}", str(src)], check=True)
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{inc}", "-x", "c++", str(src)], check=True)
