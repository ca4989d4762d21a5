"""MoE router oracle (test infrastructure only).

Pinned: tests/golden/router_reference.npz holds outputs of the reference's own compute_routing executed on CPU
(tests/golden/make_router_golden.py); tests/test_oracle_kats.py checks this module against them.

Restates python/krasis/layer.py:526-560 (TransformerLayer.compute_routing; the same
code is inlined at layer.py:583-616) and the tie-break of the Rust twin
src/moe.rs:3116-3128 (strict '>' scan => the LOWER expert index wins a tie;
torch.topk's tie order is unspecified, SURVEY.md §7 "hard parts").

  logits = hidden.float() @ gate.float().T (+ gate_bias)             fp32 GEMM
  softmax : scores = softmax(logits); select top-k on (scores + e_score_correction_bias),
            weights = scores[ids]; optional renorm by their sum      (Qwen / DeepSeek-V2)
  sigmoid : same with sigmoid                                          (Kimi)
  gpt_oss : top-k on raw logits, softmax over the k selected          (swiglu_limit > 0)

hidden and gate are BF16, so every product is exact in fp32/fp64; only the
summation order differs between implementations.  The oracle accumulates in
float64 and rounds once to float32 — the "infinitely careful fp32 GEMM".
Bit-exact id parity is asserted by the tests on inputs whose k-th / (k+1)-th score
gap exceeds `min_topk_gap(...)`; nearer ties are reported, not hidden.
"""
import numpy as np


def router_logits(hidden_f32: np.ndarray, gate_f32: np.ndarray, gate_bias=None) -> np.ndarray:
    """layer.py:532-534.  hidden [M,H], gate [E,H] (values are BF16-representable)."""
    lg = hidden_f32.astype(np.float64) @ gate_f32.astype(np.float64).T
    lg = lg.astype(np.float32)
    if gate_bias is not None:
        lg = (lg + np.asarray(gate_bias, np.float32)[None, :]).astype(np.float32)
    return lg


def _topk_lower_index_first(sel: np.ndarray, k: int) -> np.ndarray:
    # stable sort on the negated score keeps the lower expert index first among equals
    return np.argsort(-sel, axis=1, kind="stable")[:, :k]


def route_from_logits(logits: np.ndarray, top_k: int, scoring_func: str = "softmax",
                      norm_topk_prob: bool = False, e_score_correction_bias=None,
                      gpt_oss: bool = False):
    """layer.py:536-560.  Returns (ids int32 [M,k] in descending score order, weights f32 [M,k])."""
    lg = np.asarray(logits, np.float32)
    if gpt_oss:
        ids = _topk_lower_index_first(lg, top_k)
        v = np.take_along_axis(lg, ids, axis=1)
        e = np.exp((v - v.max(axis=1, keepdims=True)).astype(np.float32)).astype(np.float32)
        w = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
        return ids.astype(np.int32), w
    if scoring_func == "sigmoid":
        scores = (np.float32(1.0) / (np.float32(1.0) + np.exp(-lg).astype(np.float32))).astype(np.float32)
    else:
        e = np.exp((lg - lg.max(axis=1, keepdims=True)).astype(np.float32)).astype(np.float32)
        scores = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    sel = scores
    if e_score_correction_bias is not None:
        sel = (scores + np.asarray(e_score_correction_bias, np.float32)[None, :]).astype(np.float32)
    ids = _topk_lower_index_first(sel, top_k)
    w = np.take_along_axis(scores, ids, axis=1).astype(np.float32)
    if norm_topk_prob:
        w = (w / w.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return ids.astype(np.int32), w


def compute_routing(hidden_f32, gate_f32, top_k, **kw):
    gate_bias = kw.pop("gate_bias", None)
    return route_from_logits(router_logits(hidden_f32, gate_f32, gate_bias), top_k, **kw)


def min_topk_gap(logits: np.ndarray, top_k: int, bias=None) -> np.ndarray:
    """Per row: the smallest gap between consecutive entries of the sorted selection key
    among the top k+1 (softmax/sigmoid are monotone, so gaps in logits are what
    decide the id set AND its order).  Rows with a tiny gap are near-ties."""
    key = np.asarray(logits, np.float64)
    if bias is not None:
        raise ValueError("with a correction bias compute the gap on the biased scores")
    s = -np.sort(-key, axis=1)[:, : top_k + 1]
    return np.min(s[:, :-1] - s[:, 1:], axis=1)
