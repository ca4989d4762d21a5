"""krasis_b200 — B200-native (sm_100a) MoE prefill path behind the Krasis engine surface.

Public surface mirrors the reference (python/krasis/__init__.py:9-13 exports KrasisEngine; the prefill
entry is GpuPrefillManager.forward, python/krasis/gpu_prefill.py:4374):
    KrasisEngine        weight store + hand-off (host side of the C ABI)
    GpuPrefillManager   forward(moe_layer_idx, hidden_states, topk_ids, topk_weights, routed_only)
    compute_routing     TransformerLayer.compute_routing
There is no CPU / eager fallback anywhere in this package.
"""
from .engine import KrasisEngine, GpuPrefillManager, QuantizedExperts  # noqa: F401

__all__ = ["KrasisEngine", "GpuPrefillManager", "QuantizedExperts"]
__version__ = "0.1.0"
