"""Generates tests/golden/kvcache_reference.json by EXECUTING THE REFERENCE's own python/krasis/kv_cache.py on CPU:
PagedKVCache sizing + free list and SequenceKVState page bookkeeping (ensure_capacity / advance / free / kv_indices /
kv_indptr / last_page_len) over a scripted trace.  krasis.config is stubbed (only ModelConfig's name is imported).
Run:  python tests/golden/make_kvcache_golden.py      (build container only: needs /root/reference)
"""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python/krasis"

TRACE = [("new", 0), ("ensure", 0, 5), ("advance", 0, 5), ("ensure", 0, 40), ("advance", 0, 40), ("new", 1), ("ensure", 1, 16),
         ("advance", 1, 16), ("ensure", 0, 3), ("advance", 0, 3), ("free", 0), ("ensure", 1, 100), ("advance", 1, 100),
         ("new", 2), ("ensure", 2, 1), ("advance", 2, 1), ("free", 1), ("ensure", 2, 31), ("advance", 2, 31)]


def main():
    pkg = types.ModuleType("krasis")
    pkg.__path__ = [REF]
    sys.modules["krasis"] = pkg
    cfgm = types.ModuleType("krasis.config")
    cfgm.ModelConfig = object
    sys.modules["krasis.config"] = cfgm
    import krasis.kv_cache as kv
    cfg = types.SimpleNamespace(attention_type="gqa", is_mla=False, is_gqa=True, num_key_value_heads=2, gqa_head_dim=256)
    sized = kv.PagedKVCache(cfg, 12, torch.device("cpu"), max_mb=3)             # sizing rule
    cache = kv.PagedKVCache(cfg, 1, torch.device("cpu"), max_pages=24)
    seqs, states = {}, []
    for op in TRACE:
        if op[0] == "new":
            seqs[op[1]] = kv.SequenceKVState(cache, op[1])
        elif op[0] == "ensure":
            seqs[op[1]].ensure_capacity(op[2])
        elif op[0] == "advance":
            seqs[op[1]].advance(op[2])
        else:
            seqs[op[1]].free()
        states.append({str(i): dict(pages=list(s.pages), seq_len=s.seq_len, last_page_len=s.last_page_len(),
                                    kv_indices=s.kv_indices(torch.device("cpu")).tolist(),
                                    kv_indptr=s.kv_indptr(torch.device("cpu")).tolist()) for i, s in seqs.items()}
                      | {"free": cache.free_page_count})
    json.dump(dict(trace=TRACE, states=states, sized_pages=sized.max_pages, sized_tokens=sized.max_context_tokens,
                   k_cache_shape=list(cache.k_cache.shape), k_cache_dtype=str(cache.k_cache.dtype)),
              open(os.path.join(HERE, "kvcache_reference.json"), "w"), indent=0)
    print("wrote kvcache_reference.json", sized.max_pages)


if __name__ == "__main__":
    main()
