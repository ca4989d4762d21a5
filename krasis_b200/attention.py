"""Attention blocks behind the reference's call surface (python/krasis/linear_attention.py, attention.py).

  GatedDeltaNetAttention(cfg, layer_idx, weights, device).forward(hidden, is_decode) / reset_state()
      mirrors python/krasis/linear_attention.py:118-214,393-470 for the prefill path (M > 1, chunked).
  linear(x, weight)  — torch.nn.functional.linear for BF16 weights on tcgen05 (attention.py:526-529,672).
All arithmetic runs inside libkrasis_b200.so; torch supplies tensors and streams.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import capi


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _bf16_host(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        return t.detach().to(torch.bfloat16).cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    return np.ascontiguousarray(t, dtype=np.uint16)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16):
    """out = x @ weight.T (+ bias): x [M,K] bf16 cuda, weight [N,K] bf16 cuda."""
    if not (x.is_cuda and weight.is_cuda) or x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("linear: x and weight must be CUDA bf16 tensors (krasis_b200 has no CPU path)")
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1] or not x.is_contiguous() or not weight.is_contiguous():
        raise ValueError(f"linear: shape mismatch {tuple(x.shape)} x {tuple(weight.shape)}")
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    b = bias.float().contiguous() if bias is not None else None
    capi.check(capi.load().kb2_linear_bf16(x.data_ptr(), weight.data_ptr(), b.data_ptr() if b is not None else None,
                                           out.data_ptr(), M, N, K, int(out_dtype == torch.float32),
                                           x.device.index or 0, _stream(x.device)))
    return out


class GatedDeltaNetAttention:
    """Drop-in for python/krasis/linear_attention.py:GatedDeltaNetAttention on the prefill path.

    `cfg` needs: hidden_size, linear_num_key_heads, linear_num_value_heads, linear_key_head_dim,
    linear_value_head_dim, linear_conv_kernel_dim, rms_norm_eps (python/krasis/linear_attention.py:148-158).
    `weights` keys as in the reference (:167-176): in_proj_qkvz, in_proj_ba, out_proj, conv1d_weight, A_log,
    dt_bias, norm_weight — BF16 tensors (INT8 attention weights are disabled in the reference, config.py:209)."""

    _MAX_SHARED = 128      # layers one native handle (= one set of scratch buffers) can serve

    def __init__(self, cfg, layer_idx: int, weights: dict, device, max_tokens: int = 8192, share_scratch_with=None):
        self.cfg, self.layer_idx = cfg, layer_idx
        self.device = torch.device(device)
        self._lib = capi.load()
        for k, v in weights.items():
            if isinstance(v, tuple):
                raise ValueError(f"{k}: INT8 attention weights are not supported (the reference disables them, config.py:209)")
        if share_scratch_with is None:
            c = capi.GdnConfig(cfg.hidden_size, cfg.linear_num_key_heads, cfg.linear_num_value_heads,
                               cfg.linear_key_head_dim, cfg.linear_value_head_dim, cfg.linear_conv_kernel_dim,
                               float(cfg.rms_norm_eps), max_tokens, self._MAX_SHARED, self.device.index or 0)
            self._h = C.c_void_p()
            capi.check(self._lib.kb2_gdn_create(C.byref(c), C.byref(self._h)))
            self._owner, self._slot, self._n_slots = None, 0, [1]
            self._c = c
        else:                                   # same native handle: layers of one model share the scratch buffers
            o = share_scratch_with
            self._h, self._owner, self._c, self._n_slots = o._h, o, o._c, o._n_slots
            self._slot = self._n_slots[0]
            self._n_slots[0] += 1
            if self._slot >= self._MAX_SHARED:
                raise ValueError("too many layers share one GDN handle")
        arrs = [_bf16_host(weights[k]) for k in ("in_proj_qkvz", "in_proj_ba", "conv1d_weight", "A_log", "dt_bias",
                                                 "norm_weight", "out_proj")]
        capi.check(self._lib.kb2_gdn_set_weights_host(self._h, self._slot, *[a.ctypes.data for a in arrs]))

    def __del__(self):
        try:
            if self._owner is None and self._h:
                self._lib.kb2_gdn_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def reset_state(self):
        capi.check(self._lib.kb2_gdn_reset_state(self._h, self._slot, _stream(self.device)))

    def set_output_scatter(self, peer_ptrs=None, src_rank: int = 0) -> None:
        """Head-parallel attention under token sharding: send the out_proj rows straight to the token owners' receive buffers
        (Communicator.peer_alloc pointers, [rows][world][hidden] bf16) instead of returning them; None restores the local output.
        The setting lives in the handle shared by all layers of the model, so the caller sets it before every forward."""
        n = len(peer_ptrs) if peer_ptrs else 0
        arr = (C.c_void_p * max(n, 1))(*(peer_ptrs or [None]))
        capi.check(self._lib.kb2_gdn_set_output_scatter(self._h, arr, n, src_rank))

    def forward(self, hidden: torch.Tensor, is_decode: bool = False) -> torch.Tensor:
        if is_decode:
            raise NotImplementedError("M=1 recurrent decode is out of scope (SURVEY.md §8: prefill path only)")
        if not hidden.is_cuda or hidden.dtype != torch.bfloat16 or hidden.dim() != 2 or not hidden.is_contiguous() \
                or hidden.shape[1] != self._c.hidden_size:
            raise ValueError(f"hidden: expected contiguous CUDA bf16 [M, {self._c.hidden_size}]")
        out = torch.empty_like(hidden)
        capi.check(self._lib.kb2_gdn_forward(self._h, self._slot, hidden.data_ptr(), out.data_ptr(), hidden.shape[0],
                                             _stream(hidden.device)))
        return out

    def state(self):
        """(conv_state [C,K] bf16-as-float32, recurrent_state [nv,dk,dv] float32) on host — tests only."""
        c = self._c
        Cc = 2 * c.num_k_heads * c.k_head_dim + c.num_v_heads * c.v_head_dim
        conv = np.zeros((Cc, c.conv_kernel), np.uint16)
        rec = np.zeros((c.num_v_heads, c.k_head_dim, c.v_head_dim), np.float32)
        capi.check(self._lib.kb2_gdn_get_state_host(self._h, self._slot, conv.ctypes.data, rec.ctypes.data))
        return (conv.astype(np.uint32) << 16).view(np.float32), rec


PAGE_SIZE = 16   # python/krasis/kv_cache.py:26


class PagedKVCache:
    """GQA part of python/krasis/kv_cache.py:PagedKVCache (:31-186): FP8-E4M3 K/V pools
    [num_layers, max_pages, 16, num_kv_heads, head_dim] and a free-list page allocator."""

    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, device, max_pages: int,
                 kv_dtype=torch.float8_e4m3fn, page_size: int = PAGE_SIZE):
        if page_size != PAGE_SIZE or kv_dtype != torch.float8_e4m3fn:
            raise ValueError("only page_size=16 and float8_e4m3fn are supported (the reference defaults)")
        self.num_layers, self.num_kv_heads, self.gqa_head_dim = num_layers, num_kv_heads, head_dim
        self.page_size, self.kv_dtype, self.max_pages = page_size, kv_dtype, max_pages
        self.device = torch.device(device)
        shape = (num_layers, max_pages, page_size, num_kv_heads, head_dim)
        self.k_cache = torch.zeros(shape, dtype=kv_dtype, device=self.device)
        self.v_cache = torch.zeros(shape, dtype=kv_dtype, device=self.device)
        self._free = list(range(max_pages - 1, -1, -1))

    @classmethod
    def from_config(cls, cfg, num_layers: int, device, max_pages: Optional[int] = None, kv_dtype=torch.float8_e4m3fn,
                    page_size: int = PAGE_SIZE, max_mb: Optional[int] = None):
        """The reference constructor's signature and sizing rule (kv_cache.py:38-82): `max_pages` explicit, else
        max(64, max_mb MiB // bytes_per_page) with max_mb defaulting to 2000."""
        per_page = page_size * cfg.num_key_value_heads * cfg.gqa_head_dim * 2 * (1 if kv_dtype == torch.float8_e4m3fn else 2) * num_layers
        if max_pages is None:
            max_pages = max(64, (2000 if max_mb is None else max_mb) * 1024 * 1024 // per_page)
        return cls(num_layers, cfg.num_key_value_heads, cfg.gqa_head_dim, device, max_pages, kv_dtype, page_size)

    def get_gqa_layer_caches(self, layer_offset: int):
        return self.k_cache[layer_offset], self.v_cache[layer_offset]

    @property
    def max_context_tokens(self) -> int:
        return self.max_pages * self.page_size

    @property
    def free_page_count(self) -> int:
        return len(self._free)

    def alloc_page(self) -> int:
        if not self._free:
            raise RuntimeError("KV cache out of pages")
        return self._free.pop()

    def alloc_pages(self, n: int):
        if n > len(self._free):
            raise RuntimeError(f"KV cache exhausted: need {n} pages, have {len(self._free)}")      # kv_cache.py:149-156
        return [self._free.pop() for _ in range(n)]

    def free_pages(self, pages):
        self._free.extend(pages)


class SequenceKVState:
    """python/krasis/kv_cache.py:SequenceKVState (:189-272): page list + seq_len of one sequence."""

    def __init__(self, cache: PagedKVCache, seq_id: int = 0):
        self.cache, self.seq_id, self.pages, self.seq_len = cache, seq_id, [], 0
        self._idx = None

    def ensure_capacity(self, num_new_tokens: int):
        need = (self.seq_len + num_new_tokens + self.cache.page_size - 1) // self.cache.page_size
        while len(self.pages) < need:
            self.pages.append(self.cache.alloc_page())
            self._idx = None

    def advance(self, num_new_tokens: int):
        self.seq_len += num_new_tokens

    def kv_indices(self, device) -> torch.Tensor:
        if self._idx is None or self._idx.numel() != len(self.pages):
            self._idx = torch.tensor(self.pages, dtype=torch.int32, device=device)
        return self._idx

    def kv_indptr(self, device) -> torch.Tensor:
        return torch.tensor([0, len(self.pages)], dtype=torch.int32, device=device)

    def kv_len_arr(self, device) -> torch.Tensor:
        return torch.tensor([self.seq_len], dtype=torch.int32, device=device)

    def last_page_len(self) -> int:
        if self.seq_len == 0:                      # kv_cache.py:236-241
            return 0
        r = self.seq_len % self.cache.page_size
        return r if r else self.cache.page_size

    def last_page_len_tensor(self, device) -> torch.Tensor:
        return torch.tensor([self.last_page_len()], dtype=torch.int32, device=device)

    def free(self):
        self.cache.free_pages(self.pages)
        self.pages, self.seq_len, self._idx = [], 0, None


class GQAAttention:
    """Drop-in for python/krasis/attention.py:GQAAttention (prefill): forward(hidden, positions, kv_cache,
    seq_state, layer_offset, num_new_tokens) -> [M, hidden] bf16.  `cfg` needs hidden_size, num_attention_heads,
    num_key_value_heads, gqa_head_dim, rotary_dim, rope_theta, rms_norm_eps.  weights: q_proj, k_proj, v_proj, o_proj
    (+ optional q_norm, k_norm) BF16 tensors; gated attention is detected from q_proj's row count (:398-406)."""

    _MAX_SHARED = 128

    def __init__(self, cfg, layer_idx: int, weights: dict, device, max_tokens: int = 8192, share_scratch_with=None):
        self.cfg, self.layer_idx, self.device = cfg, layer_idx, torch.device(device)
        for k in ("q_proj_bias", "k_proj_bias", "v_proj_bias", "o_proj_bias", "sinks"):
            if weights.get(k) is not None:
                raise NotImplementedError(f"{k}: GLM / GPT-OSS attention variants are out of scope (SURVEY.md §8a)")
        nh, nkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.gqa_head_dim
        q_rows = weights["q_proj"].shape[0]
        self.gated_attention = q_rows == 2 * nh * d
        if not self.gated_attention and q_rows != nh * d:
            raise ValueError(f"q_proj has {q_rows} rows, expected {nh * d} or {2 * nh * d}")
        self._lib = capi.load()
        if share_scratch_with is None:
            c = capi.GqaConfig(cfg.hidden_size, nh, nkv, d, cfg.rotary_dim, int(self.gated_attention), float(cfg.rope_theta),
                               float(cfg.rms_norm_eps), PAGE_SIZE, max_tokens, self._MAX_SHARED, self.device.index or 0)
            self._h = C.c_void_p()
            capi.check(self._lib.kb2_gqa_create(C.byref(c), C.byref(self._h)))
            self._owner, self._slot, self._n_slots, self._c = None, 0, [1], c
        else:
            o = share_scratch_with
            if o.gated_attention != self.gated_attention:
                raise ValueError("layers sharing a GQA handle must agree on gated attention")
            self._h, self._owner, self._c, self._n_slots = o._h, o, o._c, o._n_slots
            self._slot = self._n_slots[0]
            self._n_slots[0] += 1
        arrs = [_bf16_host(weights[k]) for k in ("q_proj", "k_proj", "v_proj", "o_proj")]
        qn = _bf16_host(weights["q_norm"]) if weights.get("q_norm") is not None else None
        kn = _bf16_host(weights["k_norm"]) if weights.get("k_norm") is not None else None
        capi.check(self._lib.kb2_gqa_set_weights_host(self._h, self._slot, *[a.ctypes.data for a in arrs],
                                                      qn.ctypes.data if qn is not None else None,
                                                      kn.ctypes.data if kn is not None else None))

    def __del__(self):
        try:
            if self._owner is None and self._h:
                self._lib.kb2_gqa_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def set_output_scatter(self, peer_ptrs=None, src_rank: int = 0) -> None:
        """Head-parallel attention under token sharding: send the out_proj rows straight to the token owners' receive buffers
        (Communicator.peer_alloc pointers, [rows][world][hidden] bf16) instead of returning them; None restores the local output.
        The setting lives in the handle shared by all layers of the model, so the caller sets it before every forward."""
        n = len(peer_ptrs) if peer_ptrs else 0
        arr = (C.c_void_p * max(n, 1))(*(peer_ptrs or [None]))
        capi.check(self._lib.kb2_gqa_set_output_scatter(self._h, arr, n, src_rank))

    def forward(self, hidden: torch.Tensor, positions: torch.Tensor, kv_cache: PagedKVCache, seq_state: SequenceKVState,
                layer_offset: int, num_new_tokens: int = 0) -> torch.Tensor:
        if not hidden.is_cuda or hidden.dtype != torch.bfloat16 or hidden.dim() != 2 or not hidden.is_contiguous() \
                or hidden.shape[1] != self._c.hidden_size:
            raise ValueError(f"hidden: expected contiguous CUDA bf16 [M, {self._c.hidden_size}]")
        M = hidden.shape[0]
        if num_new_tokens not in (0, M):
            raise ValueError("num_new_tokens must equal the number of rows of hidden")
        seq_state.ensure_capacity(M)
        k_layer, v_layer = kv_cache.get_gqa_layer_caches(layer_offset)
        pos = positions.to(device=hidden.device, dtype=torch.int32).contiguous()
        out = torch.empty_like(hidden)
        capi.check(self._lib.kb2_gqa_forward(self._h, self._slot, hidden.data_ptr(), pos.data_ptr(), seq_state.seq_len,
                                             k_layer.data_ptr(), v_layer.data_ptr(),
                                             seq_state.kv_indices(hidden.device).data_ptr(), seq_state.seq_len + M,
                                             out.data_ptr(), M, _stream(hidden.device)))
        return out


# ------------------------------------------------------------------------------------------------ MLA

class MLAPagedKVCache(PagedKVCache):
    """MLA part of python/krasis/kv_cache.py:PagedKVCache (:99-117, split layout): FP8-E4M3 latent pools
    ckv [num_layers, max_pages, 16, kv_lora_rank] and kpe [num_layers, max_pages, 16, qk_rope_head_dim]."""

    def __init__(self, num_layers: int, kv_lora_rank: int, qk_rope_head_dim: int, device, max_pages: int,
                 kv_dtype=torch.float8_e4m3fn, page_size: int = PAGE_SIZE):
        if page_size != PAGE_SIZE or kv_dtype != torch.float8_e4m3fn:
            raise ValueError("only page_size=16 and float8_e4m3fn are supported (the reference defaults)")
        self.num_layers, self.ckv_dim, self.kpe_dim = num_layers, kv_lora_rank, qk_rope_head_dim
        self.page_size, self.kv_dtype, self.max_pages = page_size, kv_dtype, max_pages
        self.device = torch.device(device)
        self.ckv_cache = torch.zeros((num_layers, max_pages, page_size, kv_lora_rank), dtype=kv_dtype, device=self.device)
        self.kpe_cache = torch.zeros((num_layers, max_pages, page_size, qk_rope_head_dim), dtype=kv_dtype, device=self.device)
        self._free = list(range(max_pages - 1, -1, -1))

    def get_layer_caches(self, layer_offset: int):
        return self.ckv_cache[layer_offset], self.kpe_cache[layer_offset]


def mla_rope_inv_freq(qk_rope_dim: int, rope_theta: float, rope_scaling) -> "np.ndarray":
    """Inverse frequencies of python/krasis/attention.py:_get_rope_cos_sin (:119-157), YaRN blend included, in fp32
    with torch's own ops so the table the kernel multiplies positions with is the reference's, bit for bit."""
    import math
    dim = qk_rope_dim
    freqs = 1.0 / (rope_theta ** (torch.arange(0, dim, 2).float() / dim))
    cfg = rope_scaling or {}
    factor = cfg.get("factor", 1.0)
    if factor > 1.0:
        original_max = cfg.get("original_max_position_embeddings", 4096)
        beta_fast, beta_slow = cfg.get("beta_fast", 32.0), cfg.get("beta_slow", 1.0)
        low = math.floor(dim * math.log(original_max / (beta_fast * 2 * math.pi)) / (2 * math.log(rope_theta)))
        high = math.ceil(dim * math.log(original_max / (beta_slow * 2 * math.pi)) / (2 * math.log(rope_theta)))
        low, high = max(low, 0), min(high, dim // 2 - 1)
        ramp = torch.clamp((torch.arange(dim // 2).float() - low) / max(high - low, 0.001), 0, 1)
        mask = 1.0 - ramp
        freqs = (freqs / factor) * (1 - mask) + freqs * mask
    return freqs.float().numpy()


def mla_sm_scale(qk_nope_dim: int, qk_rope_dim: int, rope_scaling) -> float:
    """attention.py:79-88: 1/sqrt(head_dim) * yarn mscale(mscale_all_dim)^2."""
    import math
    s = 1.0 / math.sqrt(qk_nope_dim + qk_rope_dim)
    cfg = rope_scaling or {}
    factor = cfg.get("factor", 1.0)
    if factor > 1.0:
        m = 0.1 * cfg.get("mscale_all_dim", 0) * math.log(factor) + 1.0
        s *= m * m
    return s


class MLAAttention:
    """Drop-in for python/krasis/attention.py:MLAAttention (prefill): forward(hidden, positions, kv_cache, seq_state,
    layer_offset, num_new_tokens) -> [M, hidden] bf16.  `cfg` needs hidden_size, num_attention_heads, qk_nope_head_dim,
    qk_rope_head_dim, v_head_dim, kv_lora_rank, q_lora_rank (None/0 = direct q_proj), rope_theta, rope_scaling,
    rms_norm_eps.  weights as the reference (:90-107): q_proj | (q_a_proj, q_a_layernorm, q_b_proj), kv_a_proj_with_mqa,
    kv_a_layernorm, w_kc [H,nope,lora], w_vc [H,v,lora], o_proj — BF16 tensors."""

    _MAX_SHARED = 64

    def __init__(self, cfg, layer_idx: int, weights: dict, device, max_tokens: int = 8192, max_kv_len: Optional[int] = None,
                 share_scratch_with=None):
        self.cfg, self.layer_idx, self.device = cfg, layer_idx, torch.device(device)
        for k, v in weights.items():
            if isinstance(v, tuple):
                raise ValueError(f"{k}: INT8 attention weights are not supported (the reference disables them, config.py:209)")
        self.q_lora_rank = int(getattr(cfg, "q_lora_rank", 0) or 0)
        self.sm_scale = mla_sm_scale(cfg.qk_nope_head_dim, cfg.qk_rope_head_dim, getattr(cfg, "rope_scaling", None))
        self._lib = capi.load()
        if share_scratch_with is None:
            c = capi.MlaConfig(cfg.hidden_size, cfg.num_attention_heads, cfg.qk_nope_head_dim, cfg.qk_rope_head_dim,
                               cfg.v_head_dim, cfg.kv_lora_rank, self.q_lora_rank, float(cfg.rms_norm_eps), float(self.sm_scale),
                               PAGE_SIZE, max_tokens, max_kv_len or max_tokens, self._MAX_SHARED, self.device.index or 0)
            self._h = C.c_void_p()
            capi.check(self._lib.kb2_mla_create(C.byref(c), C.byref(self._h)))
            self._owner, self._slot, self._n_slots, self._c = None, 0, [1], c
        else:
            o = share_scratch_with
            self._h, self._owner, self._c, self._n_slots = o._h, o, o._c, o._n_slots
            self._slot = self._n_slots[0]
            self._n_slots[0] += 1
            if self._slot >= self._MAX_SHARED:
                raise ValueError("too many layers share one MLA handle")
        inv = np.ascontiguousarray(mla_rope_inv_freq(cfg.qk_rope_head_dim, float(cfg.rope_theta), getattr(cfg, "rope_scaling", None)),
                                   dtype=np.float32)
        h = lambda k: _bf16_host(weights[k])
        if self.q_lora_rank:
            q, qa, qan = h("q_b_proj"), h("q_a_proj"), h("q_a_layernorm")
        else:
            q, qa, qan = h("q_proj"), None, None
        rest = [h(k) for k in ("kv_a_proj_with_mqa", "kv_a_layernorm", "w_kc", "w_vc", "o_proj")]
        p = lambda a: a.ctypes.data if a is not None else None
        capi.check(self._lib.kb2_mla_set_weights_host(self._h, self._slot, p(q), p(qa), p(qan), *[p(a) for a in rest], p(inv)))

    def __del__(self):
        try:
            if self._owner is None and self._h:
                self._lib.kb2_mla_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def set_output_scatter(self, peer_ptrs=None, src_rank: int = 0) -> None:
        """Head-parallel attention under token sharding: send the out_proj rows straight to the token owners' receive buffers
        (Communicator.peer_alloc pointers, [rows][world][hidden] bf16) instead of returning them; None restores the local output.
        The setting lives in the handle shared by all layers of the model, so the caller sets it before every forward."""
        n = len(peer_ptrs) if peer_ptrs else 0
        arr = (C.c_void_p * max(n, 1))(*(peer_ptrs or [None]))
        capi.check(self._lib.kb2_mla_set_output_scatter(self._h, arr, n, src_rank))

    def forward(self, hidden: torch.Tensor, positions: torch.Tensor, kv_cache: MLAPagedKVCache, seq_state: SequenceKVState,
                layer_offset: int, num_new_tokens: int = 0) -> torch.Tensor:
        if not hidden.is_cuda or hidden.dtype != torch.bfloat16 or hidden.dim() != 2 or not hidden.is_contiguous() \
                or hidden.shape[1] != self._c.hidden_size:
            raise ValueError(f"hidden: expected contiguous CUDA bf16 [M, {self._c.hidden_size}]")
        M = hidden.shape[0]
        if num_new_tokens not in (0, M):
            raise ValueError("num_new_tokens must equal the number of rows of hidden")
        seq_state.ensure_capacity(M)
        ckv_layer, kpe_layer = kv_cache.get_layer_caches(layer_offset)
        pos = positions.to(device=hidden.device, dtype=torch.int32).contiguous()
        out = torch.empty_like(hidden)
        capi.check(self._lib.kb2_mla_forward(self._h, self._slot, hidden.data_ptr(), pos.data_ptr(), seq_state.seq_len,
                                             ckv_layer.data_ptr(), kpe_layer.data_ptr(),
                                             seq_state.kv_indices(hidden.device).data_ptr(), seq_state.seq_len + M,
                                             out.data_ptr(), M, _stream(hidden.device)))
        return out
