"""Host-side mirror of the reference's engine surface for the prefill path.

Reference interfaces mirrored (names, argument meaning and error behaviour kept):
  KrasisEngine                      src/moe.rs:1377-3296 (#[pyclass]); introspection methods :1874-1965,
                                    set_routing_config/weights :2959,2989
  GpuPrefillManager.__init__        python/krasis/gpu_prefill.py:326-346 (rank/num_ranks expert slicing :353-359)
  GpuPrefillManager.forward         python/krasis/gpu_prefill.py:4374-4484
  TransformerLayer.compute_routing  python/krasis/layer.py:526-560
torch is used only for device memory, streams and dtype plumbing; all arithmetic happens inside
libkrasis_b200.so (hand-written sm_100a kernels).  No fallback path exists.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import capi

_SCORING = {"softmax": capi.SCORE_SOFTMAX, "sigmoid": capi.SCORE_SIGMOID, "topk_softmax": capi.SCORE_TOPK_SOFTMAX}


@dataclass
class QuantizedExperts:
    """One MoE layer's LOCAL experts in the reference quantiser's output format
    (src/weights/marlin.rs:145-207 / 65-114), numpy host arrays:
      w13_q [E, 2I, H/8] uint32 (INT4) or [E, 2I, H] int8 (INT8);  w13_s [E, 2I, H/128] uint16 (raw BF16)
      w2_q  [E, H, I/8]  uint32          or [E, H, I]  int8;        w2_s  [E, H, I/128]  uint16
    w13 = [gate rows ; up rows] (src/weights/mod.rs:346-349)."""
    w13_q: np.ndarray
    w13_s: np.ndarray
    w2_q: np.ndarray
    w2_s: np.ndarray


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class KrasisEngine:
    """Weight store + routing config behind the C ABI (the B200 counterpart of `#[pyclass] KrasisEngine`).

    The reference constructor is KrasisEngine(parallel, num_threads, skip_shared_experts) followed by
    load(model_dir, ...) (src/moe.rs:1482,1538).  Here the geometry is explicit (the C ABI takes it in
    kb2_config) and weights arrive either as the quantiser's arrays (`load_quantized_layer`) or as
    device-resident tiles (`attach_tiled_layer`)."""

    def __init__(self, parallel: bool = True, num_threads: Optional[int] = None, skip_shared_experts: bool = False, *,
                 hidden_size: Optional[int] = None, moe_intermediate_size: int = 0, n_routed_experts: int = 0,
                 num_experts_per_tok: int = 0, num_moe_layers: int = 0, num_bits: int = 4, group_size: int = 128,
                 gguf_gate_up_type: Optional[str] = None, gguf_down_type: Optional[str] = None,
                 rank: int = 0, num_ranks: int = 1, scoring_func: str = "softmax", norm_topk_prob: bool = False,
                 routed_scaling_factor: float = 1.0, max_tokens: int = 8192, device: int = 0):
        """Two ways in, both kept:
          * the reference's: `KrasisEngine(parallel=True, num_threads=None, skip_shared_experts=False)` then
            `load(model_dir, ...)` (src/moe.rs:1482,1538) — geometry comes from config.json;
          * explicit geometry keywords (what kb2_config takes), weights through load_quantized_layer / load_bf16_layer /
            load_gguf_layer / attach_tiled_layer.
        `parallel` / `num_threads` size the reference's rayon pool (moe.rs:1497-1500); the GPU engine has no host worker
        threads, so they are accepted and recorded only."""
        self._parallel, self._num_threads, self._skip_shared = bool(parallel), num_threads, bool(skip_shared_experts)
        self._lib = capi.load()
        self._h = C.c_void_p()
        self._keep = {}          # device tensors attached by the caller (kept alive)
        self._export_cache = None
        self._model_dir = None
        if hidden_size is not None:
            self._create(hidden_size, moe_intermediate_size, n_routed_experts, num_experts_per_tok, num_moe_layers, num_bits,
                         group_size, gguf_gate_up_type, gguf_down_type, rank, num_ranks, scoring_func, norm_topk_prob,
                         routed_scaling_factor, max_tokens, device)

    def _create(self, hidden_size, moe_intermediate_size, n_routed_experts, num_experts_per_tok, num_moe_layers, num_bits,
                group_size, gguf_gate_up_type, gguf_down_type, rank, num_ranks, scoring_func, norm_topk_prob,
                routed_scaling_factor, max_tokens, device):
        if num_bits not in (4, 8):
            raise ValueError(f"num_bits must be 4 or 8, got {num_bits}")           # gpu_prefill.py:259
        if group_size != 128:
            raise ValueError("only group_size=128 is supported (src/weights/marlin.rs:12)")
        if scoring_func not in _SCORING:
            raise ValueError(f"unknown scoring_func {scoring_func!r}")
        gg = capi.GGUF_FORMATS
        if gguf_gate_up_type is not None:        # gguf_native=True in KrasisEngine.load (src/moe.rs:1538): keep GGUF blocks
            if gguf_gate_up_type not in gg or (gguf_down_type or gguf_gate_up_type) not in gg:
                raise ValueError(f"GGUF expert types supported on the GPU path: {sorted(gg)}")
            f13, f2 = gg[gguf_gate_up_type], gg[gguf_down_type or gguf_gate_up_type]
        else:
            f13, f2 = (capi.FMT_INT4_G128 if num_bits == 4 else capi.FMT_INT8_G128), -1
        self._gguf = gguf_gate_up_type is not None
        self._cfg = capi.Config(hidden_size, moe_intermediate_size, n_routed_experts, num_experts_per_tok,
                                num_moe_layers, f13, rank, num_ranks, _SCORING[scoring_func], int(bool(norm_topk_prob)),
                                float(routed_scaling_factor), max_tokens, device, f2)
        self._h = C.c_void_p()
        capi.check(self._lib.kb2_create(C.byref(self._cfg), C.byref(self._h)))
        self._num_bits, self._group_size = num_bits, group_size
        self.device = torch.device("cuda", device)
        s, t = C.c_int32(), C.c_int32()
        capi.check(self._lib.kb2_expert_range(self._h, C.byref(s), C.byref(t)))
        self.expert_start, self.expert_end = s.value, t.value

    def _loaded(self):
        if not self._h:
            raise RuntimeError("Model not loaded")                                  # PyRuntimeError, src/moe.rs:1790-1803

    def load(self, model_dir: str, group_size: Optional[int] = None, max_layers: Optional[int] = None,
             start_layer: Optional[int] = None, num_bits: Optional[int] = None, cpu_num_bits: Optional[int] = None,
             gpu_num_bits: Optional[int] = None, gguf_path: Optional[str] = None, gguf_native: bool = False, *,
             device: int = 0, max_tokens: int = 8192, rank: int = 0, num_ranks: int = 1, use_cache: bool = True,
             cache_dir: Optional[str] = None):
        """KrasisEngine.load (src/moe.rs:1538; python/krasis/model.py:1428-1440): config.json -> geometry, then the routed
        experts of MoE layers [start_layer, start_layer + max_layers) — from the KB2 tile cache if a valid one exists
        (`~/.krasis/cache/<model>/experts_kb2_int{b}_g128.bin`, header / hash / size checked like mod.rs:2382-2423), else from
        the BF16 safetensors through the device quantiser (and the cache is then written, single rank only), or, with
        gguf_path + gguf_native, native GGUF expert blocks.  cpu_num_bits is accepted for signature parity (no CPU experts)."""
        import json
        import os
        from . import loader, tile_cache
        from .model import HybridMoEConfig
        if self._h:
            raise RuntimeError("engine is already loaded")
        raw_bytes = open(os.path.join(model_dir, "config.json"), "rb").read()
        cfg = HybridMoEConfig.from_hf_config(json.loads(raw_bytes))
        bits = gpu_num_bits or num_bits or 4
        s0 = start_layer or 0
        n_layers = cfg.num_moe_layers - s0
        if max_layers is not None:
            n_layers = min(n_layers, max_layers)
        if n_layers < 1:
            raise ValueError("no MoE layers in the requested range")
        gg = {}
        g = None
        if gguf_path is not None and gguf_native:
            g = loader.GgufFile(gguf_path)
            _, _, _, t13, t2 = loader.gguf_expert_blocks(g, cfg.first_k_dense_replace + s0, 0, 1, cfg.hidden_size, cfg.moe_intermediate_size)
            gg = dict(gguf_gate_up_type=t13, gguf_down_type=t2)
        self._create(cfg.hidden_size, cfg.moe_intermediate_size, cfg.n_routed_experts, cfg.num_experts_per_tok, n_layers, bits,
                     group_size or 128, gg.get("gguf_gate_up_type"), gg.get("gguf_down_type"), rank, num_ranks, cfg.scoring_func,
                     cfg.norm_topk_prob, cfg.routed_scaling_factor, max_tokens, device)
        self._model_dir = model_dir
        first = cfg.first_k_dense_replace + s0
        if g is not None:
            for m in range(n_layers):
                gate, up, down, _, _ = loader.gguf_expert_blocks(g, first + m, self.expert_start, self.expert_end, cfg.hidden_size,
                                                                 cfg.moe_intermediate_size)
                self.load_gguf_layer(m, gate, up, down)
            return
        cdir = cache_dir or os.path.join(os.path.expanduser("~"), ".krasis", "cache", os.path.basename(os.path.normpath(model_dir)))
        cpath = os.path.join(cdir, tile_cache.cache_file_name(bits))
        whole = s0 == 0 and n_layers == cfg.num_moe_layers              # the cache always describes the whole model
        if use_cache and whole and os.path.exists(cpath):
            try:
                tile_cache.load_tile_cache(self, cpath, raw_bytes)
                return
            except ValueError:
                pass                                                    # stale / foreign cache: rebuild (mod.rs:2417-2423)
        loader.load_experts_from_safetensors(self, model_dir, first_k_dense=first, num_bits=bits)
        if use_cache and whole and num_ranks == 1:
            os.makedirs(cdir, exist_ok=True)
            tile_cache.write_tile_cache(self, cpath, raw_bytes, n_shared_experts=cfg.n_shared_experts)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.kb2_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection: all METHODS, as in the reference (src/moe.rs:1874-1965)
    def num_moe_layers(self): self._loaded(); return self._cfg.num_moe_layers
    def hidden_size(self): self._loaded(); return self._cfg.hidden_size
    def intermediate_size(self): self._loaded(); return self._cfg.moe_intermediate_size
    def num_experts(self): self._loaded(); return self._cfg.n_routed_experts
    def top_k(self): self._loaded(); return self._cfg.num_experts_per_tok
    def group_size(self): self._loaded(); return self._group_size
    def gpu_num_bits(self): self._loaded(); return self._num_bits
    def cpu_num_bits(self): self._loaded(); return self._num_bits
    def is_parallel(self): return self._parallel
    def has_unified(self): return False            # no CPU-format experts are kept (src/moe.rs:1930)
    def has_gguf(self): return bool(self._h) and self._gguf
    def is_marlin_format(self): return True       # gpu_prefill.py:851 reads this without calling it: keep truthy
    def marlin_w2_padded_n(self): return self._cfg.hidden_size
    def launch_count(self): return int(self._lib.kb2_launch_count(self._h))

    def tiled_bytes(self, which: int) -> int:
        return int(self._lib.kb2_tiled_bytes(self._h, which))

    # ---- weights
    def load_quantized_layer(self, moe_layer_idx: int, q: QuantizedExperts):
        E = self.expert_end - self.expert_start
        H, I = self._cfg.hidden_size, self._cfg.moe_intermediate_size
        shapes = ({4: (E, 2 * I, H // 8), 8: (E, 2 * I, H)}[self._num_bits], (E, 2 * I, H // 128),
                  {4: (E, H, I // 8), 8: (E, H, I)}[self._num_bits], (E, H, I // 128))
        arrs = []
        for name, a, shp in zip(("w13_q", "w13_s", "w2_q", "w2_s"), (q.w13_q, q.w13_s, q.w2_q, q.w2_s), shapes):
            a = np.ascontiguousarray(a)
            if tuple(a.shape) != shp:
                raise ValueError(f"{name}: expected shape {shp}, got {tuple(a.shape)}")   # moe.rs:2285-2300
            arrs.append(a)
        capi.check(self._lib.kb2_load_experts_host(self._h, moe_layer_idx, *[a.ctypes.data for a in arrs]))

    def load_quantized_layer_dev(self, moe_layer_idx: int, w13_q: torch.Tensor, w13_s: torch.Tensor,
                                 w2_q: torch.Tensor, w2_s: torch.Tensor):
        """`load_quantized_layer` for arrays that already live on the device (same reference-quantiser layout and
        shapes; int32 words for INT4, int8 for INT8, int16 raw-BF16 scales): re-tiled on the device, no host copy."""
        E = self.expert_end - self.expert_start
        H, I = self._cfg.hidden_size, self._cfg.moe_intermediate_size
        shapes = ({4: (E, 2 * I, H // 8), 8: (E, 2 * I, H)}[self._num_bits], (E, 2 * I, H // 128),
                  {4: (E, H, I // 8), 8: (E, H, I)}[self._num_bits], (E, H, I // 128))
        qdt = torch.int32 if self._num_bits == 4 else torch.int8
        for name, t, shp, dt in zip(("w13_q", "w13_s", "w2_q", "w2_s"), (w13_q, w13_s, w2_q, w2_s), shapes,
                                    (qdt, torch.int16, qdt, torch.int16)):
            if tuple(t.shape) != shp or t.dtype != dt or not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{name}: expected contiguous CUDA {dt} {shp}, got {t.dtype} {tuple(t.shape)}")
        capi.check(self._lib.kb2_load_experts_dev(self._h, moe_layer_idx, w13_q.data_ptr(), w13_s.data_ptr(),
                                                  w2_q.data_ptr(), w2_s.data_ptr(), _stream_ptr(self.device)))
        torch.cuda.synchronize(self.device)

    def quantize_group(self, w_bf16: torch.Tensor, num_bits: Optional[int] = None):
        """Krasis symmetric g128 quantiser on the device (src/weights/marlin.rs:145-207 / :65-114), bit-exact.
        w [..., K] bf16 -> (q [..., K/8] int32 | [..., K] int8, scales [..., K/128] raw-bf16 int16)."""
        bits = num_bits or self._num_bits
        if w_bf16.dtype != torch.bfloat16 or not w_bf16.is_cuda or not w_bf16.is_contiguous():
            raise ValueError("quantize_group: expected a contiguous CUDA bf16 tensor")
        K = w_bf16.shape[-1]
        rows = w_bf16.numel() // K
        lead = tuple(w_bf16.shape[:-1])
        q = (torch.empty(lead + (K // 8,), dtype=torch.int32, device=w_bf16.device) if bits == 4
             else torch.empty(lead + (K,), dtype=torch.int8, device=w_bf16.device))
        sc = torch.empty(lead + (K // 128,), dtype=torch.int16, device=w_bf16.device)
        capi.check(self._lib.kb2_quantize_group_dev(w_bf16.data_ptr(), bits, q.data_ptr(), sc.data_ptr(), rows, K,
                                                    w_bf16.device.index or 0, _stream_ptr(w_bf16.device)))
        return q, sc

    def load_bf16_layer(self, moe_layer_idx: int, w13_bf16: torch.Tensor, w2_bf16: torch.Tensor):
        """HF BF16 experts ([E_local,2I,H] gate rows first, [E_local,H,I]) -> quantise on device -> B200 tiles
        (the load_from_hf step of src/weights/mod.rs:5049-5091 without the Marlin repack)."""
        E = self.expert_end - self.expert_start
        H, I = self._cfg.hidden_size, self._cfg.moe_intermediate_size
        if tuple(w13_bf16.shape) != (E, 2 * I, H) or tuple(w2_bf16.shape) != (E, H, I):
            raise ValueError(f"load_bf16_layer: expected [{E},{2*I},{H}] and [{E},{H},{I}]")
        q13, s13 = self.quantize_group(w13_bf16.to(self.device).contiguous())
        q2, s2 = self.quantize_group(w2_bf16.to(self.device).contiguous())
        capi.check(self._lib.kb2_load_experts_dev(self._h, moe_layer_idx, q13.data_ptr(), s13.data_ptr(), q2.data_ptr(),
                                                  s2.data_ptr(), _stream_ptr(self.device)))
        torch.cuda.synchronize(self.device)

    def load_gguf_layer(self, moe_layer_idx: int, gate: np.ndarray, up: np.ndarray, down: np.ndarray):
        """Native GGUF expert blocks of the LOCAL experts: gate/up uint8 [E, I, row_bytes(H)], down uint8 [E, H, row_bytes(I)]
        (blk.{L}.ffn_{gate,up,down}_exps.weight sliced per expert, src/weights/mod.rs:3439-3488)."""
        if not self._gguf:
            raise capi.Kb2Error("engine was not created with gguf_gate_up_type")
        E = self.expert_end - self.expert_start
        H, I = self._cfg.hidden_size, self._cfg.moe_intermediate_size
        rb = capi.GGUF_ROW_BYTES
        f13 = self._cfg.weight_format
        f2 = self._cfg.w2_weight_format if self._cfg.w2_weight_format >= 0 else f13
        want = ((E, I, rb[f13](H)), (E, I, rb[f13](H)), (E, H, rb[f2](I)))
        arrs = []
        for name, a, shp in zip(("gate", "up", "down"), (gate, up, down), want):
            a = np.ascontiguousarray(a, dtype=np.uint8)
            if tuple(a.shape) != shp:
                raise ValueError(f"{name}: expected uint8 {shp}, got {tuple(a.shape)}")
            arrs.append(a)
        capi.check(self._lib.kb2_load_experts_gguf_host(self._h, moe_layer_idx, *[a.ctypes.data for a in arrs]))

    def attach_tiled_layer(self, moe_layer_idx: int, w13_q: torch.Tensor, w13_s: torch.Tensor,
                           w2_q: torch.Tensor, w2_s: torch.Tensor):
        ts = (w13_q, w13_s, w2_q, w2_s)
        for i, t in enumerate(ts):
            if not t.is_cuda or not t.is_contiguous():
                raise ValueError("tiled weights must be contiguous CUDA tensors")
            if t.numel() * t.element_size() != self.tiled_bytes(i):
                raise ValueError(f"tiled buffer {i}: expected {self.tiled_bytes(i)} bytes, got {t.numel() * t.element_size()}")
        capi.check(self._lib.kb2_attach_experts_tiled_dev(self._h, moe_layer_idx, *[t.data_ptr() for t in ts]))
        self._keep[moe_layer_idx] = ts

    # ---- GPU-weight hand-off in the reference's own (Marlin) byte order — src/moe.rs:1972-2097,2431
    def _marlin_layer(self, moe_layer_idx: int):
        """The four Marlin-order arrays of every LOCAL expert of one layer, rebuilt from the device tiles: export -> host
        untile (krasis_b200/tiles.py) -> marlin_repack (krasis_b200/marlin_cache.py).  Cached per layer: the reference's
        callers ask for the four buffers one after the other (gpu_prefill.py:1059-1062)."""
        self._loaded()
        if self._gguf:
            raise RuntimeError("GPU weights not available: this engine holds native GGUF blocks")       # moe.rs:1978
        if self._export_cache is not None and self._export_cache[0] == moe_layer_idx:
            return self._export_cache[1]
        from . import marlin_cache as mc, tiles
        E = self.expert_end - self.expert_start
        H, I = self._cfg.hidden_size, self._cfg.moe_intermediate_size
        bufs = []
        for which in range(4):
            a = np.empty(self.tiled_bytes(which), np.uint8)
            capi.check(self._lib.kb2_export_experts_tiled_host(self._h, moe_layer_idx, which, a.ctypes.data, a.size))
            bufs.append(a)
        un = tiles.untile_int4 if self._num_bits == 4 else tiles.untile_int8
        fw = mc.rowmajor_to_marlin_int4 if self._num_bits == 4 else mc.rowmajor_to_marlin_int8
        q13, s13 = un(bufs[0], bufs[1], E, 2 * I, H)
        q2, s2 = un(bufs[2], bufs[3], E, H, I)
        out = fw(q13, s13, self._group_size) + fw(q2, s2, self._group_size)
        self._export_cache = (moe_layer_idx, out)
        return out

    def _marlin_slice(self, moe_layer_idx: int, which: int, start, end) -> np.ndarray:
        self._loaded()
        s = self.expert_start if start is None else start
        e = self.expert_end if end is None else end
        if not (self.expert_start <= s <= e <= self.expert_end):
            raise ValueError(f"expert range [{s}, {e}) outside this engine's experts [{self.expert_start}, {self.expert_end})")
        return np.ascontiguousarray(self._marlin_layer(moe_layer_idx)[which][s - self.expert_start:e - self.expert_start])

    def get_expert_w13_packed(self, moe_layer_idx: int, start=None, end=None) -> bytes:
        return self._marlin_slice(moe_layer_idx, 0, start, end).tobytes()

    def get_expert_w13_scales(self, moe_layer_idx: int, start=None, end=None) -> bytes:
        return self._marlin_slice(moe_layer_idx, 1, start, end).tobytes()

    def get_expert_w2_packed(self, moe_layer_idx: int, start=None, end=None) -> bytes:
        return self._marlin_slice(moe_layer_idx, 2, start, end).tobytes()

    def get_expert_w2_scales(self, moe_layer_idx: int, start=None, end=None) -> bytes:
        return self._marlin_slice(moe_layer_idx, 3, start, end).tobytes()

    def write_experts_range_into_pinned(self, moe_layer_idx: int, start: int, end: int, w13p_ptr: int, w13p_len: int,
                                        w13s_ptr: int, w13s_len: int, w2p_ptr: int, w2p_len: int, w2s_ptr: int, w2s_len: int):
        """src/moe.rs:2431: raw (address, length) pairs of caller-owned (pinned) buffers; lengths must match exactly
        (PyValueError otherwise, :2285-2300)."""
        for which, (ptr, ln) in enumerate(((w13p_ptr, w13p_len), (w13s_ptr, w13s_len), (w2p_ptr, w2p_len), (w2s_ptr, w2s_len))):
            a = self._marlin_slice(moe_layer_idx, which, start, end)
            if a.nbytes != ln:
                raise ValueError(f"buffer {which}: expected {a.nbytes} bytes, got {ln}")
            C.memmove(ptr, a.ctypes.data, ln)

    # ---- routing (src/moe.rs:2959-3050 set_routing_config / set_routing_weights)
    def set_routing_weights(self, moe_layer_idx: int, gate_bf16, bias_f32=None, e_score_correction_bias=None):
        g = gate_bf16
        if isinstance(g, torch.Tensor):
            g = g.detach().to(torch.bfloat16).cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
        g = np.ascontiguousarray(g, dtype=np.uint16)
        if g.shape != (self._cfg.n_routed_experts, self._cfg.hidden_size):
            raise ValueError(f"gate weight: expected {(self._cfg.n_routed_experts, self._cfg.hidden_size)}, got {g.shape}")

        def f32(a):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                a = a.detach().float().cpu().numpy()
            a = np.ascontiguousarray(a, dtype=np.float32)
            if a.shape != (self._cfg.n_routed_experts,):
                raise ValueError("bias must have shape [n_routed_experts]")
            return a
        b, cb = f32(bias_f32), f32(e_score_correction_bias)
        capi.check(self._lib.kb2_set_router_host(self._h, moe_layer_idx, g.ctypes.data,
                                                 b.ctypes.data if b is not None else None,
                                                 cb.ctypes.data if cb is not None else None))

    def compute_routing(self, moe_layer_idx: int, hidden: torch.Tensor):
        """TransformerLayer.compute_routing (layer.py:526-560): -> (topk_ids int32 [M,k], topk_weights f32 [M,k])."""
        self._check_act(hidden, "hidden")
        M, k = hidden.shape[0], self._cfg.num_experts_per_tok
        ids = torch.empty((M, k), dtype=torch.int32, device=hidden.device)
        w = torch.empty((M, k), dtype=torch.float32, device=hidden.device)
        capi.check(self._lib.kb2_route(self._h, moe_layer_idx, hidden.data_ptr(), M, ids.data_ptr(), w.data_ptr(),
                                       _stream_ptr(hidden.device)))
        return ids, w

    def _check_act(self, x: torch.Tensor, name: str):
        if not x.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor: krasis_b200 has no CPU path")
        if x.dtype != torch.bfloat16 or x.dim() != 2 or x.shape[1] != self._cfg.hidden_size or not x.is_contiguous():
            raise ValueError(f"{name}: expected contiguous bf16 [M, {self._cfg.hidden_size}], got {x.dtype} {tuple(x.shape)}")
        if x.device.index != self._cfg.device:
            raise ValueError(f"{name} is on {x.device}, engine is on cuda:{self._cfg.device}")

    def moe_forward(self, moe_layer_idx, hidden_states, topk_ids, topk_weights, routed_only=False, shared=None):
        self._check_act(hidden_states, "hidden_states")
        M, k = hidden_states.shape[0], self._cfg.num_experts_per_tok
        if tuple(topk_ids.shape) != (M, k) or topk_ids.dtype != torch.int32 or not topk_ids.is_contiguous():
            raise ValueError(f"topk_ids: expected contiguous int32 [{M}, {k}]")
        if tuple(topk_weights.shape) != (M, k) or topk_weights.dtype != torch.float32 or not topk_weights.is_contiguous():
            raise ValueError(f"topk_weights: expected contiguous float32 [{M}, {k}]")
        if shared is not None:
            self._check_act(shared, "shared")
        out = torch.empty_like(hidden_states)
        capi.check(self._lib.kb2_moe_forward(self._h, moe_layer_idx, hidden_states.data_ptr(), topk_ids.data_ptr(),
                                             topk_weights.data_ptr(), out.data_ptr(), M, int(bool(routed_only)),
                                             shared.data_ptr() if shared is not None else None,
                                             _stream_ptr(hidden_states.device)))
        return out

    def moe_forward_scatter(self, moe_layer_idx, hidden_states, topk_ids, topk_weights, peer_ptrs, src_rank: int) -> None:
        """Local expert slice over ALL tokens with the expert-parallel reduce-scatter fused into the combine kernel: the partial rows
        go straight into the token owners' receive buffers (peer_ptrs from Communicator.peer_alloc, [rows][world][H] bf16 each)."""
        self._check_act(hidden_states, "hidden_states")
        M, k, R = hidden_states.shape[0], self._cfg.num_experts_per_tok, len(peer_ptrs)
        if tuple(topk_ids.shape) != (M, k) or topk_ids.dtype != torch.int32 or not topk_ids.is_contiguous():
            raise ValueError(f"topk_ids: expected contiguous int32 [{M}, {k}]")
        if tuple(topk_weights.shape) != (M, k) or topk_weights.dtype != torch.float32 or not topk_weights.is_contiguous():
            raise ValueError(f"topk_weights: expected contiguous float32 [{M}, {k}]")
        arr = (C.c_void_p * R)(*peer_ptrs)
        capi.check(self._lib.kb2_moe_forward_scatter(self._h, moe_layer_idx, hidden_states.data_ptr(), topk_ids.data_ptr(),
                                                     topk_weights.data_ptr(), arr, R, src_rank, M, _stream_ptr(hidden_states.device)))

    def finish_slots(self, slots_ptr: int, world: int, rows: int, shared: Optional[torch.Tensor], like: torch.Tensor) -> torch.Tensor:
        """bf16(rsf * bf16(sum over the `world` partial rows of every token)) + shared, from this rank's receive buffer."""
        out = torch.empty((rows, self._cfg.hidden_size), dtype=torch.bfloat16, device=like.device)
        if shared is not None:
            self._check_act(shared, "shared")
        capi.check(self._lib.kb2_finish_routed_slots(self._h, slots_ptr, world, shared.data_ptr() if shared is not None else None,
                                                     out.data_ptr(), rows, _stream_ptr(like.device)))
        return out

    def finish(self, routed: torch.Tensor, shared: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bf16(rsf * routed) + shared, in place on `routed` (the tail of gpu_prefill.py:4467-4482 after an EP reduction)."""
        self._check_act(routed, "routed")
        if shared is not None:
            self._check_act(shared, "shared")
        capi.check(self._lib.kb2_finish_routed(self._h, routed.data_ptr(), shared.data_ptr() if shared is not None else None,
                                               routed.data_ptr(), routed.shape[0], _stream_ptr(routed.device)))
        return routed

    def moe_forward_host(self, moe_layer_idx, x_host: torch.Tensor, topk_ids=None, topk_weights=None,
                         routed_only=False, out_host: Optional[torch.Tensor] = None):
        """Host-buffer entry (bytes in / bytes out like submit_forward+sync_forward, src/moe.rs:2722,2809)."""
        if x_host.is_cuda or x_host.dtype != torch.bfloat16 or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous CPU bf16 tensor")
        M = x_host.shape[0]
        if out_host is None:
            out_host = torch.empty_like(x_host)
        ids_p = topk_ids.data_ptr() if topk_ids is not None else None
        w_p = topk_weights.data_ptr() if topk_weights is not None else None
        capi.check(self._lib.kb2_moe_forward_host(self._h, moe_layer_idx, x_host.data_ptr(), ids_p, w_p,
                                                  out_host.data_ptr(), M, int(bool(routed_only)),
                                                  _stream_ptr(self.device)))
        return out_host

    def prefill_moe_stack_host(self, x_host: torch.Tensor, out_host: torch.Tensor, first_layer=0, n_layers=None):
        """H2D x, route + MoE forward through layers [first_layer, first_layer+n_layers), D2H last output."""
        if x_host.is_cuda or out_host.is_cuda or x_host.dtype != torch.bfloat16 or out_host.dtype != torch.bfloat16:
            raise ValueError("x_host/out_host must be CPU bf16 tensors")
        n_layers = self._cfg.num_moe_layers - first_layer if n_layers is None else n_layers
        capi.check(self._lib.kb2_prefill_moe_stack_host(self._h, x_host.data_ptr(), out_host.data_ptr(),
                                                        x_host.shape[0], first_layer, n_layers,
                                                        _stream_ptr(self.device)))
        return out_host

    def profile(self, on: bool):
        capi.check(self._lib.kb2_profile_enable(self._h, int(bool(on))))

    def profile_collect(self):
        """-> {kernel class: (total_ms, launches)} since profile(True); synchronises the device."""
        ms = (C.c_double * len(capi.PROF_NAMES))()
        n = (C.c_int64 * len(capi.PROF_NAMES))()
        capi.check(self._lib.kb2_profile_collect(self._h, ms, n))
        return {name: (ms[i], n[i]) for i, name in enumerate(capi.PROF_NAMES)}

    def last_expert_counts(self) -> np.ndarray:
        out = np.zeros(self.expert_end - self.expert_start, np.int32)
        capi.check(self._lib.kb2_last_expert_counts(self._h, out.ctypes.data, _stream_ptr(self.device)))
        return out


class GpuPrefillManager:
    """Drop-in for python/krasis/gpu_prefill.py:GpuPrefillManager on the prefill path.

    Constructor keeps the reference's argument names (gpu_prefill.py:326-346); arguments that only
    steer expert streaming on 16 GB GPUs (chunk_size, layer_group_size, model_path) are accepted and
    ignored — on a 180 GB B200 every expert is resident."""

    def __init__(self, model_path: str = "", device=None, num_experts: int = 0, hidden_size: int = 0,
                 intermediate_size: int = 0, params_dtype=torch.bfloat16, n_shared_experts: int = 0,
                 routed_scaling_factor: float = 1.0, first_k_dense: int = 0, chunk_size=None, num_bits: int = 4,
                 krasis_engine: Optional[KrasisEngine] = None, num_moe_layers: int = 0, layer_group_size: int = 1,
                 skip_shared_experts: bool = False, swiglu_limit: float = 0.0, rank: int = 0, num_ranks: int = 1,
                 top_k: int = 0, max_tokens: int = 8192, scoring_func: str = "softmax", norm_topk_prob: bool = False):
        if params_dtype != torch.bfloat16:
            raise ValueError("only bfloat16 activations are supported")
        if swiglu_limit > 0:
            raise NotImplementedError("GPT-OSS activation (swiglu_limit > 0) is out of scope (SURVEY.md §8a)")
        device = torch.device(device if device is not None else "cuda:0")
        if krasis_engine is None:
            krasis_engine = KrasisEngine(hidden_size=hidden_size, moe_intermediate_size=intermediate_size,
                                         n_routed_experts=num_experts, num_experts_per_tok=top_k,
                                         num_moe_layers=num_moe_layers, num_bits=num_bits, rank=rank,
                                         num_ranks=num_ranks, scoring_func=scoring_func,
                                         norm_topk_prob=norm_topk_prob,
                                         routed_scaling_factor=routed_scaling_factor, max_tokens=max_tokens,
                                         device=device.index or 0)
        self._engine = krasis_engine
        self.device = krasis_engine.device
        self.num_experts = krasis_engine.num_experts()
        self.rank, self.num_ranks = rank, num_ranks
        self.expert_start, self.expert_end = krasis_engine.expert_start, krasis_engine.expert_end
        self.num_local_experts = self.expert_end - self.expert_start
        self.hidden_size, self.intermediate_size = krasis_engine.hidden_size(), krasis_engine.intermediate_size()
        self.n_shared_experts = 0 if skip_shared_experts else n_shared_experts
        self.routed_scaling_factor = routed_scaling_factor
        self.num_bits = num_bits
        self._max_tokens = max_tokens
        self._shared_engine: Optional[KrasisEngine] = None
        self._shared_loaded = set()

    def load_shared_expert(self, moe_layer_idx: int, q: QuantizedExperts):
        """Shared expert(s) of one layer in the manager's own INT4/INT8 format (gpu_prefill.py:4803-4840,
        `get_shared_expert_weights`): q holds ONE fused expert, w13 [1, 2*n_shared*I, H/8|H], w2 [1, H, n_shared*I/8|...].
        The reference runs it as a one-expert Marlin MoE with weight 1 (:4738-4801); so does this manager, through a second
        engine with E = 1, k = 1 and intermediate size n_shared * I."""
        if self.n_shared_experts <= 0:
            raise RuntimeError("manager was built without shared experts")
        if self._shared_engine is None:
            self._shared_engine = KrasisEngine(hidden_size=self.hidden_size,
                                               moe_intermediate_size=self.n_shared_experts * self.intermediate_size,
                                               n_routed_experts=1, num_experts_per_tok=1,
                                               num_moe_layers=self._engine.num_moe_layers(), num_bits=self.num_bits,
                                               max_tokens=self._max_tokens, device=self.device.index or 0)
        self._shared_engine.load_quantized_layer(moe_layer_idx, q)
        self._shared_loaded.add(moe_layer_idx)

    def _shared_expert_forward(self, moe_layer_idx: int, hidden_states: torch.Tensor) -> torch.Tensor:
        M = hidden_states.shape[0]
        ids = torch.zeros((M, 1), dtype=torch.int32, device=hidden_states.device)      # all tokens -> expert 0, weight 1
        w = torch.ones((M, 1), dtype=torch.float32, device=hidden_states.device)
        return self._shared_engine.moe_forward(moe_layer_idx, hidden_states, ids, w, routed_only=True)

    def forward(self, moe_layer_idx: int, hidden_states: torch.Tensor, topk_ids: torch.Tensor,
                topk_weights: torch.Tensor, routed_only: bool = False, shared_output: Optional[torch.Tensor] = None):
        """gpu_prefill.py:4374-4484.  `shared_output` (optional, [M,H] bf16) is added after the rsf scaling,
        which is what the reference does with its own shared-expert result (gpu_prefill.py:4471-4480)."""
        torch.cuda.set_device(self.device)                     # gpu_prefill.py:4401
        if not routed_only and shared_output is None and moe_layer_idx in self._shared_loaded:
            shared_output = self._shared_expert_forward(moe_layer_idx, hidden_states)      # gpu_prefill.py:4471-4480
        return self._engine.moe_forward(moe_layer_idx, hidden_states, topk_ids, topk_weights,
                                        routed_only=routed_only, shared=shared_output)
