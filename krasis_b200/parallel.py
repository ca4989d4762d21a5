"""Expert-parallel prefill across GPUs: one process per GPU, torch.distributed (NCCL) for the exchange.

The reference's "EP" (python/krasis/model.py:3086-3211) replicates every token to every GPU through pinned host
memory, zeroes non-local routing weights and adds the partial sums on GPU0.  Here tokens are SHARDED across
ranks and only routed rows travel:

    route(x_local) -> bin rows by global expert id (=> by owner rank) -> all-to-all(v) rows/weights/ids
    -> owner: grouped expert GEMMs on the received rows -> all-to-all(v) back -> weighted combine at home.

All arithmetic and all row movement inside a rank is done by libkrasis_b200 kernels (kb2_ep_bin_rows,
kb2_moe_forward_rows, kb2_ep_combine_rows); torch supplies buffers, streams and the collective only.
The split-size bookkeeping below is pure host logic and is tested on CPU with the gloo backend.
"""
import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import capi


def expert_range(rank: int, num_ranks: int, num_experts: int):
    """python/krasis/gpu_prefill.py:353-359: contiguous ranges, the last rank takes the remainder."""
    per = num_experts // num_ranks
    start = rank * per
    end = num_experts if rank == num_ranks - 1 else (rank + 1) * per
    return start, end


def send_splits_from_counts(counts: List[int], num_ranks: int) -> List[int]:
    """Rows grouped by global expert id are also grouped by owner rank: split sizes = per-owner sums."""
    E = len(counts)
    out = []
    for r in range(num_ranks):
        s, e = expert_range(r, num_ranks, E)
        out.append(int(sum(counts[s:e])))
    return out


def exchange_splits(send_splits: List[int], group=None, device="cpu") -> List[int]:
    """all-to-all of the split sizes: recv_splits[src] = rows rank `src` sends to this rank."""
    world = dist.get_world_size(group)
    t_in = torch.tensor(send_splits, dtype=torch.int64, device=device)
    t_out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(t_out, t_in, group=group)
    return [int(v) for v in t_out.tolist()]


def all_to_all_rows(src: torch.Tensor, send_splits: List[int], recv_splits: List[int], group=None) -> torch.Tensor:
    """Variable-size all-to-all of rows (dim 0)."""
    out = torch.empty((sum(recv_splits),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    dist.all_to_all_single(out, src[: sum(send_splits)], output_split_sizes=recv_splits,
                           input_split_sizes=send_splits, group=group)
    return out


def splits_from_count_matrix(count_matrix: List[List[int]], rank: int):
    """count_matrix[src][e] = rows rank `src` routes to global expert e (all-gathered).  Returns
    (send_splits, recv_splits) of `rank` — all split bookkeeping from ONE gathered matrix, one host sync."""
    world = len(count_matrix)
    send = send_splits_from_counts(count_matrix[rank], world)
    s, e = expert_range(rank, world, len(count_matrix[rank]))
    recv = [int(sum(count_matrix[src][s:e])) for src in range(world)]
    return send, recv


class Communicator:
    """The expert-parallel communicator behind the C ABI (kb2_comm_*, include/krasis_b200.h): NCCL over NVLink owned by
    libkrasis_b200.  torch.distributed (any backend) is used ONCE, to hand rank 0's 128-byte unique id to the other ranks —
    the only thing a Rust / C host would have to move between its processes."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        self._lib = capi.load()
        self.rank, self.world, self.device = rank, world, torch.device("cuda", device)
        self._peer_allocs = []
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        capi.check(self._lib.kb2_comm_init(buf, rank, world, device, C.byref(self._h)))

    @staticmethod
    def new_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.load().kb2_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int, group=None):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(box[0], rank, world, device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kb2_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def all_gather_rows(self, x: torch.Tensor) -> torch.Tensor:
        """[rows, ...] on every rank -> [world * rows, ...] in rank order."""
        x = x.contiguous()
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        capi.check(self._lib.kb2_comm_all_gather(self._h, x.data_ptr(), out.data_ptr(), x.numel() * x.element_size(), self._stream()))
        return out

    def reduce_scatter_rows(self, x: torch.Tensor) -> torch.Tensor:
        """BF16 [world * rows, H] partial sums -> this rank's [rows, H] slice of the sum over ranks."""
        if x.dtype != torch.bfloat16 or x.shape[0] % self.world or not x.is_contiguous():
            raise ValueError("reduce_scatter_rows: contiguous bf16 [world * rows, ...] expected")
        out = torch.empty((x.shape[0] // self.world,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        capi.check(self._lib.kb2_comm_reduce_scatter_bf16(self._h, x.data_ptr(), out.data_ptr(), out.numel(), self._stream()))
        return out

    def all_reduce(self, x: torch.Tensor) -> torch.Tensor:
        capi.check(self._lib.kb2_comm_all_reduce_bf16(self._h, x.data_ptr(), x.data_ptr(), x.numel(), self._stream()))
        return x

    def peer_alloc(self, nbytes: int):
        """Collective: `nbytes` of zero-filled device memory on every rank, mapped into every process through CUDA IPC.  Returns the
        list of device pointers [world] through which THIS rank addresses each rank's buffer (entry `rank` is the local one)."""
        arr = (C.c_void_p * self.world)()
        capi.check(self._lib.kb2_comm_peer_alloc(self._h, nbytes, arr))
        ptrs = [int(arr[r]) for r in range(self.world)]
        self._peer_allocs.append(arr)
        return ptrs

    def barrier(self):
        """Stream-ordered cross-rank barrier (4-byte all-reduce): peers' kernels enqueued before it — and their peer-memory stores —
        are complete when it returns on this rank's stream."""
        capi.check(self._lib.kb2_comm_barrier(self._h, self._stream()))

    def reduce_rows(self, x: torch.Tensor, root: int, out: torch.Tensor = None) -> torch.Tensor:
        """BF16 sum over ranks of `x`, delivered to `root` (returns `out` there, None elsewhere)."""
        if x.dtype != torch.bfloat16 or not x.is_contiguous():
            raise ValueError("reduce_rows: contiguous bf16 tensor expected")
        if self.rank == root and out is None:
            out = torch.empty_like(x)
        capi.check(self._lib.kb2_comm_reduce_bf16(self._h, x.data_ptr(), out.data_ptr() if self.rank == root else None, x.numel(), root,
                                                  self._stream()))
        return out if self.rank == root else None

    def broadcast(self, x: torch.Tensor, root: int) -> torch.Tensor:
        capi.check(self._lib.kb2_comm_broadcast(self._h, x.data_ptr(), x.numel() * x.element_size(), root, self._stream()))
        return x


def token_shard(num_tokens: int, rank: int, world: int):
    """Rows [lo, hi) of the residual stream owned by `rank` (token-sharded prefill); num_tokens % world == 0."""
    if num_tokens % world:
        raise ValueError(f"{num_tokens} tokens do not divide over {world} ranks")
    per = num_tokens // world
    return rank * per, (rank + 1) * per


class ExpertParallelMoE:
    """Prefill MoE forward for a token shard, experts sharded over the ranks of `group`."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        cfg = engine._cfg
        if cfg.num_ranks != self.world or cfg.rank != self.rank:
            raise ValueError("engine rank/num_ranks must match the process group")
        self.k, self.H, self.E = cfg.num_experts_per_tok, cfg.hidden_size, cfg.n_routed_experts
        dev = engine.device
        cap = cfg.max_tokens * self.k
        self.cap = cap
        bf, f32, i32 = torch.bfloat16, torch.float32, torch.int32
        self._xs = torch.empty((cap, self.H), dtype=bf, device=dev)      # rows grouped by expert (send order)
        self._meta = torch.empty((cap, 2), dtype=i32, device=dev)        # [:,0] expert id, [:,1] weight bits
        self._slot = torch.empty(cap, dtype=i32, device=dev)
        self._counts = torch.empty(self.E, dtype=i32, device=dev)
        self._all_counts = torch.empty((self.world, self.E), dtype=i32, device=dev)
        self._rows = torch.empty((cap, self.H), dtype=bf, device=dev)    # received rows
        self._rmeta = torch.empty((cap, 2), dtype=i32, device=dev)
        self._orows = torch.empty((cap, self.H), dtype=bf, device=dev)   # expert outputs for received rows
        self._yrows = torch.empty((cap, self.H), dtype=bf, device=dev)   # returned rows (send order)

    def forward(self, moe_layer_idx: int, x_local: torch.Tensor, topk_ids: Optional[torch.Tensor] = None,
                topk_weights: Optional[torch.Tensor] = None, routed_only: bool = False,
                shared: Optional[torch.Tensor] = None) -> torch.Tensor:
        eng, lib, h = self.engine, self.engine._lib, self.engine._h
        eng._check_act(x_local, "x_local")
        M = x_local.shape[0]
        stream = torch.cuda.current_stream(x_local.device).cuda_stream
        if topk_ids is None:
            topk_ids, topk_weights = eng.compute_routing(moe_layer_idx, x_local)
        capi.check(lib.kb2_ep_bin_rows(h, x_local.data_ptr(), topk_ids.data_ptr(), topk_weights.data_ptr(), M,
                                       self._xs.data_ptr(), self._ws_buf().data_ptr(), self._is_buf().data_ptr(),
                                       self._slot.data_ptr(), self._counts.data_ptr(), stream))
        dist.all_gather_into_tensor(self._all_counts.view(-1), self._counts, group=self.group)
        cm = self._all_counts.cpu().tolist()                                                 # the one host sync per layer
        send, recv = splits_from_count_matrix(cm, self.rank)
        n_send, n_recv = sum(send), sum(recv)
        # every rank holds the whole count matrix: check EVERY rank's receive total so that all ranks raise together
        # (a rank that raised alone would leave the others blocked inside the all-to-all)
        worst = max(sum(splits_from_count_matrix(cm, r)[1]) for r in range(self.world))
        if worst > self.cap:
            raise ValueError(f"a rank would receive {worst} rows > capacity {self.cap}; raise max_tokens")
        rows, ids, wts = self._rows[:n_recv], self._rid_buf()[:n_recv], self._rw_buf()[:n_recv]
        dist.all_to_all_single(rows, self._xs[:n_send], output_split_sizes=recv, input_split_sizes=send, group=self.group)
        dist.all_to_all_single(ids, self._is_buf()[:n_send], output_split_sizes=recv, input_split_sizes=send, group=self.group)
        dist.all_to_all_single(wts, self._ws_buf()[:n_send], output_split_sizes=recv, input_split_sizes=send, group=self.group)
        out_rows = self._orows[:n_recv]
        capi.check(lib.kb2_moe_forward_rows(h, moe_layer_idx, rows.data_ptr(), ids.data_ptr(), wts.data_ptr(),
                                            out_rows.data_ptr(), n_recv, stream))
        back = self._yrows[:n_send]
        dist.all_to_all_single(back, out_rows, output_split_sizes=send, input_split_sizes=recv, group=self.group)
        out = torch.empty_like(x_local)
        capi.check(lib.kb2_ep_combine_rows(h, back.data_ptr(), self._slot.data_ptr(), M, int(bool(routed_only)),
                                           shared.data_ptr() if shared is not None else None, out.data_ptr(), stream))
        return out

    # contiguous metadata buffers (the [cap,2] tensor is split into two contiguous halves)
    def _is_buf(self): return self._meta.view(-1)[: self.cap]
    def _ws_buf(self): return self._meta.view(-1)[self.cap:].view(torch.float32)
    def _rid_buf(self): return self._rmeta.view(-1)[: self.cap]
    def _rw_buf(self): return self._rmeta.view(-1)[self.cap:].view(torch.float32)
