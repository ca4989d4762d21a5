"""Host-side logic of the expert-parallel path on CPU with the gloo backend, world_size 2
(the reference has no fake backend at all; tests/test_parallel_prefill.py needs >= 2 real GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from krasis_b200 import parallel as P


def test_expert_range_matches_reference_slicing():
    # python/krasis/gpu_prefill.py:353-359
    assert [P.expert_range(r, 3, 512) for r in range(3)] == [(0, 170), (170, 340), (340, 512)]
    assert [P.expert_range(r, 8, 512) for r in range(8)][-1] == (448, 512)
    assert P.expert_range(0, 1, 64) == (0, 64)


def test_send_splits_from_counts():
    counts = list(range(12))                     # 12 experts, 5 ranks -> 2,2,2,2,4 experts
    assert P.send_splits_from_counts(counts, 5) == [0 + 1, 2 + 3, 4 + 5, 6 + 7, 8 + 9 + 10 + 11]
    assert sum(P.send_splits_from_counts(counts, 5)) == sum(counts)


def test_splits_from_count_matrix_is_consistent():
    rng = np.random.default_rng(0)
    cm = rng.integers(0, 50, (4, 16)).tolist()
    sends, recvs = zip(*(P.splits_from_count_matrix(cm, r) for r in range(4)))
    for a in range(4):
        for b in range(4):
            assert sends[a][b] == recvs[b][a]          # what a sends to b is what b expects from a
    assert [sum(s) for s in sends] == [sum(row) for row in cm]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        E, k, H = 8, 2, 4
        rng = np.random.default_rng(100 + rank)
        M = 5 + rank
        ids = np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int64)
        # rows grouped by global expert id (what kb2_ep_bin_rows produces), payload = (src rank, token, expert)
        order = np.argsort(ids.reshape(-1), kind="stable")
        flat_e = ids.reshape(-1)[order]
        rows = torch.tensor(np.stack([np.full_like(flat_e, rank), order // k, flat_e, flat_e * 0], axis=1), dtype=torch.float32)
        counts = np.bincount(flat_e, minlength=E).tolist()
        send = P.send_splits_from_counts(counts, world)
        recv = P.exchange_splits(send)
        got = P.all_to_all_rows(rows, send, recv)
        s, e = P.expert_range(rank, world, E)
        assert got.shape[0] == sum(recv)
        assert ((got[:, 2] >= s) & (got[:, 2] < e)).all(), "received a row for an expert this rank does not own"
        # rows arrive grouped by source rank, in source order
        assert (got[:, 0].numpy() == np.repeat(np.arange(world), recv)).all()
        # send the rows back: reverse splits restore the original order
        back = torch.empty_like(rows)
        dist.all_to_all_single(back, got, output_split_sizes=send, input_split_sizes=recv)
        assert torch.equal(back, rows)
        ret[rank] = (sum(send), sum(recv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_all_to_all_dispatch_roundtrip_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    assert sum(v[0] for v in ret.values()) == sum(v[1] for v in ret.values())


@pytest.mark.parametrize("world", [2, 4, 8])
def test_head_parallel_shards_partition_the_layer(world):
    """shard_gdn_weights / shard_gqa_weights: the ranks' slices tile the projections exactly once (KV heads are replicated
    when there are fewer KV heads than ranks), for the Qwen3-Coder-Next geometry at 2, 4 and 8 ranks."""
    from krasis_b200.model import HybridMoEConfig, shard_gdn_weights, shard_gqa_weights
    cfg = HybridMoEConfig(hidden_size=64, num_hidden_layers=4, linear_key_head_dim=8, linear_value_head_dim=8, gqa_head_dim=8)
    nk, nv, dk, dv = cfg.linear_num_key_heads, cfg.linear_num_value_heads, cfg.linear_key_head_dim, cfg.linear_value_head_dim
    nh, nkv, d, H = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.gqa_head_dim, cfg.hidden_size
    kd, vd = nk * dk, nv * dv
    g = torch.Generator().manual_seed(0)
    w = dict(in_proj_qkvz=torch.randn(2 * kd + 2 * vd, H, generator=g), in_proj_ba=torch.randn(2 * nv, H, generator=g),
             out_proj=torch.randn(H, vd, generator=g), conv1d_weight=torch.randn(2 * kd + vd, 1, 4, generator=g),
             A_log=torch.randn(nv, generator=g), dt_bias=torch.randn(nv, generator=g), norm_weight=torch.randn(dv, generator=g))
    parts = [shard_gdn_weights(w, cfg, r, world) for r in range(world)]
    assert torch.equal(torch.cat([p["in_proj_qkvz"] for p in parts]), w["in_proj_qkvz"])       # key-head groups are contiguous
    assert torch.equal(torch.cat([p["in_proj_ba"] for p in parts]), w["in_proj_ba"])
    assert torch.equal(torch.cat([p["out_proj"] for p in parts], dim=1), w["out_proj"])
    assert torch.equal(torch.cat([p["A_log"] for p in parts]), w["A_log"])
    assert sum(p["conv1d_weight"].shape[0] for p in parts) == 2 * kd + vd
    qw = d * (2 if cfg.gated_attention else 1)
    wq = dict(q_proj=torch.randn(nh * qw, H, generator=g), k_proj=torch.randn(nkv * d, H, generator=g),
              v_proj=torch.randn(nkv * d, H, generator=g), o_proj=torch.randn(H, nh * d, generator=g))
    grp = nh // nkv
    seen_q = []
    for r in range(world):
        p, nh_l, nkv_l = shard_gqa_weights(wq, cfg, r, world)
        assert nh_l == nh // world and p["q_proj"].shape[0] == nh_l * qw and p["k_proj"].shape[0] == nkv_l * d
        h0 = r * nh // world
        assert torch.equal(p["k_proj"], wq["k_proj"][(h0 // grp) * d:(h0 // grp + nkv_l) * d])  # the KV heads of this rank's groups
        assert nh_l % nkv_l == 0                                                                 # local GQA grouping stays uniform
        seen_q.append(p["q_proj"])
    assert torch.equal(torch.cat(seen_q), wq["q_proj"])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_token_shards_and_mla_head_shards_partition(world):
    """Token-sharded prefill: the ranks' row ranges tile [0, M); MLA head shards tile q / w_kc / w_vc / o_proj exactly once
    (DeepSeek-V2-Lite geometry, 16 heads)."""
    from krasis_b200.model import DEEPSEEK_V2_LITE as cfg, shard_mla_weights
    M = 8192
    spans = [P.token_shard(M, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == M and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        P.token_shard(M + 1, 0, world)
    nh, qd, dv, lora, H = cfg.num_attention_heads, cfg.qk_nope_head_dim + cfg.qk_rope_head_dim, cfg.v_head_dim, 64, 32
    g = torch.Generator().manual_seed(0)
    w = dict(q_proj=torch.randn(nh * qd, H, generator=g), kv_a_proj_with_mqa=torch.randn(lora + 64, H, generator=g),
             kv_a_layernorm=torch.randn(lora, generator=g), w_kc=torch.randn(nh, 128, lora, generator=g),
             w_vc=torch.randn(nh, dv, lora, generator=g), o_proj=torch.randn(H, nh * dv, generator=g))
    parts = [shard_mla_weights(w, cfg, r, world) for r in range(world)]
    assert all(n == nh // world for _, n in parts)
    assert torch.equal(torch.cat([p["q_proj"] for p, _ in parts]), w["q_proj"])
    assert torch.equal(torch.cat([p["w_kc"] for p, _ in parts]), w["w_kc"]) and torch.equal(torch.cat([p["w_vc"] for p, _ in parts]), w["w_vc"])
    assert torch.equal(torch.cat([p["o_proj"] for p, _ in parts], dim=1), w["o_proj"])
    assert all(torch.equal(p["kv_a_proj_with_mqa"], w["kv_a_proj_with_mqa"]) for p, _ in parts)      # latent projection replicated


def _sharded_schedule_worker(rank, world, port, ret):
    """The token-sharded layer schedule with gloo collectives standing in for kb2_comm_*: all-gather rows -> per-rank
    partial of a column-sharded linear map -> reduce-scatter == the unsharded map on this rank's rows."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, H, N = 8 * world, 16, 4 * world
        g = torch.Generator().manual_seed(5)
        x, w = torch.randn(M, H, generator=g), torch.randn(N, H, generator=g)
        wo = torch.randn(H, N, generator=g)
        lo, hi = P.token_shard(M, rank, world)
        parts = [torch.empty(hi - lo, H) for _ in range(world)]
        dist.all_gather(parts, x[lo:hi].contiguous())
        full = torch.cat(parts)
        assert torch.equal(full, x)                                   # rank order == token order
        n0, n1 = rank * N // world, (rank + 1) * N // world           # "heads" of this rank
        partial = (full @ w[n0:n1].T) @ wo[:, n0:n1].T                 # [M, H] partial output of the head-parallel block
        out = torch.empty(hi - lo, H)
        dist.reduce_scatter(out, list(partial.split(M // world)))
        want = ((x @ w.T) @ wo.T)[lo:hi]
        ret[rank] = float((out - want).abs().max() / want.abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_token_sharded_schedule_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_schedule_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and max(ret.values()) < 1e-5


def test_chunk_kv_state_view_offsets_the_live_sequence():
    """The view a layer gets while it processes a later token chunk of the same prefill call (pipelined attention collectives):
    seq_len is shifted, capacity requests cover the shifted range, the page table is the live sequence's."""
    from krasis_b200.model import _ChunkKVState

    class _Seq:
        def __init__(self):
            self.seq_len, self.asked = 48, []

        def ensure_capacity(self, n):
            self.asked.append(n)

        def kv_indices(self, device):
            return ("pages", device)

    st = _Seq()
    v = _ChunkKVState(st, 2048)
    assert v.seq_len == 48 + 2048
    v.ensure_capacity(1024)
    assert st.asked == [2048 + 1024]
    assert v.kv_indices("cuda:0") == ("pages", "cuda:0")
    st.seq_len = 100                     # the view follows the live state
    assert v.seq_len == 2148
