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
