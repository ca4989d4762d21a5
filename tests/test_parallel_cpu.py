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


def test_all_to_all_dispatch_roundtrip_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == 2
    assert sum(v[0] for v in ret.values()) == sum(v[1] for v in ret.values())
