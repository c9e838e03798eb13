"""CPU, world_size 2 over gloo: the N > 1 host logic (query sharding + the one gather of placement
records). No device work: every rank fabricates the records of its shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_queries, stride, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = helpers.pkg().shard
    lo, hi = shard.shard_range(n_queries, rank, world)
    part = -(-n_queries // world)
    rec = torch.zeros((part, stride), dtype=torch.float64)
    cnt = torch.zeros(part, dtype=torch.int32)
    # record of global query g: every field = g + field index / 100, count = g % 7 + 1
    g = torch.arange(lo, hi, dtype=torch.float64)
    rec[: hi - lo] = g[:, None] + torch.arange(stride, dtype=torch.float64)[None, :] / 100.0
    cnt[: hi - lo] = (torch.arange(lo, hi) % 7 + 1).to(torch.int32)
    all_rec, all_cnt = shard.gather_records(rec, cnt, n_queries, dst=0)
    # the compacted form of the same gather (only the filled records travel)
    fmax = stride // 5
    cnt_c = torch.clamp(cnt, max=fmax)
    parts, cnt_g = shard.gather_compact(rec, cnt_c, n_queries, dst=0)
    if rank == 0:
        full = shard.expand_compact(parts, cnt_g, fmax)
        want = all_rec.view(n_queries, fmax, 5) * (torch.arange(fmax)[None, :] < cnt_g[:, None]).unsqueeze(-1)
        assert torch.equal(full.view(n_queries, fmax, 5), want.to(full.dtype))
        assert sum(p.shape[0] for p in parts) == int(cnt_g.sum())
        np.savez(out_path, rec=all_rec.numpy(), cnt=all_cnt.numpy())
    else:
        assert all_rec is None and all_cnt is None and parts is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_queries", [10, 11, 1])
def test_shard_and_gather_world2(tmp_path, n_queries):
    helpers.pkg()
    world, stride = 2, 35
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), n_queries, stride, out), nprocs=world, join=True)
    d = np.load(out)
    g = np.arange(n_queries, dtype=np.float64)
    assert d["rec"].shape == (n_queries, stride)
    assert np.array_equal(d["rec"], g[:, None] + np.arange(stride)[None, :] / 100.0)
    assert np.array_equal(d["cnt"], (np.arange(n_queries) % 7 + 1).astype(np.int32))


def test_shard_range_matches_reference_partition():
    shard = helpers.pkg().shard
    # src/net/epa_mpi_util.cpp:10-30: part = ceil(Q / ranks); rank r gets [r*part, min(Q, (r+1)*part))
    for q in (0, 1, 7, 8, 9, 1000001):
        for w in (1, 2, 3, 8):
            spans = [shard.shard_range(q, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == q
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            part = -(-q // w)
            assert all(hi - lo <= part for lo, hi in spans)


def test_fixed_capacity_compaction_matches_masked_compaction():
    """gather_compact_fixed (no host synchronisation) packs the same records in the same order as compact_records"""
    shard = helpers.pkg().shard
    g = torch.Generator().manual_seed(5)
    q, fmax = 257, 7
    counts = torch.randint(0, fmax + 1, (q,), generator=g, dtype=torch.int32)
    rec = torch.rand((q, fmax * 5), generator=g, dtype=torch.float64)
    want = shard.compact_records(rec, counts)
    cap = int(counts.sum()) + 11
    got, cnt = shard.gather_compact_fixed(rec, counts, cap)
    assert got.shape == (1, cap, 5) and torch.equal(cnt[0], counts)
    assert torch.equal(got[0][: want.shape[0]], want) and not got[0][want.shape[0]:].any()
    # too small a capacity drops the tail instead of writing out of bounds
    got2, _ = shard.gather_compact_fixed(rec, counts, want.shape[0] - 5)
    assert torch.equal(got2[0], want[: want.shape[0] - 5])
