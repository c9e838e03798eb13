"""Peer-memory gather (shard.PeerRecords, epa_peer_*): two processes, one GPU each, on one box. Rank 1's collect
kernel writes its records into rank 0's buffer over NVLink; the buffer must equal an NCCL gather of the same records.
Needs two GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, helpers.ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cpu = dist.new_group(backend="gloo")
    case = helpers.synth64_case()
    ctx = helpers.make_context(case, device=rank)
    ctx.build_lookup()
    opts = pkg.capi.default_options()
    rows = case.query_rows
    part = -(-rows.shape[0] // world)
    lo, hi = pkg.shard.shard_range(rows.shape[0], rank, world)
    mine = np.ascontiguousarray(rows[lo:hi])
    fmax = opts.filter_max
    dev = torch.device("cuda", rank)
    rec = torch.zeros((part, fmax * 5), dtype=torch.float64, device=dev)
    cnt = torch.zeros(part, dtype=torch.int32, device=dev)
    peer = pkg.shard.PeerRecords(pkg.capi, part, fmax, rank, dst=0, group=cpu)
    for target in ("local", "peer"):
        ctx.upload_queries(mine)
        ctx.preplace()
        ctx.select(opts)
        ctx.place_pairs(opts)
        if target == "local":
            ctx.collect_dev(opts, rec.data_ptr(), cnt.data_ptr())
        else:
            ctx.collect_dev(opts, *peer.slice_ptrs(0))
    torch.cuda.synchronize()
    peer.complete()
    g_rec, g_cnt = pkg.shard.gather_records(rec, cnt, part * world, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        n = rows.shape[0]
        ok = bool(torch.equal(g_rec[:n], peer.records[:n]) and torch.equal(g_cnt[:n], peer.counts[:n]))
        filled = int(peer.counts[:n].sum().item())
        open(out_path, "w").write("%d %d" % (int(ok), filled))
    dist.barrier()
    if rank != 0:
        peer.close()
    dist.barrier()
    if rank == 0:
        peer.close()
    ctx.close()
    dist.destroy_process_group()


def test_peer_written_records_equal_the_nccl_gather(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, 29577, out), nprocs=2, join=True)
    ok, filled = open(out).read().split()
    assert ok == "1" and int(filled) >= 64
