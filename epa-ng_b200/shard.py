"""Multi-GPU plumbing: query sharding and the single gather of placement records.

The reference's only parallelism across processes is a static block partition of the query file
(local_seq_package, /root/reference/src/net/epa_mpi_util.cpp:10-30): rank r takes the r-th block of
ceil(Q / ranks) sequences, every rank holds the whole reference tree, and no data moves between
ranks while placing. Here one process drives one GPU; the per-query fixed-stride placement records
(filter_max x 40 bytes + a count) of all ranks are gathered ONCE to rank 0 with torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests) and rank 0 formats the jplace in global query order.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_queries: int, rank: int, world: int):
    """[lo, hi) of the queries rank `rank` places: blocks of ceil(Q / world), the last may be short
    or empty (same rule as the reference's MPI partition)."""
    part = -(-n_queries // world)
    lo = min(n_queries, rank * part)
    return lo, min(n_queries, lo + part)


def gather_records(records: torch.Tensor, counts: torch.Tensor, n_queries: int, dst: int = 0):
    """records: [part, stride] float64 view of this rank's 40-byte records (padded to the common
    part size), counts: [part] int32. Returns (records[n_queries], counts[n_queries]) on `dst`
    in global query order, (None, None) elsewhere. One gather per tensor, equal-sized buffers."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return records[:n_queries], counts[:n_queries]
    rank = dist.get_rank()
    part = -(-n_queries // world)
    assert records.shape[0] == part and counts.shape[0] == part, "pad every shard to ceil(Q / world) rows"
    rec_list = [torch.empty_like(records) for _ in range(world)] if rank == dst else None
    cnt_list = [torch.empty_like(counts) for _ in range(world)] if rank == dst else None
    dist.gather(records, rec_list, dst=dst)
    dist.gather(counts, cnt_list, dst=dst)
    if rank != dst:
        return None, None
    return torch.cat(rec_list)[:n_queries], torch.cat(cnt_list)[:n_queries]
