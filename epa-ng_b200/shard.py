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


def compact_records(records: torch.Tensor, counts: torch.Tensor, fields: int = 5):
    """records: [Q, filter_max * fields] fixed-stride rows of which only the first counts[q] records of query q
    are filled (on average 1.2 of 7). Returns the filled records as one [n, fields] tensor in query order."""
    q = records.shape[0]
    fmax = records.shape[1] // fields
    keep = torch.arange(fmax, device=records.device)[None, :] < counts[:, None].to(torch.int64)
    return records.view(q, fmax, fields)[keep]


def gather_compact(records: torch.Tensor, counts: torch.Tensor, n_queries: int, dst: int = 0, fields: int = 5):
    """The gather of gather_records with the records compacted first (counts-prefix form): every rank sends only
    its filled records, padded to the largest shard's number, plus its counts. On `dst` returns
    (list of per-rank [n_r, fields] record tensors, counts[n_queries]); (None, None) elsewhere. The fixed-stride
    rows of query g are rebuilt from the counts' prefix sums (expand_compact)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    comp = compact_records(records, counts, fields)
    if world == 1:
        return [comp], counts[:n_queries]
    rank = dist.get_rank()
    n_mine = torch.tensor([comp.shape[0]], dtype=torch.int64, device=records.device)
    n_all = [torch.zeros_like(n_mine) for _ in range(world)]
    dist.all_gather(n_all, n_mine)
    sizes = [int(x.item()) for x in n_all]
    cap = max(1, max(sizes))
    padded = torch.zeros((cap, fields), dtype=records.dtype, device=records.device)
    padded[: comp.shape[0]] = comp
    rec_list = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    cnt_list = [torch.empty_like(counts) for _ in range(world)] if rank == dst else None
    dist.gather(padded, rec_list, dst=dst)
    dist.gather(counts, cnt_list, dst=dst)
    if rank != dst:
        return None, None
    return [r[:n] for r, n in zip(rec_list, sizes)], torch.cat(cnt_list)[:n_queries]


def gather_compact_fixed(records: torch.Tensor, counts: torch.Tensor, cap: int, dst: int = 0, fields: int = 5):
    """gather_compact without a host synchronisation (the caller can queue the next step's kernels behind it): the
    filled records are packed to the front of a [cap, fields] buffer by their counts' prefix sums (records beyond `cap`
    land in a spill row and are dropped: the caller checks counts.sum() <= cap afterwards), then one gather of the
    equal-sized buffers and one of the counts. On `dst` returns (records [world, cap, fields], counts [world, part]);
    (None, None) elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    q = records.shape[0]
    fmax = records.shape[1] // fields
    c64 = counts.to(torch.int64)
    first = torch.cumsum(c64, 0) - c64
    slot = torch.arange(fmax, device=records.device)[None, :]
    dest = torch.where(slot < c64[:, None], first[:, None] + slot, cap)          # unfilled slots -> spill row
    dest = torch.clamp(dest, max=cap).reshape(-1)
    comp = torch.zeros((cap + 1, fields), dtype=records.dtype, device=records.device)
    comp.index_copy_(0, dest, records.view(q * fmax, fields))
    comp = comp[:cap]
    if world == 1:
        return comp[None], counts[None]
    rank = dist.get_rank()
    rec_list = [torch.empty_like(comp) for _ in range(world)] if rank == dst else None
    cnt_list = [torch.empty_like(counts) for _ in range(world)] if rank == dst else None
    dist.gather(comp.contiguous(), rec_list, dst=dst)
    dist.gather(counts, cnt_list, dst=dst)
    if rank != dst:
        return None, None
    return torch.stack(rec_list), torch.stack(cnt_list)


def expand_compact(parts, counts: torch.Tensor, filter_max: int, fields: int = 5):
    """Inverse of the compaction on the receiving side: fixed-stride [Q, filter_max * fields] rows in global query order."""
    comp = torch.cat(parts) if len(parts) > 1 else parts[0]
    q = counts.shape[0]
    out = torch.zeros((q, filter_max, fields), dtype=comp.dtype, device=comp.device)
    keep = torch.arange(filter_max, device=comp.device)[None, :] < counts[:, None].to(torch.int64)
    out[keep] = comp
    return out.view(q, filter_max * fields)


class _CudaArray:
    """zero-copy view of raw device memory for torch.as_tensor (__cuda_array_interface__)"""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerRecords:
    """The gather without a collective, for one process per GPU on one NVLink box: rank `dst` owns the fixed-stride
    records [world * part, filter_max * 5] (float64 view of the 40-byte records) and the counts [world * part] of ALL
    shards in its device memory (epa_peer_alloc) and sends the 64-byte CUDA IPC handle to the other ranks; they map the
    buffer (epa_peer_open) and give `slice_ptrs()` - their block of it - to epa_collect_dev, so that the collect kernel
    writes the records straight into the owner's memory over NVLink while it runs. `complete()` is the barrier that ends
    the gather: a one-element all-reduce, stream-ordered behind every rank's collect kernels.
    (`group`: a process group for the handle broadcast, e.g. a gloo group; default group otherwise.)"""

    def __init__(self, capi, part: int, filter_max: int, device: int, dst: int = 0, group=None):
        import ctypes as C
        self.capi, self.device, self.dst = capi, int(device), dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.part, self.fmax = int(part), int(filter_max)
        self.rec_bytes = self.world * self.part * self.fmax * 40
        self.cnt_off = (self.rec_bytes + 255) & ~255
        total = self.cnt_off + self.world * self.part * 4
        lib = capi.load()
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self.owner = self.rank == dst
        box = [None]
        if self.owner:
            rc = lib.epa_peer_alloc(self.device, total, C.byref(ptr), handle)
            box = [bytes(handle.raw) if rc == 0 else "error %d: %s" % (rc, lib.epa_last_error(None).decode())]
        dist.broadcast_object_list(box, src=dst, group=group)          # (an allocation failure reaches every rank)
        if isinstance(box[0], str):
            raise capi.EpaError(-2, box[0])
        if not self.owner:
            rc = lib.epa_peer_open(self.device, box[0], C.byref(ptr))
            if rc != 0:
                raise capi.EpaError(rc, lib.epa_last_error(None).decode())
        self.ptr = int(ptr.value)
        dev = torch.device("cuda", self.device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.records = self.counts = None
        if self.owner:
            self.records = torch.as_tensor(_CudaArray(self.ptr, (self.world * self.part, self.fmax * 5), "<f8"), device=dev)
            self.counts = torch.as_tensor(_CudaArray(self.ptr + self.cnt_off, (self.world * self.part,), "<i4"), device=dev)

    def slice_ptrs(self, first_query: int = 0):
        """device pointers (records, counts) of query `first_query` of this rank's block"""
        q = self.rank * self.part + first_query
        return self.ptr + q * self.fmax * 40, self.ptr + self.cnt_off + q * 4

    def complete(self):
        dist.all_reduce(self.flag)

    def close(self):
        """Unmaps (owner: frees) the buffer. The caller makes sure that no rank still writes to it (a barrier)."""
        if self.ptr:
            lib = self.capi.load()
            torch.cuda.synchronize(self.device)
            (lib.epa_peer_free if self.owner else lib.epa_peer_close)(self.device, self.ptr)
            self.ptr = 0
