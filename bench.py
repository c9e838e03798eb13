#!/usr/bin/env python
"""bench.py - query-seqs/s placed on the BASELINE.json cfg2 workload (1k-taxon DNA tree, GTR+G4,
1000-site MSA, 200-bp window queries, default heuristic), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--queries Q] [--impl ours|reference]

A step = one pass of the whole hot path (encode -> preplacement -> candidate selection ->
branch-length optimisation -> LWR/filter) over this rank's Q queries.
  value  device-resident: the query chunk already lives in HBM when the timed region starts
  e2e    through the host-layer C ABI (epa_session_place) from PINNED HOST buffers, H2D of the
         queries and D2H of the placement records inside the timed region
Timing: CUDA events on the stream the library launches on (the context is switched to torch's
current stream), barrier + synchronize on both sides, max over ranks. Inputs (1 GB of queries,
256 MB of lookup tables, 0.5 GB of CLVs per step) are far larger than the 126 MB L2.
`--impl reference` times the UNMODIFIED reference binary (oracle/_ref/epa-ng, all host threads) on
a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

T_TAXA, N_SITES, WINDOW = 1000, 1000, 200
METRIC = "query-seqs/sec placed (1k-taxon tree)"
HBM_FALLBACK_GBS = 6650.0

# BASELINE.json configs[1..4]. `queries` = per GPU (weak) or in total (strong); `default_queries` bounds the
# default run of the configurations whose full size takes minutes per step.
CONFIGS = {
    "cfg2": dict(T=1000, sites=1000, window=200, kind="dna", heur=True, scaling="weak", queries=1000000,
                 what="cfg2: 1k-taxon DNA tree (GTR+G4, 1000-site MSA), synthetic 200bp window queries, "
                      "preplacement heuristic -g 0.99999"),
    "cfg3": dict(T=1000, sites=1000, window=200, kind="dna", heur=True, scaling="strong", queries=10000000,
                 what="cfg3: 1k-taxon DNA tree (GTR+G4, 1000-site MSA), 10M synthetic 200bp window queries in total, "
                      "preplacement heuristic on, block-sharded over the GPUs (strong scaling)"),
    "cfg4": dict(T=512, sites=300, window=300, kind="aa", heur=True, scaling="weak", queries=100000,
                 what="cfg4: 512-taxon amino-acid tree (LG+G4, 300-site MSA), full-length queries, 20-state path, "
                      "preplacement heuristic on"),
    "cfg5": dict(T=10000, sites=1000, window=200, kind="dna", heur=False, scaling="strong", queries=1000000,
                 default_queries=2048,
                 what="cfg5: 10k-taxon DNA tree (19 997 edges, per-rate scalers as the reference's auto mode), 200bp window "
                      "queries, --no-heur (every edge evaluated thoroughly)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
#  reference CPU arm (also the cpu_baseline of our arm)
# ------------------------------------------------------------------------------------------------
def ref_binary():
    return os.path.join(ROOT, "oracle", "_ref", "epa-ng")


def run_reference_sample(ds, n_sample, threads, keep_jplace=False, extra=(), also=None):
    """Times the unmodified reference on the first n_sample queries. The fixed start-up cost
    (file parsing, reference CLV precompute) is removed by subtracting a 64-query run."""
    synth = ge.load_package().synth
    tmp = tempfile.mkdtemp(prefix="epa_ref_")
    try:
        tf, sf, _ = synth.write_dataset(dict(ds, queries=ds["queries"][:1], qnames=ds["qnames"][:1]), tmp)

        def run(nq, sub):
            qf = os.path.join(tmp, "q_full.fasta" if sub == "full" else f"q_{sub}.fasta")
            synth.write_fasta(qf, ds["qnames"][:nq], ds["queries"][:nq])
            out = os.path.join(tmp, sub)
            os.makedirs(out, exist_ok=True)
            cmd = [ref_binary(), "-t", tf, "-s", sf, "-q", qf, "-m", ds["model"], "-w", out, "-T", str(threads), "--redo", *extra]
            t0 = time.perf_counter()
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            return time.perf_counter() - t0, os.path.join(out, "epa_result.jplace")

        n_small = min(64, n_sample) if n_sample >= 256 else 1
        run(n_small, "small")                                # page the binary and the files in
        t_small, _ = run(n_small, "small")
        t_full, jp = run(n_sample, "full")
        if t_full <= t_small:
            t_small = 0.0
        placements = None
        if keep_jplace:
            doc = json.load(open(jp))
            placements = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
        dt = max(t_full - t_small, 1e-6)
        if also is not None:
            # our arm on EXACTLY the same files (the column pre-mask depends on the query set)
            placements = (placements, also(tf, sf, os.path.join(tmp, "q_full.fasta"), tmp))
        return (n_sample - n_small) / dt, t_full, t_small, placements
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_sample_size(name, cores):
    """Queries of the bounded CPU sample: about 10-30 s of the reference on `cores` host threads."""
    if name in ("cfg2", "cfg3"):
        return int(min(200000, max(4000, 1500 * cores)))
    if name == "cfg4":
        return int(min(20000, max(500, 100 * cores)))
    return max(4, cores // 2)            # cfg5: ~2 s per query and core


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(ref_binary()):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/epa-ng not built (oracle/Makefile.ref)"}))
        return
    synth = ge.load_package().synth
    cores = os.cpu_count() or 1
    cfg = CONFIGS[args.config]
    n_sample = args.ref_queries or ref_sample_size(args.config, cores)
    ds = synth.dataset(T=cfg["T"], n_sites=cfg["sites"], n_queries=n_sample, window=cfg["window"], kind=cfg["kind"])
    extra = () if cfg["heur"] else ("--no-heur",)
    if args.model:
        ds["model"] = args.model
    vals, times = [], []
    for it in range(args.warmup + args.steps):
        qps, t_full, t_small, _ = run_reference_sample(ds, n_sample, cores, extra=extra)
        if it >= args.warmup:
            vals.append(qps)
            times.append(t_full)
    v = float(np.median(vals))
    sample = (f"first {n_sample} of the {args.config} queries, oracle/_ref/epa-ng -T {cores}, wall clock of the whole run minus "
              f"a 64-query run (start-up removed); median of {len(vals)} runs")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "query-seqs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * float(np.median(times)),
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_sample, args.gpus, args.config, args.model),
        "cpu_baseline": {"value": v, "unit": "query-seqs/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "query-seqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ncu_traffic(kernel, q, chunk):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed ncu --set full summary
    (profiles/r1_ncu_<kernel>.txt, captured on a 131 072-query chunk of this workload); None when
    the launch of this run is not that size or the summary is missing."""
    if min(q, chunk) != 131072:
        return None
    path = ncu_profile(kernel)
    if not path:
        return None
    tot, scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(path):
        t = line.split()
        if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and t[2] in scale:
            tot += float(t[1]) * scale[t[2]]
    return tot or None


def ncu_profile(kernel):
    """Newest committed ncu --set full summary of `kernel` (profiles/r<N>_ncu_<kernel>.txt)."""
    for rnd in ("r2", "r1"):
        path = os.path.join(ROOT, "profiles", "%s_ncu_%s.txt" % (rnd, kernel.replace("_kernel", "")))
        if os.path.exists(path):
            return path
    return None


def ncu_fp64_per_pair(variant="blo_site"):
    """Executed fp64 warp instructions per (query, edge) pair of blo_site_kernel, from the committed capture's
    'fp64_warp_instructions_per_pair' line: profiles/r2_ncu_blo_site.txt (one eigenvalue group, the BASELINE model),
    ..._general_gtr.txt (three groups), ..._per_rate.txt (per-rate scalers, cfg5)."""
    path = ncu_profile(variant + "_kernel")
    if not path:
        return None
    for line in open(path):
        t = line.split()
        if len(t) >= 2 and t[0] == "fp64_warp_instructions_per_pair":
            return float(t[1])
    return None


def ncu_metric(kernel, metric):
    """One number of the committed ncu --set full summary of `kernel`, or None."""
    path = ncu_profile(kernel)
    if not path:
        return None
    for line in open(path):
        t = line.split()
        if len(t) >= 2 and t[0] == metric:
            try:
                return float(t[1])
            except ValueError:
                return None
    return None


def workload_config(q_per_gpu, gpus, name="cfg2", model=""):
    c = CONFIGS[name]
    out = {"workload": c["what"], "name": name, "model": model or ("LG+G4{0.8}" if c["kind"] == "aa" else "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}"),
           "taxa": c["T"], "edges": 2 * c["T"] - 3, "sites": c["sites"], "window": c["window"],
           "queries_per_gpu": q_per_gpu, "queries_total": q_per_gpu * gpus,
           "sharding": f"queries block-sharded over {gpus} GPU(s), reference state replicated",
           "l2": "inputs larger than L2 (no flush needed)"}
    if q_per_gpu * gpus != c["queries"] and c["scaling"] == "strong":
        out["bounded"] = (f"{q_per_gpu * gpus} of the configuration's {c['queries']} queries (throughput per query does not "
                          f"depend on the count; pass --queries {c['queries'] // max(1, gpus)} for the full size)")
    return out


# ------------------------------------------------------------------------------------------------
#  clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
#  files -> jplace through the C++ host pipeline (epa_run_files_multi)
# ------------------------------------------------------------------------------------------------
class quiet_stdout:
    """The host layer logs to stdout like the reference does; bench.py prints one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.devnull = os.open(os.devnull, os.O_WRONLY)
        self.saved = os.dup(1)
        os.dup2(self.devnull, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.devnull)


def head_placements(jplace_path, n_first):
    """name -> placements of the first n_first pqueries of a (large) jplace file, without parsing all of it."""
    want = n_first
    chunks, seen = [], 0
    with open(jplace_path, "r") as fh:
        while seen < want:
            block = fh.read(1 << 22)
            if not block:
                break
            chunks.append(block)
            seen += block.count('"n": [')
    text = "".join(chunks)
    start = text.index("[", text.index('"placements"'))
    pos, out = start + 1, {}
    for _ in range(min(want, seen)):
        a = text.find('{"p"', pos)
        b = text.find("}", a)
        if a < 0 or b < 0:
            break
        pq = json.loads(text[a:b + 1])
        for nm in pq["n"]:
            out[nm] = pq["p"]
        pos = b + 1
    return out


def files_leg(pkg, ds, n_files, devices, chunk, recs, counts, fmax):
    """Wall clock of files -> jplace (FASTA and bfast query files in a tmpfs directory) through
    epa_run_files_multi on `devices`, with a 64-query run of the same files as the start-up figure."""
    synth, session = pkg.synth, pkg.session
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="epa_files_", dir=base)
    try:
        tf, sf, _ = synth.write_dataset(dict(ds, queries=ds["queries"][:64], qnames=ds["qnames"][:64]), tmp)
        small = os.path.join(tmp, "query.fasta")
        qf = os.path.join(tmp, "q_full.fasta")
        synth.write_fasta(qf, ds["qnames"][:n_files], ds["queries"][:n_files])
        bf = session.fasta_to_bfast(qf, tmp)

        def run(qfile, sub):
            out = os.path.join(tmp, sub)
            t0 = time.perf_counter()
            st = session.run_files_multi(tf, sf, qfile, ds["model"], out, devices=devices, chunk_size=chunk,
                                         invocation="bench.py files leg")
            return time.perf_counter() - t0, st, os.path.join(out, "epa_result.jplace")

        with quiet_stdout():
            run(small, "warm")
            t_small, _, _ = run(small, "small")
            res = {}
            # the first full-size run page-locks the staging pool (kept for later runs in the process): reported as cold.
            # Session set-up and tear-down allocate and free several GB of device memory, which now and then stalls for
            # a few hundred ms: every kind is run three times, the median is the value, all three are listed.
            def measure(kind, qfile, repeats):
                runs = []
                for i in range(repeats):
                    t, st, jp = run(qfile, kind)
                    runs.append((t, st, jp))
                runs.sort(key=lambda x: x[0])
                t, st, jp = runs[len(runs) // 2]
                res[kind] = {"value": n_files / t, "unit": "query-seqs/s", "seconds": t,
                             "seconds_all_runs": [round(x[0], 4) for x in runs],
                             "value_startup_removed": (n_files - 64) / max(t - t_small, 1e-9),
                             "query_file_bytes": os.path.getsize(qfile), "jplace_bytes": os.path.getsize(jp), "stats": st}
                res[kind]["jplace"] = jp

            measure("fasta_cold", qf, 1)
            measure("fasta", qf, 3)
            measure("bfast", bf, 3)
        same = subprocess.run(["cmp", "-s", res["fasta"]["jplace"], res["bfast"]["jplace"]]).returncode == 0
        n_cmp = min(20000, n_files) if recs is not None else 0
        got = head_placements(res["fasta"]["jplace"], n_cmp) if n_cmp else {}
        bad = 0
        r = recs.reshape(len(counts), fmax, 5) if n_cmp else None
        for qi in range(n_cmp):
            nm = ds["qnames"][qi]
            c = int(counts[qi])
            want = [[int(np.float64(x[0]).view(np.uint64)), x[1], x[2], x[4], x[3]] for x in r[qi, :c]]
            g = got.get(nm)
            if g is None or len(g) != len(want) or any(
                    a[0] != b[0] or any(abs(a[k] - b[k]) > 5.1e-11 * max(1.0, abs(b[k])) for k in range(1, 5)) for a, b in zip(g, want)):
                bad += 1
        for kind in ("fasta_cold", "fasta", "bfast"):
            del res[kind]["jplace"]
        res.update({"queries": n_files, "devices": list(devices), "host_threads": os.cpu_count(),
                    "startup_seconds_64_queries": t_small,
                    "value": res["fasta"]["value"], "unit": "query-seqs/s", "fasta_and_bfast_jplace_identical": same,
                    "vs_device_records": {"queries_compared": n_cmp, "mismatches": bad,
                                          "note": "jplace text (10 decimals) against the records of the timed e2e step"},
                    "what": "epa_run_files_multi: tree + reference MSA + query file -> epa_result.jplace, wall clock of the "
                            "whole call (CUDA context, reference CLVs and lookup tables, memory-mapped query file indexed and "
                            "decoded by host threads, placement, jplace formatting and writing), files on tmpfs"})
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------
#  our arm
# ------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist

    pkg = ge.load_package()
    capi, synth, session = pkg.capi, pkg.synth, pkg.session
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")       # host-side barrier around the single-process files leg

    cfg = CONFIGS[args.config]
    if args.queries:
        Q = args.queries
    elif cfg["scaling"] == "strong":
        Q = -(-cfg.get("default_queries", cfg["queries"]) // world) if "default_queries" in cfg else -(-cfg["queries"] // world)
    else:
        Q = cfg["queries"]
    # unique synthetic queries per rank (at most 1M are generated; larger counts repeat them)
    Q_gen = min(Q, 1000000)
    ds = synth.dataset(T=cfg["T"], n_sites=cfg["sites"], n_queries=Q_gen, window=cfg["window"], kind=cfg["kind"], seed_q=2 + rank)
    if args.model:
        # same data, another model string (e.g. a general GTR: three distinct eigenvalues, the general kernel variant)
        ds["model"] = args.model
    if Q > Q_gen:
        reps = -(-Q // Q_gen)
        ds["queries"] = np.tile(ds["queries"], (reps, 1))[:Q]
        ds["qnames"] = ["q%08d" % i for i in range(Q)]
    sess = session.Session(ds["newick"], ds["names"], ds["ref"], ds["model"], device=local)
    ctx = sess.ctx
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    opts = capi.default_options(prescoring=1 if cfg["heur"] else 0)
    fmax = opts.filter_max
    chunk = args.chunk
    if not cfg["heur"]:
        chunk = int(min(chunk, max(1, (1 << 28) // sess.n_edges)))      # all-pairs mode: the session's own clamp
    B, n = sess.n_edges, sess.sites

    host_q = torch.from_numpy(ds["queries"]).pin_memory()
    dev_q = host_q.to(dev)
    rec_dev = torch.zeros((Q, fmax * 5), dtype=torch.float64, device=dev)       # 40-byte records
    cnt_dev = torch.zeros(Q, dtype=torch.int32, device=dev)
    rec_host = torch.zeros((Q, fmax * 5), dtype=torch.float64).pin_memory()
    cnt_host = torch.zeros(Q, dtype=torch.int32).pin_memory()

    # all ranks upload their queries at the same time: what the host gives each GPU when all of them pull
    h2d_gbs = None
    if world > 1:
        tmp_q = torch.empty_like(dev_q)
        tmp_q.copy_(host_q, non_blocking=True)
        torch.cuda.synchronize()
        dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            tmp_q.copy_(host_q, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        g = torch.tensor([3 * host_q.numel() / (c0.elapsed_time(c1) / 1e3) / 1e9], device=dev, dtype=torch.float64)
        gl = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(gl, g)
        h2d_gbs = [round(float(x.item()), 1) for x in gl]
        del tmp_q

    stage_ms = {"upload_encode": 0.0, "preplace": 0.0, "select": 0.0, "thorough": 0.0, "collect": 0.0}
    pairs_total = [0]

    # N > 1, device-resident step: rank 0 owns the records of all shards and every rank's collect kernel writes its block
    # straight into that buffer over NVLink (CUDA IPC peer memory, shard.PeerRecords) - the gather happens inside the
    # kernel; --nccl-gather (or a box without peer access) uses one NCCL gather of the fixed-stride records instead
    peer, peer_note = None, None
    if world > 1 and not args.nccl_gather:
        try:
            peer = pkg.shard.PeerRecords(capi, Q, fmax, local, dst=0, group=cpu_group)
        except Exception as ex:                  # noqa: BLE001 - any failure falls back to the NCCL gather
            peer_note = "peer memory unavailable (%s): NCCL gather" % str(ex)[:120]
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if peer is not None:
                peer.close()
            peer = None

    def step_resident(record=False, local_only=False):
        use_peer = peer is not None and not local_only
        for lo in range(0, Q, chunk):
            nq = min(chunk, Q - lo)
            ctx.encode_queries_dev(dev_q.data_ptr() + lo * n, nq, True)
            if cfg["heur"]:
                ctx.hint_selection(opts)        # dynamic heuristic: candidates are selected in the preplacement epilogue
                ctx.preplace()
            npairs = ctx.select(opts)
            ctx.place_pairs(opts)
            if use_peer:
                rp, cp = peer.slice_ptrs(lo)
                ctx.collect_dev(opts, rp, cp)
            else:
                ctx.collect_dev(opts, rec_dev.data_ptr() + lo * fmax * 40, cnt_dev.data_ptr() + lo * 4)
            if record:
                for k, v in ctx.timings().items():
                    stage_ms[k] += v
                pairs_total[0] += npairs
        if use_peer:
            peer.complete()                     # every rank's records are in rank 0's memory
        elif world > 1 and not local_only:
            # the single gather of placement records (NCCL over NVLink): rank 0 ends up with the
            # records of all shards in global query order
            pkg.shard.gather_records(rec_dev, cnt_dev, Q * world, dst=0)

    comp_all_host = cnt_all_host = None
    e2e_records = [0]
    if world > 1 and rank == 0:
        comp_all_host = torch.zeros((Q * world * fmax, 5), dtype=torch.float64).pin_memory()     # compacted records of all shards
        cnt_all_host = torch.zeros(Q * world, dtype=torch.int32).pin_memory()

    # N > 1, end to end: the gather of a step and rank 0's copy-out of all shards (0.4 GB over one PCIe link at N = 8)
    # run on a side stream beside the kernels of the next step; the device records alternate between two buffers
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    rec_alt = [rec_dev, torch.zeros_like(rec_dev)] if world > 1 else [rec_dev]
    cnt_alt = [cnt_dev, torch.zeros_like(cnt_dev)] if world > 1 else [cnt_dev]
    gathered = [None, None]                 # event: the gather of the step that used buffer k has read it
    e2e_no = [0]
    e2e_last = [0]

    def step_e2e():
        if world == 1:
            sess.place((host_q.data_ptr(), Q), opts, chunk, out=rec_host.data_ptr(), counts=cnt_host.data_ptr())
            return
        # N > 1: queries from pinned host memory (the copy of the next chunk overlaps this one), records
        # stay on the device, ONE gather brings every shard to rank 0 (NCCL), rank 0 copies all of them out
        k = e2e_no[0] & 1
        e2e_no[0] += 1
        rec_dev, cnt_dev = rec_alt[k], cnt_alt[k]
        if gathered[k] is not None:
            stream.wait_event(gathered[k])
        for lo in range(0, Q, chunk):
            nq = min(chunk, Q - lo)
            if lo + nq < Q:
                ctx.hint_next_chunk(host_q.data_ptr() + (lo + nq) * n, min(chunk, Q - lo - nq))
            ctx.upload_queries_ptr(host_q.data_ptr() + lo * n, nq, True)
            if cfg["heur"]:
                ctx.hint_selection(opts)
                ctx.preplace()
            ctx.select(opts)
            ctx.place_pairs(opts)
            ctx.collect_dev(opts, rec_dev.data_ptr() + lo * fmax * 40, cnt_dev.data_ptr() + lo * 4)
        # compacted: only the filled records travel (1.2 of 7 per query), NCCL gather, then rank 0's copy-out
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            parts, cnts = pkg.shard.gather_compact(rec_dev, cnt_dev, Q * world, dst=0)
            gathered[k] = torch.cuda.Event()
            gathered[k].record(side)
            if rank == 0:
                at = 0
                for p_ in parts:
                    comp_all_host[at:at + p_.shape[0]].copy_(p_, non_blocking=True)
                    p_.record_stream(side)
                    at += p_.shape[0]
                e2e_records[0] = at
                cnt_all_host.copy_(cnts, non_blocking=True)
                cnts.record_stream(side)
        e2e_last[0] = k

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, **kw):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn(**kw)
        if side is not None:
            stream.wait_stream(side)          # the last step's gather and copy-out belong to the timed region
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms_res = timed(step_resident, args.steps, record=True)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    peer_same = None
    if peer is not None:
        # the peer-written buffer against one NCCL gather of the same records
        step_resident(local_only=True)
        g_rec, g_cnt = pkg.shard.gather_records(rec_dev, cnt_dev, Q * world, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            peer_same = bool(torch.equal(g_rec, peer.records) and torch.equal(g_cnt, peer.counts))
        del g_rec, g_cnt
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # device-resident and host paths must agree bit for bit
    if world == 1:
        same = bool(np.array_equal(rec_host.numpy(), rec_dev.cpu().numpy())
                    and np.array_equal(cnt_host.numpy(), cnt_dev.cpu().numpy()))
    else:
        rec_host.copy_(rec_alt[e2e_last[0]])
        cnt_host.copy_(cnt_alt[e2e_last[0]])
        same = True
        if rank == 0:
            # rank 0's own shard comes first in the gathered, compacted records
            n0 = int(cnt_all_host[:Q].sum())
            mine = pkg.shard.expand_compact([comp_all_host[:n0]], cnt_all_host[:Q], fmax)
            filled = (torch.arange(fmax)[None, :] < cnt_host[:, None]).unsqueeze(-1)
            same = bool(torch.equal(mine.view(Q, fmax, 5), rec_host.view(Q, fmax, 5) * filled)
                        and torch.equal(cnt_all_host[:Q], cnt_host))

    # files -> jplace through the C++ pipeline, all GPUs of the job driven by rank 0's process
    files = None
    if not args.no_files and args.config in ("cfg2", "cfg3"):
        if cpu_group is not None:
            dist.barrier(group=cpu_group)
        if rank == 0:
            # 10^6 queries per GPU of the job, at most 4 x 10^6 (the file is written here, outside the timed part)
            n_files = int(args.files_queries or min(4000000, Q * world))
            fds = ds
            if n_files > len(ds["qnames"]):
                reps = -(-n_files // len(ds["qnames"]))
                fds = dict(ds, queries=np.tile(ds["queries"], (reps, 1))[:n_files],
                           qnames=["r%d%s" % (i // len(ds["qnames"]), ds["qnames"][i % len(ds["qnames"])]) for i in range(n_files)])
            files = files_leg(pkg, fds, min(n_files, len(fds["qnames"])), list(range(world)), chunk, rec_host.numpy(), cnt_host.numpy(), fmax)
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

    if rank == 0:
        total_q = Q * world
        value = total_q * args.steps / (ms_res / 1e3)
        e2e = total_q * args.steps / (ms_e2e / 1e3)
        peak, peak_src = peaks()
        pairs = pairs_total[0] / args.steps           # candidate pairs per step (this rank)
        w = cfg["window"]
        S = 4 if cfg["kind"] == "dna" else 20
        sr = 4 if (cfg["kind"] == "dna" and cfg["T"] > 2000) else 1      # per-rate scalers above 2000 tips (reference auto mode)
        per_unit = {
            "preplace": 9.0 * w,                                   # w lookup doubles + w query bytes per (query, edge)
            "thorough": 2.0 * w * 4 * S * 8 + 2.0 * w * 4 * sr + w + 40,   # two CLV windows + scalers + query + record
        }
        units = {"preplace": float(Q) * B if cfg["heur"] else 0.0, "thorough": pairs}
        kernels = {}
        for k in ("preplace", "thorough"):
            ms = stage_ms[k] / args.steps
            ach = units[k] * per_unit[k] / (ms / 1e3) / 1e9 if ms > 0 else 0.0
            kernels[k] = {"ms_per_step": ms, "units_per_step": units[k], "bytes_per_unit": per_unit[k],
                          "achieved_gbs": ach, "frac": ach / peak}
        ctx.build_lookup()                 # re-run warm: the first build pays module loading
        lk_ms = ctx.lookup_ms()
        K = 16 if S == 4 else 24
        lk_bytes = 2.0 * n * 4 * S * 8 + 2.0 * n * 4 * sr + n * K * 8            # two CLVs + scalers in, table out
        # DNA: a tip edge reads the tip's n state masks instead of a CLV (as the reference reads tipchars)
        lk_tip = lk_bytes - (n * 4 * S * 8 - n if cfg["kind"] == "dna" else 0)
        T_ = cfg["T"]
        lk_total = (B - T_) * lk_bytes + T_ * lk_tip
        kernels["lookup_build"] = {"ms": lk_ms, "units": B, "bytes_per_unit": lk_total / B,
                                   "bytes_inner_edge": lk_bytes, "bytes_tip_edge": lk_tip,
                                   "achieved_gbs": lk_total / (lk_ms / 1e3) / 1e9 if lk_ms > 0 else 0.0}
        kernels["lookup_build"]["frac"] = kernels["lookup_build"]["achieved_gbs"] / peak
        for k in ("upload_encode", "select", "collect"):
            kernels[k] = {"ms_per_step": stage_ms[k] / args.steps}
        dom = max(("preplace", "thorough"), key=lambda k: kernels[k]["ms_per_step"])
        kname = {"preplace": "preplace_mma_kernel" if S == 4 else "preplace_kernel",
                 "thorough": "blo_site_kernel" if S == 4 else "blo_generic_kernel"}[dom]
        # the preplacement kernel is an exact u8 x u8 -> s32 digit GEMM on the tensor cores: useful integer
        # operations = 2 * queries * (edges * 6 digits) * (window * 4 one-hot columns)
        pre_ms = kernels["preplace"]["ms_per_step"]
        if pre_ms > 0 and S == 4:
            kernels["preplace"]["tensor"] = {
                "kind": "tcgen05.mma kind::i8 (u8 x u8 -> s32), 128 x 192 x 32",
                "useful_tops": 2.0 * Q * B * 6 * (4 * w) / (pre_ms / 1e3) / 1e12,
                "note": "bf16 dense peak in MEASURED_PEAKS.json; the 8-bit integer rate is nominally twice that"}
        n_launch = max(1, (Q + chunk - 1) // chunk)
        roofline = {"kernel": kname, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": ncu_traffic(kname, Q, chunk), "peak_source": peak_src,
                    "launch_ms": kernels[dom]["ms_per_step"] / n_launch,
                    "algorithmic_bytes_per_launch": units[dom] * per_unit[dom] / n_launch}
        if dom == "thorough" and S == 4:
            # the branch-length optimisation re-reads its CLV windows from L1/L2 for every pass and spends its
            # time in fp64 arithmetic: the HBM fraction is the contract's figure, the fp64 pipe is what limits it.
            # fp64 roofline: measured DFMA peak of this GPU (epa_measure_fp64_peak) against the kernel's executed fp64
            # warp instructions per pair from the committed ncu capture (x 32 lanes x 2 flops).
            fp64_peak = capi.measure_fp64_peak(local)
            ev = sorted(session.parse_model(ds["model"])["eigenvals"])
            groups = 1 + sum(abs(a - b) > 1e-13 * abs(ev[0]) for a, b in zip(ev[:-1], ev[1:-1]))     # distinct non-zero eigenvalues
            variant = "blo_site_per_rate" if sr > 1 else {1: "blo_site", 3: "blo_site_general_gtr"}.get(groups)
            per_pair = ncu_fp64_per_pair(variant) if variant else None
            roofline["fp64"] = {"peak_tflops_measured": fp64_peak, "fp64_warp_instructions_per_pair_ncu": per_pair,
                                "kernel_variant": variant}
            if per_pair:
                ach = pairs * per_pair * 64.0 / (kernels[dom]["ms_per_step"] / 1e3) / 1e12
                roofline["fp64"].update({"achieved_tflops": ach, "frac": ach / fp64_peak})
            roofline["fp64_pipe_active_pct_ncu"] = ncu_metric(kname, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
            roofline["note"] = ("fp64-issue bound (Newton-Raphson derivative sums and CLV passes, DESIGN 3.2): fp64 figures "
                                "next to the HBM figure")

        cpu = None
        parity = None
        if world == 1 and not args.no_cpu and os.path.exists(ref_binary()):
            cores = os.cpu_count() or 1
            n_sample = int(min(Q, args.ref_queries or (min(100000, max(2000, 1000 * cores)) if args.config in ("cfg2", "cfg3")
                                                        else ref_sample_size(args.config, cores))))
            via_files = sr > 1      # per-rate scalers: the reference's scaler window offset depends on the column pre-mask

            def ours_on_files(tf, sf, qf, tmp):
                out_dir = os.path.join(tmp, "ours")
                with quiet_stdout():
                    session.run_files(tf, sf, qf, ds["model"], out_dir, opts)
                doc = json.load(open(os.path.join(out_dir, "epa_result.jplace")))
                return {nm: pq["p"] for pq in doc["placements"] for nm in pq["n"]}

            qps, t_full, t_small, ref_pl = run_reference_sample(ds, n_sample, cores, keep_jplace=True,
                                                                extra=() if cfg["heur"] else ("--no-heur",),
                                                                also=ours_on_files if via_files else None)
            cpu = {"value": qps, "unit": "query-seqs/s", "cores": cores, "kind": "reference",
                   "sample": f"first {n_sample} of this run's queries, oracle/_ref/epa-ng -T {cores}: {t_full:.2f}s wall "
                             f"minus {t_small:.2f}s for a short run of the same files (start-up removed)"}
            if via_files:
                parity = compare_placement_dicts(ref_pl[1], ref_pl[0])
                parity["note"] = "both arms on the same files (epa_run_files): the pre-mask depends on the query set"
            else:
                parity = compare_with_reference(rec_host.numpy(), cnt_host.numpy(), ds["qnames"], ref_pl, fmax)

        out = {
            "metric": METRIC, "value": value, "unit": "query-seqs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(Q, world, args.config, args.model),
            "e2e": {"value": e2e, "unit": "query-seqs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(Q) * n,
                    "d2h_bytes_per_step": int(Q) * (fmax * 40 + 4) if world == 1 else int(e2e_records[0]) * 40 + int(Q) * world * 4,
                    "note": None if world == 1 else "per rank: upload from pinned host memory, records stay on the device; one NCCL gather of "
                                                    "the compacted records + counts to rank 0, which copies all shards out (d2h bytes = rank 0's); "
                                                    "gather and copy-out of a step run on a side stream beside the next step's kernels, the last "
                                                    "step's are waited for inside the timed region"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "candidate_pairs_per_query": pairs / Q, "chunk": chunk, "resident_equals_e2e": same,
        }
        if world > 1:
            out["host"] = {"h2d_gbs_per_rank_all_ranks_uploading": h2d_gbs}
            out["gather"] = ({"kind": "peer memory", "what": "rank 0 owns the records of all shards (CUDA IPC); every rank's collect kernel writes "
                              "its block straight into it over NVLink, one 4-byte all-reduce completes the step",
                              "equals_nccl_gather": peer_same} if peer is not None
                             else {"kind": "nccl", "what": "one NCCL gather of the fixed-stride records + one of the counts", "note": peer_note})
        if files:
            out["e2e_files"] = files
        if cpu:
            out["cpu_baseline"] = cpu
        if parity:
            out["parity_vs_reference"] = parity
        print(json.dumps(out))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if peer is not None and rank != 0:
            peer.close()                        # the importers unmap ...
        dist.barrier()
        if peer is not None and rank == 0:
            peer.close()                        # ... before the owner frees
        dist.destroy_process_group()
    sess.close()


def compare_placement_dicts(got, want):
    n_edges = n_bad = 0
    worst = 0.0
    for name, w in want.items():
        g = got.get(name, [])
        if [int(p[0]) for p in g] != [int(p[0]) for p in w]:
            n_edges += 1
            continue
        for a, p in zip(g, w):
            rel = abs(a[1] - p[1]) / abs(p[1])
            worst = max(worst, rel)
            if rel > 1e-6 or abs(a[2] - p[2]) > 1e-6 or abs(a[3] - p[3]) > 1e-4 or abs(a[4] - p[4]) > 1e-4:
                n_bad += 1
                break
    return {"queries_compared": len(want), "edge_list_mismatches": n_edges, "value_mismatches": n_bad, "worst_logl_rel": worst}


def compare_with_reference(recs, counts, names, ref_pl, fmax):
    """Name-keyed comparison with the reference's jplace: identical ordered edge lists, logl rel
    1e-6, LWR abs 1e-6, lengths abs 1e-4 (the north_star tolerances)."""
    recs = recs.reshape(len(counts), fmax, 5)
    n_cmp = n_edges = n_bad = 0
    worst = 0.0
    for qi, name in enumerate(names):
        if name not in ref_pl:
            continue
        want = ref_pl[name]
        c = int(counts[qi])
        got = recs[qi, :c]
        n_cmp += 1
        edges = [int(np.float64(x).view(np.uint64)) for x in got[:, 0]]
        if edges != [int(p[0]) for p in want]:
            n_edges += 1
            continue
        for g, p in zip(got, want):
            rel = abs(g[1] - p[1]) / abs(p[1])
            worst = max(worst, rel)
            if rel > 1e-6 or abs(g[2] - p[2]) > 1e-6 or abs(g[4] - p[3]) > 1e-4 or abs(g[3] - p[4]) > 1e-4:
                n_bad += 1
                break
    return {"queries_compared": n_cmp, "edge_list_mismatches": n_edges, "value_mismatches": n_bad,
            "worst_logl_rel": worst}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configuration [cfg2]")
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step [the configuration's size]")
    ap.add_argument("--chunk", type=int, default=131072)
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather the records of the resident step with NCCL instead of peer-memory writes")
    ap.add_argument("--ref-queries", type=int, default=0, help="size of the CPU reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--model", default="", help="place under this model string instead of the configuration's")
    ap.add_argument("--no-files", action="store_true", help="skip the files -> jplace leg")
    ap.add_argument("--files-queries", type=int, default=0, help="queries of the files -> jplace leg [min(Q, 2M)]")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
