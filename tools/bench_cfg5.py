"""cfg5 of BASELINE.json on ONE GPU, bounded: 10k-taxon DNA tree (19 997 edges), full-edge thorough
evaluation of every query (--no-heur). Reports pairs/s and query-seqs/s for Q queries and compares
the first NREF queries with the unmodified reference (oracle/_ref/epa-ng --no-heur)."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
import torch
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
NREF = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
t0 = time.time()
ds = pkg.synth.dataset(T=T, n_sites=1000, n_queries=Q, window=200)
t_gen = time.time() - t0
t0 = time.time()
sess = pkg.session.Session(ds["newick"], ds["names"], ds["ref"], ds["model"], device=0)
t_setup = time.time() - t0
opts = pkg.capi.default_options(prescoring=0)
fmax = opts.filter_max
host_q = torch.from_numpy(ds["queries"]).pin_memory()
rec = torch.zeros((Q, fmax * 5), dtype=torch.float64).pin_memory()
cnt = torch.zeros(Q, dtype=torch.int32).pin_memory()
sess.place((host_q.data_ptr(), min(Q, 256)), opts, 0, out=rec.data_ptr(), counts=cnt.data_ptr())   # warm-up
torch.cuda.synchronize(); t0 = time.time()
sess.place((host_q.data_ptr(), Q), opts, 0, out=rec.data_ptr(), counts=cnt.data_ptr())
torch.cuda.synchronize(); dt = time.time() - t0
out = {"workload": "cfg5 (bounded): %d-taxon DNA tree, %d edges, %d queries, no heuristic, 1 GPU" % (T, sess.n_edges, Q),
       "dataset_s": t_gen, "setup_s": t_setup, "seconds": dt, "query_seqs_per_s": Q / dt, "pairs_per_s": Q * sess.n_edges / dt}
ref = os.path.join(ROOT, "oracle", "_ref", "epa-ng")
if NREF and os.path.exists(ref):
    tmp = tempfile.mkdtemp(prefix="cfg5_")
    tf, sf, _ = pkg.synth.write_dataset(dict(ds, queries=ds["queries"][:1], qnames=ds["qnames"][:1]), tmp)
    qf = os.path.join(tmp, "q.fasta")
    pkg.synth.write_fasta(qf, ds["qnames"][:NREF], ds["queries"][:NREF])
    t0 = time.time()
    subprocess.run([ref, "-t", tf, "-s", sf, "-q", qf, "-m", ds["model"], "-w", tmp, "-T", str(os.cpu_count()), "--redo", "--no-heur"],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out["reference_s_for_%d_queries" % NREF] = time.time() - t0
    doc = json.load(open(os.path.join(tmp, "epa_result.jplace")))
    want = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
    # our side on EXACTLY the same files (the column pre-mask depends on the query set, and with
    # per-rate scalers the reference's window offset makes the results depend on the mask)
    ours_dir = os.path.join(tmp, "ours"); os.makedirs(ours_dir, exist_ok=True)
    pkg.session.run_files(tf, sf, qf, ds["model"], ours_dir, opts)
    mine = {n: pq["p"] for pq in json.load(open(os.path.join(ours_dir, "epa_result.jplace")))["placements"] for n in pq["n"]}
    bad_edges = bad_vals = 0; worst = 0.0
    for name, w in want.items():
        g = mine[name]
        if [int(p[0]) for p in g] != [int(p[0]) for p in w]:
            bad_edges += 1
            if bad_edges <= 2: print("MISMATCH", name, "ours", g[:3], "ref", w[:3], file=sys.stderr)
            continue
        for a, p in zip(g, w):
            rel = abs(a[1] - p[1]) / abs(p[1]); worst = max(worst, rel)
            if rel > 1e-6 or abs(a[2] - p[2]) > 1e-6 or abs(a[3] - p[3]) > 1e-4 or abs(a[4] - p[4]) > 1e-4:
                bad_vals += 1; break
    out["parity_vs_reference"] = {"queries_compared": len(want), "edge_list_mismatches": bad_edges, "value_mismatches": bad_vals, "worst_logl_rel": worst}
print(json.dumps(out))
