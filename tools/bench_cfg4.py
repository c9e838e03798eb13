"""cfg4 of BASELINE.json on one GPU: 512-taxon AA tree (LG+G4, 300-site MSA), 100k full-length queries.
Prints stage times and query-seqs/s (device-resident and from pinned host memory)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
import torch
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ds = pkg.synth.dataset(T=512, n_sites=300, n_queries=Q, window=300, kind="aa")
t0 = time.time()
sess = pkg.session.Session(ds["newick"], ds["names"], ds["ref"], ds["model"], device=0)
t_setup = time.time() - t0
ctx = sess.ctx
opts = pkg.capi.default_options()
fmax = opts.filter_max
host_q = torch.from_numpy(ds["queries"]).pin_memory()
dev_q = host_q.cuda()
rec = torch.zeros((Q, fmax * 5), dtype=torch.float64).pin_memory()
cnt = torch.zeros(Q, dtype=torch.int32).pin_memory()
chunk = 32768
def resident():
    tot = {}
    pairs = 0
    for lo in range(0, Q, chunk):
        nq = min(chunk, Q - lo)
        ctx.encode_queries_dev(dev_q.data_ptr() + lo * sess.sites, nq, True)
        ctx.preplace(); pairs += ctx.select(opts); ctx.place_pairs(opts)
        r = torch.zeros((nq, fmax * 5), dtype=torch.float64, device="cuda"); c = torch.zeros(nq, dtype=torch.int32, device="cuda")
        ctx.collect_dev(opts, r.data_ptr(), c.data_ptr())
        for k, v in ctx.timings().items(): tot[k] = tot.get(k, 0.0) + v
    return tot, pairs
resident()
torch.cuda.synchronize(); t0 = time.time(); tot, pairs = resident(); torch.cuda.synchronize(); dt = time.time() - t0
sess.place((host_q.data_ptr(), Q), opts, chunk, out=rec.data_ptr(), counts=cnt.data_ptr())
torch.cuda.synchronize(); t0 = time.time()
sess.place((host_q.data_ptr(), Q), opts, chunk, out=rec.data_ptr(), counts=cnt.data_ptr())
torch.cuda.synchronize(); dt2 = time.time() - t0
print(json.dumps({"workload": "cfg4: 512-taxon AA tree (LG+G4, 300 sites), %d queries" % Q, "edges": sess.n_edges,
                  "setup_s": t_setup, "resident_qps": Q / dt, "e2e_qps": Q / dt2, "stage_ms": tot,
                  "pairs_per_query": pairs / Q, "mean_placements": float(cnt.numpy().mean())}))
