"""Thorough-kernel throughput for longer query windows (developer tool): python tools/bench_window.py W [Q]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
import torch
W = int(sys.argv[1]); Q = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
ds = pkg.synth.dataset(T=1000, n_sites=1000, n_queries=Q, window=W)
sess = pkg.session.Session(ds["newick"], ds["names"], ds["ref"], ds["model"], device=0)
ctx = sess.ctx; opts = pkg.capi.default_options()
q = torch.from_numpy(ds["queries"]).cuda()
rec = torch.zeros((Q, opts.filter_max * 5), dtype=torch.float64, device="cuda"); cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
for it in range(2):
    ctx.encode_queries_dev(q.data_ptr(), Q, True); ctx.preplace(); npairs = ctx.select(opts); ctx.place_pairs(opts)
    ctx.collect_dev(opts, rec.data_ptr(), cnt.data_ptr())
t = ctx.timings()
print(json.dumps({"window": W, "queries": Q, "pairs": npairs, "stage_ms": t, "pairs_per_s": npairs / (t["thorough"] / 1e3)}))
