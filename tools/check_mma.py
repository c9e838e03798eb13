"""GPU check: tensor-core preplacement vs the shared-memory kernels vs the oracle (prescores)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
pkg = helpers.pkg()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
ds = pkg.synth.dataset(T=T, n_sites=1000, n_queries=NQ, window=200)
q = ds["queries"].copy()
# interior gaps / N in some queries exercise the fully-ambiguous column
rng = np.random.default_rng(5)
for i in range(0, NQ, 7):
    cols = np.flatnonzero(q[i] != ord("-"))
    pick = rng.choice(cols[1:-1], size=5, replace=False)
    q[i, pick[:3]] = ord("-"); q[i, pick[3:]] = ord("N")
case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], q, ds["model"])
seqs = case.query_rows
def run():
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    ctx.upload_queries(seqs)
    t0 = time.time(); ctx.preplace(); P = ctx.get_prescores(); dt = time.time() - t0
    tm = ctx.timings(); ctx.close()
    return P, tm
P1, t1 = run()
os.environ["EPA_B200_NO_MMA"] = "1"
P2, t2 = run()
print("timings mma", t1, "smem", t2)
d = np.abs(P1 - P2)
print("max abs diff mma vs smem: %.3e  max rel %.3e  (|pre| range %.1f..%.1f)" % (d.max(), (d / np.abs(P2)).max(), np.abs(P2).min(), np.abs(P2).max()))
bad = np.argwhere(d > 1e-8)
print("entries off by > 1e-8:", len(bad), bad[:10].tolist())
for qi in (0, 1, 7, NQ - 1):
    want = case.placer.preplace(case.qseqs[qi])
    print("query %d: oracle vs mma max abs %.3e, vs smem %.3e" % (qi, np.abs(want - P1[qi]).max(), np.abs(want - P2[qi]).max()))
