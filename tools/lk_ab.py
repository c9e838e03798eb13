import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
import helpers
ds = pkg.synth.dataset(T=1000, n_sites=1000, n_queries=64, window=200)
s = pkg.session.Session(ds["newick"], ds["names"], ds["ref"], ds["model"])
ctx = s.ctx
ts = []
for i in range(6):
    ctx.build_lookup(); ts.append(ctx.lookup_ms())
print("lookup ms", ts)
lk = np.stack([ctx.get_lookup(e) for e in range(0, 1997, 50)])
np.save("/root/repo/gpurun_out/lk_%s.npy" % ("clv" if os.environ.get("EPA_B200_LOOKUP_TIP_CLV") else "mask"), lk)
