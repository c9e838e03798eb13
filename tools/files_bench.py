#!/usr/bin/env python
"""Files -> jplace leg of bench.py on its own:  python tools/files_bench.py [queries] [n_devices] [repeat]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = bench.ge.load_package()
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ds = pkg.synth.dataset(T=bench.T_TAXA, n_sites=bench.N_SITES, n_queries=Q, window=bench.WINDOW)
for _ in range(rep):
    r = bench.files_leg(pkg, ds, Q, list(range(nd)), 131072, None, None, 7)
    print(json.dumps({k: (dict(v, stats={a: round(b, 3) for a, b in v["stats"].items()}) if isinstance(v, dict) and "stats" in v else v)
                      for k, v in r.items() if k != "what"}))
