"""Golden vectors for per-rate scalers: the unmodified reference (oracle/_ref/epa-ng --rate-scalers on)
on a seeded 300-taxon synthetic data set whose CLVs do get rescaled. Run in the build container:
    python tests/golden/make_golden_rate.py
"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle()
ds = pkg.synth.dataset(T=300, n_sites=400, n_queries=24, window=120, seed_tree=7, seed_q=8)
tmp = tempfile.mkdtemp(prefix="rate300_")
tf, sf, qf = pkg.synth.write_dataset(ds, tmp)
ref, _ = orc.run_reference(tf, sf, qf, ds["model"], os.path.join(tmp, "ref"), threads=4, extra=("--rate-scalers", "on"))
out = os.path.join(ROOT, "tests", "golden", "rate300", "reference_placements.json")
json.dump({"dataset": dict(T=300, n_sites=400, n_queries=24, window=120, seed_tree=7, seed_q=8), "model": ds["model"],
           "flags": "--rate-scalers on", "placements": ref}, open(out, "w"), indent=0)
print("wrote", out, len(ref), "queries")
