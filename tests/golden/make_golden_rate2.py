"""Golden vectors for per-rate scalers beyond DNA with <= 4 rate categories: the unmodified reference
(oracle/_ref/epa-ng --rate-scalers on) on
  dna8   the seeded 300-taxon DNA data set of make_golden_rate.py under GTR+G8
  aa     a seeded 300-taxon amino-acid data set (LG+G4{0.8}, 120 sites, 16 queries of 80 residues)
whose CLVs do get rescaled. Run in the build container:
    python tests/golden/make_golden_rate2.py
"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle()

DNA8 = dict(T=300, n_sites=400, n_queries=24, window=120, seed_tree=7, seed_q=8)
DNA8_MODEL = "GTR{1/2/1/1/2/1}+FU{0.3/0.2/0.2/0.3}+G8{0.5}"
AA = dict(T=300, n_sites=120, n_queries=16, window=80, seed_tree=11, seed_q=12, kind="aa")

out = {}
for key, spec, model in (("dna8", DNA8, DNA8_MODEL), ("aa", AA, None)):
    ds = pkg.synth.dataset(**spec)
    model = model or ds["model"]
    tmp = tempfile.mkdtemp(prefix="rate2_")
    tf, sf, qf = pkg.synth.write_dataset(ds, tmp)
    ref, _ = orc.run_reference(tf, sf, qf, model, os.path.join(tmp, "ref"), threads=4, extra=("--rate-scalers", "on"))
    out[key] = {"dataset": spec, "model": model, "flags": "--rate-scalers on", "placements": ref}
out["aa"]["note"] = ("the reference's generic tip-inner CLV update (libpll core_partials.c:461-506) rescales whole sites and bumps entry "
                     "[site index] of the [site][rate] counter array under per-rate scalers; the oracle restates that and "
                     "tests/test_oracle_rate_scalers.py pins it on these placements")
path = os.path.join(ROOT, "tests", "golden", "rate300", "reference_placements_rate2.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: len(v["placements"]) for k, v in out.items()})
