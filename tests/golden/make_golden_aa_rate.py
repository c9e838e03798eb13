"""Golden vectors for amino acids under per-rate scalers (the reference's --rate-scalers on, and its automatic
choice above 2000 tips): the unmodified reference (oracle/_ref/epa-ng --rate-scalers on) on a seeded 300-taxon
caterpillar-like LG+G4 data set (120 sites, 16 queries of 80 residues).

With PLL_ATTRIB_RATE_SCALERS the reference runs libpll's generic kernels, and the generic tip-inner CLV update
(libpll core_partials.c:461-506) tests and rescales WHOLE sites and counts the rescaling in entry [site index] of
the [site][rate] counter array. On a deep ladder most updates are tip-inner, so this data set exercises that
behaviour in the reference tree's CLVs, in the lookup tables and inside the tiny trees of the thorough phase.
Run in the build container:
    python tests/golden/make_golden_aa_rate.py
"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle()

SPEC = dict(T=300, n_sites=120, n_queries=16, window=80, seed_tree=21, seed_q=22, kind="aa", ladder=0.95, brlen=0.1)

ds = pkg.synth.dataset(**SPEC)
tmp = tempfile.mkdtemp(prefix="aa_rate_")
tf, sf, qf = pkg.synth.write_dataset(ds, tmp)
out = {"dataset": SPEC, "model": ds["model"], "flags": "--rate-scalers on"}
out["placements"], _ = orc.run_reference(tf, sf, qf, ds["model"], os.path.join(tmp, "ref"), threads=4, extra=("--rate-scalers", "on"))
out["placements_no_heur"], _ = orc.run_reference(tf, sf, qf, ds["model"], os.path.join(tmp, "ref2"), threads=4,
                                                 extra=("--rate-scalers", "on", "--no-heur"))
out["placements_raxml_blo"], _ = orc.run_reference(tf, sf, qf, ds["model"], os.path.join(tmp, "ref3"), threads=4,
                                                   extra=("--rate-scalers", "on", "--raxml-blo"))
out["model_pinv"] = "LG+G4{0.8}+IU{0.2}"
out["placements_pinv"], _ = orc.run_reference(tf, sf, qf, out["model_pinv"], os.path.join(tmp, "ref4"), threads=4,
                                              extra=("--rate-scalers", "on"))
path = os.path.join(ROOT, "tests", "golden", "rate300", "reference_placements_aa_ladder.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, len(out["placements"]), len(out["placements_no_heur"]))
