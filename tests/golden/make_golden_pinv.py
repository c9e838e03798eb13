"""Golden vectors for +I models (proportion of invariant sites): the unmodified reference
(oracle/_ref/epa-ng) with +IU{p} model strings on the committed data sets. Run in the build
container:
    python tests/golden/make_golden_pinv.py
cfg1      test/data of the reference, GTR+IU{0.2}+G4: default options and --no-heur unfiltered
synth64   the 64-taxon DNA fixture, +IU{0.15}: default options (200 queries)
synthaa   the 32-taxon amino-acid fixture, LG+IU{0.1}+G4{0.8}: default options
rate300   the seeded 300-taxon data set of make_golden_rate.py (CLVs get rescaled), +IU{0.1}, with
          per-site scalers (--rate-scalers off) and per-rate scalers (on)
"""
import json, os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle()

CFG1_PINV = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+IU{0.2}+G4{1.0}"
SYNTH64_PINV = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+IU{0.15}+G4{0.5}"
SYNTHAA_PINV = "LG+IU{0.1}+G4{0.8}"
RATE300_PINV = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+IU{0.1}+G4{0.5}"


def run(t, s, q, model, extra=(), threads=1):
    tmp = tempfile.mkdtemp(prefix="pinv_")
    pl, tree = orc.run_reference(t, s, q, model, tmp, threads=threads, extra=extra)
    return {"model": model, "extra": list(extra), "placements": pl}


out = {}
d = os.path.join(HERE, "cfg1")
t, s, q = (os.path.join(d, f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
out["cfg1_default"] = run(t, s, q, CFG1_PINV)
out["cfg1_noheur_all"] = run(t, s, q, CFG1_PINV, ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "13"))
d = os.path.join(HERE, "synth64")
out["synth64_default"] = run(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), SYNTH64_PINV)
d = os.path.join(HERE, "synthaa")
out["synthaa_default"] = run(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), SYNTHAA_PINV)
ds = pkg.synth.dataset(T=300, n_sites=400, n_queries=24, window=120, seed_tree=7, seed_q=8)
tmp = tempfile.mkdtemp(prefix="rate300_")
tf, sf, qf = pkg.synth.write_dataset(ds, tmp)
out["rate300_site"] = run(tf, sf, qf, RATE300_PINV, ("--rate-scalers", "off"), threads=4)
out["rate300_rate"] = run(tf, sf, qf, RATE300_PINV, ("--rate-scalers", "on"), threads=4)
path = os.path.join(HERE, "pinv", "reference_placements.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: len(v["placements"]) for k, v in out.items()})
