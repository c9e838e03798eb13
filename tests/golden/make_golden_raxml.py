"""Golden vectors for --raxml-blo (branch-length optimisation "the way RAxML-EPA did it":
optimize_branch_triplet with sliding == false -> pllmod_opt_optimize_branch_lengths_local): the
unmodified reference (oracle/_ref/epa-ng --raxml-blo) on the committed data sets. Run in the build
container:
    python tests/golden/make_golden_raxml.py
cfg1      default options and --no-heur unfiltered (GTR+G4)
cfg1_pinv default options with the +IU{0.2} model
synth64   default options (200 window queries)
synthaa   default options (LG+G4{0.8})
rate300   the seeded 300-taxon data set of make_golden_rate.py, per-site and per-rate scalers
"""
import json, os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle()

GTRG = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+G4{1.0}"
CFG1_PINV = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+IU{0.2}+G4{1.0}"
SYNTH = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}"


def run(t, s, q, model, extra=(), threads=1):
    tmp = tempfile.mkdtemp(prefix="raxml_")
    pl, tree = orc.run_reference(t, s, q, model, tmp, threads=threads, extra=("--raxml-blo",) + tuple(extra))
    return {"model": model, "extra": ["--raxml-blo"] + list(extra), "placements": pl}


out = {}
d = os.path.join(HERE, "cfg1")
t, s, q = (os.path.join(d, f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
out["cfg1_default"] = run(t, s, q, GTRG)
out["cfg1_noheur_all"] = run(t, s, q, GTRG, ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "13"))
out["cfg1_pinv_default"] = run(t, s, q, CFG1_PINV)
d = os.path.join(HERE, "synth64")
out["synth64_default"] = run(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), SYNTH)
d = os.path.join(HERE, "synthaa")
out["synthaa_default"] = run(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), "LG+G4{0.8}")
ds = pkg.synth.dataset(T=300, n_sites=400, n_queries=24, window=120, seed_tree=7, seed_q=8)
tmp = tempfile.mkdtemp(prefix="rate300_")
tf, sf, qf = pkg.synth.write_dataset(ds, tmp)
out["rate300_site"] = run(tf, sf, qf, ds["model"], ("--rate-scalers", "off"), threads=4)
out["rate300_rate"] = run(tf, sf, qf, ds["model"], ("--rate-scalers", "on"), threads=4)
path = os.path.join(HERE, "raxml_blo", "reference_placements.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: len(v["placements"]) for k, v in out.items()})
