"""Golden vectors for free-rate models (+R[n]{rates}{weights}, src/core/raxml/Model.cpp:405-455): the unmodified
reference on cfg1. Run in the build container:   python tests/golden/make_golden_freerates.py
"""
import json, os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
orc = ge.load_oracle()
B = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}"
MODELS = {"r4_default": B + "+R4", "r4_user": B + "+R4{0.1/0.5/1.2/3.0}{0.4/0.3/0.2/0.1}", "r2_rates_only": B + "+R2{0.3/2.0}",
          "r4_pinv": B + "+IU{0.1}+R4{0.2/0.6/1.0/2.5}{1/2/2/1}"}
d = os.path.join(HERE, "cfg1")
t, s, q = (os.path.join(d, f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
out = {}
for key, model in MODELS.items():
    tmp = tempfile.mkdtemp(prefix="freerates_")
    pl, _ = orc.run_reference(t, s, q, model, tmp, threads=1, extra=("--no-heur", "--filter-min-lwr", "0", "--filter-max", "13"))
    out[key] = {"model": model, "extra": ["--no-heur", "--filter-min-lwr", "0", "--filter-max", "13"], "placements": pl}
path = os.path.join(d, "reference_freerates.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, len(out))
