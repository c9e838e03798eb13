"""Golden vectors for empirical base frequencies (+F / +FC: compute_and_set_empirical_frequencies,
src/core/pll/optimize.cpp:457-472) and for +IC (which the reference leaves at 0): the unmodified reference
on the committed data sets. Run in the build container:
    python tests/golden/make_golden_freqs.py
"""
import json, os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
orc = ge.load_oracle()

CFG1_FC = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FC+G4{1.0}"
CFG1_F_IC = "GTR{1/2/1/1/2/1}+F+IC+G4{0.7}"
AA_F = "LG+F+G4{0.8}"


def run(t, s, q, model, extra=()):
    tmp = tempfile.mkdtemp(prefix="freqs_")
    pl, tree = orc.run_reference(t, s, q, model, tmp, threads=1, extra=extra)
    return {"model": model, "extra": list(extra), "placements": pl}


out = {}
d = os.path.join(HERE, "cfg1")
t, s, q = (os.path.join(d, f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
out["cfg1_fc_default"] = run(t, s, q, CFG1_FC)
out["cfg1_fc_noheur_all"] = run(t, s, q, CFG1_FC, ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "13"))
out["cfg1_f_ic_default"] = run(t, s, q, CFG1_F_IC)
d = os.path.join(HERE, "synthaa")
out["synthaa_f_default"] = run(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), AA_F)
# frequencies the reference prints for cfg1 (6 digits): "Base frequencies (empirical): ..."
out["cfg1_printed_freqs"] = [0.317154, 0.274246, 0.148892, 0.259707]
path = os.path.join(HERE, "cfg1", "reference_empirical.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: (len(v["placements"]) if isinstance(v, dict) else v) for k, v in out.items()})
