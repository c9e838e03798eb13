"""Golden vectors for every empirical amino-acid matrix of the reference (PM/util/models_aa.c:28-57): the
unmodified reference on the first 6 queries of the synthaa fixture, model <NAME>+G4{0.8}. Run in the build
container:   python tests/golden/make_golden_aa_models.py
"""
import json, os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
orc = ge.load_oracle()
pkg = ge.load_package()

names = sorted(json.load(open(os.path.join(ROOT, "oracle", "protein_models.json"))).keys())
d = os.path.join(HERE, "synthaa")
qn, qs = orc.read_fasta(os.path.join(d, "query.fasta"))
tmp = tempfile.mkdtemp(prefix="aamodels_")
q6 = os.path.join(d, "query6.fasta")          # committed: the column mask depends on the query set
with open(q6, "w") as fh:
    for n, s in list(zip(qn, qs))[:6]:
        fh.write(">%s\n%s\n" % (n, s))
out = {}
for name in names:
    model = name + "+G4{0.8}"
    pl, _ = orc.run_reference(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), q6, model, os.path.join(tmp, name), threads=1)
    out[name] = {"model": model, "placements": pl}
path = os.path.join(d, "reference_models.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, len(out), "models")
