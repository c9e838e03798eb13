"""Golden vectors for the named DNA models of the reference (PM/util/models_dna.c:40-125): the unmodified reference
on cfg1, each model once with its ML-mode default rates (0.5 ... 1.0 over the six rates, Model.cpp:484-490) and once
with user rates for its symmetry classes. Run in the build container:
    python tests/golden/make_golden_dna_models.py
"""
import json, os, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
orc = ge.load_oracle()
NUNIQ = {"JC": 1, "K80": 2, "F81": 1, "HKY": 2, "TN93ef": 3, "TN93": 3, "K81": 3, "K81uf": 3, "TPM2": 3, "TPM2uf": 3,
         "TPM3": 3, "TPM3uf": 3, "TIM1": 4, "TIM1uf": 4, "TIM2": 4, "TIM2uf": 4, "TIM3": 4, "TIM3uf": 4, "TVMef": 5,
         "TVM": 5, "SYM": 6, "GTR": 6, "TrN": 3, "TPM1": 3, "TIM2ef": 4}
VALS = [0.7, 2.9, 1.3, 0.6, 3.4, 1.0]
FREQ = "+FU{0.31/0.19/0.22/0.28}"
d = os.path.join(HERE, "cfg1")
t, s, q = (os.path.join(d, f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
out = {}
for name, k in NUNIQ.items():
    equal_f = name in ("JC", "K80", "SYM") or name.endswith("ef") or name in ("K81", "TPM1", "TPM2", "TPM3", "TIM1", "TIM2", "TIM3")
    f = "" if equal_f else FREQ
    models = {name + "_default": name + f + "+G4{0.9}"}
    if k > 1:
        models[name + "_user"] = name + "{" + "/".join(str(v) for v in VALS[:k - 1] + [1.0]) + "}" + f + "+G4{0.9}"
    for key, model in models.items():
        tmp = tempfile.mkdtemp(prefix="dnamodels_")
        pl, _ = orc.run_reference(t, s, q, model, tmp, threads=1)
        out[key] = {"model": model, "placements": pl}
path = os.path.join(d, "reference_dna_models.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, len(out))
