"""Generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/epa-ng and oracle/_ref/libpllref.so, built by oracle/Makefile.ref) in THIS container.

    python tests/golden/make_golden.py

cfg1/      the reference's own test data (test/data/{ref.tre,aln.fasta,query.fasta}; data files,
           not sources) and the reference's placements on them for two model strings and three
           option sets (default heuristic, --no-heur unfiltered, heuristic unfiltered).
synthaa/   a 32-taxon synthetic amino-acid data set (LG+G4) with 60 queries, a few ambiguity codes.
synth64/   a 64-taxon synthetic DNA data set (epa-ng_b200/synth.py, seeds fixed) with the
           reference's placements of 200 window queries, default options and --no-heur for
           the first 5 queries.
Every jplace is reduced to {name: [[edge, logl, lwr, distal, pendant], ...]} + the tree string.
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pyoracle  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

REF_DATA = "/root/reference/test/data"
GTR_B = ("GTR{0.676278/2.012275/0.478487/0.753965/2.406436/1.0}"
            "+FU{0.245629/0.235012/0.253054/0.266305}+G4{1.078763}")


def run(tree, msa, query, model, extra):
    out = tempfile.mkdtemp(prefix="golden_")
    try:
        pl, tree_str = pyoracle.run_reference(tree, msa, query, model, out, threads=1, extra=extra)
    finally:
        shutil.rmtree(out, ignore_errors=True)
    return {"model": model, "extra": list(extra), "tree": tree_str, "placements": pl}


def main():
    cfg1 = os.path.join(HERE, "cfg1")
    os.makedirs(cfg1, exist_ok=True)
    for f in ("ref.tre", "aln.fasta", "query.fasta"):
        shutil.copy(os.path.join(REF_DATA, f), os.path.join(cfg1, f))
    t, s, q = (os.path.join(cfg1, f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
    runs = {}
    for mname, model in (("gtrg", "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+G4{1.0}"), ("gtrb", GTR_B)):
        runs[mname + "_default"] = run(t, s, q, model, ())
        runs[mname + "_noheur_all"] = run(t, s, q, model, ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "13"))
        runs[mname + "_heur_all"] = run(t, s, q, model, ("--filter-min-lwr", "0", "--filter-max", "13"))
        runs[mname + "_acc"] = run(t, s, q, model, ("--filter-acc-lwr", "0.999", "--filter-max", "5"))
    json.dump(runs, open(os.path.join(cfg1, "reference_placements.json"), "w"), indent=1)

    # rooted input trees of the reference's test data (6 of the 8 taxa): placements and tree string
    # on the ROOTED tree (default --preserve-rooting on) and on the unrooted one (off)
    rooted = {}
    for f in ("ref_rooted.tre", "ref_rooted_2.tre", "ref_rooted_3.tre", "ref_rooted_innerlabels.tre"):
        shutil.copy(os.path.join(REF_DATA, f), os.path.join(cfg1, f))
        tr = os.path.join(cfg1, f)
        model = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+G4{1.0}"
        rooted[f] = {"default": run(tr, s, q, model, ()),
                     "noheur_all": run(tr, s, q, model, ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "10")),
                     "unrooted": run(tr, s, q, model, ("--preserve-rooting", "off"))}
    json.dump(rooted, open(os.path.join(cfg1, "reference_rooted.json"), "w"), indent=1)

    synth = ge.load_package().synth
    d = os.path.join(HERE, "synth64")
    ds = synth.dataset(T=64, n_sites=300, n_queries=200, window=100)
    tf, sf, qf = synth.write_dataset(ds, d)
    runs = {"default": run(tf, sf, qf, ds["model"], ()),
            "fix_heur": run(tf, sf, qf, ds["model"], ("-G", "0.05")),
            "baseball": run(tf, sf, qf, ds["model"], ("--baseball-heur",))}
    q5 = os.path.join(d, "query5.fasta")
    synth.write_fasta(q5, ds["qnames"][:5], ds["queries"][:5])
    runs["noheur_all_first5"] = run(tf, sf, q5, ds["model"], ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "125"))
    os.remove(q5)
    json.dump(runs, open(os.path.join(d, "reference_placements.json"), "w"), indent=1)

    # amino acids (LG+G4): 32 taxa, 120 sites, 60 queries of 80 residues; a few ambiguity codes
    # (X is scored on the column of N in the reference's preplacement)
    d = os.path.join(HERE, "synthaa")
    ds = synth.dataset(T=32, n_sites=120, n_queries=60, window=80, kind="aa")
    q = ds["queries"]
    for i, ch in enumerate("XBZ-X"):
        row = q[i]
        inside = [k for k in range(len(row)) if row[k] != ord('-')]
        row[inside[7 + i]] = ord(ch)
        row[inside[31 + 2 * i]] = ord(ch)
    tf, sf, qf = synth.write_dataset(ds, d)
    runs = {"default": run(tf, sf, qf, ds["model"], ())}
    q3 = os.path.join(d, "query3.fasta")
    synth.write_fasta(q3, ds["qnames"][:3], ds["queries"][:3])
    runs["noheur_all_first3"] = run(tf, sf, q3, ds["model"], ("--no-heur", "--filter-min-lwr", "0", "--filter-max", "61"))
    os.remove(q3)
    json.dump(runs, open(os.path.join(d, "reference_placements.json"), "w"), indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
