"""cfg2-shaped fixture (BASELINE.json configs[1]: 1k-taxon DNA tree, GTR+G4, 1000 sites, 200-bp window queries):
the UNMODIFIED reference's placements of the first 10 000 queries of the bench data set (epa-ng_b200/synth.py,
the seeds bench.py uses), and of 24 queries on a second tree with a general GTR model (three distinct non-zero
eigenvalues: the general variant of the thorough DNA kernel at full scale).

    python tests/golden/make_golden_cfg2.py          (needs oracle/_ref/epa-ng; ~1 minute on 8 cores)

Writes tests/golden/cfg2/reference_10k.json.gz = {model, n_queries, dataset_sha1, placements{name: [[edge, logl,
lwr, distal, pendant], ...]}} (numbers as the jplace prints them)."""
import gzip
import hashlib
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

GTR_GENERAL = "GTR{0.676278/2.012275/0.478487/0.753965/2.406436/1.0}+FU{0.245629/0.235012/0.253054/0.266305}+G4{1.078763}"


def dataset_sha1(ds):
    h = hashlib.sha1()
    h.update(ds["newick"].encode())
    h.update(ds["ref"].tobytes())
    h.update(ds["queries"].tobytes())
    return h.hexdigest()


def run(synth, ds, model, threads=8):
    tmp = tempfile.mkdtemp(prefix="golden_cfg2_")
    try:
        tf, sf, qf = synth.write_dataset(ds, tmp)
        pl, _ = pyoracle.run_reference(tf, sf, qf, model, os.path.join(tmp, "out"), threads=threads)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return pl


def main():
    synth = ge.load_package().synth
    out_dir = os.path.join(HERE, "cfg2")
    os.makedirs(out_dir, exist_ok=True)
    ds = synth.dataset(T=1000, n_sites=1000, n_queries=10000, window=200)
    doc = {"model": ds["model"], "n_queries": 10000, "dataset_sha1": dataset_sha1(ds), "placements": run(synth, ds, ds["model"])}
    with gzip.open(os.path.join(out_dir, "reference_10k.json.gz"), "wt") as fh:
        json.dump(doc, fh, separators=(",", ":"))
    ds2 = synth.dataset(T=1000, n_sites=1000, n_queries=2000, window=200, seed_tree=11, seed_q=12)
    doc2 = {"model": GTR_GENERAL, "n_queries": 2000, "dataset_sha1": dataset_sha1(ds2), "placements": run(synth, ds2, GTR_GENERAL)}
    with gzip.open(os.path.join(out_dir, "reference_gtr_2k.json.gz"), "wt") as fh:
        json.dump(doc2, fh, separators=(",", ":"))
    print("wrote", out_dir, len(doc["placements"]), len(doc2["placements"]))


if __name__ == "__main__":
    main()
