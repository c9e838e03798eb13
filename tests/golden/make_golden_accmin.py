"""--filter-acc-lwr with --filter-min > 1 (until_accumulated_reached keeps max(summed, min - 1) entries,
src/set_manipulators.cpp:90-113): placements of the reference's test data recorded from the unmodified reference.
    python tests/golden/make_golden_accmin.py  ->  tests/golden/cfg1/reference_accmin.json"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402

MODEL = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+G4{1.0}"
RUNS = {"acc05_min3": ("--no-heur", "--filter-acc-lwr", "0.5", "--filter-min", "3", "--filter-max", "5"),
        "acc0999_min2": ("--no-heur", "--filter-acc-lwr", "0.999", "--filter-min", "2", "--filter-max", "6"),
        "acc09_min1": ("--no-heur", "--filter-acc-lwr", "0.9", "--filter-max", "4")}


def main():
    d = os.path.join(HERE, "cfg1")
    out = {}
    for key, extra in RUNS.items():
        tmp = tempfile.mkdtemp(prefix="golden_acc_")
        try:
            pl, _ = pyoracle.run_reference(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), os.path.join(d, "query.fasta"),
                                           MODEL, tmp, threads=1, extra=extra)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        out[key] = {"model": MODEL, "extra": list(extra), "placements": pl}
        print(key, {k: len(v) for k, v in pl.items()})
    json.dump(out, open(os.path.join(d, "reference_accmin.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
