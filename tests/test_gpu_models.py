"""GPU: the model strings, model files and options whose ORACLE side is pinned on recorded reference runs, now
through the device path (files -> jplace, epa_run_files): named DNA models (JC/F81 run the one-group and K80 the
two-group variant of the lane = site kernel, HKY/TN93 the general one), free-rate models, protein matrices,
PROTGTR from a RAxML 8 info file, -m <file>, and --no-pre-mask."""
import json
import os

import pytest

import helpers

pytestmark = pytest.mark.gpu


def _opts(capi, extra):
    kw = {}
    extra = list(extra or [])
    i = 0
    while i < len(extra):
        a = extra[i]
        if a == "--no-heur":
            kw["prescoring"] = 0
        elif a == "--no-pre-mask":
            kw["premasking"] = 0
        elif a == "--filter-min-lwr":
            i += 1
            kw["support_threshold"] = float(extra[i])
        elif a == "--filter-max":
            i += 1
            kw["filter_max"] = int(extra[i])
        elif a == "--filter-min":
            i += 1
            kw["filter_min"] = int(extra[i])
        elif a == "--filter-acc-lwr":
            i += 1
            kw["filter_acc_lwr"] = 1
            kw["support_threshold"] = float(extra[i])
        else:
            raise AssertionError(f"option {a} not mapped")
        i += 1
    return capi.default_options(**kw)


def _run(built, tmp_path, key, files, model, extra, want, **tol):
    out = str(tmp_path / key.replace("/", "_"))
    built.session.run_files(*files, model, out, opts=_opts(built.capi, extra))
    doc = json.load(open(os.path.join(out, "epa_result.jplace")))
    got = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
    bad = []
    for name, w in want.items():
        try:
            helpers.assert_placements_close(got[name], w, f"{key}/{name}", **tol)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{key}: {len(bad)} of {len(want)} queries differ: {bad[:2]}"


CFG1 = tuple(os.path.join(helpers.GOLDEN, "cfg1", f) for f in ("ref.tre", "aln.fasta", "query.fasta"))
AA6 = tuple(os.path.join(helpers.GOLDEN, "synthaa", f) for f in ("tree.nwk", "ref.fasta", "query6.fasta"))
DNA = json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_dna_models.json")))
RATES = json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_freerates.json")))
AAM = json.load(open(os.path.join(helpers.GOLDEN, "synthaa", "reference_models.json")))


@pytest.mark.parametrize("key", sorted(DNA))
def test_named_dna_models(built, tmp_path, key):
    _run(built, tmp_path, key, CFG1, DNA[key]["model"], DNA[key].get("extra"), DNA[key]["placements"])


def test_eigenvalue_groups_are_found(built):
    """JC and F81 have one distinct non-zero eigenvalue, K80 (and every model at the reference's default rates
    0.5 ... 1.0) two, models with user rates three: the three variants of the thorough DNA kernel all run above."""
    groups = {}
    for key in ("JC_default", "F81_default", "K80_default", "HKY_default", "HKY_user", "GTR_user"):
        ev = sorted(built.session.parse_model(DNA[key]["model"])["eigenvals"])
        nz = [e for e in ev if abs(e) > 1e-9]
        distinct = 1 + sum(abs(a - b) > 1e-13 * abs(nz[0]) for a, b in zip(nz, nz[1:]))
        groups[key] = distinct
    assert groups == {"JC_default": 1, "F81_default": 1, "K80_default": 2, "HKY_default": 2, "HKY_user": 3, "GTR_user": 3}


@pytest.mark.parametrize("key", sorted(RATES))
def test_free_rate_models(built, tmp_path, key):
    _run(built, tmp_path, key, CFG1, RATES[key]["model"], RATES[key].get("extra"), RATES[key]["placements"])


@pytest.mark.parametrize("name", ["BLOSUM62", "WAG", "MTZOA", "Q.PFAM_GB", "HIVB"])
def test_protein_matrices(built, tmp_path, name):
    _run(built, tmp_path, name, AA6, AAM[name]["model"], AAM[name].get("extra"), AAM[name]["placements"])


def test_protgtr_and_model_file(built, tmp_path):
    g = json.load(open(os.path.join(helpers.GOLDEN, "synthaa", "reference_protgtr.json")))
    _run(built, tmp_path, "protgtr_string", AA6, g["model"], None, g["placements"])
    # -m <file>: the RAxML 8 info file itself (src/main.cpp:433-436)
    _run(built, tmp_path, "protgtr_file", AA6, os.path.join(helpers.GOLDEN, "modelfiles", "rax8_prot"), None, g["placements"])


def test_no_pre_mask(built, tmp_path):
    d = os.path.join(helpers.GOLDEN, "synth64")
    g = json.load(open(os.path.join(d, "reference_nopremask.json")))
    files = (os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"))
    _run(built, tmp_path, "nopremask", files, g["model"], g["extra"], g["placements"])


ACC = json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_accmin.json")))


@pytest.mark.parametrize("key", sorted(ACC))
def test_accumulated_filter_with_minimum(built, tmp_path, key):
    """--filter-acc-lwr with --filter-min: the reference keeps max(summed, min - 1) entries
    (until_accumulated_reached, src/set_manipulators.cpp:90-113; tests/golden/make_golden_accmin.py)."""
    _run(built, tmp_path, key, CFG1, ACC[key]["model"], ACC[key]["extra"], ACC[key]["placements"])
