"""GPU: --raxml-blo (epa_options.sliding_blo = 0) through the C ABI - the RAXML variants of the three
thorough kernels (lane = site with first-round tables / full first pass / per-rate scalers / +I, the
8-category DNA kernel, both amino-acid mappings) against the oracle, and the placements against the
reference's recorded ones (tests/golden/raxml_blo, tests/golden/make_golden_raxml.py; the oracle's
restatement is pinned on the same vectors in tests/test_oracle_raxml_blo.py)."""
import json
import os
import subprocess

import pytest

import helpers
from test_gpu_pinv import _check_pairs, _check_placements, _file_case
from test_oracle_pinv import CFG1_PINV, RATE300
from test_oracle_raxml_blo import gold

pytestmark = pytest.mark.gpu


def _raxml_case(case):
    o = helpers.oracle()
    case.placer = o.Placer(case.ref, o.Options(sliding_blo=False))
    return case


@pytest.mark.parametrize("no_first", [False, True], ids=["first_tables", "full_first_pass"])
def test_cfg1_raxml_blo(built, monkeypatch, no_first):
    if no_first:
        monkeypatch.setenv("EPA_B200_NO_FIRST", "1")
    g = gold()
    case = _raxml_case(helpers.cfg1_case())
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_pairs(case, ctx, built.capi.default_options(prescoring=0, sliding_blo=0), logl_rel=1e-9, len_abs=1e-6)
    _check_placements(case, ctx, built.capi.default_options(sliding_blo=0), g["cfg1_default"]["placements"])
    _check_placements(case, ctx, built.capi.default_options(sliding_blo=0, prescoring=0, support_threshold=0.0, filter_max=13),
                      g["cfg1_noheur_all"]["placements"])
    ctx.close()


def test_cfg1_pinv_raxml_blo(built):
    case = _raxml_case(helpers.cfg1_case(CFG1_PINV))
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_pairs(case, ctx, built.capi.default_options(prescoring=0, sliding_blo=0), logl_rel=1e-9, len_abs=1e-6)
    _check_placements(case, ctx, built.capi.default_options(sliding_blo=0), gold()["cfg1_pinv_default"]["placements"])
    ctx.close()


def test_synth64_raxml_blo(built):
    case = _raxml_case(helpers.synth64_case())
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_pairs(case, ctx, built.capi.default_options(sliding_blo=0), every=5)
    _check_placements(case, ctx, built.capi.default_options(sliding_blo=0), gold()["synth64_default"]["placements"])
    ctx.close()


def test_synth64_raxml_blo_eight_rate_categories(built):
    case = _raxml_case(_file_case("synth64", "GTR{1/2/1/1/2/1}+FU{0.3/0.2/0.2/0.3}+G8{0.5}"))
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_pairs(case, ctx, built.capi.default_options(sliding_blo=0), every=9)
    ctx.close()


@pytest.mark.parametrize("per_rate", [False, True], ids=["site_scalers", "rate_scalers"])
def test_rate300_raxml_blo(built, per_rate):
    ds = built.synth.dataset(**RATE300)
    case = _raxml_case(helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], ds["model"],
                                                per_rate=per_rate, bugcompat=per_rate, column_mask=True))
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_pairs(case, ctx, built.capi.default_options(sliding_blo=0))
    _check_placements(case, ctx, built.capi.default_options(sliding_blo=0),
                      gold()["rate300_rate" if per_rate else "rate300_site"]["placements"])
    ctx.close()


@pytest.mark.parametrize("old_aa", [False, True], ids=["unit_mapped", "site_rate_threads"])
def test_synthaa_raxml_blo(built, monkeypatch, old_aa):
    if old_aa:
        monkeypatch.setenv("EPA_B200_OLD_AA", "1")
    case = _raxml_case(helpers.synthaa_case())
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_pairs(case, ctx, built.capi.default_options(sliding_blo=0), every=3)
    _check_placements(case, ctx, built.capi.default_options(sliding_blo=0), gold()["synthaa_default"]["placements"])
    ctx.close()


def test_cli_raxml_blo(built, tmp_path):
    exe = os.path.join(helpers.ROOT, "epa-ng_b200", "epa-ng-b200")
    d = os.path.join(helpers.GOLDEN, "synth64")
    out = str(tmp_path / "cli")
    subprocess.run([exe, "-t", os.path.join(d, "tree.nwk"), "-s", os.path.join(d, "ref.fasta"), "-q",
                    os.path.join(d, "query.fasta"), "-m", "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}", "-w", out,
                    "--redo", "--raxml-blo"], check=True, stdout=subprocess.DEVNULL)
    doc = json.load(open(os.path.join(out, "epa_result.jplace")))
    got = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
    bad = []
    for name, want in gold()["synth64_default"]["placements"].items():
        try:
            helpers.assert_placements_close(got[name], want, name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} differ: {bad[:3]}"
