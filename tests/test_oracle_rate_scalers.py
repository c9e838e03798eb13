"""Per-rate scalers (PLL_ATTRIB_RATE_SCALERS, the reference's --rate-scalers on / auto above 2000 tips):
pins the oracle's per-rate path - including the reference's scaler window offset in the thorough
phase - against placements recorded from the unmodified reference on a seeded
300-taxon data set whose CLVs do get rescaled (tests/golden/make_golden_rate.py)."""
import json
import os

import numpy as np
import pytest

import helpers


@pytest.fixture(scope="module")
def rate300(built):
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements.json")))
    ds = built.synth.dataset(**g["dataset"])
    assert ds["model"] == g["model"]
    return ds, g["placements"]


def _case(ds, bugcompat):
    return helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], ds["model"],
                                    per_rate=True, bugcompat=bugcompat, column_mask=True)


def test_scalers_are_exercised(rate300):
    ds, _ = rate300
    case = _case(ds, True)
    counts = [int(side.scaler.max()) for side in case.ref.sides.values() if side.scaler is not None]
    assert max(counts) >= 1, "no CLV of the fixture was rescaled: the test would not pin anything"


def test_oracle_per_rate_bugcompat_matches_reference(rate300):
    ds, want = rate300
    case = _case(ds, True)
    for name, seq in zip(case.qnames, case.qseqs):
        got = case.placer.place(seq)
        w = want[name]
        assert [p.edge for p in got] == [int(x[0]) for x in w], name
        for p, x in zip(got, w):
            assert abs(p.logl - x[1]) <= 1e-9 * abs(x[1])
            assert abs(p.lwr - x[2]) <= 1e-6 and abs(p.distal - x[3]) <= 1e-5 and abs(p.pendant - x[4]) <= 1e-5


def test_corrected_focus_differs_from_reference(rate300):
    """The offset is a real effect on this fixture: reading the scalers of the site itself gives
    different log-likelihoods (documented deviation switch, not the default)."""
    ds, want = rate300
    case = _case(ds, False)
    diff = 0
    for name, seq in zip(case.qnames[:8], case.qseqs[:8]):
        got = case.placer.place(seq)
        if [p.edge for p in got] != [int(x[0]) for x in want[name]] or abs(got[0].logl - want[name][0][1]) > 1e-6 * abs(got[0].logl):
            diff += 1
    assert diff > 0


def test_oracle_per_rate_eight_categories_matches_reference(built):
    """GTR+G8 with per-rate scalers (tests/golden/make_golden_rate2.py): the 4x4 kernels of the reference keep
    proper per-rate counts for any number of categories."""
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_rate2.json")))["dna8"]
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model"],
                                    per_rate=True, bugcompat=True, column_mask=True)
    assert max(int(s.scaler.max()) for s in case.ref.sides.values() if s.scaler is not None) >= 1
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seq)]
        helpers.assert_placements_close(got, g["placements"][name], name, logl_rel=1e-9, len_abs=1e-5)


def _ti_rescaled(built):
    """oracle statistic: sites rescaled by the generic tip-inner update under per-rate scalers"""
    import ctypes
    return ctypes.c_ulong.in_dll(helpers.oracle().lib(), "orc_stat_ti_rescaled")


@pytest.mark.parametrize("fixture", ["rate2", "ladder"])
def test_oracle_amino_acids_per_rate_matches_reference(built, fixture):
    """Amino acids under --rate-scalers on: the reference runs libpll's generic kernels, whose tip-inner CLV update
    (LP/core_partials.c:461-506) tests and rescales whole sites and counts that in entry [site index] of the
    [site][rate] array. 'ladder' (tests/golden/make_golden_aa_rate.py) is a caterpillar-like tree on which those
    rescalings do happen - in the reference CLVs and inside the tiny trees; 'rate2' is a random tree where the
    difference to a per-rate computation is that tip-inner updates do not rescale single rates."""
    if fixture == "rate2":
        g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_rate2.json")))["aa"]
    else:
        g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_aa_ladder.json")))
    ds = built.synth.dataset(**g["dataset"])
    cnt = _ti_rescaled(built)
    c0 = cnt.value
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model"],
                                    per_rate=True, bugcompat=True, column_mask=True)
    c1 = cnt.value
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seq)]
        helpers.assert_placements_close(got, g["placements"][name], name, logl_rel=1e-9, len_abs=1e-5)
    if fixture == "ladder":
        assert c1 - c0 > 0, "no whole-site rescaling in the reference tree: the fixture would not pin the counter placement"
        assert cnt.value - c1 > 0, "no whole-site rescaling inside a tiny tree"


def test_oracle_amino_acids_per_rate_all_edges(built):
    """--no-heur on the ladder data set: every edge goes through the thorough phase (first four queries)."""
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_aa_ladder.json")))
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model"],
                                    per_rate=True, bugcompat=True, column_mask=True,
                                    opts=helpers.oracle().Options(prescoring=False))
    for name, seq in zip(case.qnames[:4], case.qseqs[:4]):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seq)]
        helpers.assert_placements_close(got, g["placements_no_heur"][name], name, logl_rel=1e-9, len_abs=1e-5)


def test_oracle_amino_acids_per_rate_raxml_blo(built):
    """--raxml-blo on the ladder data set (proximal step on a tip edge = tip-tip update, never rescaled)"""
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_aa_ladder.json")))
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model"],
                                    per_rate=True, bugcompat=True, column_mask=True,
                                    opts=helpers.oracle().Options(sliding_blo=False))
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seq)]
        helpers.assert_placements_close(got, g["placements_raxml_blo"][name], name, logl_rel=1e-9, len_abs=1e-5)


def test_oracle_amino_acids_per_rate_with_invariant_sites(built):
    """+IU{0.2} on top: the invariant term joins sums built from CLVs that the generic tip-inner update rescaled as
    whole sites (LP/core_derivatives.c:676-687, LP/core_likelihood.c:524-556)."""
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_aa_ladder.json")))
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model_pinv"],
                                    per_rate=True, bugcompat=True, column_mask=True)
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seq)]
        helpers.assert_placements_close(got, g["placements_pinv"][name], name, logl_rel=1e-9, len_abs=1e-5)
