"""--raxml-blo (optimize_branch_triplet with sliding == false, src/core/pll/optimize.cpp:274-278 ->
pllmod_opt_optimize_branch_lengths_local PM/optimize/pll_optimize.c:778-1097 with the older Newton
variant PM/optimize/opt_algorithms.c:281-384): pins the oracle's restatement against placements
recorded from the unmodified reference (tests/golden/make_golden_raxml.py)."""
import json
import os

import pytest

import helpers
from test_oracle_pinv import CFG1_PINV, RATE300

SYNTH = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}"


def gold():
    return json.load(open(os.path.join(helpers.GOLDEN, "raxml_blo", "reference_placements.json")))


def _check(case, want, logl_rel=1e-9, **kw):
    o = helpers.oracle()
    placer = o.Placer(case.ref, o.Options(sliding_blo=False, **kw))
    bad = []
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in placer.place(seq)]
        try:
            helpers.assert_placements_close(got, want[name], name, logl_rel=logl_rel, len_abs=1e-5)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(want)} queries differ: {bad[:3]}"


def test_cfg1_raxml_blo_matches_reference():
    g = gold()
    case = helpers.cfg1_case()
    _check(case, g["cfg1_default"]["placements"])
    _check(case, g["cfg1_noheur_all"]["placements"], prescoring=False, support_threshold=0.0, filter_max=13)
    # the mode is a different optimiser: its lengths differ from the default mode's
    name = case.qnames[0]
    default = helpers.golden("cfg1")["gtrg_default"]["placements"][name][0]
    assert abs(default[3] - g["cfg1_default"]["placements"][name][0][3]) > 1e-3


def test_cfg1_pinv_raxml_blo_matches_reference():
    _check(helpers.cfg1_case(CFG1_PINV), gold()["cfg1_pinv_default"]["placements"])


def test_synth64_raxml_blo_matches_reference():
    _check(helpers.synth64_case(), gold()["synth64_default"]["placements"])


def test_synthaa_raxml_blo_matches_reference():
    _check(helpers.synthaa_case(), gold()["synthaa_default"]["placements"])


@pytest.mark.parametrize("per_rate", [False, True])
def test_rate300_raxml_blo_matches_reference(built, per_rate):
    ds = built.synth.dataset(**RATE300)
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], ds["model"],
                                    per_rate=per_rate, bugcompat=per_rate, column_mask=True)
    _check(case, gold()["rate300_rate" if per_rate else "rate300_site"]["placements"])
