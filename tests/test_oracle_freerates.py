"""Free-rate models (+R[n]{rates}{weights}): the oracle - including the reference's quirk that tiny-tree
likelihoods keep the default weights 1/R (src/tree/tiny_util.cpp:110-111) - against placements recorded from
the unmodified reference (tests/golden/make_golden_freerates.py); the host parser against the oracle."""
import json
import os

import numpy as np
import pytest

import helpers

GOLD = json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_freerates.json")))


@pytest.mark.parametrize("key", sorted(GOLD))
def test_oracle_matches_reference(key):
    o = helpers.oracle()
    case = helpers.cfg1_case(GOLD[key]["model"])
    placer = o.Placer(case.ref, o.Options(prescoring=False, support_threshold=0.0, filter_max=13))
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in placer.place(seq)]
        helpers.assert_placements_close(got, GOLD[key]["placements"][name], f"{key}/{name}", logl_rel=1e-9, len_abs=1e-5)


def test_user_weights_normalise_the_rates_but_do_not_weight_the_placement():
    o = helpers.oracle()
    m = o.parse_model(GOLD["r4_user"]["model"])
    assert np.allclose(m.weights, [0.4, 0.3, 0.2, 0.1]) and abs(float((m.rates * m.weights).sum()) - 1.0) < 1e-14
    case = helpers.cfg1_case(GOLD["r4_user"]["model"])
    assert np.allclose(case.placer.pmodel.weights, 0.25)


@pytest.mark.parametrize("key", sorted(GOLD))
def test_host_parser_matches_oracle(built, key):
    o = helpers.oracle()
    want = o.parse_model(GOLD[key]["model"])
    got = built.session.parse_model(GOLD[key]["model"])
    assert got["rate_cats"] == want.rate_cats
    assert np.allclose(got["rates"], want.rates, rtol=1e-13, atol=0)
    assert np.allclose(got["weights"], want.weights, rtol=1e-15, atol=0)
