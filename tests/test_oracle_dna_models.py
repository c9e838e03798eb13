"""The 22 named DNA models of the reference and their aliases (PM/util/models_dna.c:40-125), with the ML-mode
default rates (0.5 ... 1.0 over the six rates whatever the symmetry, src/core/raxml/Model.cpp:484-490) and with
user rates per symmetry class: the oracle against placements recorded from the unmodified reference
(tests/golden/make_golden_dna_models.py), and the host layer's parser against the oracle."""
import json
import os

import numpy as np
import pytest

import helpers

GOLD = json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_dna_models.json")))
KEYS = sorted(GOLD)


@pytest.mark.parametrize("key", KEYS)
def test_oracle_matches_reference(key):
    case = helpers.cfg1_case(GOLD[key]["model"])
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seq)]
        helpers.assert_placements_close(got, GOLD[key]["placements"][name], f"{key}/{name}", logl_rel=1e-9, len_abs=1e-5)


@pytest.mark.parametrize("key", KEYS)
def test_host_parser_matches_oracle(built, key):
    o = helpers.oracle()
    want = o.parse_model(GOLD[key]["model"])
    got = built.session.parse_model(GOLD[key]["model"])
    assert got["states"] == 4 and got["rate_cats"] == want.rate_cats
    assert np.allclose(got["freqs"], want.freqs, rtol=1e-15, atol=0)
    S = 4
    V, Vi = got["eigenvecs"].reshape(S, S), got["inv_eigenvecs"].reshape(S, S)
    Vw, Viw = want.eigenvecs.reshape(S, S), want.inv_eigenvecs.reshape(S, S)
    for t in (1e-3, 0.1, 2.0):
        P = np.eye(S) + (Vi * np.expm1(got["eigenvals"] * t)[None, :]) @ V
        Pw = np.eye(S) + (Viw * np.expm1(want.eigenvals * t)[None, :]) @ Vw
        assert np.allclose(P, Pw, rtol=0, atol=1e-14)
