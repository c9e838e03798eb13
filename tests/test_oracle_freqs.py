"""Empirical base frequencies (+F / +FC) and +IC: pins the oracle against placements recorded from the
unmodified reference (tests/golden/make_golden_freqs.py)."""
import json
import os

import numpy as np

import helpers

CFG1_FC = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FC+G4{1.0}"
CFG1_F_IC = "GTR{1/2/1/1/2/1}+F+IC+G4{0.7}"
AA_F = "LG+F+G4{0.8}"


def gold():
    return json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_empirical.json")))


def _check(case, want, opts=None, logl_rel=1e-9):
    o = helpers.oracle()
    placer = o.Placer(case.ref, opts) if opts is not None else case.placer
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in placer.place(seq)]
        helpers.assert_placements_close(got, want[name], name, logl_rel=logl_rel, len_abs=1e-5)


def test_cfg1_empirical_frequencies_match_reference():
    g = gold()
    case = helpers.cfg1_case(CFG1_FC)
    assert np.allclose(case.model.freqs, g["cfg1_printed_freqs"], atol=1e-6) and abs(case.model.freqs.sum() - 1) < 1e-12
    _check(case, g["cfg1_fc_default"]["placements"])
    o = helpers.oracle()
    _check(case, g["cfg1_fc_noheur_all"]["placements"], o.Options(prescoring=False, support_threshold=0.0, filter_max=13))


def test_cfg1_plus_f_plus_ic_matches_reference():
    # +F = +FC; +IC stays at 0 in the reference ("P-inv (empirical): 0")
    case = helpers.cfg1_case(CFG1_F_IC)
    assert case.model.pinv == 0.0
    _check(case, gold()["cfg1_f_ic_default"]["placements"])


def test_synthaa_empirical_frequencies_match_reference():
    d = os.path.join(helpers.GOLDEN, "synthaa")
    case = helpers.load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), AA_F)
    _check(case, gold()["synthaa_f_default"]["placements"])
