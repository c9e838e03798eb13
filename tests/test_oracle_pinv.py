"""+I models (proportion of invariant sites, +IU{p}): pins the oracle's invariant-site path
(LP/models.c:495-760, LP/core_likelihood.c:524-549, LP/core_derivatives.c:676-687,757-772,
LP/core_pmatrix.c:209-220) against placements recorded from the unmodified reference
(tests/golden/make_golden_pinv.py)."""
import json
import os

import numpy as np
import pytest

import helpers

CFG1_PINV = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+IU{0.2}+G4{1.0}"
SYNTH64_PINV = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+IU{0.15}+G4{0.5}"
SYNTHAA_PINV = "LG+IU{0.1}+G4{0.8}"
RATE300_PINV = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+IU{0.1}+G4{0.5}"
RATE300 = dict(T=300, n_sites=400, n_queries=24, window=120, seed_tree=7, seed_q=8)


def gold():
    return json.load(open(os.path.join(helpers.GOLDEN, "pinv", "reference_placements.json")))


def _check(case, want, opts=None, logl_rel=1e-9):
    o = helpers.oracle()
    placer = o.Placer(case.ref, opts) if opts is not None else case.placer
    bad = []
    for name, seq in zip(case.qnames, case.qseqs):
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in placer.place(seq)]
        try:
            helpers.assert_placements_close(got, want[name], name, logl_rel=logl_rel, len_abs=1e-5)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(want)} queries differ: {bad[:3]}"


def test_invariant_sites_follow_libpll():
    case = helpers.cfg1_case(CFG1_PINV)
    inv = case.model.invariant
    masks = case.tip_masks()
    for s in range(case.n):
        a = np.bitwise_and.reduce(masks[:, s])
        want = -1 if a == 0 or (a & (a - 1)) else int(a).bit_length() - 1
        assert inv[s] == want
    assert (inv >= 0).sum() > 100 and (inv < 0).sum() > 100


def test_cfg1_pinv_matches_reference():
    g = gold()
    assert g["cfg1_default"]["model"] == CFG1_PINV
    case = helpers.cfg1_case(CFG1_PINV)
    _check(case, g["cfg1_default"]["placements"])
    o = helpers.oracle()
    _check(case, g["cfg1_noheur_all"]["placements"], o.Options(prescoring=False, support_threshold=0.0, filter_max=13))


def test_pinv_changes_the_result():
    # the fixture pins something: the same data without +I gives other likelihoods
    g = gold()
    case = helpers.cfg1_case(helpers.GTRG)
    name, seq = case.qnames[0], case.qseqs[0]
    got = case.placer.place(seq)
    assert abs(got[0].logl - g["cfg1_default"]["placements"][name][0][1]) > 1.0


def test_synth64_pinv_matches_reference():
    g = gold()
    d = os.path.join(helpers.GOLDEN, "synth64")
    case = helpers.load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), SYNTH64_PINV)
    assert (case.model.invariant >= 0).any()
    _check(case, g["synth64_default"]["placements"])


def test_synthaa_pinv_matches_reference():
    g = gold()
    d = os.path.join(helpers.GOLDEN, "synthaa")
    case = helpers.load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"), SYNTHAA_PINV)
    assert (case.model.invariant >= 0).any()
    _check(case, g["synthaa_default"]["placements"])


@pytest.mark.parametrize("per_rate", [False, True])
def test_rate300_pinv_matches_reference(built, per_rate):
    g = gold()
    ds = built.synth.dataset(**RATE300)
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], RATE300_PINV,
                                    per_rate=per_rate, bugcompat=per_rate, column_mask=True)
    _check(case, g["rate300_rate" if per_rate else "rate300_site"]["placements"])
