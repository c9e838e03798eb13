"""Every empirical amino-acid matrix of the reference (PM/util/models_aa.c:28-57; 28 tables): the oracle against
placements recorded from the unmodified reference (tests/golden/make_golden_aa_models.py), and the host
layer's model parser against the oracle."""
import json
import os

import numpy as np
import pytest

import helpers

GOLD = json.load(open(os.path.join(helpers.GOLDEN, "synthaa", "reference_models.json")))
NAMES = sorted(GOLD)


def test_all_reference_tables_are_present():
    o = helpers.oracle()
    assert len(NAMES) == 28 and set(NAMES) == set(o.protein_tables())


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    d = os.path.join(helpers.GOLDEN, "synthaa")
    case = helpers.load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query6.fasta"),
                             GOLD[name]["model"])
    want = GOLD[name]["placements"]
    seqs = dict(zip(case.qnames, case.qseqs))
    for qname, w in want.items():
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seqs[qname])]
        helpers.assert_placements_close(got, w, f"{name}/{qname}", logl_rel=1e-9, len_abs=1e-5)


@pytest.mark.parametrize("name", NAMES)
def test_host_parser_matches_oracle(built, name):
    o = helpers.oracle()
    want = o.parse_model(GOLD[name]["model"])
    got = built.session.parse_model(GOLD[name]["model"])
    S = 20
    assert got["states"] == S and got["rate_cats"] == 4
    assert np.allclose(got["freqs"], want.freqs, rtol=1e-15, atol=0)
    V, Vi = got["eigenvecs"].reshape(S, S), got["inv_eigenvecs"].reshape(S, S)
    Vw, Viw = want.eigenvecs.reshape(S, S), want.inv_eigenvecs.reshape(S, S)
    for t in (1e-3, 0.1, 2.0):
        P = np.eye(S) + (Vi * np.expm1(got["eigenvals"] * t)[None, :]) @ V
        Pw = np.eye(S) + (Viw * np.expm1(want.eigenvals * t)[None, :]) @ Vw
        assert np.allclose(P, Pw, rtol=0, atol=1e-13)


def test_protgtr_from_a_raxml8_info_file(built):
    """PROTGTR{190 rates}+FU{..}+G4{..} as the reference derives it from its own RAxML 8 protein info file
    (test/data/modelfiles/rax8_prot): oracle against the recorded reference run, host parser against the oracle."""
    d = os.path.join(helpers.GOLDEN, "synthaa")
    g = json.load(open(os.path.join(d, "reference_protgtr.json")))
    assert g["model"] == built.session.model_from_file(os.path.join(helpers.GOLDEN, "modelfiles", "rax8_prot"))
    case = helpers.load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query6.fasta"), g["model"])
    seqs = dict(zip(case.qnames, case.qseqs))
    for qname, w in g["placements"].items():
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seqs[qname])]
        helpers.assert_placements_close(got, w, qname, logl_rel=1e-9, len_abs=1e-5)
    want = helpers.oracle().parse_model(g["model"])
    got = built.session.parse_model(g["model"])
    assert np.allclose(got["freqs"], want.freqs, rtol=1e-14, atol=0)
    S = 20
    V, Vi = got["eigenvecs"].reshape(S, S), got["inv_eigenvecs"].reshape(S, S)
    Vw, Viw = want.eigenvecs.reshape(S, S), want.inv_eigenvecs.reshape(S, S)
    for t in (1e-3, 0.1, 2.0):
        P = np.eye(S) + (Vi * np.expm1(got["eigenvals"] * t)[None, :]) @ V
        Pw = np.eye(S) + (Viw * np.expm1(want.eigenvals * t)[None, :]) @ Vw
        assert np.allclose(P, Pw, rtol=0, atol=1e-13)
