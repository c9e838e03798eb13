"""CPU: pins the oracle (oracle/epa_oracle.c + oracle/pyoracle.py) against outputs of the
UNMODIFIED reference recorded in tests/golden/ (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import helpers


def _oracle_run(case, **opt_kw):
    o = helpers.oracle()
    opts = o.Options(**opt_kw)
    placer = o.Placer(case.ref, opts)
    placer.lookup = case.placer.lookup if case.placer.lookup is not None else placer.build_lookup()
    case.placer.lookup = placer.lookup
    return {name: [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in placer.place(seq)]
            for name, seq in zip(case.qnames, case.qseqs)}


@pytest.mark.parametrize("mname,model", [("gtrg", helpers.GTRG), ("gtrb", helpers.GTR_B)])
def test_cfg1_matches_reference(mname, model):
    gold = helpers.golden("cfg1")
    case = helpers.cfg1_case(model)
    runs = {
        "default": dict(),
        "noheur_all": dict(prescoring=False, support_threshold=0.0, filter_max=13),
        "heur_all": dict(support_threshold=0.0, filter_max=13),
    }
    for rname, kw in runs.items():
        got = _oracle_run(case, **kw)
        want = gold[f"{mname}_{rname}"]["placements"]
        assert set(got) == set(want)
        for name in want:
            helpers.assert_placements_close(got[name], want[name], f"{mname}/{rname}/{name}")


def test_cfg1_numbered_newick_matches_reference():
    gold = helpers.golden("cfg1")
    case = helpers.cfg1_case()
    assert helpers.oracle().numbered_newick(case.tree) == gold["gtrg_default"]["tree"]


def test_cfg1_tree_logl_equal_on_every_edge():
    # the reference's strongest invariant for directional CLVs (test/src/epa_pll_util.cpp:82-121)
    case = helpers.cfg1_case()
    vals = [case.ref.tree_logl(e) for e in range(case.tree.num_branches)]
    assert np.allclose(vals, vals[0], rtol=1e-12, atol=0)


def test_synth64_matches_reference():
    gold = helpers.golden("synth64")
    case = helpers.synth64_case()
    got = _oracle_run(case)
    want = gold["default"]["placements"]
    assert set(got) == set(want)
    bad = []
    for name in want:
        try:
            helpers.assert_placements_close(got[name], want[name], name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(want)} queries differ: {bad[:3]}"
    assert helpers.oracle().numbered_newick(case.tree) == gold["default"]["tree"]


def test_range_restriction_invariance():
    # test/src/pll_util.cpp:325-335: when the flanks are gaps the focused evaluation equals the
    # full one minus the (constant) contribution of the fully ambiguous sites; here: the thorough
    # result with premasking must equal placing the window cut out as its own alignment columns.
    o = helpers.oracle()
    case = helpers.synth64_case()
    seq = case.qseqs[0]
    b, w = o.valid_range(seq)
    pl = case.placer.thorough(seq, 5)
    # same query, gaps replaced by '-' already; evaluate by hand on the window only
    import ctypes as C
    d, p, length = case.ref.edges[5]
    m = case.placer.mask_tab[np.frombuffer(seq.encode(), dtype=np.uint8)].astype(np.uint32)
    res = o.OrcBlo()
    o.lib().orc_place_thorough(C.byref(case.model.c()), case.n, C.byref(d.c), C.byref(p.c), length,
                               m.ctypes.data_as(C.POINTER(C.c_uint32)), b, w, C.byref(res))
    assert res.logl == pl.logl and res.pendant == pl.pendant


def test_synthaa_matches_reference():
    """Amino acids (LG+G4, 20-state path) incl. the X -> N preplacement quirk of the reference."""
    gold = helpers.golden("synthaa")
    case = helpers.synthaa_case()
    got = _oracle_run(case)
    want = gold["default"]["placements"]
    assert set(got) == set(want)
    bad = []
    for name in want:
        try:
            # 1e-9: the table frequencies are normalised as pll_set_frequencies does (LG sums to 1.000001 as published)
            helpers.assert_placements_close(got[name], want[name], name, logl_rel=1e-9)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(want)} queries differ: {bad[:3]}"


@pytest.mark.parametrize("rname,kw", [("fix_heur", dict(heuristic=1, prescoring_threshold=0.05)),
                                      ("baseball", dict(heuristic=2))])
def test_synth64_other_heuristics_match_reference(rname, kw):
    # -G / --baseball-heur (src/core/heuristics.hpp:66-117)
    gold = helpers.golden("synth64")[rname]["placements"]
    case = helpers.synth64_case()
    got = _oracle_run(case, **kw)
    bad = []
    for name in gold:
        try:
            helpers.assert_placements_close(got[name], gold[name], name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{rname}: {len(bad)} of {len(gold)} queries differ: {bad[:3]}"


def test_synth64_no_pre_mask_matches_reference():
    """--no-pre-mask (no column masking, every query scored over the whole alignment with its gaps as fully
    ambiguous sites): first 40 synth64 queries against the reference's recorded run (reference_nopremask.json,
    written with oracle.run_reference(..., extra=("--no-pre-mask",)))."""
    import json
    o = helpers.oracle()
    d = os.path.join(helpers.GOLDEN, "synth64")
    g = json.load(open(os.path.join(d, "reference_nopremask.json")))
    case = helpers.load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"),
                             g["model"], premasking=False)
    seqs = dict(zip(case.qnames, case.qseqs))
    assert case.n == 300
    for name, want in g["placements"].items():
        got = [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in case.placer.place(seqs[name])]
        helpers.assert_placements_close(got, want, name, logl_rel=1e-9, len_abs=1e-5)
