"""GPU: parity at the BASELINE.json cfg2 shape (1k-taxon DNA tree, 1000 sites, 200-bp window queries) against
placements recorded from the unmodified reference (tests/golden/make_golden_cfg2.py): 10 000 queries of the bench
data set (one-group variant of the thorough kernel, tensor-core preplacement, TMEM sumtables) and 2 000 queries
under a general GTR model (three distinct eigenvalues); plus long windows that take the global-scratch variant."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
D = os.path.join(helpers.GOLDEN, "cfg2")


def _sha1(ds):
    h = hashlib.sha1()
    h.update(ds["newick"].encode())
    h.update(ds["ref"].tobytes())
    h.update(ds["queries"].tobytes())
    return h.hexdigest()


def _place(built, ds, model, n):
    sess = built.session.Session(ds["newick"], ds["names"], ds["ref"], model)
    try:
        out, counts = sess.place(ds["queries"][:n])
    finally:
        sess.close()
    return helpers.records_to_lists(out, counts)


@pytest.mark.parametrize("fname,kw", [("reference_10k.json.gz", dict()),
                                      ("reference_gtr_2k.json.gz", dict(seed_tree=11, seed_q=12))])
def test_cfg2_shape_matches_reference(built, fname, kw):
    g = json.load(gzip.open(os.path.join(D, fname), "rt"))
    n = g["n_queries"]
    ds = built.synth.dataset(T=1000, n_sites=1000, n_queries=n, window=200, **kw)
    assert _sha1(ds) == g["dataset_sha1"], "the synthetic data set is not the one the fixture was recorded on"
    got = _place(built, ds, g["model"], n)
    bad = []
    worst = 0.0
    for qi, name in enumerate(ds["qnames"][:n]):
        want = g["placements"][name]
        try:
            helpers.assert_placements_close(got[qi], want, name)          # north_star tolerances: 1e-6 rel, 1e-4 lengths
            worst = max(worst, max(abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(got[qi], want)))
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {n} queries differ: {bad[:3]}"
    assert worst < 1e-9          # the jplace prints 10 decimals of values around -1000


def test_long_windows_global_scratch_variant(built):
    """Windows of more than 512 sites (the sumtables no longer fit tensor or shared memory for six warps): the
    global-scratch variant of the thorough kernel, several queries per edge, against the oracle."""
    ds = built.synth.dataset(T=20, n_sites=1400, n_queries=48, window=1100, seed_tree=8, seed_q=9)
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], ds["model"])
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    opts = built.capi.default_options()
    out, counts = ctx.place_chunk(case.query_rows, opts)
    for qi in range(0, len(case.qseqs), 4):
        want = case.placer.place(case.qseqs[qi])
        got = out[qi][:counts[qi]]
        assert [int(g["branch_id"]) for g in got] == [p.edge for p in want], (qi, got, want)
        for g, p in zip(got, want):
            assert abs(g["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (qi, g, p)
            assert abs(g["lwr"] - p.lwr) <= 1e-6
            assert abs(g["pendant_length"] - p.pendant) <= 1e-5 and abs(g["distal_length"] - p.distal) <= 1e-5
    ctx.close()
