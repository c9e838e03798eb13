"""GPU parity tests: every stage of the CUDA path, called through the C ABI (capi.py), against the
CPU oracle on the same inputs, and the end result against the reference's recorded placements.

Tolerances (floating point, north_star): log-likelihoods 1e-6 relative (we check much tighter
where the arithmetic is a straight restatement), LWR 1e-6 absolute, branch lengths 1e-4 absolute,
identical edge rankings.
"""
import ctypes as C

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg1(built):
    case = helpers.cfg1_case()
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    yield case, ctx
    ctx.close()


@pytest.fixture(scope="module")
def synth64(built):
    case = helpers.synth64_case()
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    case.placer.build_lookup()
    yield case, ctx
    ctx.close()


def _check_clvs(case, ctx):
    T = len(case.tree.tips)
    for uid, side in case.ref.sides.items():
        node = ctx.ids[uid]
        clv, sc = ctx.get_clv(node)
        if node < T:
            S, R = case.model.states, case.model.rate_cats
            want = ((side.tip[:, None] >> np.arange(S)[None, :]) & 1).astype(float)
            assert np.array_equal(clv.reshape(case.n, R, S), np.repeat(want[:, None, :], R, axis=1))
            assert not sc.any()
        else:
            assert np.allclose(clv, side.clv, rtol=1e-12, atol=0), f"clv of node {node}"
            assert np.array_equal(sc, side.scaler)


def test_clv_precompute_matches_oracle(cfg1):
    _check_clvs(*cfg1)


def test_clv_precompute_matches_oracle_synth(synth64):
    _check_clvs(*synth64)


def test_tree_logl_equal_on_every_edge(synth64):
    case, ctx = synth64
    want = case.ref.tree_logl(0)
    vals = [ctx.edge_loglikelihood(e) for e in range(case.tree.num_branches)]
    assert np.allclose(vals, want, rtol=1e-11, atol=0)


def _check_lookup(case, ctx):
    o = helpers.oracle()
    lk = case.placer.lookup if case.placer.lookup is not None else case.placer.build_lookup()
    for e in range(case.tree.num_branches):
        got = ctx.get_lookup(e)
        assert np.allclose(got, lk[e], rtol=1e-11, atol=1e-11), f"lookup of edge {e}"


def test_lookup_matches_oracle(cfg1):
    _check_lookup(*cfg1)


def test_lookup_matches_oracle_synth(synth64):
    _check_lookup(*synth64)


def _check_preplace(case, ctx):
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    got = ctx.get_prescores()
    want = np.stack([case.placer.preplace(s) for s in case.qseqs])
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    return got, want


def test_preplace_matches_oracle(cfg1):
    _check_preplace(*cfg1)


def test_preplace_select_matches_oracle_synth(synth64, built):
    case, ctx = synth64
    got, want = _check_preplace(case, ctx)
    opts = built.capi.default_options()
    n_pairs = ctx.select(opts)
    q, e, _ = ctx.get_pairs(raw=False)
    assert len(q) == n_pairs
    mine = {}
    for qi, ei in zip(q, e):
        mine.setdefault(int(qi), set()).add(int(ei))
    for qi in range(len(case.qseqs)):
        assert mine[qi] == set(case.placer.candidates(want[qi])), f"candidates of query {qi}"


def test_thorough_all_pairs_matches_oracle(cfg1, built):
    case, ctx = cfg1
    opts = built.capi.default_options(prescoring=0)
    ctx.upload_queries(case.query_rows)
    n_pairs = ctx.select(opts)
    assert n_pairs == len(case.qseqs) * case.tree.num_branches
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    for qi, ei, r in zip(q, e, raw):
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-9 * abs(p.logl), (qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= 1e-6, (qi, ei, r, p)
        assert abs(r["distal_length"] - p.distal) <= 1e-6, (qi, ei, r, p)


def test_thorough_candidates_match_oracle_synth(synth64, built):
    case, ctx = synth64
    opts = built.capi.default_options()
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    ctx.select(opts)
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    worst = 0.0
    for qi, ei, r in zip(q, e, raw):
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= 1e-5, (qi, ei, r, p)
        assert abs(r["distal_length"] - p.distal) <= 1e-5, (qi, ei, r, p)
        worst = max(worst, abs(r["likelihood"] - p.logl) / abs(p.logl))
    print("worst relative logl difference", worst)


@pytest.mark.parametrize("mname,model", [("gtrg", helpers.GTRG), ("gtrb", helpers.GTR_B)])
def test_cfg1_placements_match_reference(built, mname, model):
    gold = helpers.golden("cfg1")
    case = helpers.cfg1_case(model)
    ctx = helpers.make_context(case)
    runs = {
        "default": dict(),
        "noheur_all": dict(prescoring=0, support_threshold=0.0, filter_max=13),
        "heur_all": dict(support_threshold=0.0, filter_max=13),
        "acc": dict(filter_acc_lwr=1, support_threshold=0.999, filter_max=5),
    }
    for rname, kw in runs.items():
        opts = built.capi.default_options(**kw)
        out, counts = ctx.place_chunk(case.query_rows, opts)
        got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
        want = gold[f"{mname}_{rname}"]["placements"]
        for name in want:
            helpers.assert_placements_close(got[name], want[name], f"{mname}/{rname}/{name}")
    ctx.close()


def test_synth64_placements_match_reference(synth64, built):
    case, ctx = synth64
    gold = helpers.golden("synth64")["default"]["placements"]
    out, counts = ctx.place_chunk(case.query_rows, built.capi.default_options())
    got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
    bad = []
    for name in gold:
        try:
            helpers.assert_placements_close(got[name], gold[name], name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(gold)} queries differ from the reference: {bad[:3]}"


@pytest.mark.parametrize("rname,kw", [("fix_heur", dict(heuristic=1, prescoring_threshold=0.05)),
                                      ("baseball", dict(heuristic=2))])
def test_other_heuristics_match_reference(synth64, built, rname, kw):
    # -G / --baseball-heur: candidate sets against the oracle, placements against the reference
    case, ctx = synth64
    o = helpers.oracle()
    opts = built.capi.default_options(**kw)
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    ctx.select(opts)
    q, e, _ = ctx.get_pairs(raw=False)
    mine = {}
    for qi, ei in zip(q, e):
        mine.setdefault(int(qi), set()).add(int(ei))
    placer = o.Placer(case.ref, o.Options(**{k: v for k, v in kw.items()}))
    placer.lookup = case.placer.lookup
    for qi, seq in enumerate(case.qseqs):
        assert mine[qi] == set(placer.candidates(placer.preplace(seq))), f"{rname}: candidates of query {qi}"
    gold = helpers.golden("synth64")[rname]["placements"]
    out, counts = ctx.place_chunk(case.query_rows, opts)
    got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
    bad = []
    for name in gold:
        try:
            helpers.assert_placements_close(got[name], gold[name], name)
        except AssertionError as err:
            bad.append(str(err))
    assert not bad, f"{rname}: {len(bad)} of {len(gold)} queries differ from the reference: {bad[:3]}"


def test_uploaded_clvs_equal_computed(cfg1, built):
    # drop-in for Tree::get_clv: host-owned CLVs (here the oracle's) handed over unchanged
    case, ctx0 = cfg1
    ctx = helpers.make_context(case, compute=False)
    T = len(case.tree.tips)
    ctx.upload_clvs([(ctx.ids[uid], s.clv, s.scaler) for uid, s in case.ref.sides.items() if ctx.ids[uid] >= T])
    out, counts = ctx.place_chunk(case.query_rows, built.capi.default_options())
    out0, counts0 = ctx0.place_chunk(case.query_rows, built.capi.default_options())
    assert np.array_equal(counts, counts0)
    a, b = helpers.records_to_lists(out, counts), helpers.records_to_lists(out0, counts0)
    for x, y in zip(a, b):
        helpers.assert_placements_close(x, y, "uploaded vs computed", logl_rel=1e-10, lwr_abs=1e-9, len_abs=1e-7)
    ctx.close()


def test_edge_cases(synth64, built):
    case, ctx = synth64
    capi = built.capi
    opts = capi.default_options()
    # empty chunk
    out, counts = ctx.place_chunk(np.zeros((0, case.n), dtype=np.uint8), opts)
    assert out.shape[0] == 0
    # invalid character and all-gap query are errors (Tiny_Tree.cpp:145-156)
    rows = case.query_rows[:3].copy()
    rows[1, 10] = ord('!')
    with pytest.raises(capi.EpaError) as ei:
        ctx.place_chunk(rows, opts)
    assert ei.value.code == capi.EPA_ERR_QUERY and "query 1" in str(ei.value)
    rows = case.query_rows[:3].copy()
    rows[2, :] = ord('-')
    with pytest.raises(capi.EpaError) as ei:
        ctx.place_chunk(rows, opts)
    assert ei.value.code == capi.EPA_ERR_QUERY and "query 2" in str(ei.value)
    # ragged windows: one site, full width, lower case, ambiguity codes
    rows = np.full((4, case.n), ord('-'), dtype=np.uint8)
    rows[0, 17] = ord('A')
    rows[1, :] = np.frombuffer(case.qseqs[0].replace('-', 'N').lower().encode(), dtype=np.uint8)
    rows[2, 5:60] = np.frombuffer(("ACGTRYKMSWBDHVN" * 4)[:55].encode(), dtype=np.uint8)
    rows[3, case.n - 40:] = np.frombuffer(case.qseqs[1].replace('-', 'A')[:40].encode(), dtype=np.uint8)
    out, counts = ctx.place_chunk(rows, opts)
    got = helpers.records_to_lists(out, counts)
    for qi in range(4):
        want = case.placer.place(bytes(rows[qi]).decode().upper())
        helpers.assert_placements_close(got[qi], [(p.edge, p.logl, p.lwr, p.distal, p.pendant) for p in want],
                                        f"ragged query {qi}")
    # state machine
    ctx2 = helpers.make_context(case, compute=False)
    with pytest.raises(capi.EpaError) as ei:
        ctx2.build_lookup()
    assert ei.value.code == capi.EPA_ERR_STATE
    ctx2.close()


# ---------------------------------------------------------------------------------------------
#  amino acids (20-state path, LG+G4)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def synthaa(built):
    case = helpers.synthaa_case()
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    case.placer.build_lookup()
    yield case, ctx
    ctx.close()


def test_aa_clvs_and_tree_logl(synthaa):
    case, ctx = synthaa
    _check_clvs(case, ctx)
    want = case.ref.tree_logl(0)
    vals = [ctx.edge_loglikelihood(e) for e in range(case.tree.num_branches)]
    assert np.allclose(vals, want, rtol=1e-11, atol=0)


def test_aa_lookup_and_preplace(synthaa, built):
    case, ctx = synthaa
    _check_lookup(case, ctx)
    got, want = _check_preplace(case, ctx)
    n_pairs = ctx.select(built.capi.default_options())
    q, e, _ = ctx.get_pairs(raw=False)
    mine = {}
    for qi, ei in zip(q, e):
        mine.setdefault(int(qi), set()).add(int(ei))
    for qi in range(len(case.qseqs)):
        assert mine[qi] == set(case.placer.candidates(want[qi])), f"candidates of query {qi}"


def test_aa_thorough_matches_oracle(synthaa, built):
    case, ctx = synthaa
    opts = built.capi.default_options(prescoring=0)
    ctx.upload_queries(case.query_rows[:4])
    ctx.select(opts)
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    for qi, ei, r in zip(q, e, raw):
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-9 * abs(p.logl), (qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= 1e-6, (qi, ei, r, p)
        assert abs(r["distal_length"] - p.distal) <= 1e-6, (qi, ei, r, p)


def test_aa_tensor_core_kernel_and_dfma_kernel_agree(synthaa, built):
    """The amino-acid thorough placement runs on the fp64 tensor-core kernel (DMMA.884, sumtable in tensor memory,
    kernels_blo_aa.cuh); EPA_B200_AA_DFMA routes a context to the DFMA kernel (kernels_blo_generic.cuh). Both are
    checked against the oracle above / here, they must not be bit-identical (different summation orders), and the
    launch counters show that two different kernels ran."""
    case, ctx = synthaa
    import os
    os.environ["EPA_B200_AA_DFMA"] = "1"
    try:
        ctx2 = helpers.make_context(case)
        ctx2.build_lookup()
    finally:
        del os.environ["EPA_B200_AA_DFMA"]
    opts = built.capi.default_options(prescoring=0)
    res = []
    for c in (ctx, ctx2):
        c.upload_queries(case.query_rows[:6])
        c.select(opts)
        c.place_pairs(opts)
        res.append(c.get_pairs())
    (q1, e1, r1), (q2, e2, r2) = res
    assert np.array_equal(q1, q2) and np.array_equal(e1, e2)
    assert np.allclose(r1["likelihood"], r2["likelihood"], rtol=1e-10, atol=0)
    assert np.allclose(r1["pendant_length"], r2["pendant_length"], rtol=0, atol=1e-7)
    assert np.allclose(r1["distal_length"], r2["distal_length"], rtol=0, atol=1e-7)
    assert not np.array_equal(r1["likelihood"], r2["likelihood"]), "both contexts took the same kernel"
    for qi, ei, r in list(zip(q2, e2, r2))[::7]:
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-9 * abs(p.logl), (qi, ei, r, p)
    ctx2.close()


def test_aa_placements_match_reference(synthaa, built):
    case, ctx = synthaa
    gold = helpers.golden("synthaa")
    out, counts = ctx.place_chunk(case.query_rows, built.capi.default_options())
    got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
    bad = []
    for name, want in gold["default"]["placements"].items():
        try:
            helpers.assert_placements_close(got[name], want, name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(gold['default']['placements'])} AA queries differ from the reference: {bad[:3]}"


# ---------------------------------------------------------------------------------------------
# Dense chunk: enough begin-sorted queries per tile for the tensor-core preplacement (tcgen05
# digit GEMM), the single-scan selection with staged candidates, the first-round BLO tables and
# the TMEM-backed sumtables to be the paths that run. Seeded synthetic data, oracle as the checker.
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dense(built):
    ds = built.synth.dataset(T=24, n_sites=400, n_queries=6000, window=120, seed_tree=3, seed_q=4)
    q = ds["queries"].copy()
    rng = np.random.default_rng(11)
    for i in range(0, len(q), 5):                      # interior gaps and N: the fully ambiguous column
        cols = np.flatnonzero(q[i] != ord("-"))
        pick = rng.choice(cols[1:-1], size=4, replace=False)
        q[i, pick[:2]] = ord("-")
        q[i, pick[2:]] = ord("N")
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], q, ds["model"])
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    case.placer.build_lookup()
    yield case, ctx
    ctx.close()


def test_dense_chunk_tensor_core_preplace_matches_oracle(dense, built):
    case, ctx = dense
    ctx.upload_queries(case.query_rows)
    l0 = ctx.launch_count()
    ctx.preplace()
    got = ctx.get_prescores()
    sample = list(range(0, len(case.qseqs), 97)) + [1, 5, len(case.qseqs) - 1]
    want = {qi: case.placer.preplace(case.qseqs[qi]) for qi in sample}
    for qi in sample:
        # fixed-point table: 2^-38 per site, see kernels_preplace_mma.cuh (tolerance 1e-11 relative)
        assert np.allclose(got[qi], want[qi], rtol=1e-11, atol=0), f"prescores of query {qi}"
    # the same chunk through the shared-memory kernels (tensor-core path switched off per context)
    import os
    os.environ["EPA_B200_NO_MMA"] = "1"
    try:
        ctx2 = helpers.make_context(case)
        ctx2.build_lookup()
    finally:
        del os.environ["EPA_B200_NO_MMA"]
    ctx2.upload_queries(case.query_rows)
    ctx2.preplace()
    ref = ctx2.get_prescores()
    assert not np.array_equal(got, ref), "both contexts took the same kernel: the tensor-core path did not run"
    assert np.allclose(got, ref, rtol=1e-11, atol=0)
    # candidate sets: staged single-scan selection vs the oracle on its own prescores
    opts = built.capi.default_options()
    n_pairs = ctx.select(opts)
    q, e, _ = ctx.get_pairs(raw=False)
    assert len(q) == n_pairs
    mine = {}
    for qi, ei in zip(q, e):
        mine.setdefault(int(qi), set()).add(int(ei))
    for qi in sample:
        assert mine[qi] == set(case.placer.candidates(want[qi])), f"candidates of query {qi}"
    ctx2.select(opts)
    q2, e2, _ = ctx2.get_pairs(raw=False)
    assert np.array_equal(q, q2) and np.array_equal(e, e2), "candidate lists differ between the two preplacement paths"
    ctx2.close()


def test_fused_epilogue_selection_equals_the_two_kernel_path(dense, built):
    """The tensor-core preplacement selecting in its epilogue (epa_hint_selection: running maximum, rescaled sum of
    exponentials and the best eight scores of each half row, no score matrix) against the unfused path (score
    matrix + select_count): identical pair lists, for a sharp threshold and for one that makes many queries
    overflow the in-register lists and take the unfused kernels."""
    case, ctx0 = dense
    import os
    os.environ["EPA_B200_FUSED_SELECT"] = "1"          # opt-in path (read at context creation)
    try:
        ctx = helpers.make_context(case)
        ctx.build_lookup()
    finally:
        del os.environ["EPA_B200_FUSED_SELECT"]
    for thresh in (0.99999, 0.9999999999):
        opts = built.capi.default_options(prescoring_threshold=thresh)
        ctx0.upload_queries(case.query_rows)
        ctx0.hint_selection(opts)                       # ignored without the switch
        ctx0.preplace()
        ctx0.get_prescores()
        n0 = ctx0.select(opts)
        q0, e0, _ = ctx0.get_pairs(raw=False)
        ctx.upload_queries(case.query_rows)
        ctx.hint_selection(opts)
        ctx.preplace()
        with pytest.raises(built.capi.EpaError, match="not materialised"):
            ctx.get_prescores()
        n1 = ctx.select(opts)
        q1, e1, _ = ctx.get_pairs(raw=False)
        assert n0 == n1 and np.array_equal(q0, q1) and np.array_equal(e0, e1), thresh
        per_query = np.bincount(q1, minlength=len(case.qseqs))
        if thresh > 0.999999:
            assert (per_query > 8).any(), "no query needed more than eight candidates: the left-over path did not run"
        # a select with other options than announced falls back to the unfused kernels
        ctx.upload_queries(case.query_rows)
        ctx.hint_selection(opts)
        ctx.preplace()
        other = built.capi.default_options(heuristic=1, prescoring_threshold=0.1)
        n2 = ctx.select(other)
        ctx.upload_queries(case.query_rows)
        ctx.preplace()
        assert ctx.select(other) == n2
    ctx.close()


def test_ambiguity_codes_take_the_tensor_core_path(built):
    """Queries with a few IUPAC ambiguity codes (R, Y, K, M, S, W, B, D, H, V) are scored by the tensor-core kernel with
    those sites as N plus a per-site correction; queries with more than 16 of them take the shared-memory kernel. Both
    against the oracle's scores and candidate sets, and against a context without the tensor-core path."""
    ds = built.synth.dataset(T=24, n_sites=400, n_queries=3000, window=150, seed_tree=13, seed_q=14)
    q = ds["queries"].copy()
    rng = np.random.default_rng(5)
    codes = np.frombuffer(b"RYKMSWBDHV", dtype=np.uint8)
    for i in range(len(q)):
        cols = np.flatnonzero(q[i] != ord("-"))
        k = 0 if i % 3 == 0 else (1 + i % 4 if i % 50 else 30)          # none / 1..4 / 30 (beyond the cap)
        if k:
            pick = rng.choice(cols, size=k, replace=False)
            q[i, pick] = codes[rng.integers(0, len(codes), size=k)]
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], q, ds["model"])
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    case.placer.build_lookup()
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    got = ctx.get_prescores()
    import os
    os.environ["EPA_B200_NO_MMA"] = "1"
    try:
        ctx2 = helpers.make_context(case)
        ctx2.build_lookup()
    finally:
        del os.environ["EPA_B200_NO_MMA"]
    ctx2.upload_queries(case.query_rows)
    ctx2.preplace()
    ref = ctx2.get_prescores()
    assert np.allclose(got, ref, rtol=1e-11, atol=0)
    amb_rows = [i for i in range(len(q)) if i % 3 and i % 50]
    assert not np.array_equal(got[amb_rows], ref[amb_rows]), "the ambiguous queries did not take the tensor-core kernel"
    beyond = [i for i in range(len(q)) if i % 3 and i % 50 == 0]
    assert np.array_equal(got[beyond], ref[beyond]), "queries beyond the cap must take the shared-memory kernel"
    sample = amb_rows[::97] + beyond[:2]
    want = {qi: case.placer.preplace(case.qseqs[qi]) for qi in sample}
    for qi in sample:
        assert np.allclose(got[qi], want[qi], rtol=1e-11, atol=0), f"prescores of query {qi}"
    opts = built.capi.default_options()
    ctx.select(opts)
    ctx2.select(opts)
    q1, e1, _ = ctx.get_pairs(raw=False)
    q2, e2, _ = ctx2.get_pairs(raw=False)
    assert np.array_equal(q1, q2) and np.array_equal(e1, e2)
    out, counts = ctx.place_chunk(case.query_rows[amb_rows[:40]], opts)
    for k, qi in enumerate(amb_rows[:40:8]):
        wantp = case.placer.place(case.qseqs[qi])
        gotp = out[8 * k][:counts[8 * k]]
        assert [int(g["branch_id"]) for g in gotp] == [p.edge for p in wantp]
        for g, p in zip(gotp, wantp):
            assert abs(g["likelihood"] - p.logl) <= 1e-8 * abs(p.logl)
    ctx.close(); ctx2.close()


def test_dense_chunk_placements_match_oracle(dense, built):
    case, ctx = dense
    opts = built.capi.default_options()
    out, counts = ctx.place_chunk(case.query_rows, opts)
    for qi in list(range(0, len(case.qseqs), 211)) + [5, 10]:
        want = case.placer.place(case.qseqs[qi])
        got = out[qi][:counts[qi]]
        assert [int(g["branch_id"]) for g in got] == [p.edge for p in want], (qi, got, want)
        for g, p in zip(got, want):
            assert abs(g["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (qi, g, p)
            assert abs(g["lwr"] - p.lwr) <= 1e-6
            assert abs(g["pendant_length"] - p.pendant) <= 1e-5 and abs(g["distal_length"] - p.distal) <= 1e-5


# ---------------------------------------------------------------------------------------------
# Per-rate scalers (EPA_FLAG_RATE_SCALERS; the reference's --rate-scalers on / auto above 2000 tips)
# with and without the reference's scaler window offset (EPA_FLAG_BUGCOMPAT_FOCUS). The oracle's
# per-rate path is pinned against the reference in tests/test_oracle_rate_scalers.py.
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=[True, False], ids=["bugcompat", "corrected"])
def rate300(built, request):
    import json
    import os
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements.json")))
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], ds["model"],
                                    per_rate=True, bugcompat=request.param, column_mask=True)
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    yield case, ctx, (g["placements"] if request.param else None)
    ctx.close()


def test_per_rate_clvs_lookup_and_tree_logl(rate300):
    case, ctx, _ = rate300
    _check_clvs(case, ctx)
    want = case.ref.tree_logl(0)
    vals = [ctx.edge_loglikelihood(e) for e in range(0, case.tree.num_branches, 37)]
    assert np.allclose(vals, want, rtol=1e-11, atol=0)
    lk = case.placer.build_lookup()
    for e in range(0, case.tree.num_branches, 23):
        assert np.allclose(ctx.get_lookup(e), lk[e], rtol=1e-11, atol=1e-11), f"lookup of edge {e}"


def test_per_rate_thorough_and_placements(rate300, built):
    case, ctx, ref = rate300
    opts = built.capi.default_options()
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    ctx.select(opts)
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    for qi, ei, r in zip(q, e, raw):
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= 1e-5 and abs(r["distal_length"] - p.distal) <= 1e-5, (qi, ei, r, p)
    out, counts = ctx.place_chunk(case.query_rows, opts)
    for qi, name in enumerate(case.qnames):
        got = out[qi][:counts[qi]]
        want = case.placer.place(case.qseqs[qi])
        assert [int(g["branch_id"]) for g in got] == [p.edge for p in want], name
        if ref is not None:
            # bug-compatible mode: the reference's own recorded placements
            w = ref[name]
            assert [int(g["branch_id"]) for g in got] == [int(x[0]) for x in w], name
            for g, x in zip(got, w):
                assert abs(g["likelihood"] - x[1]) <= 1e-6 * abs(x[1])
                assert abs(g["lwr"] - x[2]) <= 1e-6 and abs(g["distal_length"] - x[3]) <= 1e-4 and abs(g["pendant_length"] - x[4]) <= 1e-4


def test_mid_length_windows_tmem16_rows(built):
    """Windows of 257..512 sites: the thorough kernel gives 16 tensor-memory rows to four warps and the
    tensor-core preplacement splits K into two passes. Checked against the oracle on a sample."""
    ds = built.synth.dataset(T=24, n_sites=640, n_queries=900, window=350, seed_tree=5, seed_q=6)
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], ds["model"])
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    opts = built.capi.default_options()
    out, counts = ctx.place_chunk(case.query_rows, opts)
    for qi in range(0, len(case.qseqs), 60):
        want = case.placer.place(case.qseqs[qi])
        got = out[qi][:counts[qi]]
        assert [int(g["branch_id"]) for g in got] == [p.edge for p in want], (qi, got, want)
        for g, p in zip(got, want):
            assert abs(g["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (qi, g, p)
            assert abs(g["lwr"] - p.lwr) <= 1e-6
            assert abs(g["pendant_length"] - p.pendant) <= 1e-5 and abs(g["distal_length"] - p.distal) <= 1e-5
    ctx.close()


@pytest.mark.parametrize("bugcompat", [True, False], ids=["bugcompat", "corrected"])
def test_per_rate_eight_categories(built, bugcompat):
    """GTR+G8 with per-rate scalers: the 4-lanes-per-site DNA kernel and the generic lookup build, against the
    oracle in both scaler-window modes and (bug-compatible mode) against the reference's recorded placements."""
    import json
    import os
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_rate2.json")))["dna8"]
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model"],
                                    per_rate=True, bugcompat=bugcompat, column_mask=True)
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_clvs(case, ctx)
    lk = case.placer.build_lookup()
    for e in range(0, case.tree.num_branches, 23):
        assert np.allclose(ctx.get_lookup(e), lk[e], rtol=1e-11, atol=1e-11), f"lookup of edge {e}"
    for sliding in (1, 0):
        o = helpers.oracle()
        case.placer = o.Placer(case.ref, o.Options(sliding_blo=bool(sliding)))
        case.placer.lookup = lk
        opts = built.capi.default_options(sliding_blo=sliding)
        ctx.upload_queries(case.query_rows)
        ctx.preplace()
        ctx.select(opts)
        ctx.place_pairs(opts)
        q, e, raw = ctx.get_pairs()
        for qi, ei, r in zip(q, e, raw):
            p = case.placer.thorough(case.qseqs[qi], int(ei))
            assert abs(r["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (sliding, qi, ei, r, p)
            assert abs(r["pendant_length"] - p.pendant) <= 1e-5 and abs(r["distal_length"] - p.distal) <= 1e-5, (sliding, qi, ei, r, p)
    if bugcompat:
        out, counts = ctx.place_chunk(case.query_rows, built.capi.default_options())
        got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
        for name, want in g["placements"].items():
            helpers.assert_placements_close(got[name], want, name)
    ctx.close()


@pytest.mark.parametrize("fixture,bugcompat", [("ladder", True), ("ladder", False), ("rate2", True)],
                         ids=["ladder-bugcompat", "ladder-corrected", "random-bugcompat"])
def test_per_rate_amino_acids(built, fixture, bugcompat):
    """Amino acids with per-rate scalers (--rate-scalers on, auto above 2000 tips): the reference runs libpll's
    generic kernels, whose tip-inner update rescales whole sites and counts them in entry [site index] of the
    [site][rate] array (oracle pinned in tests/test_oracle_rate_scalers.py). On the ladder tree those rescalings
    happen in the reference CLVs, the lookup tables and the tiny trees. Every stage against the oracle, the end
    result (bug-compatible scaler window) against the reference's recorded placements, incl. --no-heur."""
    import json
    import os
    if fixture == "ladder":
        g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_aa_ladder.json")))
    else:
        g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_rate2.json")))["aa"]
    ds = built.synth.dataset(**g["dataset"])
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], g["model"],
                                    per_rate=True, bugcompat=bugcompat, column_mask=True)
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_clvs(case, ctx)
    # (with misplaced counters the tree log-likelihood is not the same on every edge: edge by edge against the oracle)
    for e in range(0, case.tree.num_branches, 37):
        assert np.isclose(ctx.edge_loglikelihood(e), case.ref.tree_logl(e), rtol=1e-11, atol=0), f"tree logl at edge {e}"
    lk = case.placer.build_lookup()
    for e in range(0, case.tree.num_branches, 7):
        assert np.allclose(ctx.get_lookup(e), lk[e], rtol=1e-11, atol=1e-11), f"lookup of edge {e}"
    opts = built.capi.default_options()
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    ctx.select(opts)
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    for qi, ei, r in zip(q, e, raw):
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), (qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= 1e-5 and abs(r["distal_length"] - p.distal) <= 1e-5, (qi, ei, r, p)
    if bugcompat:
        out, counts = ctx.place_chunk(case.query_rows, opts)
        got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
        for name, w in g["placements"].items():
            helpers.assert_placements_close(got[name], w, name)
        if "placements_no_heur" in g:
            out, counts = ctx.place_chunk(case.query_rows, built.capi.default_options(prescoring=0))
            got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
            for name, w in g["placements_no_heur"].items():
                helpers.assert_placements_close(got[name], w, name)
    # --raxml-blo: every pair against the oracle, placements against the reference's
    o = helpers.oracle()
    case.placer = o.Placer(case.ref, o.Options(sliding_blo=False))
    case.placer.lookup = lk
    opts = built.capi.default_options(sliding_blo=0)
    ctx.upload_queries(case.query_rows)
    ctx.preplace()
    ctx.select(opts)
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    for qi, ei, r in zip(q, e, raw):
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= 1e-8 * abs(p.logl), ("raxml", qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= 1e-5 and abs(r["distal_length"] - p.distal) <= 1e-5, ("raxml", qi, ei, r, p)
    if bugcompat and "placements_raxml_blo" in g:
        out, counts = ctx.place_chunk(case.query_rows, opts)
        got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
        for name, w in g["placements_raxml_blo"].items():
            helpers.assert_placements_close(got[name], w, name)
    ctx.close()


def test_per_rate_amino_acids_files_to_jplace(built, tmp_path):
    """The command-line program with --rate-scalers on on the amino-acid ladder data set: files -> jplace against
    the placements the reference wrote for the same files and flags."""
    import json
    import os
    import subprocess
    g = json.load(open(os.path.join(helpers.GOLDEN, "rate300", "reference_placements_aa_ladder.json")))
    ds = built.synth.dataset(**g["dataset"])
    tf, sf, qf = built.synth.write_dataset(ds, str(tmp_path / "in"))
    exe = os.path.join(helpers.ROOT, "epa-ng_b200", "epa-ng-b200")
    out = str(tmp_path / "out")
    subprocess.run([exe, "-t", tf, "-s", sf, "-q", qf, "-m", g["model"], "-w", out, "--redo", "--rate-scalers", "on"],
                   check=True, stdout=subprocess.DEVNULL)
    doc = json.load(open(os.path.join(out, "epa_result.jplace")))
    got = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
    for name, want in g["placements"].items():
        helpers.assert_placements_close(got[name], want, name)


def test_two_contexts_with_different_models_alternate_on_one_device(cfg1, synth64, built):
    """The model tables of the thorough kernels live in one constant-memory symbol per device: a context rebinds it
    before it launches when another context used the device in between."""
    (case_a, ctx_a), (case_b, ctx_b) = cfg1, synth64
    opts = built.capi.default_options()
    a0, ca0 = ctx_a.place_chunk(case_a.query_rows, opts)
    b0, cb0 = ctx_b.place_chunk(case_b.query_rows, opts)
    for _ in range(2):
        a1, ca1 = ctx_a.place_chunk(case_a.query_rows, opts)
        b1, cb1 = ctx_b.place_chunk(case_b.query_rows, opts)
        assert np.array_equal(ca0, ca1) and a0.tobytes() == a1.tobytes()
        assert np.array_equal(cb0, cb1) and b0.tobytes() == b1.tobytes()
    want = case_a.placer.place(case_a.qseqs[0])
    assert [int(g["branch_id"]) for g in a1[0][:ca1[0]]] == [p.edge for p in want]
