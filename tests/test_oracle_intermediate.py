"""The oracle's stages against vectors dumped from the UNMODIFIED libpll (oracle/_ref/libpllref.so through ctypes,
tests/golden/make_golden_intermediate.py): eigen system, transition matrices, CLV updates (tip-tip and tip-inner,
scalers), sumtables (tip|inner and inner|inner), first/second derivatives at three branch lengths, edge
log-likelihoods. Pins rows 7a-7f of SURVEY 8a stage by stage, not only through final placements."""
import ctypes as C
import json
import os

import numpy as np

import helpers

G = json.load(open(os.path.join(helpers.GOLDEN, "intermediate", "libpll_vectors.json")))


def _setup():
    o = helpers.oracle()
    n, S, R = G["sites"], G["states"], G["rate_cats"]
    m = o.Model(states=S, subst=np.array(G["subst"]), freqs=np.array(G["freqs"]), alpha=G["alpha"], rate_cats=R,
                rates=o.gamma_rates(G["alpha"], R), weights=np.full(R, 1.0 / R)).finalize()
    tab = o.state_mask_table(S)
    tips = [o.SideData(tip=np.ascontiguousarray(tab[np.frombuffer(s.encode(), dtype=np.uint8)].astype(np.uint32)))
            for s in G["sequences"]]
    return o, m, tips, n, S, R


def test_model_tables_match_libpll():
    o, m, _, n, S, R = _setup()
    assert np.allclose(m.rates, G["rates"], rtol=1e-14, atol=0)
    # eigenvectors are defined up to sign and order: compare what they are used for
    V, Vi, ev = np.array(G["eigenvecs"]).reshape(S, S), np.array(G["inv_eigenvecs"]).reshape(S, S), np.array(G["eigenvals"])
    assert np.allclose(np.sort(m.eigenvals), np.sort(ev), rtol=1e-13, atol=1e-15)
    for t, want in zip(G["lengths"], G["pmatrix"]):
        assert np.allclose(m.pmatrix(t), want, rtol=1e-13, atol=1e-16), t
        # and libpll's own tables reproduce its matrices the way the kernels use them
        for r in range(R):
            P = np.eye(S) + (Vi * np.expm1(ev * G["rates"][r] * t)[None, :]) @ V
            assert np.allclose(P.ravel(), want[r * S * S:(r + 1) * S * S], rtol=1e-12, atol=1e-15)


def _clvs(o, m, tips, n, S, R):
    L = o.lib()
    pm = [m.pmatrix(t) for t in G["lengths"]]
    mc = m.c()
    clv3, sc3 = np.zeros(n * R * S), np.zeros(n, dtype=np.uint32)
    L.orc_update_partial(C.byref(mc), n, o._dp(clv3), o._up(sc3), C.byref(tips[0].c), o._dp(pm[0]), C.byref(tips[1].c), o._dp(pm[1]))
    in3 = o.SideData(clv=clv3, scaler=sc3)
    clv4, sc4 = np.zeros(n * R * S), np.zeros(n, dtype=np.uint32)
    L.orc_update_partial(C.byref(mc), n, o._dp(clv4), o._up(sc4), C.byref(tips[2].c), o._dp(pm[2]), C.byref(in3.c), o._dp(pm[3]))
    in4 = o.SideData(clv=clv4, scaler=sc4)
    return L, mc, pm, in3, in4


def test_clv_updates_match_libpll():
    o, m, tips, n, S, R = _setup()
    _, _, _, in3, in4 = _clvs(o, m, tips, n, S, R)
    assert np.allclose(in3.clv, G["clv3"], rtol=1e-13, atol=0) and list(in3.scaler) == G["scaler3"]
    assert np.allclose(in4.clv, G["clv4"], rtol=1e-13, atol=0) and list(in4.scaler) == G["scaler4"]


def test_sumtables_derivatives_and_logl_match_libpll():
    o, m, tips, n, S, R = _setup()
    L, mc, pm, in3, in4 = _clvs(o, m, tips, n, S, R)
    for key, (a, b) in {"tip2_inner3": (tips[2], in3), "inner4_inner3": (in4, in3)}.items():
        st = np.zeros(n * R * S)
        L.orc_sumtable(C.byref(mc), n, C.byref(a.c), C.byref(b.c), o._dp(st))
        want = np.array(G["edges"][key]["sumtable"])
        # eigenvector signs cancel inside one entry (left and right factor), their order does not matter to the sums
        # over states below; the site sums over the eigen index are what the derivatives read
        got_sites = np.sort(st.reshape(n, R, S), axis=2)
        want_sites = np.sort(want.reshape(n, R, S), axis=2)
        scale = np.abs(want_sites).max(axis=2, keepdims=True)
        assert np.all(np.abs(got_sites - want_sites) <= 1e-11 * scale), key
        for t, df_w, ddf_w in G["edges"][key]["derivatives"]:
            df, ddf = C.c_double(), C.c_double()
            L.orc_derivatives(C.byref(mc), n, o._dp(st), t, C.byref(df), C.byref(ddf))
            assert abs(df.value - df_w) <= 1e-10 * abs(df_w) and abs(ddf.value - ddf_w) <= 1e-10 * abs(ddf_w), (key, t)
    l1 = L.orc_edge_logl(C.byref(mc), n, C.byref(in3.c), C.byref(tips[2].c), o._dp(pm[1]), None)
    l2 = L.orc_edge_logl(C.byref(mc), n, C.byref(in4.c), C.byref(in3.c), o._dp(pm[0]), None)
    assert abs(l1 - G["logl_tip2_inner3_matrix1"]) <= 1e-12 * abs(l1)
    assert abs(l2 - G["logl_inner4_inner3_matrix0"]) <= 1e-12 * abs(l2)
