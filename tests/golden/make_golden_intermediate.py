"""Intermediate vectors of the hot path from the UNMODIFIED libpll (oracle/_ref/libpllref.so, built by
oracle/Makefile.ref), called through ctypes in THIS container: transition matrices at three lengths, an inner
CLV (pll_update_partials on two tips and on tip + inner), the sumtable of an edge (pll_update_sumtable, tip|inner
and inner|inner), first and second derivatives at three lengths (pll_compute_likelihood_derivatives) and the edge
log-likelihood. Scalar kernels (PLL_ATTRIB_ARCH_CPU), DNA, 4 rate categories, per-site scaling.

    python tests/golden/make_golden_intermediate.py  ->  tests/golden/intermediate/libpll_vectors.json
"""
import ctypes as C
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(ROOT, "oracle", "_ref", "libpllref.so")

u, dp, upp = C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint)


class Partition(C.Structure):               # pll_partition_t, LP/pll.h:241-288 (the fields read here)
    _fields_ = [("tips", u), ("clv_buffers", u), ("nodes", u), ("states", u), ("sites", u), ("pattern_weight_sum", u),
                ("rate_matrices", u), ("prob_matrices", u), ("rate_cats", u), ("scale_buffers", u), ("attributes", u),
                ("alignment", C.c_size_t), ("states_padded", u),
                ("clv", C.POINTER(dp)), ("pmatrix", C.POINTER(dp)), ("rates", dp), ("rate_weights", dp),
                ("subst_params", C.POINTER(dp)), ("scale_buffer", C.POINTER(upp)), ("frequencies", C.POINTER(dp)),
                ("prop_invar", dp), ("invariant", C.POINTER(C.c_int)), ("pattern_weights", upp),
                ("eigen_decomp_valid", C.POINTER(C.c_int)), ("eigenvecs", C.POINTER(dp)), ("inv_eigenvecs", C.POINTER(dp)),
                ("eigenvals", C.POINTER(dp))]


class Operation(C.Structure):               # pll_operation_t, LP/pll.h:325-335
    _fields_ = [("parent_clv_index", u), ("parent_scaler_index", C.c_int), ("child1_clv_index", u), ("child1_matrix_index", u),
                ("child1_scaler_index", C.c_int), ("child2_clv_index", u), ("child2_matrix_index", u), ("child2_scaler_index", C.c_int)]


def main():
    L = C.CDLL(LIB)
    L.pll_partition_create.restype = C.POINTER(Partition)
    L.pll_partition_create.argtypes = [u] * 9
    L.pll_compute_edge_loglikelihood.restype = C.c_double
    L.pll_compute_edge_loglikelihood.argtypes = [C.POINTER(Partition), u, C.c_int, u, C.c_int, u, upp, dp]
    L.pll_compute_likelihood_derivatives.argtypes = [C.POINTER(Partition), C.c_int, C.c_int, C.c_double, upp, dp, dp, dp]
    L.pll_update_sumtable.argtypes = [C.POINTER(Partition), u, u, C.c_int, C.c_int, upp, dp]
    L.pll_update_prob_matrices.argtypes = [C.POINTER(Partition), upp, upp, dp, u]
    L.pll_set_tip_states.argtypes = [C.POINTER(Partition), u, C.c_void_p, C.c_char_p]
    rng = random.Random(11)
    n, S, R = 40, 4, 4
    seqs = ["".join(rng.choice("ACGTACGTACGTACGTRYN-") for _ in range(n)) for _ in range(3)]
    subst = [0.676278, 2.012275, 0.478487, 0.753965, 2.406436, 1.0]
    freqs = [0.245629, 0.235012, 0.253054, 0.266305]
    alpha = 0.7
    lengths = [0.07, 0.31, 1.2, 0.004]
    p = L.pll_partition_create(3, 2, S, n, 1, 4, R, 2, 0)            # 3 tips, 2 inner CLVs, scalar kernels
    assert p
    L.pll_set_subst_params(p, 0, (C.c_double * 6)(*subst))
    L.pll_set_frequencies(p, 0, (C.c_double * 4)(*freqs))
    rates = (C.c_double * R)()
    L.pll_compute_gamma_cats(C.c_double(alpha), R, rates, 0)            # PLL_GAMMA_RATES_MEAN
    L.pll_set_category_rates(p, rates)
    nt_map = C.c_void_p.in_dll(L, "pll_map_nt")
    for i, s in enumerate(seqs):
        assert L.pll_set_tip_states(p, i, C.addressof(nt_map), s.encode())
    params = (u * R)(*[0] * R)
    L.pll_update_prob_matrices(p, params, (u * 4)(0, 1, 2, 3), (C.c_double * 4)(*lengths), 4)
    NONE = -1
    # inner 3 = tips 0, 1 (tip-tip); inner 4 = tip 2 + inner 3 (tip-inner)
    ops = (Operation * 2)(Operation(3, 0, 0, 0, NONE, 1, 1, NONE), Operation(4, 1, 2, 2, NONE, 3, 3, 0))
    L.pll_update_partials(p, ops, 2)
    pc = p.contents
    Sp = pc.states_padded
    assert Sp == S

    def arr(ptr, k):
        return [ptr[i] for i in range(k)]

    out = {"sites": n, "states": S, "rate_cats": R, "sequences": seqs, "subst": subst, "freqs": freqs, "alpha": alpha,
           "lengths": lengths, "rates": arr(rates, R),
           "eigenvals": arr(pc.eigenvals[0], S), "eigenvecs": arr(pc.eigenvecs[0], S * S), "inv_eigenvecs": arr(pc.inv_eigenvecs[0], S * S),
           "pmatrix": [arr(pc.pmatrix[i], R * S * S) for i in range(4)],
           "clv3": arr(pc.clv[3], n * R * S), "scaler3": arr(pc.scale_buffer[0], n),
           "clv4": arr(pc.clv[4], n * R * S), "scaler4": arr(pc.scale_buffer[1], n)}
    # sumtables: tip 2 | inner 3 (the pendant edge of a tiny tree) and inner 4 | inner 3
    derivs = {}
    for key, (a, sa, b, sb) in {"tip2_inner3": (2, NONE, 3, 0), "inner4_inner3": (4, 1, 3, 0)}.items():
        sumtable = (C.c_double * (n * R * Sp))()
        assert L.pll_update_sumtable(p, a, b, sa, sb, params, sumtable)
        rec = {"sumtable": list(sumtable), "derivatives": []}
        for t in (0.01, 0.3, 2.5):
            df, ddf = C.c_double(), C.c_double()
            assert L.pll_compute_likelihood_derivatives(p, sa, sb, C.c_double(t), params, sumtable, C.byref(df), C.byref(ddf))
            rec["derivatives"].append([t, df.value, ddf.value])
        derivs[key] = rec
    out["edges"] = derivs
    out["logl_tip2_inner3_matrix1"] = L.pll_compute_edge_loglikelihood(p, 3, 0, 2, NONE, 1, params, None)
    out["logl_inner4_inner3_matrix0"] = L.pll_compute_edge_loglikelihood(p, 4, 1, 3, 0, 0, params, None)
    os.makedirs(os.path.join(HERE, "intermediate"), exist_ok=True)
    json.dump(out, open(os.path.join(HERE, "intermediate", "libpll_vectors.json"), "w"))
    print("wrote", len(json.dumps(out)), "bytes; logl", out["logl_tip2_inner3_matrix1"], out["logl_inner4_inner3_matrix0"])
    L.pll_partition_destroy(p)


if __name__ == "__main__":
    main()
