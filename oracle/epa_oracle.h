/*
 * oracle/epa_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar, single-threaded restatement of the arithmetic on EPA-ng's per-query
 * placement hot path. It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg have a checker that travels to the GPU box. Nothing in the
 * product (epa-ng_b200/) may include, link or call this file.
 *
 * Parity of this oracle is PINNED against the unmodified reference built by
 * oracle/Makefile.ref (oracle/_ref/epa-ng, oracle/_ref/libpllref.so): see
 * tests/test_oracle_vs_reference.py and the committed vectors in tests/golden/.
 *
 * Layouts follow libpll (all citations relative to /root/reference):
 *   CLV      [site][rate][state]                libs/pll-modules/libs/libpll/src/core_likelihood.c:1419-1459
 *   pmatrix  [rate][i][j] = P(i->j | t*rate)    .../core_pmatrix.c:185-249
 *   eigenvecs / inv_eigenvecs [i*S+j]           .../models.c:394-404
 *   scaler   uint32[site] or uint32[site][rate] .../core_partials.c:690-766
 *   tip      uint32 state mask per site         .../pll.c:875-957 (tipchars resolved through tipmap)
 */
#ifndef EPA_ORACLE_H
#define EPA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int states;                  /* S: 4 or 20 */
  int rate_cats;               /* R */
  int per_rate_scalers;        /* 0: uint32 scaler[site]; 1: scaler[site][rate] (PLL_ATTRIB_RATE_SCALERS) */
  int bugcompat_focus;         /* 1: reproduce the reference's per-rate scaler focus shift */
  const double * eigenvals;    /* [S] */
  const double * eigenvecs;    /* [S*S] */
  const double * inv_eigenvecs;/* [S*S] */
  const double * freqs;        /* [S] */
  const double * rates;        /* [R] */
  const double * weights;      /* [R] */
  double pinv;                 /* proportion of invariant sites (+I), 0 = none (LP/models.c:495-544) */
  const int * invariant;       /* [n] state index of an invariant site or -1 (LP/models.c:651-760);
                                  may be NULL when pinv == 0 */
} orc_model_t;

/* one side of an edge / one child of an update: either a CLV (+ optional scaler) or tip masks */
typedef struct {
  const double * clv;          /* [n][R][S] or NULL when tip != NULL */
  const uint32_t * scaler;     /* NULL = no scaling recorded */
  const uint32_t * tip;        /* [n] state masks, or NULL */
} orc_side_t;

typedef struct {
  double logl;
  double pendant;
  double distal;               /* already rescaled to the original edge (Tiny_Tree.cpp:183-185) */
  int rounds;                  /* smoothing rounds executed (diagnostic) */
  int restored;                /* 1 if the "worse -> restore" exit fired (optimize.cpp:224-232) */
} orc_blo_result_t;

/* ---- model ------------------------------------------------------------------------------- */
/* libpll gamma.c:220-292, mode 0 = mean, 1 = median */
int  orc_gamma_rates(double alpha, int ncat, int median, double * out_rates);
/* libpll models.c:182-410: eigen system of sqrt(pi) Q sqrt(pi)^-1, mean rate 1 */
int  orc_eigen(int S, const double * subst /*[S(S-1)/2]*/, const double * freqs,
               double * eigenvals, double * eigenvecs, double * inv_eigenvecs);
/* libpll models.c:651-760 pll_update_invariant_sites: AND of all tip state masks per site;
   a single remaining state -> its index, otherwise -1 */
void orc_invariant_sites(int S, int n_tips, int n, const uint32_t * tip_masks /*[n_tips][n]*/, int * invariant);
/* libpll core_pmatrix.c:185-249 */
void orc_pmatrix(const orc_model_t * m, double t, double * pmat /*[R][S][S]*/);

/* ---- CLV kernels ------------------------------------------------------------------------- */
/* libpll partials.c:237-291 / core_partials.c (tt :82, ti :202-508, ii :612-766) */
void orc_update_partial(const orc_model_t * m, int n,
                        double * parent_clv, uint32_t * parent_scaler,
                        const orc_side_t * left, const double * lmat,
                        const orc_side_t * right, const double * rmat);
/* libpll core_likelihood.c:351-921 (tip|inner) and :1191-1496 (inner|inner) */
double orc_edge_logl(const orc_model_t * m, int n, const orc_side_t * parent,
                     const orc_side_t * child, const double * pmat, double * persite /*or NULL*/);
/* libpll core_derivatives.c:116-641 */
void orc_sumtable(const orc_model_t * m, int n, const orc_side_t * parent,
                  const orc_side_t * child, double * sumtable /*[n][R][S]*/);
/* libpll core_derivatives.c:643-858 */
void orc_derivatives(const orc_model_t * m, int n, const double * sumtable, double t,
                     double * df, double * ddf);

/* ---- tiny tree / placement --------------------------------------------------------------- */
/* Tiny_Tree.cpp:48-129 + tiny_util.cpp:234-306: inner CLV looking toward the new tip */
void orc_tiny_inner(const orc_model_t * m, int n, const orc_side_t * distal,
                    const orc_side_t * proximal, double orig_length,
                    double * inner_clv, uint32_t * inner_scaler);
/* Tiny_Tree.cpp:18-46,114-128 + Lookup_Store.hpp:73-81: lookup[site][k] for K char masks */
void orc_lookup_build(const orc_model_t * m, int n, const orc_side_t * distal,
                      const orc_side_t * proximal, double orig_length,
                      const uint32_t * char_masks, int K, double * lookup /*[n][K]*/);
/* Lookup_Store.hpp:110-141 (summation order preserved) */
double orc_preplace_score(const double * lookup, int K, const uint8_t * cols /*[n] column per site*/,
                          int begin, int span);
/* Tiny_Tree.cpp:131-218 + optimize.cpp:60-286 + opt_algorithms.c:133-262 */
void orc_place_thorough(const orc_model_t * m, int n, const orc_side_t * distal,
                        const orc_side_t * proximal, double orig_length,
                        const uint32_t * query_tip /*[n]*/, int begin, int span,
                        orc_blo_result_t * out);

/* the same with --raxml-blo (optimize.cpp:274-278 -> pllmod_opt_optimize_branch_lengths_local,
   PM/optimize/pll_optimize.c:778-1097; Newton variant PM/optimize/opt_algorithms.c:281-384) */
void orc_place_thorough_raxml(const orc_model_t * m, int n, const orc_side_t * distal,
                              const orc_side_t * proximal, double orig_length,
                              const uint32_t * query_tip /*[n]*/, int begin, int span,
                              orc_blo_result_t * out);

/* ---- candidate selection / output stage --------------------------------------------------- */
/* set_manipulators.cpp:43-69 */
void orc_lwr(const double * logl, int n, double * lwr);
/* heuristics.hpp:40-64 + set_manipulators.cpp:90-113: indices of candidates, sorted desc by LWR;
   ties broken by lower index (the reference's std::sort leaves ties unspecified) */
int  orc_select_accumulated(const double * lwr, int n, double thresh, int * out_idx);
/* set_manipulators.cpp:131-163 on an already LWR-sorted list: number kept */
int  orc_filter_support(const double * lwr_sorted, int n, double thresh, int min, int max);

#ifdef __cplusplus
}
#endif
#endif
