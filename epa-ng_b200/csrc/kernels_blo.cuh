// kernels_blo.cuh - HOT LOOP B: branch-length optimisation of one (query, edge) pair on the
// three-taxon "tiny tree" (distal D, proximal X, new tip T).
//
// Reference behaviour restated (paths relative to /root/reference, LP = libs/pll-modules/libs/libpll/src,
// PM = libs/pll-modules/src):
//   Tiny_Tree::place (opt)          src/tree/Tiny_Tree.cpp:131-218
//   optimize_branch_triplet         src/core/pll/optimize.cpp:253-286
//   opt_branch_lengths_pplacer      src/core/pll/optimize.cpp:60-248
//   pllmod_opt_minimize_newton      PM/optimize/opt_algorithms.c:86-262
//   sumtable / derivatives          LP/core_derivatives.c:321-471,473-641,643-858
//   CLV update / edge logl          LP/core_partials.c:202-352,612-766, LP/core_likelihood.c:351-578
//   range focus                     src/core/pll/pll_util.cpp:388-418
//
// This file holds the shared definitions of the thorough kernels (constants, BloArgs, BloResult,
// c_model) and the FIRST DNA kernel, which today only runs for 8 rate categories: DNA with 1, 2 or 4
// categories takes the lane = site kernel of kernels_blo_site.cuh, amino acids the CTA-per-pair
// kernel of kernels_blo_generic.cuh.
//
// DNA kernel (S = 4): ONE WARP PER PAIR, two lane mappings.
//  * CLV passes (inner CLV, edge log-likelihood, sumtable build): R lanes share a site
//    (lane % R = rate category), a warp sweeps 32/R sites per step and every lane reads exactly one
//    32-byte sector of each CLV, so the CLV windows stream fully coalesced from L2/HBM. The rate sum
//    is log2(R) xor-shuffles; the per-site log is rotated over the R lanes of a site (one log per
//    lane per R sites instead of one per site).
//  * Newton iterations: lane = site. The sumtable lives in the warp's shared-memory slice as one
//    row of 1 + 3R doubles per site (the eigenvalue-0 component of all rates is folded into one
//    t-independent entry; odd row length = conflict-free lane-per-site reads at immediate offsets),
//    the whole site evaluates in registers without any cross-lane traffic, and only the final
//    (f, f') pair is butterfly-reduced.
// Transition matrices and per-mask tip vectors are rebuilt by the warp whenever a length changes.
// Work items come from an edge-major, window-sorted list in blocks of 32 per CTA, so the warps of a
// CTA work on overlapping CLV windows (L1 reuse). Nothing but 24 bytes per pair leaves the SM.
#pragma once
#include "common.cuh"

namespace epa {

// model tables as constant-bank operands (static indices fold into the DFMA; one translation unit)
__constant__ DevModel c_model;

#define EPA_DEFAULT_PENDANT 0.10536051565782630123   /* -ln(0.9), src/util/constants.hpp:12 */
#define EPA_MIN_BRLEN 1.0e-4                          /* PM/optimize/pll_optimize.h:57 */
#define EPA_MAX_BRLEN 100.0                           /* :58 */
#define EPA_DEFAULT_BRLEN 0.1                         /* :54 */
#define EPA_BLO_EPSILON 1e-1                          /* src/core/pll/optimize.hpp:9 */
#define EPA_NR_MAX_ITERS 30
#define EPA_SMOOTHINGS 32
#ifndef EPA_WORK_BLOCK
#define EPA_WORK_BLOCK 32u
#endif

struct BloResult { double logl, pendant, distal; };

struct BloArgs {
  DevTree tree;
  int n;
  const EdgeDev * edges;
  const uint8_t * codes;          // [nq][n]
  const int * begin;
  const int * span;
  const uint32_t * work;          // explicit mode: (edge, window)-sorted permutation of pair ids
  const uint32_t * pair_q;        // NULL = implicit all-pairs mode
  const uint32_t * pair_e;
  const uint32_t * perm;          // implicit mode: queries sorted by window start
  uint32_t n_pairs;
  uint32_t nq, n_edges;           // implicit mode: item i -> edge i / nq, query perm[i % nq], pair id q*n_edges+e
  unsigned long long * counter;   // dynamic work counter (zeroed before launch)
  BloResult * out;                // [pair id]
  double * scratch;               // global sumtable scratch (GS variant): [total warps][1 + 3R planes]
  int wcap;                       // sites the shared-memory sumtable can hold
  int bugcompat;                  // per-rate scalers: the reference's window offset, R = 8 kernel
  int raxml;                      // 1 = --raxml-blo: the three edges optimised one after the other, unconstrained
                                  // (pllmod_opt_optimize_branch_lengths_local, PM/optimize/pll_optimize.c:778-1097)
};

// --raxml-blo takes the older Newton-Raphson variant with a bisection fallback
// (pllmod_opt_minimize_newton_old, PM/optimize/opt_algorithms.c:281-384). `deriv(x, f, df)` evaluates
// the derivative sums of the current sumtable; `failed` is set where the reference sets pll_errno
// (non-finite derivatives, iteration limit) - the optimiser then gives up on the pair.
template <class Deriv>
__device__ __forceinline__ double newton_old(Deriv && deriv, double x1, double xguess, double x2, double tol, bool & failed)
{
  double df, dx, f, xh, xl, rts, rts_old = 0.0;
  failed = false;
  rts = fmax(fmin(xguess, x2), x1);
  deriv(rts, f, df);
  if (!isfinite(f) || !isfinite(df)) { failed = true; return 0.0; }
  if (df >= 0.0 && fabs(f) < tol) return rts;
  if (f < 0.0) { xl = rts; xh = x2; }
  else { xh = rts; xl = x1; }
  #pragma unroll 1
  for (int i = 1; i <= EPA_NR_MAX_ITERS; ++i)
  {
    rts_old = rts;
    if (df <= 0.0 || ((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0)
    {
      dx = 0.5 * (xh - xl);
      rts = xl + dx;
      if (xl == rts) return rts;
    }
    else
    {
      dx = f / df;
      const double temp = rts;
      rts -= dx;
      if (temp == rts) return rts;
    }
    if (fabs(dx) < tol) return rts_old;
    if (i == EPA_NR_MAX_ITERS) break;
    if (rts < x1) rts = x1;
    deriv(rts, f, df);
    if (!isfinite(f) || !isfinite(df)) { failed = true; return 0.0; }
    if (df > 0.0 && fabs(f) < tol) return rts;
    if (f < 0.0) xl = rts; else xh = rts;
  }
  failed = true;
  return rts_old;
}

// Sumtable row of one site: [stationary part, 3R decaying components], padded to an ODD number of
// doubles so that the lane = site reads of the Newton loop (stride = one row) touch 16 distinct
// bank pairs per half-warp, while every component sits at a compile-time offset inside the row.
__host__ __device__ constexpr int blo_row(int R) { return (1 + 3 * R) | 1; }

template <int R>
struct BloWarpSmem {
  // offsets in doubles inside one warp's slice
  static constexpr int P_D = 0;
  static constexpr int P_P = R * 16;
  static constexpr int P_E = 2 * R * 16;
  static constexpr int TV = 3 * R * 16;              // [4][16 mask positions][R]
  static constexpr int EX = TV + R * 64;             // [R*4] expm1 scratch; [3][3R] diag tables (R = 8)
  static constexpr int SUM = EX + 3 * R * 4;         // [wcap][blo_row(R)]
  __host__ __device__ static constexpr size_t doubles(int wcap)
  {
    return (size_t) SUM + (size_t) wcap * blo_row(R);
  }
};

struct BloCtaSmem {
  double V[16], Vinv[16];             // lane-divergent indexing in warp_pmatrix
  __align__(16) double tipleft[64];   // [mask][j] = sum_{k in mask} pi_k Vinv[k][j]
  unsigned long long q_next, q_end;   // current block of work items
  int q_lock;
};

// P[r][i][j] = delta_ij + sum_k Vinv[i][k] expm1(lambda_k rate_r t) V[k][j]   (LP/core_pmatrix.c:185-249)
template <int R>
__device__ __forceinline__ void warp_pmatrix(const BloCtaSmem & cs, double t, double * P, double * ex, int lane)
{
  for (int idx = lane; idx < R * 4; idx += 32)
    ex[idx] = expm1(c_model.eigenvals[idx & 3] * c_model.rates[idx >> 2] * t);
  __syncwarp();
  for (int idx = lane; idx < R * 16; idx += 32)
  {
    const int r = idx >> 4, i = (idx >> 2) & 3, j = idx & 3;
    double acc = (i == j) ? 1.0 : 0.0;
    #pragma unroll
    for (int k = 0; k < 4; ++k) acc += (cs.Vinv[i * 4 + k] * ex[r * 4 + k]) * cs.V[k * 4 + j];
    P[(i * 4 + j) * R + r] = acc;          // rate-minor: the R lanes of a site read consecutive words
  }
  __syncwarp();
}

// Row position of a state mask inside the tip-vector table: A, C, G, T take positions 0..3 so that
// (for R = 4) their 32-byte rate blocks fall into the four disjoint quarters of the 128-byte bank
// window: a warp's lookups of unambiguous characters are conflict-free and broadcast.
__device__ __forceinline__ int tv_pos(int mask)
{
  // mask:  0  1  2  3  4  5  6  7  8  9 10 11 12 13 14 15
  // pos : 15  0  1  5  2  6  7  8  3  9 10 11 12 13 14  4
  return (int) ((0x4EDCBA938762510Full >> (mask * 4)) & 15ull);
}

// tv[i][pos(mask)][r] = sum_{j in mask} P[r][i][j]  (the pendant matrix applied to a tip state set)
template <int R>
__device__ __forceinline__ void warp_tipvec(const double * P, double * tv, int lane)
{
  for (int idx = lane; idx < R * 64; idx += 32)
  {
    const int r = idx >> 6, mask = (idx >> 2) & 15, i = idx & 3;
    double acc = 0.0;
    #pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((mask >> j) & 1) acc += P[(i * 4 + j) * R + r];
    tv[(i * 16 + tv_pos(mask)) * R + r] = acc;
  }
  __syncwarp();
}

template <int R>
__device__ __forceinline__ double rate_sum(double v)
{
  #pragma unroll
  for (int o = 1; o < R; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int R>
__device__ __forceinline__ uint32_t rate_min(uint32_t v)
{
  #pragma unroll
  for (int o = 1; o < R; o <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// true when all R lanes of the site group report `small`
template <int R>
__device__ __forceinline__ bool group_all(bool small, int lane)
{
  if (R == 1) return small;
  const unsigned ballot = __ballot_sync(0xffffffffu, small);
  const unsigned gm = (1u << R) - 1u;
  return ((ballot >> (lane & ~(R - 1))) & gm) == gm;
}

// first and second derivative sums over the window (LP/core_derivatives.c:643-858), lane = site.
// Plane 0 holds the stationary (lambda = 0) part of every site, sum_r w_r x[r][0], which does not
// depend on t; planes 1.. hold the 3R decaying components x[r][j], j = 1..3.
template <int R>
__device__ __forceinline__ void warp_derivatives(const double * sum, double * ex, int w, double t,
                                              int lane, double & f, double & df)
{
  constexpr int NK = 3 * R, ROW = blo_row(R);
  // diag tables: lane k < 3R computes exp(lambda_j rate_r t), weights folded in
  double e = 0.0, lk = 0.0;
  {
    const int k = lane % NK;
    lk = c_model.eigenvals[1 + k % 3] * c_model.rates[k / 3];
    e = exp(lk * t) * c_model.weights[k / 3];
  }
  double a1 = 0.0, a2 = 0.0;
  if constexpr (R <= 4)
  {
    double d0[NK], d1[NK], d2[NK];
    #pragma unroll
    for (int k = 0; k < NK; ++k)
    {
      const double ek = __shfl_sync(0xffffffffu, e, k);
      const double lkk = __shfl_sync(0xffffffffu, lk, k);
      d0[k] = ek; d1[k] = lkk * ek; d2[k] = lkk * d1[k];
    }
    #pragma unroll 4
    for (int s = lane; s < w; s += 32)
    {
      const double * row = sum + s * ROW;
      double c0 = row[0], c1 = 0.0, c2 = 0.0;
      #pragma unroll
      for (int k = 0; k < NK; ++k)
      {
        const double x = row[k + 1];
        c0 += x * d0[k]; c1 += x * d1[k]; c2 += x * d2[k];
      }
      const double inv = 1.0 / c0;
      const double g1 = -c1 * inv;
      a1 += g1;
      a2 += g1 * g1 - c2 * inv;
    }
  }
  else
  {
    // many rate categories: diag tables through shared memory (broadcast reads)
    __syncwarp();
    if (lane < NK) { ex[lane] = e; ex[NK + lane] = lk * e; ex[2 * NK + lane] = lk * lk * e; }
    __syncwarp();
    for (int s = lane; s < w; s += 32)
    {
      const double * row = sum + s * ROW;
      double c0 = row[0], c1 = 0.0, c2 = 0.0;
      #pragma unroll 8
      for (int k = 0; k < NK; ++k)
      {
        const double x = row[k + 1];
        c0 += x * ex[k]; c1 += x * ex[NK + k]; c2 += x * ex[2 * NK + k];
      }
      const double inv = 1.0 / c0;
      const double g1 = -c1 * inv;
      a1 += g1;
      a2 += g1 * g1 - c2 * inv;
    }
    __syncwarp();
  }
  f = warp_sum(a1);
  df = warp_sum(a2);
}

// One sumtable row (site s, rate r): the stationary component is weighted and summed over the R
// lanes of the site, the decaying ones go to their slots of the site's row.
template <int R>
__device__ __forceinline__ void store_sum_row(double * sum, int s, int r, bool act, double wr,
                                              double st0, double st1, double st2, double st3, double inv)
{
  // +I: the invariant term of the site joins the t-independent entry (LP/core_derivatives.c:676-687)
  const double base = rate_sum<R>(st0 * wr) + inv;
  if (act)
  {
    double * row = sum + s * blo_row(R);
    if (r == 0) row[0] = base;
    row[1 + r * 3] = st1;
    row[2 + r * 3] = st2;
    row[3 + r * 3] = st3;
  }
}

// bounded Newton-Raphson, PM/optimize/opt_algorithms.c:133-262; returns 0.0 on failure
template <int R>
__device__ __forceinline__ double warp_newton(const double * sum, double * ex, int w, int lane,
                                              double xmin, double xguess, double xmax, double tol)
{
  double x = fmax(fmin(xguess, xmax), xmin);
  double xl = xmin, xh = xmax;
  const double dxmax = xmax / EPA_NR_MAX_ITERS;
  int iter = 0;
  for (;;)
  {
    if (iter++ > EPA_NR_MAX_ITERS) return 0.0;
    double f, df;
    warp_derivatives<R>(sum, ex, w, x, lane, f, df);
    if (!isfinite(f) || !isfinite(df)) return 0.0;
    double dx;
    if (df > 0.0)
    {
      if (fabs(f) < tol) return x;
      if (f < 0.0) xl = x; else xh = x;
      dx = -1.0 * f / df;
    }
    else
      dx = -1.0 * f / fabs(df);
    dx = fmax(fmin(dx, dxmax), -dxmax);
    if (x + dx < xl) dx = xl - x;
    if (x + dx > xh) dx = xh - x;
    if (fabs(dx) < tol) return x;
    x += dx;
    x = fmax(fmin(x, xmax), xmin);
  }
}

// Inputs of one (site, rate) unit, loaded one unit ahead of their use (software prefetch: with
// 8-9 resident warps per SM the L2 latency of the CLV stream is not hidden by other warps).
struct UnitIn { double dv[4], xv[4]; int mask; uint32_t scal; double inv; };

// SCALERS: 0 = none, 1 = per-site count of the window site, 2 = per-rate scalers: the lane's own rate
// (sD/sX then address the node's whole [n][R] block; `pr_begin` and `bugcompat` select the entry as in
// kernels_blo_site.cuh: rate_weights)
template <int R, int SCALERS>
__device__ __forceinline__ UnitIn load_unit(const double * __restrict__ D, const double * __restrict__ X,
                                            const uint32_t * __restrict__ sD, const uint32_t * __restrict__ sX,
                                            const uint8_t * __restrict__ qc, int s, int w, int r,
                                            const double * __restrict__ inv_w, int pr_begin = 0, int bugcompat = 0)
{
  const int sc = s < w ? s : w - 1;            // tail lanes re-read the last site; results are discarded
  UnitIn u;
  u.inv = inv_w ? __ldg(inv_w + sc) : 0.0;
  load_vec<4>(D + ((size_t) sc * R + r) * 4, u.dv);
  load_vec<4>(X + ((size_t) sc * R + r) * 4, u.xv);
  u.mask = qc[sc] & 15;
  if (SCALERS == 2)
  {
    const size_t base = (bugcompat ? (size_t) pr_begin + (size_t) sc * R : ((size_t) pr_begin + sc) * R) + r;
    u.scal = __ldg(sD + base) + __ldg(sX + base);
  }
  else
    u.scal = SCALERS == 1 ? __ldg(sD + sc) + __ldg(sX + sc) : 0u;
  return u;
}

// Pass A: inner CLV toward the new tip from (D, X); returns the edge log-likelihood
// new_tip | inner over the window and leaves the pendant sumtable (inner vs tip) in `sum`.
template <int R, bool PR>
__device__ __forceinline__ double warp_pass_tip(const BloCtaSmem & cs, const double * ws, double * sum,
                                             const double * __restrict__ D, const double * __restrict__ X,
                                             const uint32_t * __restrict__ sD, const uint32_t * __restrict__ sX,
                                             const uint8_t * __restrict__ qc, int w, int lane,
                                             const double * __restrict__ inv_w, int pr_begin, int bugcompat)
{
  constexpr int SPW = 32 / R;
  const int r = lane % R, so = lane / R;
  double pd[16], pp[16];
  #pragma unroll
  for (int k = 0; k < 16; ++k) { pd[k] = ws[BloWarpSmem<R>::P_D + k * R + r]; pp[k] = ws[BloWarpSmem<R>::P_P + k * R + r]; }
  const double * tv = ws + BloWarpSmem<R>::TV + r;
  const double wr = c_model.weights[r];
  double acc = 0.0;
  double mine = 1.0, minv = 0.0;
  uint32_t mscal = 0;
  // whole trips of R unit steps: the R lanes of a site take turns at the logarithm
  const int n_units = ((w + SPW * R - 1) / (SPW * R)) * R;
  UnitIn nxt = load_unit<R, PR ? 2 : 1>(D, X, sD, sX, qc, so, w, r, inv_w, pr_begin, bugcompat);
  #pragma unroll 1
  for (int i = 0; i < n_units; ++i)
  {
    const UnitIn cur = nxt;
    const int s = i * SPW + so;
    nxt = load_unit<R, PR ? 2 : 1>(D, X, sD, sX, qc, s + SPW, w, r, inv_w, pr_begin, bugcompat);
    const bool act = s < w;
    double in[4];
    uint32_t scal = cur.scal;
    bool small = true;
    #pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      const double ta = pd[k * 4] * cur.dv[0] + pd[k * 4 + 1] * cur.dv[1] + pd[k * 4 + 2] * cur.dv[2] + pd[k * 4 + 3] * cur.dv[3];
      const double tb = pp[k * 4] * cur.xv[0] + pp[k * 4 + 1] * cur.xv[1] + pp[k * 4 + 2] * cur.xv[2] + pp[k * 4 + 3] * cur.xv[3];
      in[k] = ta * tb;
      small = small && (in[k] < EPA_SCALE_THRESHOLD);
    }
    if constexpr (PR)
    {
      // per-rate scalers: a rate whose count is d above the site's minimum weighs 2^(-256 d); no
      // rescaling inside the tiny tree (it would only move factors of 2^256 between values and counts)
      const uint32_t kmin = rate_min<R>(scal);
      const double f = rate_scale_factor(min(scal - kmin, EPA_RATE_MAXDIFF));
      #pragma unroll
      for (int k = 0; k < 4; ++k) in[k] *= f;
      scal = kmin;
    }
    else if (group_all<R>(small, lane))
    {
      #pragma unroll
      for (int k = 0; k < 4; ++k) in[k] *= EPA_SCALE_FACTOR;
      scal += 1;
    }
    const double * tp = tv + tv_pos(cur.mask) * R;
    double term = (in[0] * c_model.freqs[0]) * tp[0] + (in[1] * c_model.freqs[1]) * tp[16 * R]
                + (in[2] * c_model.freqs[2]) * tp[32 * R] + (in[3] * c_model.freqs[3]) * tp[48 * R];
    term = rate_sum<R>(term * wr);
    {
      // pendant sumtable: tip side takes pi*Vinv, inner side takes V
      double st[4];
      #pragma unroll
      for (int j = 0; j < 4; ++j)
      {
        const double right = c_model.eigenvecs[j * 4] * in[0] + c_model.eigenvecs[j * 4 + 1] * in[1]
                           + c_model.eigenvecs[j * 4 + 2] * in[2] + c_model.eigenvecs[j * 4 + 3] * in[3];
        st[j] = cs.tipleft[cur.mask * 4 + j] * right;
      }
      store_sum_row<R>(sum, s, r, act, wr, st[0], st[1], st[2], st[3], cur.inv);
    }
    if ((i % R) == r) { mine = act ? term : 1.0; mscal = act ? scal : 0u; minv = act ? cur.inv : 0.0; }
    if ((i % R) == R - 1)
      acc += site_loglk(mine, mscal, minv);
  }
  __syncwarp();
  return warp_sum(acc);
}

// Pass B: inner CLV toward the distal node from (T, X); leaves the distal sumtable (D vs inner)
template <int R, bool PR>
__device__ __forceinline__ void warp_pass_distal(const BloCtaSmem & cs, const double * ws, double * sum,
                                              const double * __restrict__ D, const double * __restrict__ X,
                                              const uint32_t * __restrict__ sD, const uint32_t * __restrict__ sX,
                                              const uint8_t * __restrict__ qc, int w, int lane,
                                              const double * __restrict__ inv_w, int pr_begin, int bugcompat,
                                              int pm_x = BloWarpSmem<R>::P_P)
{
  // pm_x: offset of X's transition matrix (--raxml-blo runs the pass with D and X swapped as well)
  constexpr int SPW = 32 / R;
  const int r = lane % R, so = lane / R;
  double pp[16];
  #pragma unroll
  for (int k = 0; k < 16; ++k) pp[k] = ws[pm_x + k * R + r];
  const double * tv = ws + BloWarpSmem<R>::TV + r;
  const double wr = c_model.weights[r];
  const int n_units = (w + SPW - 1) / SPW;
  UnitIn nxt = load_unit<R, PR ? 2 : 0>(D, X, sD, sX, qc, so, w, r, inv_w, pr_begin, bugcompat);
  #pragma unroll 1
  for (int i = 0; i < n_units; ++i)
  {
    const UnitIn cur = nxt;
    const int s = i * SPW + so;
    nxt = load_unit<R, PR ? 2 : 0>(D, X, sD, sX, qc, s + SPW, w, r, inv_w, pr_begin, bugcompat);
    const bool act = s < w;
    double in[4];
    bool small = true;
    #pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      const double tb = pp[k * 4] * cur.xv[0] + pp[k * 4 + 1] * cur.xv[1] + pp[k * 4 + 2] * cur.xv[2] + pp[k * 4 + 3] * cur.xv[3];
      in[k] = tv[(k * 16 + tv_pos(cur.mask)) * R] * tb;
      small = small && (in[k] < EPA_SCALE_THRESHOLD);
    }
    if constexpr (PR)
    {
      const double f = rate_scale_factor(min(cur.scal - rate_min<R>(cur.scal), EPA_RATE_MAXDIFF));
      #pragma unroll
      for (int k = 0; k < 4; ++k) in[k] *= f;
    }
    else if (group_all<R>(small, lane))
    {
      #pragma unroll
      for (int k = 0; k < 4; ++k) in[k] *= EPA_SCALE_FACTOR;
    }
    double st[4];
    #pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const double left = cur.dv[0] * c_model.pivinv[j] + cur.dv[1] * c_model.pivinv[4 + j]
                        + cur.dv[2] * c_model.pivinv[8 + j] + cur.dv[3] * c_model.pivinv[12 + j];
      const double right = c_model.eigenvecs[j * 4] * in[0] + c_model.eigenvecs[j * 4 + 1] * in[1]
                         + c_model.eigenvecs[j * 4 + 2] * in[2] + c_model.eigenvecs[j * 4 + 3] * in[3];
      st[j] = left * right;
    }
    store_sum_row<R>(sum, s, r, act, wr, st[0], st[1], st[2], st[3], cur.inv);
  }
  __syncwarp();
}

// next work item of the CTA's current block of EPA_WORK_BLOCK consecutive items (lane 0 only)
__device__ __forceinline__ unsigned long long cta_next_item(BloCtaSmem & cs, unsigned long long * counter)
{
  while (atomicCAS(&cs.q_lock, 0, 1) != 0) { }
  __threadfence_block();
  volatile unsigned long long * qn = &cs.q_next;
  volatile unsigned long long * qe = &cs.q_end;
  if (*qn == *qe)
  {
    const unsigned long long base = atomicAdd(counter, (unsigned long long) EPA_WORK_BLOCK);
    *qn = base;
    *qe = base + EPA_WORK_BLOCK;
  }
  const unsigned long long item = *qn;
  *qn = item + 1;
  __threadfence_block();
  atomicExch(&cs.q_lock, 0);
  return item;
}

// GS = sumtable in global scratch (windows that do not fit the shared-memory slice)
// RAXML = --raxml-blo (see newton_old above and the lane = site kernel)
// PR = per-rate scalers
template <int R, bool GS, bool RAXML = false, bool PR = false>
__global__ void __launch_bounds__(256, 1)
blo_dna_kernel(BloArgs a)
{
  extern __shared__ __align__(16) double smem_d[];
  __shared__ BloCtaSmem cs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 16; i += blockDim.x)
  {
    cs.V[i] = c_model.eigenvecs[i];
    cs.Vinv[i] = c_model.inv_eigenvecs[i];
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
  {
    const int mask = i >> 2, j = i & 3;
    double acc = 0.0;
    for (int k = 0; k < 4; ++k)
      if ((mask >> k) & 1) acc += c_model.pivinv[k * 4 + j];
    cs.tipleft[i] = acc;
  }
  if (threadIdx.x == 0) { cs.q_next = 0; cs.q_end = 0; cs.q_lock = 0; }
  __syncthreads();

  const size_t per_warp = BloWarpSmem<R>::doubles(GS ? 0 : a.wcap);
  double * ws = smem_d + (size_t) warp * per_warp;
  double * sum = GS ? a.scratch + ((size_t) blockIdx.x * n_warps + warp) * (size_t) a.n * blo_row(R)
                    : ws + BloWarpSmem<R>::SUM;
  double * ex = ws + BloWarpSmem<R>::EX;

  for (;;)
  {
    unsigned long long item = 0;
    if (lane == 0) item = cta_next_item(cs, a.counter);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= a.n_pairs) break;
    uint32_t pid, q, e;
    if (a.pair_q)
    {
      pid = a.work ? a.work[item] : (uint32_t) item;
      q = a.pair_q[pid];
      e = a.pair_e[pid];
    }
    else
    {
      e = (uint32_t) (item / a.nq);
      q = a.perm ? a.perm[item % a.nq] : (uint32_t) (item % a.nq);
      pid = q * a.n_edges + e;
    }
    const EdgeDev ed = a.edges[e];
    const int begin = a.begin[q], w = a.span[q];
    if (w <= 0 || (!GS && w > a.wcap))
    {
      if (lane == 0) a.out[pid] = BloResult{NAN, NAN, NAN};
      continue;
    }
    const int n = a.n;
    const double * D = a.tree.clv + ed.distal * a.tree.clv_stride + (size_t) begin * R * 4;
    const double * X = a.tree.clv + ed.proximal * a.tree.clv_stride + (size_t) begin * R * 4;
    const uint32_t * sD = PR ? a.tree.scaler + (size_t) ed.distal * n * R : a.tree.scaler + (size_t) ed.distal * n + begin;
    const uint32_t * sX = PR ? a.tree.scaler + (size_t) ed.proximal * n * R : a.tree.scaler + (size_t) ed.proximal * n + begin;
    const uint8_t * qc = a.codes + (size_t) q * n + begin;
    const double * inv_w = a.tree.inv ? a.tree.inv + begin : nullptr;      // +I: pll_util.cpp:413-414

    // optimize_branch_triplet: lengths orig/2, orig/2, -ln 0.9. The smoothing loop of
    // opt_branch_lengths_pplacer is unrolled into half rounds (tip phase, distal phase) so that
    // every device function has exactly one call site: the code stays small enough for the
    // instruction cache without giving up inlining.
    const double orig = ed.length;
    double len[3] = {orig / 2.0, orig / 2.0, EPA_DEFAULT_PENDANT};      // distal, proximal, pendant
    if constexpr (RAXML)
    {
      // pllmod_opt_optimize_branch_lengths_local, radius 1 (PM/optimize/pll_optimize.c:778-1097)
      using L = BloWarpSmem<R>;
      #pragma unroll 1
      for (int mi = 0; mi < 3; ++mi) warp_pmatrix<R>(cs, len[mi], ws + mi * (R * 16), ex, lane);
      warp_tipvec<R>(ws + L::P_E, ws + L::TV, lane);
      auto pass_tip = [&]() -> double { return warp_pass_tip<R, PR>(cs, ws, sum, D, X, sD, sX, qc, w, lane, inv_w, begin, a.bugcompat); };
      auto edge = [&](int mi) -> bool
      {
        double xguess = len[mi];
        if (xguess < EPA_MIN_BRLEN || xguess > EPA_MAX_BRLEN) xguess = EPA_DEFAULT_BRLEN;
        bool failed;
        const double xres = newton_old([&](double x, double & f, double & df) { warp_derivatives<R>(sum, ex, w, x, lane, f, df); },
                                       EPA_MIN_BRLEN, xguess, EPA_MAX_BRLEN, EPA_MIN_BRLEN / 10.0, failed);
        if (failed) return false;
        const bool moved = fabs(xres - len[mi]) > 1e-10;
        len[mi] = xres;
        if (moved)
        {
          warp_pmatrix<R>(cs, xres, ws + mi * (R * 16), ex, lane);
          if (mi == 2) warp_tipvec<R>(ws + L::P_E, ws + L::TV, lane);
        }
        return true;
      };
      double loglikelihood = pass_tip();
      int iters = EPA_SMOOTHINGS;
      bool ok = true;
      while (iters)
      {
        if (!(ok = edge(2))) break;
        warp_pass_distal<R, PR>(cs, ws, sum, D, X, sD, sX, qc, w, lane, inv_w, begin, a.bugcompat, L::P_P);
        if (!(ok = edge(0))) break;
        warp_pass_distal<R, PR>(cs, ws, sum, X, D, sX, sD, qc, w, lane, inv_w, begin, a.bugcompat, L::P_D);
        if (!(ok = edge(1))) break;
        double new_logl = pass_tip();
        const double pend_before = len[2];
        if (!(ok = edge(2))) break;
        if (fabs(len[2] - pend_before) > 1e-10) new_logl = pass_tip();
        if (new_logl - loglikelihood > new_logl * 1e-13)
        {
          --iters;
          if (fabs(new_logl - loglikelihood) < EPA_BLO_EPSILON) iters = 0;
          loglikelihood = new_logl;
        }
        else { loglikelihood = new_logl; break; }
      }
      if (lane == 0)
      {
        BloResult res;
        res.logl = ok ? loglikelihood : 0.0;
        res.distal = (orig / (len[0] + len[1])) * len[0];
        res.pendant = len[2];
        a.out[pid] = res;
      }
      __syncwarp();
      continue;
    }
    const double original_length = len[0] * 2;
    double old_d = len[0], old_e = len[2], loglikelihood = 0.0;
    int smoothings = EPA_SMOOTHINGS;
    unsigned rebuild = 7u;                                               // matrices to recompute
    bool first = true, distal_phase = false;
    for (;;)
    {
      #pragma unroll 1
      for (int mi = 0; mi < 3; ++mi)
        if (rebuild & (1u << mi))
        {
          const double t = mi == 0 ? len[0] : (mi == 1 ? len[1] : len[2]);
          warp_pmatrix<R>(cs, t, ws + mi * (R * 16), ex, lane);
          if (mi == 2) warp_tipvec<R>(ws + BloWarpSmem<R>::P_E, ws + BloWarpSmem<R>::TV, lane);
        }
      rebuild = 0u;
      double xmin, xmax, xguess;
      if (!distal_phase)
      {
        // score the current lengths (also builds the pendant sumtable of the coming round)
        const double new_logl = -warp_pass_tip<R, PR>(cs, ws, sum, D, X, sD, sX, qc, w, lane, inv_w, begin, a.bugcompat);
        if (first) { loglikelihood = new_logl; first = false; }
        else
        {
          if (new_logl - loglikelihood > new_logl * 1e-14)
          {
            len[2] = old_e; len[0] = old_d; len[1] = original_length - old_d;   // worse: restore and stop
            break;
          }
          --smoothings;
          if (fabs(new_logl - loglikelihood) < EPA_BLO_EPSILON) smoothings = 0;
          loglikelihood = new_logl;
        }
        if (!smoothings) break;
        old_d = len[0]; old_e = len[2];
        xmin = EPA_MIN_BRLEN; xmax = EPA_MAX_BRLEN; xguess = len[2];
        if (xguess < xmin || xguess > xmax) xguess = EPA_DEFAULT_BRLEN;
      }
      else
      {
        warp_pass_distal<R, PR>(cs, ws, sum, D, X, sD, sX, qc, w, lane, inv_w, begin, a.bugcompat);
        xmin = fmin(EPA_MIN_BRLEN / 2.0, original_length / 2.0);
        xmax = original_length - xmin / 10.0;
        xguess = len[0];
        if (xguess < xmin || xguess > xmax) xguess = original_length / 2.0;
      }
      const double xres = warp_newton<R>(sum, ex, w, lane, xmin, xguess, xmax, xmin / 10.0);
      if (xres > 0.0)
      {
        if (!distal_phase) { len[2] = xres; rebuild = 4u; }
        else { len[0] = xres; len[1] = original_length - xres; rebuild = 3u; }
      }
      distal_phase = !distal_phase;
    }
    if (lane == 0)
    {
      BloResult res;
      res.logl = -loglikelihood;
      res.distal = (orig / (len[0] + len[1])) * len[0];      // Tiny_Tree.cpp:183-185
      res.pendant = len[2];
      a.out[pid] = res;
    }
    __syncwarp();
  }
}

}  // namespace epa
