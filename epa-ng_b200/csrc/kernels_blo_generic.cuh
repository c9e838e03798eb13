// kernels_blo_generic.cuh - HOT LOOP B for any state count (amino acids: S = 20).
//
// Same algorithm and reference citations as kernels_blo.cuh (Tiny_Tree::place with branch-length
// optimisation, src/tree/Tiny_Tree.cpp:131-218, src/core/pll/optimize.cpp:60-286, libpll sumtable /
// derivatives / partials / likelihood kernels, PM/optimize/opt_algorithms.c:86-262); different
// shape: with 20 states one (site, rate) unit is 1200 FMAs and the sumtable of a 300-site window
// is 185 KB, so ONE CTA OF 256 THREADS WORKS ON ONE PAIR:
//  * CLV passes: thread = (site, rate) unit, the R threads of a site are adjacent lanes (rate sum and
//    scaling vote by shuffles / ballot); the three transition matrices, the per-character tip vectors,
//    V, V^-1 and pi V^-1 live in shared memory (rate blocks padded against bank conflicts);
//  * Newton iterations: thread = site over 1 + R(S-1) site-major planes of the sumtable (coalesced;
//    the planes live in an L2-resident per-CTA scratch), diag tables in shared memory, block reduction
//    in a fixed order.
#pragma once
#include "kernels_blo.cuh"
#include "kernels_blo_site.cuh"          // tcgen05.ld / st helpers

namespace epa {

constexpr int GEN_THREADS = 256;

// Sumtable rows in TENSOR MEMORY (unit-mapped phases). A (rate, site) unit is always handled by the same thread -
// unit u belongs to thread u % 256, as its (u / 256)-th unit - in the CLV passes that write its S - 1 decaying
// entries and in the Newton sweeps that read them, so the row can live in the thread's own TMEM lane: warps w and
// w + 4 share a 32-lane quarter and take 256 columns each, a unit takes GEN_TM_COLS of them. 6 units per thread
// hold windows of up to 6 * 256 / R sites (384 for four rate categories); longer windows keep the L2-resident planes.
constexpr int GEN_TM_COLS = 40;          // 19 doubles = 38 columns, padded to the x32 + x8 load shapes
constexpr int GEN_TM_UNITS = 6;

__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}

// the S - 1 decaying entries of one unit -> the thread's TMEM slot (S = 20: 19 doubles)
template <int S>
__device__ __forceinline__ void gen_tm_store(uint32_t taddr, const double (&st)[S])
{
  static_assert(S - 1 <= 20, "a unit's row must fit GEN_TM_COLS columns");
  uint32_t a[32], b[8];
  #pragma unroll
  for (int j = 0; j < 16; ++j)
  {
    const double v = j + 1 < S ? st[j + 1] : 0.0;
    a[2 * j] = (uint32_t) __double2loint(v); a[2 * j + 1] = (uint32_t) __double2hiint(v);
  }
  #pragma unroll
  for (int j = 0; j < 4; ++j)
  {
    const double v = 17 + j < S ? st[17 + j] : 0.0;
    b[2 * j] = (uint32_t) __double2loint(v); b[2 * j + 1] = (uint32_t) __double2hiint(v);
  }
  tc_st32(taddr, a);
  tc_st8(taddr + 32, b);
}

template <int S>
__device__ __forceinline__ void gen_tm_load(uint32_t taddr, double (&x)[S - 1])
{
  uint32_t a[32], b[8];
  tc_ld32(taddr, a);
  tc_ld8(taddr + 32, b);
  tc_wait_ld();
  #pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < S - 1) x[j] = __hiloint2double((int) a[2 * j + 1], (int) a[2 * j]);
  #pragma unroll
  for (int j = 0; j < 4; ++j)
    if (16 + j < S - 1) x[16 + j] = __hiloint2double((int) b[2 * j + 1], (int) b[2 * j]);
}

template <int S, int R>
struct GenSmem {
  static constexpr int PS = S * S + 4;                 // padded rate block of a P-matrix
  static constexpr int TVS = MAX_CODES * S + 4;        // padded rate block of the tip-vector table
  static constexpr int NK = R * (S - 1);               // decaying components of a site
  static constexpr int V = 0;
  static constexpr int VINV = V + S * S;
  static constexpr int PIVINV = VINV + S * S;
  static constexpr int TIPLEFT = PIVINV + S * S;       // [code][S]
  static constexpr int P = TIPLEFT + MAX_CODES * S;    // [3][R][PS]: distal, proximal, pendant
  static constexpr int TV = P + 3 * R * PS;            // [R][TVS]
  static constexpr int EX = TV + R * TVS;              // [R*S]
  static constexpr int DIAG = EX + R * S;              // [3][NK]
  static constexpr int RED = DIAG + 3 * NK;            // [2][8] block-reduction scratch
  static constexpr int TOTAL = RED + 16;
};

// deterministic block-wide sum of two values (fixed order over the 8 warps)
__device__ __forceinline__ void block_sum2(double & a, double & b, double * red)
{
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5;
  __syncthreads();                                     // previous readers of `red` are done
  if ((threadIdx.x & 31) == 0) { red[warp] = a; red[8 + warp] = b; }
  __syncthreads();
  double sa = 0.0, sb = 0.0;
  #pragma unroll
  for (int w = 0; w < GEN_THREADS / 32; ++w) { sa += red[w]; sb += red[8 + w]; }
  a = sa; b = sb;
}

template <int S, int R>
__device__ __forceinline__ void gen_pmatrix(double * sm, double t, int which)
{
  using L = GenSmem<S, R>;
  for (int idx = threadIdx.x; idx < R * S; idx += GEN_THREADS)
    sm[L::EX + idx] = expm1(c_model.eigenvals[idx % S] * c_model.rates[idx / S] * t);
  __syncthreads();
  double * P = sm + L::P + which * R * L::PS;
  for (int idx = threadIdx.x; idx < R * S * S; idx += GEN_THREADS)
  {
    const int r = idx / (S * S), i = (idx / S) % S, j = idx % S;
    double acc = (i == j) ? 1.0 : 0.0;
    for (int k = 0; k < S; ++k) acc += (sm[L::VINV + i * S + k] * sm[L::EX + r * S + k]) * sm[L::V + k * S + j];
    P[r * L::PS + i * S + j] = acc;
  }
  __syncthreads();
}

template <int S, int R>
__device__ __forceinline__ void gen_tipvec(double * sm)
{
  using L = GenSmem<S, R>;
  const double * P = sm + L::P + 2 * R * L::PS;
  const int ncodes = c_model.ncodes;
  for (int idx = threadIdx.x; idx < R * ncodes * S; idx += GEN_THREADS)
  {
    const int r = idx / (ncodes * S), c = (idx / S) % ncodes, i = idx % S;
    const uint32_t mask = c_model.code2mask[c];
    double acc = 0.0;
    for (int j = 0; j < S; ++j)
      if ((mask >> j) & 1u) acc += P[r * L::PS + i * S + j];
    sm[L::TV + r * L::TVS + c * S + i] = acc;
  }
  __syncthreads();
}

// sumtable row of unit (site s, rate r): stationary component summed over the rates into plane 0,
// decaying components into planes 1 + r(S-1) + (j-1)
template <int S, int R>
__device__ __forceinline__ void gen_store_row(double * sum, int wpad, int s, int r, bool act, double wr, const double (&st)[S],
                                              double inv)
{
  // +I: the invariant term of the site joins the t-independent entry (LP/core_derivatives.c:676-687)
  const double base = rate_sum<R>(st[0] * wr) + inv;
  if (act)
  {
    if (r == 0) sum[s] = base;
    #pragma unroll
    for (int j = 1; j < S; ++j) sum[(size_t) (r * (S - 1) + j) * wpad + s] = st[j];
  }
}

template <int S, int R>
__device__ __forceinline__ double gen_pass_tip(double * sm, double * sum, int wpad, const double * __restrict__ D,
                                               const double * __restrict__ X, const uint32_t * __restrict__ sD,
                                               const uint32_t * __restrict__ sX, const uint8_t * __restrict__ qc, int w,
                                               const double * __restrict__ inv_w)
{
  using L = GenSmem<S, R>;
  const int lane = threadIdx.x & 31;
  const int r = threadIdx.x % R;
  const double * Pd = sm + L::P + r * L::PS;
  const double * Pp = sm + L::P + (R + r) * L::PS;
  const double * tvr = sm + L::TV + r * L::TVS;
  const double wr = c_model.weights[r];
  double acc = 0.0, unused = 0.0;
  const int n_units = ((w * R + GEN_THREADS - 1) / GEN_THREADS) * GEN_THREADS;
  for (int u = threadIdx.x; u < n_units; u += GEN_THREADS)
  {
    const int s = u / R;
    const bool act = s < w;
    const int sc = act ? s : w - 1;
    double dv[S], xv[S], in[S];
    load_vec<S>(D + ((size_t) sc * R + r) * S, dv);
    load_vec<S>(X + ((size_t) sc * R + r) * S, xv);
    const int code = qc[sc] & (MAX_CODES - 1);
    uint32_t scal = __ldg(sD + sc) + __ldg(sX + sc);
    const double inv = inv_w ? __ldg(inv_w + sc) : 0.0;
    bool small = true;
    #pragma unroll 4
    for (int i = 0; i < S; ++i)
    {
      double ta = 0.0, tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S; ++j) { ta += Pd[i * S + j] * dv[j]; tb += Pp[i * S + j] * xv[j]; }
      in[i] = ta * tb;
      small = small && (in[i] < EPA_SCALE_THRESHOLD);
    }
    if (group_all<R>(small, lane))
    {
      #pragma unroll
      for (int i = 0; i < S; ++i) in[i] *= EPA_SCALE_FACTOR;
      scal += 1;
    }
    double term = 0.0;
    #pragma unroll
    for (int i = 0; i < S; ++i) term += (in[i] * c_model.freqs[i]) * tvr[code * S + i];
    term = rate_sum<R>(term * wr);
    if (act && r == 0) acc += site_loglk(term, scal, inv);
    double st[S];
    #pragma unroll 4
    for (int j = 0; j < S; ++j)
    {
      double right = 0.0;
      #pragma unroll
      for (int k = 0; k < S; ++k) right += sm[L::V + j * S + k] * in[k];
      st[j] = sm[L::TIPLEFT + code * S + j] * right;
    }
    gen_store_row<S, R>(sum, wpad, s, r, act, wr, st, inv);
  }
  block_sum2(acc, unused, sm + L::RED);
  return acc;
}

template <int S, int R>
__device__ __forceinline__ void gen_pass_distal(double * sm, double * sum, int wpad, const double * __restrict__ D,
                                                const double * __restrict__ X, const uint8_t * __restrict__ qc, int w,
                                                const double * __restrict__ inv_w, int which_x = 1)
{
  // which_x: matrix of X's edge (1 = proximal; --raxml-blo also runs the pass with the nodes swapped, 0)
  using L = GenSmem<S, R>;
  const int lane = threadIdx.x & 31;
  const int r = threadIdx.x % R;
  const double * Pp = sm + L::P + (which_x * R + r) * L::PS;
  const double * tvr = sm + L::TV + r * L::TVS;
  const double wr = c_model.weights[r];
  const int n_units = ((w * R + GEN_THREADS - 1) / GEN_THREADS) * GEN_THREADS;
  for (int u = threadIdx.x; u < n_units; u += GEN_THREADS)
  {
    const int s = u / R;
    const bool act = s < w;
    const int sc = act ? s : w - 1;
    double dv[S], xv[S], in[S];
    load_vec<S>(D + ((size_t) sc * R + r) * S, dv);
    load_vec<S>(X + ((size_t) sc * R + r) * S, xv);
    const int code = qc[sc] & (MAX_CODES - 1);
    bool small = true;
    #pragma unroll 4
    for (int i = 0; i < S; ++i)
    {
      double tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S; ++j) tb += Pp[i * S + j] * xv[j];
      in[i] = tvr[code * S + i] * tb;
      small = small && (in[i] < EPA_SCALE_THRESHOLD);
    }
    if (group_all<R>(small, lane))
    {
      #pragma unroll
      for (int i = 0; i < S; ++i) in[i] *= EPA_SCALE_FACTOR;
    }
    double st[S];
    #pragma unroll 4
    for (int j = 0; j < S; ++j)
    {
      double left = 0.0, right = 0.0;
      #pragma unroll
      for (int k = 0; k < S; ++k) { left += dv[k] * sm[L::PIVINV + k * S + j]; right += sm[L::V + j * S + k] * in[k]; }
      st[j] = left * right;
    }
    gen_store_row<S, R>(sum, wpad, s, r, act, wr, st, inv_w ? __ldg(inv_w + sc) : 0.0);
  }
  __syncthreads();
}

// ---- lane = site CLV passes over the site-blocked CLV copy ------------------------------------
// thread = site, loop over the rate categories: the transition matrices are warp-uniform 128-bit
// shared-memory broadcasts (a (site, rate)-per-thread mapping needs one shared-memory word per FMA
// and is bound by the shared-memory pipe at a quarter of the fp64 rate), the eigenvector tables are
// constant-bank operands, the CLV components and the sumtable planes are coalesced, and there are no
// cross-lane rate sums. No 2^256 rescaling inside the tiny tree (see kernels_blo_site.cuh).
template <int S>
__device__ __forceinline__ size_t clvt_off_c(int abs_site, int C)
{
  return (size_t) (abs_site >> 5) * (size_t) (C * 32) + (size_t) (abs_site & 31);
}

// Work unit = (rate, site) with the SITE index fastest: the threads of a warp share the rate (except
// where a warp straddles two rates), R * w units keep 256 threads busy for any window length. The
// per-rate contributions to the site likelihood and to the stationary sumtable entry go through
// two small shared-memory arrays and are added over the rates in a fixed order afterwards.
// Per-rate scalers (PR kernels; amino acids run libpll's generic kernels then). A site's rate r weighs
// 2^(-256 d_r), d_r = its counter minus the site's minimum (LP/core_likelihood.c:474-491, core_derivatives.c:404-423).
// The counters of the two edge CLVs are read where the reference reads them (rate_weights, kernels_blo_site.cuh).
// An inner CLV that comes from a tip-inner update (`ti`: the distal pass always, the tip pass on a tip edge) is
// rescaled as a WHOLE site when all its R*S entries are small, and the reference counts that in entry
// [site index] of the inner node's [site][rate] array (LP/core_partials.c:461-506): site s, rate r picks up the
// rescaling of site s*R + r, and a rescaled site keeps its factor 2^256 uncompensated. Inner-inner updates rescale
// single rates with their own counters, which cancels and is skipped here.
// fbuf[r][s] = weight of unit (r, s) for the sums of this pass and its Newton sweeps; scaled[s] = site rescaled.
struct GenPr {
  const uint32_t * sDn, * sXn;        // [n][R] counters of the distal / proximal node
  int begin, bugcompat, ti;
  double * fbuf;
  uint8_t * scaled;
};

template <int R>
__device__ __forceinline__ uint32_t gen_pr_weights(const GenPr & pr, int wpad, int w, int s, double (&F)[R])
{
  const size_t base = pr.bugcompat ? (size_t) pr.begin + (size_t) s * R : ((size_t) pr.begin + s) * R;
  uint32_t kr[R], kmin = 0xffffffffu;
  #pragma unroll
  for (int r = 0; r < R; ++r)
  {
    kr[r] = __ldg(pr.sDn + base + r) + __ldg(pr.sXn + base + r);
    if (s * R + r < w) kr[r] += pr.scaled[s * R + r];
    kmin = min(kmin, kr[r]);
  }
  const double whole = pr.scaled[s] ? EPA_SCALE_FACTOR : 1.0;
  #pragma unroll
  for (int r = 0; r < R; ++r)
  {
    F[r] = rate_scale_factor(min(kr[r] - kmin, EPA_RATE_MAXDIFF)) * whole;
    pr.fbuf[r * wpad + s] = F[r];
  }
  return kmin;
}

// scaled[s] = every unit of site s flagged small (fbuf holds the flags of the unit loop), only for tip-inner updates
template <int R>
__device__ __forceinline__ void gen_pr_sites(const GenPr & pr, int wpad, int w)
{
  for (int s = threadIdx.x; s < w; s += GEN_THREADS)
  {
    bool all = pr.ti != 0;
    #pragma unroll
    for (int r = 0; r < R; ++r) all = all && (pr.fbuf[r * wpad + s] != 0.0);
    pr.scaled[s] = all ? 1 : 0;
  }
  __syncthreads();
}

template <int S, int R, bool PR = false>
__device__ __forceinline__ double gen_pass_tip_site(double * sm, double * sum, int wpad, const double * __restrict__ DT,
                                                    const double * __restrict__ XT, const uint32_t * __restrict__ sD,
                                                    const uint32_t * __restrict__ sX, const uint8_t * __restrict__ qc,
                                                    int begin, int w, double * rbuf, const double * __restrict__ inv_w,
                                                    uint64_t tm, const GenPr & pr = GenPr{})
{
  using L = GenSmem<S, R>;
  double * termbuf = rbuf, * basebuf = rbuf + R * wpad;
  // tm != 0: rows go to this thread's tensor-memory slots; the loop then runs the same number of trips in every lane
  // (tcgen05.st is warp-wide) and a lane beyond the last unit recomputes that unit and stores it in its own slot
  const int n_units = R * w;
  const int u_end = tm ? ((n_units + GEN_THREADS - 1) / GEN_THREADS) * GEN_THREADS : n_units;
  for (int u0 = threadIdx.x, slot = 0; u0 < u_end; u0 += GEN_THREADS, ++slot)
  {
    const bool act = u0 < n_units;
    const int u = act ? u0 : n_units - 1;
    const int r = u / w, s = u - r * w;
    const size_t off = clvt_off_c<S>(begin + s, R * S) + (size_t) (r * S) * 32;
    const int code = qc[s] & (MAX_CODES - 1);
    double dv[S], xv[S], in[S];
    #pragma unroll
    for (int k = 0; k < S; ++k) { dv[k] = __ldg(DT + off + (size_t) k * 32); xv[k] = __ldg(XT + off + (size_t) k * 32); }
    const double2 * Pd = reinterpret_cast<const double2 *>(sm + L::P + r * L::PS);
    const double2 * Pp = reinterpret_cast<const double2 *>(sm + L::P + (R + r) * L::PS);
    #pragma unroll
    for (int i = 0; i < S; ++i)
    {
      double ta = 0.0, tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S / 2; ++j)
      {
        const double2 a = Pd[i * (S / 2) + j], b = Pp[i * (S / 2) + j];
        ta += a.x * dv[2 * j]; ta += a.y * dv[2 * j + 1];
        tb += b.x * xv[2 * j]; tb += b.y * xv[2 * j + 1];
      }
      in[i] = ta * tb;
    }
    if constexpr (PR)
    {
      bool small = true;
      #pragma unroll
      for (int i = 0; i < S; ++i) small = small && (in[i] < EPA_SCALE_THRESHOLD);
      if (act) pr.fbuf[r * wpad + s] = small ? 1.0 : 0.0;
    }
    const double * tvr = sm + L::TV + r * L::TVS + code * S;
    const double * tl = sm + L::TIPLEFT + code * S;
    const double wr = c_model.weights[r];
    double tr = 0.0;
    #pragma unroll
    for (int i = 0; i < S; ++i) tr += (in[i] * c_model.freqs[i]) * tvr[i];
    if (act) termbuf[r * wpad + s] = tr * wr;
    double st[S];
    #pragma unroll
    for (int j = 0; j < S; ++j)
    {
      double right = 0.0;
      #pragma unroll
      for (int k = 0; k < S; ++k) right += c_model.eigenvecs[j * S + k] * in[k];
      st[j] = tl[j] * right;
    }
    if (act) basebuf[r * wpad + s] = st[0] * wr;
    if (tm) gen_tm_store<S>((uint32_t) tm + (uint32_t) slot * GEN_TM_COLS, st);
    else
    {
      #pragma unroll
      for (int j = 1; j < S; ++j) sum[(size_t) (r * (S - 1) + j) * wpad + s] = st[j];
    }
  }
  if (tm) tc_wait_st();
  __syncthreads();
  if constexpr (PR) gen_pr_sites<R>(pr, wpad, w);
  double acc = 0.0, unused = 0.0;
  for (int s = threadIdx.x; s < w; s += GEN_THREADS)
  {
    double term = 0.0, base = 0.0;
    uint32_t scal;
    if constexpr (PR)
    {
      double F[R];
      scal = gen_pr_weights<R>(pr, wpad, w, s, F);
      #pragma unroll
      for (int r = 0; r < R; ++r) { term += termbuf[r * wpad + s] * F[r]; base += basebuf[r * wpad + s] * F[r]; }
    }
    else
    {
      #pragma unroll
      for (int r = 0; r < R; ++r) { term += termbuf[r * wpad + s]; base += basebuf[r * wpad + s]; }
      scal = __ldg(sD + s) + __ldg(sX + s);
    }
    const double inv = inv_w ? __ldg(inv_w + s) : 0.0;
    sum[s] = base + inv;
    acc += site_loglk(term, scal, inv);
  }
  block_sum2(acc, unused, sm + L::RED);
  return acc;
}

template <int S, int R, bool PR = false>
__device__ __forceinline__ void gen_pass_distal_site(double * sm, double * sum, int wpad, const double * __restrict__ DT,
                                                     const double * __restrict__ XT, const uint8_t * __restrict__ qc,
                                                     int begin, int w, double * rbuf, const double * __restrict__ inv_w,
                                                     uint64_t tm, int which_x = 1, const GenPr & pr = GenPr{})
{
  using L = GenSmem<S, R>;
  double * basebuf = rbuf;
  const int n_units = R * w;
  const int u_end = tm ? ((n_units + GEN_THREADS - 1) / GEN_THREADS) * GEN_THREADS : n_units;
  for (int u0 = threadIdx.x, slot = 0; u0 < u_end; u0 += GEN_THREADS, ++slot)
  {
    const bool act = u0 < n_units;
    const int u = act ? u0 : n_units - 1;
    const int r = u / w, s = u - r * w;
    const size_t off = clvt_off_c<S>(begin + s, R * S) + (size_t) (r * S) * 32;
    const int code = qc[s] & (MAX_CODES - 1);
    double dv[S], xv[S], in[S];
    #pragma unroll
    for (int k = 0; k < S; ++k) { dv[k] = __ldg(DT + off + (size_t) k * 32); xv[k] = __ldg(XT + off + (size_t) k * 32); }
    const double2 * Pp = reinterpret_cast<const double2 *>(sm + L::P + (which_x * R + r) * L::PS);
    const double * tvr = sm + L::TV + r * L::TVS + code * S;
    #pragma unroll
    for (int i = 0; i < S; ++i)
    {
      double tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S / 2; ++j)
      {
        const double2 b = Pp[i * (S / 2) + j];
        tb += b.x * xv[2 * j]; tb += b.y * xv[2 * j + 1];
      }
      in[i] = tvr[i] * tb;
    }
    if constexpr (PR)
    {
      bool small = true;
      #pragma unroll
      for (int i = 0; i < S; ++i) small = small && (in[i] < EPA_SCALE_THRESHOLD);
      if (act) pr.fbuf[r * wpad + s] = small ? 1.0 : 0.0;
    }
    const double wr = c_model.weights[r];
    double st[S];
    #pragma unroll
    for (int j = 0; j < S; ++j)
    {
      double left = 0.0, right = 0.0;
      #pragma unroll
      for (int k = 0; k < S; ++k) { left += dv[k] * c_model.pivinv[k * S + j]; right += c_model.eigenvecs[j * S + k] * in[k]; }
      st[j] = left * right;
    }
    if (act) basebuf[r * wpad + s] = st[0] * wr;
    if (tm) gen_tm_store<S>((uint32_t) tm + (uint32_t) slot * GEN_TM_COLS, st);
    else
    {
      #pragma unroll
      for (int j = 1; j < S; ++j) sum[(size_t) (r * (S - 1) + j) * wpad + s] = st[j];
    }
  }
  if (tm) tc_wait_st();
  __syncthreads();
  if constexpr (PR) gen_pr_sites<R>(pr, wpad, w);
  for (int s = threadIdx.x; s < w; s += GEN_THREADS)
  {
    double base = 0.0;
    if constexpr (PR)
    {
      double F[R];
      (void) gen_pr_weights<R>(pr, wpad, w, s, F);
      #pragma unroll
      for (int r = 0; r < R; ++r) base += basebuf[r * wpad + s] * F[r];
    }
    else
    {
      #pragma unroll
      for (int r = 0; r < R; ++r) base += basebuf[r * wpad + s];
    }
    sum[s] = base + (inv_w ? __ldg(inv_w + s) : 0.0);
  }
  __syncthreads();
}

// Derivative sums with (rate, site) work units (site fastest): every thread takes the S - 1 decaying
// components of one rate of one site - all 256 threads stay busy for any window, the loads of a
// unit are independent, the decay tables of the rate are warp-uniform shared-memory broadcasts -
// and the per-rate partial sums are combined per site in a fixed order.
template <int S, int R, bool PR = false>
__device__ __forceinline__ void gen_derivatives_units(double * sm, const double * sum, int wpad, int w, double t,
                                                      double & f, double & df, double * rbuf, uint64_t tm,
                                                      const double * fbuf = nullptr)
{
  using L = GenSmem<S, R>;
  constexpr int NK = L::NK;
  __syncthreads();
  for (int k = threadIdx.x; k < NK; k += GEN_THREADS)
  {
    const double lk = c_model.eigenvals[1 + k % (S - 1)] * c_model.rates[k / (S - 1)];
    const double e = exp(lk * t) * c_model.weights[k / (S - 1)];
    sm[L::DIAG + k] = e; sm[L::DIAG + NK + k] = lk * e; sm[L::DIAG + 2 * NK + k] = lk * lk * e;
  }
  __syncthreads();
  double * b0 = rbuf, * b1 = rbuf + R * wpad, * b2 = rbuf + 2 * R * wpad;
  const int n_units = R * w;
  const int u_end = tm ? ((n_units + GEN_THREADS - 1) / GEN_THREADS) * GEN_THREADS : n_units;
  for (int u0 = threadIdx.x, slot = 0; u0 < u_end; u0 += GEN_THREADS, ++slot)
  {
    const bool act = u0 < n_units;
    const int u = act ? u0 : n_units - 1;
    const int r = u / w, s = u - r * w;
    const double * dg = sm + L::DIAG + r * (S - 1);
    double x[S - 1];
    if (tm) gen_tm_load<S>((uint32_t) tm + (uint32_t) slot * GEN_TM_COLS, x);
    else
    {
      const double * col = sum + (size_t) (r * (S - 1) + 1) * wpad + s;
      #pragma unroll
      for (int j = 0; j < S - 1; ++j) x[j] = col[(size_t) j * wpad];
    }
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    #pragma unroll
    for (int j = 0; j < S - 1; ++j) { c0 += x[j] * dg[j]; c1 += x[j] * dg[NK + j]; c2 += x[j] * dg[2 * NK + j]; }
    if constexpr (PR)
    {
      const double F = fbuf[r * wpad + s];
      c0 *= F; c1 *= F; c2 *= F;
    }
    if (act) { b0[r * wpad + s] = c0; b1[r * wpad + s] = c1; b2[r * wpad + s] = c2; }
  }
  __syncthreads();
  double a1 = 0.0, a2 = 0.0;
  for (int s = threadIdx.x; s < w; s += GEN_THREADS)
  {
    double c0 = sum[s], c1 = 0.0, c2 = 0.0;
    #pragma unroll
    for (int r = 0; r < R; ++r) { c0 += b0[r * wpad + s]; c1 += b1[r * wpad + s]; c2 += b2[r * wpad + s]; }
    const double inv = 1.0 / c0;
    const double g1 = -c1 * inv;
    a1 += g1;
    a2 += g1 * g1 - c2 * inv;
  }
  block_sum2(a1, a2, sm + L::RED);
  f = a1; df = a2;
}

template <int S, int R>
__device__ __forceinline__ void gen_derivatives(double * sm, const double * sum, int wpad, int w, double t, double & f, double & df)
{
  using L = GenSmem<S, R>;
  constexpr int NK = L::NK;
  __syncthreads();
  for (int k = threadIdx.x; k < NK; k += GEN_THREADS)
  {
    const double lk = c_model.eigenvals[1 + k % (S - 1)] * c_model.rates[k / (S - 1)];
    const double e = exp(lk * t) * c_model.weights[k / (S - 1)];
    sm[L::DIAG + k] = e; sm[L::DIAG + NK + k] = lk * e; sm[L::DIAG + 2 * NK + k] = lk * lk * e;
  }
  __syncthreads();
  double a1 = 0.0, a2 = 0.0;
  for (int s = threadIdx.x; s < w; s += GEN_THREADS)
  {
    double c0 = sum[s], c1 = 0.0, c2 = 0.0;
    #pragma unroll 4
    for (int k = 0; k < NK; ++k)
    {
      const double x = sum[(size_t) (k + 1) * wpad + s];
      c0 += x * sm[L::DIAG + k]; c1 += x * sm[L::DIAG + NK + k]; c2 += x * sm[L::DIAG + 2 * NK + k];
    }
    const double inv = 1.0 / c0;
    const double g1 = -c1 * inv;
    a1 += g1;
    a2 += g1 * g1 - c2 * inv;
  }
  block_sum2(a1, a2, sm + L::RED);
  f = a1; df = a2;
}

template <int S, int R, bool PR = false>
__device__ __forceinline__ double gen_newton(double * sm, const double * sum, int wpad, int w, double xmin, double xguess,
                                             double xmax, double tol, double * rbuf, uint64_t tm, const double * fbuf = nullptr)
{
  double x = fmax(fmin(xguess, xmax), xmin);
  double xl = xmin, xh = xmax;
  const double dxmax = xmax / EPA_NR_MAX_ITERS;
  int iter = 0;
  for (;;)
  {
    if (iter++ > EPA_NR_MAX_ITERS) return 0.0;
    double f, df;
    if (rbuf) gen_derivatives_units<S, R, PR>(sm, sum, wpad, w, x, f, df, rbuf, tm, fbuf);
    else gen_derivatives<S, R>(sm, sum, wpad, w, x, f, df);
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

// RAXML = --raxml-blo (pllmod_opt_optimize_branch_lengths_local with radius 1, PM/optimize/pll_optimize.c:778-1097)
// PR = per-rate scalers (needs the site-blocked CLV copy)
template <int S, int R, bool RAXML = false, bool PR = false>
__global__ void __launch_bounds__(GEN_THREADS, 1)
blo_generic_kernel(BloArgs a, int wpad, const double * __restrict__ clvT, size_t t_stride, int use_tmem, int bugcompat = 0)
{
  using L = GenSmem<S, R>;
  extern __shared__ __align__(16) double sm[];
  __shared__ unsigned long long s_item;
  for (int i = threadIdx.x; i < S * S; i += GEN_THREADS)
  {
    sm[L::V + i] = c_model.eigenvecs[i];
    sm[L::VINV + i] = c_model.inv_eigenvecs[i];
    sm[L::PIVINV + i] = c_model.pivinv[i];
  }
  const int ncodes = c_model.ncodes;
  __syncthreads();
  for (int i = threadIdx.x; i < ncodes * S; i += GEN_THREADS)
  {
    const int c = i / S, j = i % S;
    const uint32_t mask = c_model.code2mask[c];
    double acc = 0.0;
    for (int k = 0; k < S; ++k)
      if ((mask >> k) & 1u) acc += sm[L::PIVINV + k * S + j];
    sm[L::TIPLEFT + i] = acc;
  }
  __syncthreads();
  double * sum = a.scratch + (size_t) blockIdx.x * (size_t) (1 + L::NK) * wpad;
  // tensor memory for the sumtable rows of the unit-mapped phases (one CTA per SM: 246 registers x 256 threads)
  __shared__ uint32_t tmem_slot;
  const bool have_tm = clvT != nullptr && use_tmem;
  if (have_tm)
  {
    if (threadIdx.x < 32)
    {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tm_base = have_tm ? tmem_slot + ((uint32_t) (((threadIdx.x >> 5) & 3) * 32) << 16) + (uint32_t) (threadIdx.x >> 7) * 256u : 0u;

  for (;;)
  {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(a.counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
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
    if (w <= 0 || w > wpad)
    {
      if (threadIdx.x == 0) a.out[pid] = BloResult{NAN, NAN, NAN};
      continue;
    }
    const int n = a.n;
    const double * D = a.tree.clv + ed.distal * a.tree.clv_stride + (size_t) begin * R * S;
    const double * X = a.tree.clv + ed.proximal * a.tree.clv_stride + (size_t) begin * R * S;
    const uint32_t * sD = a.tree.scaler + (size_t) ed.distal * n + begin;
    const uint32_t * sX = a.tree.scaler + (size_t) ed.proximal * n + begin;
    const uint8_t * qc = a.codes + (size_t) q * n + begin;
    const double * inv_w = a.tree.inv ? a.tree.inv + begin : nullptr;      // +I: pll_util.cpp:413-414
    // bit 32 = rows in tensor memory, low word = this thread's slot 0
    const uint64_t tm = (have_tm && R * w <= GEN_THREADS * GEN_TM_UNITS) ? ((1ull << 32) | tm_base) : 0ull;
    GenPr pr{};
    if constexpr (PR)
    {
      pr.sDn = a.tree.scaler + (size_t) ed.distal * n * R;
      pr.sXn = a.tree.scaler + (size_t) ed.proximal * n * R;
      pr.begin = begin; pr.bugcompat = bugcompat;
      pr.fbuf = sm + L::TOTAL + 3 * R * wpad;
      pr.scaled = reinterpret_cast<uint8_t *>(sm + L::TOTAL + 4 * R * wpad);
    }

    const double orig = ed.length;
    double len[3] = {orig / 2.0, orig / 2.0, EPA_DEFAULT_PENDANT};
    if constexpr (RAXML)
    {
      // Steps of one round, one CLV pass and one Newton call each (single call sites keep the code small):
      //   0: [tip pass] score the round, pendant   1: distal   2: proximal   3: tip pass, pendant from the tip's side
      // Step 0 of the next round re-scores only if step 3 moved the pendant length (otherwise the
      // log-likelihood and the sumtable of step 3 are still those of the current lengths).
      double loglikelihood = 0.0, logl_now = 0.0;
      int iters = EPA_SMOOTHINGS, step = 0;
      unsigned rebuild = 7u;
      bool first = true, need_tip = true, ok = true;
      for (;;)
      {
        #pragma unroll 1
        for (int mi = 0; mi < 3; ++mi)
          if (rebuild & (1u << mi))
          {
            gen_pmatrix<S, R>(sm, len[mi], mi);
            if (mi == 2) gen_tipvec<S, R>(sm);
          }
        rebuild = 0u;
        int target;
        if (step == 0 || step == 3)
        {
          if (step == 3 || need_tip)
          {
            if constexpr (PR) pr.ti = ed.distal < a.tree.n_tips;
            logl_now = clvT
                ? gen_pass_tip_site<S, R, PR>(sm, sum, wpad, clvT + (size_t) ed.distal * t_stride, clvT + (size_t) ed.proximal * t_stride, sD, sX, qc, begin, w, sm + L::TOTAL, inv_w, tm, pr)
                : gen_pass_tip<S, R>(sm, sum, wpad, D, X, sD, sX, qc, w, inv_w);
          }
          if (step == 0)
          {
            if (first) { loglikelihood = logl_now; first = false; }
            else
            {
              const double new_logl = logl_now;
              if (new_logl - loglikelihood > new_logl * 1e-13)
              {
                --iters;
                if (fabs(new_logl - loglikelihood) < EPA_BLO_EPSILON) iters = 0;
                loglikelihood = new_logl;
              }
              else { loglikelihood = new_logl; iters = 0; }      // a worse score is kept and ends the loop
            }
            if (!iters) break;
          }
          target = 2;
        }
        else
        {
          const bool prox = step == 2;          // proximal edge: the distal pass with the two nodes swapped
          // (proximal step on a tip edge: the inner CLV comes from a tip-tip update, which never rescales)
          if constexpr (PR) pr.ti = !(prox && ed.distal < a.tree.n_tips);
          if (clvT)
            gen_pass_distal_site<S, R, PR>(sm, sum, wpad, clvT + (size_t) (prox ? ed.proximal : ed.distal) * t_stride,
                                           clvT + (size_t) (prox ? ed.distal : ed.proximal) * t_stride, qc, begin, w, sm + L::TOTAL, inv_w, tm, prox ? 0 : 1, pr);
          else
            gen_pass_distal<S, R>(sm, sum, wpad, prox ? X : D, prox ? D : X, qc, w, inv_w, prox ? 0 : 1);
          target = prox ? 1 : 0;
        }
        double xguess = len[target];
        if (xguess < EPA_MIN_BRLEN || xguess > EPA_MAX_BRLEN) xguess = EPA_DEFAULT_BRLEN;
        bool failed;
        double * rbuf = clvT ? sm + L::TOTAL : nullptr;
        const double xres = newton_old([&](double x, double & f, double & df)
                                       {
                                         if (rbuf) gen_derivatives_units<S, R, PR>(sm, sum, wpad, w, x, f, df, rbuf, tm, pr.fbuf);
                                         else gen_derivatives<S, R>(sm, sum, wpad, w, x, f, df);
                                       }, EPA_MIN_BRLEN, xguess, EPA_MAX_BRLEN, EPA_MIN_BRLEN / 10.0, failed);
        if (failed) { ok = false; break; }
        const bool moved = fabs(xres - len[target]) > 1e-10;
        len[target] = xres;
        if (moved) rebuild = 1u << target;
        if (step == 3) need_tip = moved;
        step = (step + 1) & 3;
      }
      if (threadIdx.x == 0)
      {
        BloResult res;
        res.logl = ok ? loglikelihood : 0.0;
        res.distal = (orig / (len[0] + len[1])) * len[0];
        res.pendant = len[2];
        a.out[pid] = res;
      }
      continue;
    }
    const double original_length = len[0] * 2;
    double old_d = len[0], old_e = len[2], loglikelihood = 0.0;
    int smoothings = EPA_SMOOTHINGS;
    unsigned rebuild = 7u;
    bool first = true, distal_phase = false;
    for (;;)
    {
      #pragma unroll 1
      for (int mi = 0; mi < 3; ++mi)
        if (rebuild & (1u << mi))
        {
          const double t = mi == 0 ? len[0] : (mi == 1 ? len[1] : len[2]);
          gen_pmatrix<S, R>(sm, t, mi);
          if (mi == 2) gen_tipvec<S, R>(sm);
        }
      rebuild = 0u;
      double xmin, xmax, xguess;
      if (!distal_phase)
      {
        if constexpr (PR) pr.ti = ed.distal < a.tree.n_tips;
        const double new_logl = clvT
            ? -gen_pass_tip_site<S, R, PR>(sm, sum, wpad, clvT + (size_t) ed.distal * t_stride, clvT + (size_t) ed.proximal * t_stride, sD, sX, qc, begin, w, sm + L::TOTAL, inv_w, tm, pr)
            : -gen_pass_tip<S, R>(sm, sum, wpad, D, X, sD, sX, qc, w, inv_w);
        if (first) { loglikelihood = new_logl; first = false; }
        else
        {
          if (new_logl - loglikelihood > new_logl * 1e-14)
          {
            len[2] = old_e; len[0] = old_d; len[1] = original_length - old_d;
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
        if constexpr (PR) pr.ti = 1;
        if (clvT)
          gen_pass_distal_site<S, R, PR>(sm, sum, wpad, clvT + (size_t) ed.distal * t_stride, clvT + (size_t) ed.proximal * t_stride, qc, begin, w, sm + L::TOTAL, inv_w, tm, 1, pr);
        else
          gen_pass_distal<S, R>(sm, sum, wpad, D, X, qc, w, inv_w);
        xmin = fmin(EPA_MIN_BRLEN / 2.0, original_length / 2.0);
        xmax = original_length - xmin / 10.0;
        xguess = len[0];
        if (xguess < xmin || xguess > xmax) xguess = original_length / 2.0;
      }
      const double xres = gen_newton<S, R, PR>(sm, sum, wpad, w, xmin, xguess, xmax, xmin / 10.0, clvT ? sm + L::TOTAL : nullptr, tm, pr.fbuf);
      if (xres > 0.0)
      {
        if (!distal_phase) { len[2] = xres; rebuild = 4u; }
        else { len[0] = xres; len[1] = original_length - xres; rebuild = 3u; }
      }
      distal_phase = !distal_phase;
    }
    if (threadIdx.x == 0)
    {
      BloResult res;
      res.logl = -loglikelihood;
      res.distal = (orig / (len[0] + len[1])) * len[0];
      res.pendant = len[2];
      a.out[pid] = res;
    }
  }
  if (have_tm)
  {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_slot) : "memory");
  }
}

// host-side launcher; scratch is (re)allocated by the caller-owned buffer
template <int S, int R>
inline cudaError_t launch_blo_generic_sr(int sm_count, size_t smem_optin, int max_span, BloArgs & a, void ** scratch,
                                         size_t * scratch_cap, cudaStream_t stream, const double * clvT, size_t t_stride, int use_tmem,
                                         int per_rate = 0, int bugcompat = 0)
{
  using L = GenSmem<S, R>;
  const int wpad = (std::max(1, max_span) + 31) & ~31;
  if (per_rate && (!clvT || R == 1)) return cudaErrorNotSupported;
  // fixed tables + (unit-mapped phases) per-rate partial sums [3][R][wpad] (+ per-rate scalers: weights [R][wpad], flags [wpad])
  const size_t smem = ((size_t) L::TOTAL + (clvT ? (size_t) 3 * R * wpad : 0) + (per_rate ? (size_t) R * wpad + wpad / 8 : 0)) * sizeof(double);
  if (smem > smem_optin) return cudaErrorInvalidConfiguration;
  const int ctas_per_sm = (int) std::max<size_t>(1, std::min<size_t>(3, smem_optin / (smem + 1024)));
  const unsigned grid = (unsigned) std::min<uint64_t>((uint64_t) sm_count * ctas_per_sm, a.n_pairs);
  const size_t need = (size_t) grid * (1 + L::NK) * wpad * sizeof(double);
  if (need > *scratch_cap)
  {
    if (*scratch) dev_free(*scratch);
    *scratch = nullptr; *scratch_cap = 0;
    cudaError_t e = dev_alloc(scratch, need);
    if (e != cudaSuccess) return e;
    *scratch_cap = need;
  }
  a.scratch = static_cast<double *>(*scratch);
  if constexpr (R > 1)
  {
    if (a.raxml && per_rate)
    {
      cudaError_t e = cudaFuncSetAttribute(blo_generic_kernel<S, R, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
      if (e != cudaSuccess) return e;
      blo_generic_kernel<S, R, true, true><<<grid, GEN_THREADS, smem, stream>>>(a, wpad, clvT, t_stride, use_tmem, bugcompat);
      return cudaGetLastError();
    }
  }
  if (a.raxml)
  {
    cudaError_t e = cudaFuncSetAttribute(blo_generic_kernel<S, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    blo_generic_kernel<S, R, true><<<grid, GEN_THREADS, smem, stream>>>(a, wpad, clvT, t_stride, use_tmem);
    return cudaGetLastError();
  }
  if constexpr (R > 1)
  {
    if (per_rate)
    {
      cudaError_t e = cudaFuncSetAttribute(blo_generic_kernel<S, R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
      if (e != cudaSuccess) return e;
      blo_generic_kernel<S, R, false, true><<<grid, GEN_THREADS, smem, stream>>>(a, wpad, clvT, t_stride, use_tmem, bugcompat);
      return cudaGetLastError();
    }
  }
  cudaError_t e = cudaFuncSetAttribute(blo_generic_kernel<S, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  blo_generic_kernel<S, R><<<grid, GEN_THREADS, smem, stream>>>(a, wpad, clvT, t_stride, use_tmem);
  return cudaGetLastError();
}

inline cudaError_t launch_blo_generic(int S, int R, int sm_count, size_t smem_optin, int max_span, const DevModel *,
                                      BloArgs & a, void ** scratch, size_t * scratch_cap, cudaStream_t stream,
                                      const double * clvT = nullptr, size_t t_stride = 0, int use_tmem = 1,
                                      int per_rate = 0, int bugcompat = 0)
{
  if (S == 20 && R == 4) return launch_blo_generic_sr<20, 4>(sm_count, smem_optin, max_span, a, scratch, scratch_cap, stream, clvT, t_stride, use_tmem, per_rate, bugcompat);
  if (S == 20 && R == 1) return launch_blo_generic_sr<20, 1>(sm_count, smem_optin, max_span, a, scratch, scratch_cap, stream, clvT, t_stride, use_tmem, 0, 0);
  return cudaErrorNotSupported;
}

}  // namespace epa
