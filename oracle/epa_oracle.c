/*
 * oracle/epa_oracle.c - TEST INFRASTRUCTURE ONLY (see epa_oracle.h).
 *
 * Scalar CPU restatement of the EPA-ng placement arithmetic. Each function cites the
 * reference code it follows (paths relative to /root/reference; LP = libs/pll-modules/
 * libs/libpll/src, PM = libs/pll-modules/src). Written from the published algorithms and
 * the reference's observable behaviour; no reference source is copied.
 *
 * Not supported (documented in DESIGN.md): ascertainment bias correction, site repeats,
 * pattern weights != 1.
 */
#include "epa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SCALE_FACTOR    0x1p256          /* LP/pll.h:96 */
#define ORC_SCALE_THRESHOLD 0x1p-256         /* LP/pll.h:97 */
#define ORC_RATE_MAXDIFF    4                /* LP/pll.h:104 */
#define ORC_MAX_STATES      20
#define ORC_MAX_RATES       16

/* ========================================================================================== */
/*  Gamma rate categories (LP/gamma.c). The numerical recipes are the published ones:         */
/*  AS 32 (incomplete gamma), Pike & Hill 291 (ln gamma), AS 70 (normal quantile),            */
/*  AS 91 (chi-square quantile), combined as in Yang (1994) discrete gamma.                   */
/* ========================================================================================== */

static double orc_lngamma(double a)
{
  /* Pike & Hill (1966) algorithm 291: shift to x >= 7 then Stirling series */
  double x = a, shift = 0.0;
  if (x < 7.0)
  {
    double prod = 1.0, z = a;
    while (z < 7.0) { prod *= z; z += 1.0; }
    x = z;
    shift = -log(prod);
  }
  double z = 1.0 / (x * x);
  return shift + (x - 0.5) * log(x) - x + .918938533204673
       + (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z
          + .083333333333333) / x;
}

static double orc_incomplete_gamma(double x, double alpha, double lng)
{
  /* AS 32: series for x<=1 or x<alpha, continued fraction otherwise */
  const double accurate = 1e-8, overflow = 1e30;
  if (x == 0) return 0;
  if (x < 0 || alpha <= 0) return -1;
  double factor = exp(alpha * log(x) - x - lng);
  if (!(x > 1 && x >= alpha))
  {
    double gin = 1, term = 1, rn = alpha;
    do { rn += 1; term *= x / rn; gin += term; } while (term > accurate);
    return gin * factor / alpha;
  }
  double a = 1 - alpha, b = a + x + 1, term = 0;
  double pn[6] = {1, x, x + 1, x * b, 0, 0};
  double gin = pn[2] / pn[3];
  for (;;)
  {
    a += 1; b += 2; term += 1;
    double an = a * term;
    pn[4] = b * pn[2] - an * pn[0];
    pn[5] = b * pn[3] - an * pn[1];
    if (pn[5] != 0)
    {
      double rn = pn[4] / pn[5];
      double dif = fabs(gin - rn);
      if (dif <= accurate && dif <= accurate * rn)
        return 1 - factor * gin;
      gin = rn;
    }
    for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
    if (fabs(pn[4]) >= overflow)
      for (int i = 0; i < 4; ++i) pn[i] /= overflow;
  }
}

static double orc_point_normal(double prob)
{
  /* AS 70 (Odeh & Evans 1974) */
  const double a0 = -.322232431088, a1 = -1, a2 = -.342242088547, a3 = -.0204231210245,
               a4 = -.453642210148e-4, b0 = .0993484626060, b1 = .588581570495,
               b2 = .531103462366, b3 = .103537752850, b4 = .0038560700634;
  double p1 = prob < 0.5 ? prob : 1 - prob;
  if (p1 < 1e-20) return -9999;
  double y = sqrt(log(1 / (p1 * p1)));
  double z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0)
               / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
  return prob < 0.5 ? -z : z;
}

static double orc_point_chi2(double prob, double v)
{
  /* AS 91 (Best & Roberts 1975) */
  const double e = .5e-6, aa = .6931471805;
  double p = prob;
  if (p < .000002 || p > .999998 || v <= 0) return -1;
  double g = orc_lngamma(v / 2);
  double xx = v / 2, c = xx - 1, ch;
  int refine = 1;
  if (v < -1.24 * log(p))
  {
    ch = pow(p * xx * exp(g + xx * aa), 1 / xx);
    if (ch - e < 0) return ch;
  }
  else if (v <= .32)
  {
    ch = 0.4;
    double a = log(1 - p), q;
    do
    {
      q = ch;
      double p1 = 1 + ch * (4.67 + ch);
      double p2 = ch * (6.73 + ch * (6.66 + ch));
      double t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
      ch -= (1 - exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (fabs(q / ch - 1) - .01 > 0);
  }
  else
  {
    double x = orc_point_normal(p);
    double p1 = 0.222222 / v;
    ch = v * pow(x * sqrt(p1) + 1 - p1, 3.0);
    if (ch > 2.2 * v + 6) ch = -2 * (log(1 - p) - c * log(.5 * ch) + g);
  }
  (void) refine;
  double q;
  do
  {
    q = ch;
    double p1 = .5 * ch;
    double t = orc_incomplete_gamma(p1, xx, g);
    if (t < 0) return -1;
    double p2 = p - t;
    t = p2 * exp(xx * aa + g + p1 - c * log(ch));
    double b = t / ch, a = 0.5 * t - b * c;
    double s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
    double s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
    double s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
    double s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
    double s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
    double s6 = (120 + c * (346 + 127 * c)) / 5040;
    ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (fabs(q / ch - 1) > e);
  return ch;
}

int orc_gamma_rates(double alpha, int ncat, int median, double * out)
{
  /* LP/gamma.c:220-292 */
  if (alpha < 0.02 || ncat < 1) return 0;
  if (ncat == 1) { out[0] = 1.0; return 1; }
  const double factor = alpha / alpha * ncat, beta = alpha;
  if (median)
  {
    double middle = 1.0 / (2.0 * ncat), t = 0;
    for (int i = 0; i < ncat; ++i)
    {
      out[i] = orc_point_chi2((i * 2 + 1) * middle, 2.0 * alpha) / (2.0 * beta);
      t += out[i];
    }
    for (int i = 0; i < ncat; ++i) out[i] *= factor / t;
    return 1;
  }
  double * cut = (double *) malloc(sizeof(double) * ncat);
  double lnga1 = orc_lngamma(alpha + 1);
  for (int i = 0; i < ncat - 1; ++i)
    cut[i] = orc_point_chi2((i + 1.0) / ncat, 2.0 * alpha) / (2.0 * beta);
  for (int i = 0; i < ncat - 1; ++i)
    cut[i] = orc_incomplete_gamma(cut[i] * beta, alpha + 1, lnga1);
  out[0] = cut[0] * factor;
  out[ncat - 1] = (1 - cut[ncat - 2]) * factor;
  for (int i = 1; i < ncat - 1; ++i) out[i] = (cut[i] - cut[i - 1]) * factor;
  free(cut);
  return 1;
}

/* ========================================================================================== */
/*  Eigen system of the symmetrised rate matrix (LP/models.c:182-410).                        */
/*  The reference uses Householder tridiagonalisation + QL; here a cyclic Jacobi sweep, which */
/*  yields the same decomposition up to ordering / rotation inside degenerate eigenspaces -   */
/*  every quantity downstream (P(t), sumtable contractions) is invariant to that choice.      */
/* ========================================================================================== */

int orc_eigen(int S, const double * subst, const double * freqs,
              double * eigenvals, double * eigenvecs, double * inv_eigenvecs)
{
  if (S > ORC_MAX_STATES) return 0;
  double a[ORC_MAX_STATES][ORC_MAX_STATES], v[ORC_MAX_STATES][ORC_MAX_STATES];
  const int np = S * (S - 1) / 2;
  double last = subst[np - 1];

  /* models.c:182-256: A = sqrt(pi) Q sqrt(pi)^-1, normalised to mean rate 1 */
  for (int i = 0; i < S; ++i) for (int j = 0; j < S; ++j) a[i][j] = 0;
  int k = 0;
  for (int i = 0; i < S; ++i)
    for (int j = i + 1; j < S; ++j)
    {
      double r = subst[k++];
      if (last > 0.0) r /= last;
      a[i][j] = a[j][i] = r * sqrt(freqs[i] * freqs[j]);
      a[i][i] -= r * freqs[j];
      a[j][j] -= r * freqs[i];
    }
  double mean = 0;
  for (int i = 0; i < S; ++i) mean += freqs[i] * (-a[i][i]);
  for (int i = 0; i < S; ++i) for (int j = 0; j < S; ++j) a[i][j] /= mean;

  /* cyclic Jacobi */
  for (int i = 0; i < S; ++i) for (int j = 0; j < S; ++j) v[i][j] = (i == j);
  for (int sweep = 0; sweep < 100; ++sweep)
  {
    double off = 0;
    for (int i = 0; i < S; ++i) for (int j = i + 1; j < S; ++j) off += a[i][j] * a[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < S; ++p)
      for (int q = p + 1; q < S; ++q)
      {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < S; ++r)
        {
          double arp = a[r][p], arq = a[r][q];
          a[r][p] = c * arp - s * arq;
          a[r][q] = s * arp + c * arq;
        }
        for (int r = 0; r < S; ++r)
        {
          double apr = a[p][r], aqr = a[q][r];
          a[p][r] = c * apr - s * aqr;
          a[q][r] = s * apr + c * aqr;
        }
        for (int r = 0; r < S; ++r)
        {
          double vrp = v[r][p], vrq = v[r][q];
          v[r][p] = c * vrp - s * vrq;
          v[r][q] = s * vrp + c * vrq;
        }
      }
  }
  /* column m of v is the m-th eigenvector u_m of A.
     models.c:394-404: eigenvecs[m][j] = u_m[j]*sqrt(pi_j); inv_eigenvecs[i][m] = u_m[i]/sqrt(pi_i) */
  for (int m = 0; m < S; ++m)
  {
    eigenvals[m] = a[m][m];
    for (int j = 0; j < S; ++j)
    {
      eigenvecs[m * S + j] = v[j][m] * sqrt(freqs[j]);
      inv_eigenvecs[j * S + m] = v[j][m] / sqrt(freqs[j]);
    }
  }
  return 1;
}

void orc_invariant_sites(int S, int n_tips, int n, const uint32_t * tip_masks, int * invariant)
{
  /* LP/models.c:651-760: start from the gap state (all bits), AND every tip's mask in */
  const uint32_t gap = S >= 32 ? 0xffffffffu : ((1u << S) - 1u);
  for (int s = 0; s < n; ++s)
  {
    uint32_t st = gap;
    for (int t = 0; t < n_tips; ++t) st &= tip_masks[(size_t) t * n + s];
    invariant[s] = (st == 0 || (st & (st - 1)) != 0) ? -1 : __builtin_ctz(st);
  }
}

void orc_pmatrix(const orc_model_t * m, double t, double * pmat)
{
  /* LP/core_pmatrix.c:185-249: P = I + Vinv diag(expm1(lambda r t)) V ; t == 0 -> I */
  const int S = m->states;
  double expd[ORC_MAX_STATES], temp[ORC_MAX_STATES * ORC_MAX_STATES];
  for (int r = 0; r < m->rate_cats; ++r)
  {
    double * P = pmat + (size_t) r * S * S;
    if (t > 0.)
    {
      /* :209-220: with +I the rates are stretched by 1 / (1 - pinv) */
      if (m->pinv > 1e-8)                     /* PLL_MISC_EPSILON, LP/pll.h:106 */
        for (int j = 0; j < S; ++j) expd[j] = expm1(m->eigenvals[j] * m->rates[r] * t / (1.0 - m->pinv));
      else
        for (int j = 0; j < S; ++j) expd[j] = expm1(m->eigenvals[j] * m->rates[r] * t);
      for (int j = 0; j < S; ++j)
        for (int k = 0; k < S; ++k) temp[j * S + k] = m->inv_eigenvecs[j * S + k] * expd[k];
      for (int j = 0; j < S; ++j)
        for (int k = 0; k < S; ++k)
        {
          double acc = (j == k) ? 1.0 : 0.0;
          for (int q = 0; q < S; ++q) acc += temp[j * S + q] * m->eigenvecs[q * S + k];
          P[j * S + k] = acc;
        }
    }
    else
      for (int j = 0; j < S; ++j) for (int k = 0; k < S; ++k) P[j * S + k] = (j == k);
  }
}

/* ========================================================================================== */
/*  CLV update                                                                                */
/* ========================================================================================== */

/* sum_j M[i][j] * x[j] where x is either a CLV entry vector or the indicator of a tip mask */
static inline double orc_row_dot(const double * row, const double * clv, uint32_t mask, int S)
{
  double acc = 0;
  if (clv)
    for (int j = 0; j < S; ++j) acc += row[j] * clv[j];
  else
    for (int j = 0; j < S; ++j, mask >>= 1) if (mask & 1) acc += row[j];
  return acc;
}

/* test statistic: sites rescaled by the generic tip-inner update under per-rate scalers (see below) */
unsigned long orc_stat_ti_rescaled = 0;

void orc_update_partial(const orc_model_t * m, int n,
                        double * parent_clv, uint32_t * parent_scaler,
                        const orc_side_t * left, const double * lmat,
                        const orc_side_t * right, const double * rmat)
{
  const int S = m->states, R = m->rate_cats, span = S * R;
  const int per_rate = m->per_rate_scalers;
  const int tip_tip = left->tip && right->tip;
  /* Generic (non-4-state) tip-inner update under per-rate scalers, core_partials.c:461-506: the whole site is
     tested and rescaled as with per-site scalers, and the count goes to entry [site index] of the
     [site][rate] array. */
  const int ti_quirk = per_rate && S != 4 && !tip_tip && (left->tip || right->tip);
  const size_t scaler_size = per_rate ? (size_t) n * R : (size_t) n;

  /* core_partials.c:24-46 fill_parent_scaler: parent starts as the sum of the children's */
  if (parent_scaler)
  {
    for (size_t i = 0; i < scaler_size; ++i)
      parent_scaler[i] = (left->scaler && !left->tip ? left->scaler[i] : 0)
                       + (right->scaler && !right->tip ? right->scaler[i] : 0);
  }

  for (int s = 0; s < n; ++s)
  {
    double * out = parent_clv + (size_t) s * span;
    int site_scale = (parent_scaler && (!per_rate || ti_quirk) && !tip_tip) ? 1 : 0;
    for (int r = 0; r < R; ++r)
    {
      const double * lclv = left->clv ? left->clv + (size_t) s * span + r * S : NULL;
      const double * rclv = right->clv ? right->clv + (size_t) s * span + r * S : NULL;
      const uint32_t lmask = left->tip ? left->tip[s] : 0;
      const uint32_t rmask = right->tip ? right->tip[s] : 0;
      int rate_scale = 1;
      for (int i = 0; i < S; ++i)
      {
        double ta = orc_row_dot(lmat + ((size_t) r * S + i) * S, lclv, lmask, S);
        double tb = orc_row_dot(rmat + ((size_t) r * S + i) * S, rclv, rmask, S);
        out[r * S + i] = ta * tb;
        rate_scale &= (out[r * S + i] < ORC_SCALE_THRESHOLD);
      }
      /* tip-tip never rescales (core_partials.c:82-127) */
      if (parent_scaler && per_rate && !tip_tip && !ti_quirk)
      {
        if (rate_scale)
        {
          for (int i = 0; i < S; ++i) out[r * S + i] *= ORC_SCALE_FACTOR;
          parent_scaler[(size_t) s * R + r] += 1;
        }
      }
      else
        site_scale = site_scale && rate_scale;
    }
    if (site_scale)
    {
      for (int i = 0; i < span; ++i) out[i] *= ORC_SCALE_FACTOR;
      parent_scaler[s] += 1;
      if (ti_quirk) orc_stat_ti_rescaled += 1;
    }
  }
}

/* ========================================================================================== */
/*  Edge log-likelihood                                                                       */
/* ========================================================================================== */

/* Collects the scaler counts of one site (core_likelihood.c:1405-1437): returns the per-site
   count, and for per-rate mode fills rel[r] with the capped difference to the minimum. */
static unsigned int orc_site_scalings(const orc_model_t * m, const uint32_t * ps, const uint32_t * cs,
                                      int s, unsigned int * rel)
{
  const int R = m->rate_cats;
  if (!m->per_rate_scalers)
    return (ps ? ps[s] : 0) + (cs ? cs[s] : 0);
  unsigned int mn = UINT32_MAX;
  for (int r = 0; r < R; ++r)
  {
    rel[r] = (ps ? ps[(size_t) s * R + r] : 0) + (cs ? cs[(size_t) s * R + r] : 0);
    if (rel[r] < mn) mn = rel[r];
  }
  for (int r = 0; r < R; ++r)
  {
    unsigned int d = rel[r] - mn;
    rel[r] = d < ORC_RATE_MAXDIFF ? d : ORC_RATE_MAXDIFF;
  }
  return mn;
}

double orc_edge_logl(const orc_model_t * m, int n, const orc_side_t * parent,
                     const orc_side_t * child, const double * pmat, double * persite)
{
  /* Tip handling follows LP/likelihood.c:586-636: if either side is a tip the other one plays
     "parent" (clvp); mathematically the two roles are symmetric for reversible models. */
  const orc_side_t * P = parent, * C = child;
  if (P->tip) { P = child; C = parent; }
  const int S = m->states, R = m->rate_cats, span = S * R;
  double minlh[ORC_RATE_MAXDIFF];
  { double f = 1.0; for (int i = 0; i < ORC_RATE_MAXDIFF; ++i) { f *= ORC_SCALE_THRESHOLD; minlh[i] = f; } }
  unsigned int rel[ORC_MAX_RATES];
  double logl = 0;
  for (int s = 0; s < n; ++s)
  {
    unsigned int site_scalings =
        orc_site_scalings(m, P->scaler, C->tip ? NULL : C->scaler, s, rel);
    double terma = 0, terminv = 0;
    for (int r = 0; r < R; ++r)
    {
      const double * clvp = P->clv + (size_t) s * span + r * S;
      const double * clvc = C->clv ? C->clv + (size_t) s * span + r * S : NULL;
      const uint32_t cmask = C->tip ? C->tip[s] : 0;
      double terma_r = 0;
      for (int j = 0; j < S; ++j)
      {
        double termb = orc_row_dot(pmat + ((size_t) r * S + j) * S, clvc, cmask, S);
        terma_r += clvp[j] * m->freqs[j] * termb;
      }
      if (m->per_rate_scalers && rel[r] > 0) terma_r *= minlh[rel[r] - 1];
      /* core_likelihood.c:524-537: invariant-site mixture */
      if (m->pinv > 0)
      {
        terma += m->weights[r] * terma_r * (1. - m->pinv);
        if (m->invariant[s] != -1) terminv += m->weights[r] * m->freqs[m->invariant[s]] * m->pinv;
      }
      else
        terma += terma_r * m->weights[r];
    }
    double site_lk;
    if (site_scalings)
    {
      if (terminv > 0.)
      {
        /* :543-549: the scaling is undone for the variable term only */
        unsigned int capped = site_scalings < ORC_RATE_MAXDIFF ? site_scalings : ORC_RATE_MAXDIFF;
        site_lk = log(terma * minlh[capped - 1] + terminv);
      }
      else
        site_lk = log(terma) + site_scalings * log(ORC_SCALE_THRESHOLD);
    }
    else
      site_lk = log(terma + terminv);
    if (persite) persite[s] = site_lk;
    logl += site_lk;
  }
  return logl;
}

/* ========================================================================================== */
/*  Sumtable + derivatives                                                                    */
/* ========================================================================================== */

void orc_sumtable(const orc_model_t * m, int n, const orc_side_t * parent,
                  const orc_side_t * child, double * sumtable)
{
  /* core_derivatives.c:321-471 (ii) / :473-641 (ti: the tip always takes the pi*Vinv side) */
  const orc_side_t * L = parent, * Rr = child;
  if (child->tip) { L = child; Rr = parent; }
  const int S = m->states, R = m->rate_cats, span = S * R;
  double minlh[ORC_RATE_MAXDIFF];
  { double f = 1.0; for (int i = 0; i < ORC_RATE_MAXDIFF; ++i) { f *= ORC_SCALE_THRESHOLD; minlh[i] = f; } }
  unsigned int rel[ORC_MAX_RATES];
  for (int s = 0; s < n; ++s)
  {
    if (m->per_rate_scalers)
      orc_site_scalings(m, L->tip ? NULL : L->scaler, Rr->scaler, s, rel);
    for (int r = 0; r < R; ++r)
    {
      const double * lclv = L->clv ? L->clv + (size_t) s * span + r * S : NULL;
      const double * rclv = Rr->clv + (size_t) s * span + r * S;
      uint32_t lmask0 = L->tip ? L->tip[s] : 0;
      double * sum = sumtable + (size_t) s * span + r * S;
      for (int j = 0; j < S; ++j)
      {
        double lefterm = 0, righterm = 0;
        uint32_t lmask = lmask0;
        for (int k = 0; k < S; ++k, lmask >>= 1)
        {
          double x = lclv ? lclv[k] : (double) (lmask & 1);
          lefterm += x * m->freqs[k] * m->inv_eigenvecs[k * S + j];
          righterm += m->eigenvecs[j * S + k] * rclv[k];
        }
        sum[j] = lefterm * righterm;
        if (m->per_rate_scalers && rel[r] > 0) sum[j] *= minlh[rel[r] - 1];
      }
    }
  }
}

/* diagnostic counter (tools/blo_stats.py): derivative evaluations since the last reset */
unsigned long long orc_stat_deriv_calls = 0;
unsigned long long orc_stat_clamped = 0, orc_stat_nr_calls = 0, orc_stat_nr_hist[40] = {0};

void orc_derivatives(const orc_model_t * m, int n, const double * sumtable, double t,
                     double * df, double * ddf)
{
  ++orc_stat_deriv_calls;
  /* core_derivatives.c:643-694 (site kernel), :757-772 (diagptable), :844-847 (accumulate) */
  const int S = m->states, R = m->rate_cats, span = S * R;
  double diag[ORC_MAX_RATES * ORC_MAX_STATES][3];
  for (int r = 0; r < R; ++r)
    for (int j = 0; j < S; ++j)
    {
      double lk = m->eigenvals[j] * (m->rates[r] / (1.0 - m->pinv));      /* :757-772 ki */
      double e = exp(lk * t);
      diag[r * S + j][0] = e;
      diag[r * S + j][1] = lk * e;
      diag[r * S + j][2] = lk * lk * e;
    }
  double d1 = 0, d2 = 0;
  for (int s = 0; s < n; ++s)
  {
    const double * sum = sumtable + (size_t) s * span;
    double lk0 = 0, lk1 = 0, lk2 = 0;
    for (int r = 0; r < R; ++r)
    {
      double c0 = 0, c1 = 0, c2 = 0;
      for (int j = 0; j < S; ++j)
      {
        c0 += sum[r * S + j] * diag[r * S + j][0];
        c1 += sum[r * S + j] * diag[r * S + j][1];
        c2 += sum[r * S + j] * diag[r * S + j][2];
      }
      if (m->pinv > 0)
      {
        /* :676-687: the invariant term enters unscaled, whatever the site's scaler count */
        double inv_site_lk = m->invariant[s] == -1 ? 0 : m->freqs[m->invariant[s]] * m->pinv;
        c0 = c0 * (1. - m->pinv) + inv_site_lk;
        c1 = c1 * (1. - m->pinv);
        c2 = c2 * (1. - m->pinv);
      }
      lk0 += c0 * m->weights[r];
      lk1 += c1 * m->weights[r];
      lk2 += c2 * m->weights[r];
    }
    double deriv1 = -lk1 / lk0;
    double deriv2 = deriv1 * deriv1 - lk2 / lk0;
    d1 += deriv1;
    d2 += deriv2;
  }
  *df = d1;
  *ddf = d2;
}

/* ========================================================================================== */
/*  Bounded Newton-Raphson (PM/optimize/opt_algorithms.c:133-262, single variable)            */
/* ========================================================================================== */

typedef struct { const orc_model_t * m; int n; const double * sumtable; } orc_nr_ctx_t;

/* returns the optimum, or 0.0 on failure (pllmod_opt_minimize_newton returns PLL_FAILURE) */
static double orc_newton(double xmin, double xguess, double xmax, double tol, int max_iters,
                         const orc_nr_ctx_t * ctx)
{
  double x = fmax(fmin(xguess, xmax), xmin);
  double xl = xmin, xh = xmax;
  const double dxmax = xmax / max_iters;
  int iter = 0;
  ++orc_stat_nr_calls;
  for (;;)
  {
    if (iter < 40) orc_stat_nr_hist[iter]++;
    if (iter++ > max_iters) return 0.0;
    double f, df;
    orc_derivatives(ctx->m, ctx->n, ctx->sumtable, x, &f, &df);
    if (!isfinite(f) || !isfinite(df)) return 0.0;
    double dx;
    if (df > 0.0)
    {
      if (fabs(f) < tol) return x;
      if (f < 0.0) xl = x; else xh = x;
      dx = -1 * f / df;
    }
    else
      dx = -1 * f / fabs(df);
    if (fabs(dx) > dxmax) ++orc_stat_clamped;
    dx = fmax(fmin(dx, dxmax), -dxmax);
    if (x + dx < xl) dx = xl - x;
    if (x + dx > xh) dx = xh - x;
    if (fabs(dx) < tol) return x;
    x += dx;
    x = fmax(fmin(x, xmax), xmin);
  }
}

/* ========================================================================================== */
/*  Tiny tree                                                                                 */
/* ========================================================================================== */

#define ORC_DEFAULT_PENDANT (-log(0.9))     /* src/util/constants.hpp:12 */
#define ORC_MIN_BRLEN 1.0e-4                /* PM/optimize/pll_optimize.h:57 */
#define ORC_MAX_BRLEN 100.                  /* :58 */
#define ORC_DEFAULT_BRLEN 0.1               /* :54 */
#define ORC_BLO_EPSILON 1e-1                /* src/core/pll/optimize.hpp:9 */

void orc_tiny_inner(const orc_model_t * m, int n, const orc_side_t * distal,
                    const orc_side_t * proximal, double orig_length,
                    double * inner_clv, uint32_t * inner_scaler)
{
  /* Tiny_Tree.cpp:84-112: child1 = distal, child2 = proximal, both at orig/2 */
  const int S = m->states, R = m->rate_cats;
  double * pm = (double *) malloc(sizeof(double) * R * S * S);
  orc_pmatrix(m, orig_length / 2.0, pm);
  orc_update_partial(m, n, inner_clv, inner_scaler, distal, pm, proximal, pm);
  free(pm);
}

void orc_lookup_build(const orc_model_t * m, int n, const orc_side_t * distal,
                      const orc_side_t * proximal, double orig_length,
                      const uint32_t * char_masks, int K, double * lookup)
{
  const int S = m->states, R = m->rate_cats;
  double * inner = (double *) malloc(sizeof(double) * (size_t) n * R * S);
  uint32_t * iscal = (uint32_t *) calloc((size_t) n * (m->per_rate_scalers ? R : 1), sizeof(uint32_t));
  double * ppend = (double *) malloc(sizeof(double) * R * S * S);
  double * persite = (double *) malloc(sizeof(double) * n);
  uint32_t * tip = (uint32_t *) malloc(sizeof(uint32_t) * n);
  orc_tiny_inner(m, n, distal, proximal, orig_length, inner, iscal);
  orc_pmatrix(m, ORC_DEFAULT_PENDANT, ppend);
  orc_side_t in = {inner, iscal, NULL};
  for (int k = 0; k < K; ++k)
  {
    for (int s = 0; s < n; ++s) tip[s] = char_masks[k];
    orc_side_t q = {NULL, NULL, tip};
    orc_edge_logl(m, n, &q, &in, ppend, persite);
    for (int s = 0; s < n; ++s) lookup[(size_t) s * K + k] = persite[s];
  }
  free(inner); free(iscal); free(ppend); free(persite); free(tip);
}

double orc_preplace_score(const double * lookup, int K, const uint8_t * cols, int begin, int span)
{
  /* Lookup_Store.hpp:110-141: groups of four ((a+b)+(c+d)), then the tail one by one */
  double sum = 0;
  int site = begin;
  const int end = begin + span;
  for (; site + 3 < end; site += 4)
  {
    double one = lookup[(size_t) site * K + cols[site]] + lookup[(size_t) (site + 1) * K + cols[site + 1]];
    double two = lookup[(size_t) (site + 2) * K + cols[site + 2]] + lookup[(size_t) (site + 3) * K + cols[site + 3]];
    one += two;
    sum += one;
  }
  for (; site < end; ++site) sum += lookup[(size_t) site * K + cols[site]];
  return sum;
}

static orc_side_t orc_focus(const orc_model_t * m, const orc_side_t * s, int begin)
{
  /* pll_util.cpp:388-418 shift_partition_focus */
  const int span = m->states * m->rate_cats;
  orc_side_t f = {NULL, NULL, NULL};
  if (s->tip) f.tip = s->tip + begin;
  if (s->clv) f.clv = s->clv + (size_t) begin * span;
  if (s->scaler)
    f.scaler = s->scaler + ((m->per_rate_scalers && !m->bugcompat_focus) ? (size_t) begin * m->rate_cats
                                                                        : (size_t) begin);
  return f;
}

void orc_place_thorough(const orc_model_t * m_full, int n_full, const orc_side_t * distal_full,
                        const orc_side_t * proximal_full, double orig_length,
                        const uint32_t * query_tip, int begin, int span,
                        orc_blo_result_t * out)
{
  (void) n_full;
  /* pll_util.cpp:413-414: the invariant array moves with the window */
  orc_model_t m_focus = *m_full;
  if (m_focus.invariant) m_focus.invariant += begin;
  const orc_model_t * m = &m_focus;
  const int S = m->states, R = m->rate_cats, n = span;
  const size_t psz = (size_t) R * S * S;
  const orc_side_t distal = orc_focus(m, distal_full, begin);
  const orc_side_t proximal = orc_focus(m, proximal_full, begin);
  const orc_side_t tip = {NULL, NULL, query_tip + begin};

  double * inner = (double *) malloc(sizeof(double) * (size_t) n * R * S);
  uint32_t * iscal = (uint32_t *) calloc((size_t) n * (m->per_rate_scalers ? R : 1), sizeof(uint32_t));
  double * sumtable = (double *) malloc(sizeof(double) * (size_t) n * R * S);
  double * p_dist = (double *) malloc(sizeof(double) * psz);
  double * p_prox = (double *) malloc(sizeof(double) * psz);
  double * p_pend = (double *) malloc(sizeof(double) * psz);
  const orc_side_t in = {inner, iscal, NULL};
  const orc_nr_ctx_t nr = {m, n, sumtable};

  /* optimize.cpp:253-286 optimize_branch_triplet -> traverse_update_partials */
  double len_dist = orig_length / 2.0, len_prox = orig_length / 2.0, len_pend = ORC_DEFAULT_PENDANT;
  orc_pmatrix(m, len_dist, p_dist);
  orc_pmatrix(m, len_prox, p_prox);
  orc_pmatrix(m, len_pend, p_pend);
  orc_update_partial(m, n, inner, iscal, &distal, p_dist, &proximal, p_prox);

  /* optimize.cpp:60-248 opt_branch_lengths_pplacer */
  const int max_iters = 30;
  const double original_length = len_dist * 2;
  double loglikelihood = -orc_edge_logl(m, n, &tip, &in, p_pend, NULL);
  int smoothings = 32;
  out->rounds = 0;
  out->restored = 0;
  while (smoothings)
  {
    const double old_dist = len_dist, old_pend = len_pend;
    out->rounds++;

    /* pendant */
    double xmin = ORC_MIN_BRLEN, xmax = ORC_MAX_BRLEN, xtol = xmin / 10.0, xguess = len_pend;
    if (xguess < xmin || xguess > xmax) xguess = ORC_DEFAULT_BRLEN;
    orc_sumtable(m, n, &in, &tip, sumtable);
    double xres = orc_newton(xmin, xguess, xmax, xtol, max_iters, &nr);
    if (xres > 0.0)
    {
      len_pend = xres;
      orc_pmatrix(m, len_pend, p_pend);
    }

    /* distal: inner CLV now looks toward the distal node (new tip x proximal) */
    orc_update_partial(m, n, inner, iscal, &tip, p_pend, &proximal, p_prox);
    xguess = len_dist;
    xmin = fmin(ORC_MIN_BRLEN / 2.0, original_length / 2.0);
    xtol = xmin / 10.0;
    xmax = original_length - xtol;
    if (xguess < xmin || xguess > xmax) xguess = original_length / 2.0;
    orc_sumtable(m, n, &distal, &in, sumtable);
    xres = orc_newton(xmin, xguess, xmax, xtol, max_iters, &nr);
    if (xres > 0.0)
    {
      len_dist = xres;
      len_prox = original_length - xres;
      orc_pmatrix(m, len_dist, p_dist);
      orc_pmatrix(m, len_prox, p_prox);
    }

    /* score */
    orc_update_partial(m, n, inner, iscal, &distal, p_dist, &proximal, p_prox);
    double new_logl = -orc_edge_logl(m, n, &tip, &in, p_pend, NULL);
    if (new_logl - loglikelihood > new_logl * 1e-14)
    {
      len_pend = old_pend;
      len_dist = old_dist;
      len_prox = original_length - old_dist;
      out->restored = 1;
      break;
    }
    --smoothings;
    if (fabs(new_logl - loglikelihood) < ORC_BLO_EPSILON) smoothings = 0;
    loglikelihood = new_logl;
  }

  /* Tiny_Tree.cpp:183-185 */
  out->logl = -loglikelihood;
  out->distal = (orig_length / (len_dist + len_prox)) * len_dist;
  out->pendant = len_pend;

  free(inner); free(iscal); free(sumtable); free(p_dist); free(p_prox); free(p_pend);
}

/* ------------------------------------------------------------------------------------------ */
/*  --raxml-blo: optimize_branch_triplet with sliding == false (src/core/pll/optimize.cpp:274-278) */
/*  -> pllmod_opt_optimize_branch_lengths_local(radius 1, keep_update 1), PM/optimize/          */
/*  pll_optimize.c:778-1097, Newton variant PM/optimize/opt_algorithms.c:281-384                */
/* ------------------------------------------------------------------------------------------ */

/* pllmod_opt_minimize_newton_old: Newton-Raphson with a bisection fallback. *failed is set where
   the reference sets pll_errno (non-finite derivatives, iteration limit). */
static double orc_newton_old(double x1, double xguess, double x2, double tol, int max_iters,
                             const orc_nr_ctx_t * ctx, int * failed)
{
  double df, dx, f, temp, xh, xl, rts, rts_old = 0.0;
  *failed = 0;
  rts = xguess;
  if (rts < x1) rts = x1;
  if (rts > x2) rts = x2;
  orc_derivatives(ctx->m, ctx->n, ctx->sumtable, rts, &f, &df);
  if (!isfinite(f) || !isfinite(df)) { *failed = 1; return -INFINITY; }
  if (df >= 0.0 && fabs(f) < tol) return rts;
  if (f < 0.0) { xl = rts; xh = x2; }
  else { xh = rts; xl = x1; }
  dx = fabs(xh - xl);
  for (int i = 1; i <= max_iters; i++)
  {
    rts_old = rts;
    if ((df <= 0.0) || (((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0))
    {
      dx = 0.5 * (xh - xl);
      rts = xl + dx;
      if (xl == rts) return rts;
    }
    else
    {
      dx = f / df;
      temp = rts;
      rts -= dx;
      if (temp == rts) return rts;
    }
    if (fabs(dx) < tol) return rts_old;
    if (i == max_iters) break;
    if (rts < x1) rts = x1;
    orc_derivatives(ctx->m, ctx->n, ctx->sumtable, rts, &f, &df);
    if (!isfinite(f) || !isfinite(df)) { *failed = 1; return -INFINITY; }
    if (df > 0.0 && fabs(f) < tol) return rts;
    if (f < 0.0) xl = rts; else xh = rts;
  }
  *failed = 1;                               /* "Exceeded maximum number of iterations" */
  return rts_old;
}

/* one recomp_iterative step on a single edge (pll_optimize.c:799-833): sumtable across the edge,
   Newton, length and (if it moved by more than 1e-10) transition matrix updated. Returns 0 on failure. */
static int orc_raxml_edge(const orc_model_t * m, int n, const orc_side_t * a, const orc_side_t * b,
                          double * sumtable, double * len, double * pmat)
{
  const orc_nr_ctx_t nr = {m, n, sumtable};
  const double xmin = ORC_MIN_BRLEN, xmax = ORC_MAX_BRLEN, xtol = ORC_MIN_BRLEN / 10.0, xorig = *len;
  double xguess = *len;
  if (xguess < xmin || xguess > xmax) xguess = ORC_DEFAULT_BRLEN;
  orc_sumtable(m, n, a, b, sumtable);
  int failed;
  const double xres = orc_newton_old(xmin, xguess, xmax, xtol, 30, &nr, &failed);
  if (failed) return 0;
  *len = xres;
  if (fabs(xres - xorig) > 1e-10) orc_pmatrix(m, xres, pmat);
  return 1;
}

void orc_place_thorough_raxml(const orc_model_t * m_full, int n_full, const orc_side_t * distal_full,
                              const orc_side_t * proximal_full, double orig_length,
                              const uint32_t * query_tip, int begin, int span,
                              orc_blo_result_t * out)
{
  (void) n_full;
  orc_model_t m_focus = *m_full;
  if (m_focus.invariant) m_focus.invariant += begin;
  const orc_model_t * m = &m_focus;
  const int S = m->states, R = m->rate_cats, n = span;
  const size_t psz = (size_t) R * S * S;
  const orc_side_t distal = orc_focus(m, distal_full, begin);
  const orc_side_t proximal = orc_focus(m, proximal_full, begin);
  const orc_side_t tip = {NULL, NULL, query_tip + begin};

  double * inner = (double *) malloc(sizeof(double) * (size_t) n * R * S);
  uint32_t * iscal = (uint32_t *) calloc((size_t) n * (m->per_rate_scalers ? R : 1), sizeof(uint32_t));
  double * sumtable = (double *) malloc(sizeof(double) * (size_t) n * R * S);
  double * p_dist = (double *) malloc(sizeof(double) * psz);
  double * p_prox = (double *) malloc(sizeof(double) * psz);
  double * p_pend = (double *) malloc(sizeof(double) * psz);
  const orc_side_t in = {inner, iscal, NULL};

  double len_dist = orig_length / 2.0, len_prox = orig_length / 2.0, len_pend = ORC_DEFAULT_PENDANT;
  orc_pmatrix(m, len_dist, p_dist);
  orc_pmatrix(m, len_prox, p_prox);
  orc_pmatrix(m, len_pend, p_pend);
  orc_update_partial(m, n, inner, iscal, &distal, p_dist, &proximal, p_prox);

  /* pll_optimize.c:991-1091 */
  double loglikelihood = orc_edge_logl(m, n, &tip, &in, p_pend, NULL);
  int iters = 32, ok = 1;
  out->rounds = 0;
  out->restored = 0;
  while (iters)
  {
    out->rounds++;
    /* first edge, radius 1: pendant, then the edges behind the inner node's other two directions */
    ok = orc_raxml_edge(m, n, &in, &tip, sumtable, &len_pend, p_pend);
    if (!ok) break;
    orc_update_partial(m, n, inner, iscal, &tip, p_pend, &proximal, p_prox);     /* toward distal */
    ok = orc_raxml_edge(m, n, &distal, &in, sumtable, &len_dist, p_dist);
    if (!ok) break;
    orc_update_partial(m, n, inner, iscal, &distal, p_dist, &tip, p_pend);       /* toward proximal */
    ok = orc_raxml_edge(m, n, &proximal, &in, sumtable, &len_prox, p_prox);
    if (!ok) break;
    orc_update_partial(m, n, inner, iscal, &proximal, p_prox, &distal, p_dist);  /* back toward the tip */
    /* second edge (the new tip), radius 0: the pendant edge once more */
    ok = orc_raxml_edge(m, n, &tip, &in, sumtable, &len_pend, p_pend);
    if (!ok) break;
    const double new_logl = orc_edge_logl(m, n, &tip, &in, p_pend, NULL);
    if (new_logl - loglikelihood > new_logl * 1e-13)
    {
      --iters;
      if (fabs(new_logl - loglikelihood) < ORC_BLO_EPSILON) iters = 0;
      loglikelihood = new_logl;
    }
    else
    {
      /* PLLMOD_OPT_BLO_NEWTON_OLDFAST: a worse score is kept and ends the loop (:1084-1088) */
      loglikelihood = new_logl;
      out->restored = 1;
      break;
    }
  }
  /* a failed Newton call makes the optimiser return PLL_FAILURE (0): the placement carries logl 0 */
  out->logl = ok ? loglikelihood : 0.0;
  out->distal = (orig_length / (len_dist + len_prox)) * len_dist;
  out->pendant = len_pend;

  free(inner); free(iscal); free(sumtable); free(p_dist); free(p_prox); free(p_pend);
}

/* ========================================================================================== */
/*  Candidate selection / output stage                                                        */
/* ========================================================================================== */

void orc_lwr(const double * logl, int n, double * lwr)
{
  /* set_manipulators.cpp:43-69 */
  double mx = logl[0], total = 0;
  for (int i = 1; i < n; ++i) if (mx < logl[i]) mx = logl[i];
  for (int i = 0; i < n; ++i) { lwr[i] = exp(logl[i] - mx); total += lwr[i]; }
  for (int i = 0; i < n; ++i) lwr[i] /= total;
}

typedef struct { double lwr; int idx; } orc_rank_t;
static int orc_rank_cmp(const void * a, const void * b)
{
  const orc_rank_t * x = (const orc_rank_t *) a, * y = (const orc_rank_t *) b;
  if (x->lwr > y->lwr) return -1;
  if (x->lwr < y->lwr) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

int orc_select_accumulated(const double * lwr, int n, double thresh, int * out_idx)
{
  /* set_manipulators.cpp:90-113 with min = 1, max = unlimited */
  orc_rank_t * r = (orc_rank_t *) malloc(sizeof(orc_rank_t) * n);
  for (int i = 0; i < n; ++i) { r[i].lwr = lwr[i]; r[i].idx = i; }
  qsort(r, n, sizeof(orc_rank_t), orc_rank_cmp);
  double sum = 0;
  int k = 0;
  for (; k < n && sum < thresh; ++k) sum += r[k].lwr;
  if (k < 1) k = 1;
  for (int i = 0; i < k; ++i) out_idx[i] = r[i].idx;
  free(r);
  return k;
}

int orc_filter_support(const double * lwr_sorted, int n, double thresh, int min, int max)
{
  /* set_manipulators.cpp:131-163: keep lwr > thresh, at least min, at most max */
  int kept = 0;
  while (kept < n && lwr_sorted[kept] > thresh) kept++;
  int res = kept;
  if (kept < min) res = min;
  if (max && kept > max) res = max;
  if (res > n) res = n;
  return res;
}
