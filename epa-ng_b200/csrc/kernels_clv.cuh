// kernels_clv.cuh - reference-tree side of the hot path: transition matrices, directional CLVs,
// per-edge lookup tables, edge log-likelihood.
//
// Reference behaviour restated (paths relative to /root/reference, LP = libs/pll-modules/libs/libpll/src):
//   pmatrix            LP/core_pmatrix.c:185-249
//   CLV update         LP/core_partials.c:202-352 (tip-inner), :612-766 (inner-inner), :24-46 (scalers)
//   edge logl / lookup LP/core_likelihood.c:441-578, src/tree/Tiny_Tree.cpp:18-46,84-128,
//                      src/core/Lookup_Store.hpp:73-81
#pragma once
#include "common.cuh"

namespace epa {

// ---------------------------------------------------------------------------------------------
// P(t) = I + Vinv diag(expm1(lambda * rate * t)) V  for a list of branch lengths.
// grid = n_mats, block = 128. out[m][r][i][j]
// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void pmatrix_kernel(const DevModel * __restrict__ m, const double * __restrict__ lengths,
                               double * __restrict__ out)
{
  __shared__ double expd[MAX_RATES * S];
  const int R = m->R;
  const double t = lengths[blockIdx.x];
  for (int idx = threadIdx.x; idx < R * S; idx += blockDim.x)
    expd[idx] = expm1(m->eigenvals[idx % S] * m->rates[idx / S] * t);
  __syncthreads();
  double * P = out + (size_t) blockIdx.x * R * S * S;
  for (int idx = threadIdx.x; idx < R * S * S; idx += blockDim.x)
  {
    const int r = idx / (S * S), i = (idx / S) % S, j = idx % S;
    double acc = (i == j) ? 1.0 : 0.0;
    #pragma unroll
    for (int k = 0; k < S; ++k)
      acc += (m->inv_eigenvecs[i * S + k] * expd[r * S + k]) * m->eigenvecs[k * S + j];
    P[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Tip state masks -> 0/1 CLVs (all rate blocks identical), scaler 0. One thread per (tip, site).
// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void tip_expand_kernel(const DevModel * __restrict__ m, DevTree tree,
                                  const uint32_t * __restrict__ masks, size_t total)
{
  const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int R = m->R;
  const uint32_t mask = masks[idx];
  double * dst = tree.clv + idx * (size_t) (R * S);
  double v[S];
  #pragma unroll
  for (int j = 0; j < S; ++j) v[j] = (mask >> j) & 1u ? 1.0 : 0.0;
  for (int r = 0; r < R; ++r) store_vec<S>(dst + r * S, v);
  tree.scaler[idx] = 0;
}

// stage `count` doubles from global into shared memory (whole block)
__device__ __forceinline__ void stage_doubles(double * dst, const double * __restrict__ src, int count)
{
  for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
}

// ---------------------------------------------------------------------------------------------
// Felsenstein pruning step, one thread per site, all ops of one dependency level in one launch.
// grid = (n_ops, ceil(n / blockDim)), dynamic smem = 2*R*S*S doubles.
// ---------------------------------------------------------------------------------------------
template <int S, int R>
__global__ void __launch_bounds__(128)
clv_update_kernel(DevTree tree, int n, const ClvOpDev * __restrict__ ops, const double * __restrict__ pmats)
{
  extern __shared__ double smem[];
  double * Pl = smem;
  double * Pr = smem + R * S * S;
  const ClvOpDev op = ops[blockIdx.x];
  stage_doubles(Pl, pmats + (size_t) op.lmat * R * S * S, R * S * S);
  stage_doubles(Pr, pmats + (size_t) op.rmat * R * S * S, R * S * S);
  __syncthreads();
  const int site = blockIdx.y * blockDim.x + threadIdx.x;
  if (site >= n) return;

  const double * L = tree.clv + op.left * tree.clv_stride + (size_t) site * (R * S);
  const double * Rc = tree.clv + op.right * tree.clv_stride + (size_t) site * (R * S);
  double * out = tree.clv + op.parent * tree.clv_stride + (size_t) site * (R * S);

  double res[R][S];
  bool all_small = true;
  #pragma unroll
  for (int r = 0; r < R; ++r)
  {
    double lv[S], rv[S];
    load_vec<S>(L + r * S, lv);
    load_vec<S>(Rc + r * S, rv);
    #pragma unroll
    for (int i = 0; i < S; ++i)
    {
      double ta = 0.0, tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S; ++j)
      {
        ta += Pl[(r * S + i) * S + j] * lv[j];
        tb += Pr[(r * S + i) * S + j] * rv[j];
      }
      res[r][i] = ta * tb;
      all_small = all_small && (res[r][i] < EPA_SCALE_THRESHOLD);
    }
  }
  if (tree.sr > 1)
  {
    // per-rate scalers (PLL_ATTRIB_RATE_SCALERS, LP/core_partials.c:690-766): every rate block is
    // rescaled on its own and keeps its own count
    const uint32_t * sl = tree.scaler + ((size_t) op.left * n + site) * R;
    const uint32_t * sr = tree.scaler + ((size_t) op.right * n + site) * R;
    uint32_t * sp = tree.scaler + ((size_t) op.parent * n + site) * R;
    if (S != 4 && op.tip_tip == 2)
    {
      // The reference runs libpll's generic kernels under per-rate scalers, and the generic tip-inner update
      // (LP/core_partials.c:461-506) tests and rescales the WHOLE site and counts it in entry [site index] of the
      // parent's [site][rate] array. That entry belongs to another thread's site: the parent's counters are zeroed
      // before the launch and every contribution is an integer atomic add (order independent).
      #pragma unroll
      for (int r = 0; r < R; ++r)
      {
        if (all_small)
        {
          #pragma unroll
          for (int i = 0; i < S; ++i) res[r][i] *= EPA_SCALE_FACTOR;
        }
        store_vec<S>(out + r * S, res[r]);
        const uint32_t sc = sl[r] + sr[r];
        if (sc) atomicAdd(sp + r, sc);
      }
      if (all_small) atomicAdd(tree.scaler + (size_t) op.parent * n * R + site, 1u);
      return;
    }
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      bool small = true;
      #pragma unroll
      for (int i = 0; i < S; ++i) small = small && (res[r][i] < EPA_SCALE_THRESHOLD);
      uint32_t sc = sl[r] + sr[r];
      if (small && op.tip_tip != 1)
      {
        sc += 1;
        #pragma unroll
        for (int i = 0; i < S; ++i) res[r][i] *= EPA_SCALE_FACTOR;
      }
      store_vec<S>(out + r * S, res[r]);
      sp[r] = sc;
    }
    return;
  }
  uint32_t sc = tree.scaler[(size_t) op.left * n + site] + tree.scaler[(size_t) op.right * n + site];
  const bool scale = all_small && op.tip_tip != 1;
  if (scale) sc += 1;
  #pragma unroll
  for (int r = 0; r < R; ++r)
  {
    if (scale)
    {
      #pragma unroll
      for (int i = 0; i < S; ++i) res[r][i] *= EPA_SCALE_FACTOR;
    }
    store_vec<S>(out + r * S, res[r]);
  }
  tree.scaler[(size_t) op.parent * n + site] = sc;
}

// ---------------------------------------------------------------------------------------------
// Column table of the lookup build: M[r][c][i] = pi_i * sum_{j in colmask[c]} Ppend[r][i][j].
// One block; Ppend = P(-ln 0.9) is pmats[pend_index].
// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void lookup_coltable_kernel(const DevModel * __restrict__ m, const double * __restrict__ ppend,
                                       double * __restrict__ M)
{
  const int R = m->R, K = m->K;
  for (int idx = threadIdx.x; idx < R * K * S; idx += blockDim.x)
  {
    const int r = idx / (K * S), c = (idx / S) % K, i = idx % S;
    const uint32_t mask = m->colmask[c];
    double termb = 0.0;
    #pragma unroll
    for (int j = 0; j < S; ++j)
      if ((mask >> j) & 1u) termb += ppend[(r * S + i) * S + j];
    M[idx] = m->freqs[i] * termb;
  }
}

// ---------------------------------------------------------------------------------------------
// Per-edge lookup table, one thread per site (generic path; any S, any compiled R).
//   inner[r][i] = (P(len/2) D)[r][i] * (P(len/2) X)[r][i]          (+ per-site scaling)
//   lookup[e][s][c] = log( sum_r w_r sum_i inner[r][i] M[r][c][i] ) + scaler * log(2^-256)
// Column c with colmask 0 is the zero column (preplacement reads it for out-of-range sites).
// grid = (n_edges, ceil(n/blockDim)); dynamic smem = (R*S*S + R*K*S) doubles.
// ---------------------------------------------------------------------------------------------
// Generic kernels under per-rate scalers (amino acids): the inner CLV of a tiny tree whose distal node is a tip
// comes from libpll's generic tip-inner update, which rescales a WHOLE site when all its entries are small and
// counts that in entry [site index] of the inner node's [site][rate] array (LP/core_partials.c:461-506, see
// clv_update_kernel). flags[e][s] = 1 where site s of tip edge e is rescaled that way; other edges are skipped.
// grid = (n_edges, ceil(n/blockDim)); dynamic smem = R*S*S doubles.
template <int S, int R>
__global__ void __launch_bounds__(128)
lookup_ti_flags_kernel(DevTree tree, int n, const EdgeDev * __restrict__ edges, const double * __restrict__ pmats_half,
                       uint8_t * __restrict__ flags)
{
  extern __shared__ double smem[];
  double * P = smem;
  const EdgeDev e = edges[blockIdx.x];
  if (e.distal >= tree.n_tips) return;
  stage_doubles(P, pmats_half + (size_t) blockIdx.x * R * S * S, R * S * S);
  __syncthreads();
  const int site = blockIdx.y * blockDim.x + threadIdx.x;
  if (site >= n) return;
  const double * D = tree.clv + e.distal * tree.clv_stride + (size_t) site * (R * S);
  const double * X = tree.clv + e.proximal * tree.clv_stride + (size_t) site * (R * S);
  bool all_small = true;
  #pragma unroll 1
  for (int r = 0; r < R; ++r)
  {
    double dv[S], xv[S];
    load_vec<S>(D + r * S, dv);
    load_vec<S>(X + r * S, xv);
    #pragma unroll 4
    for (int i = 0; i < S; ++i)
    {
      double ta = 0.0, tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S; ++j)
      {
        ta += P[(r * S + i) * S + j] * dv[j];
        tb += P[(r * S + i) * S + j] * xv[j];
      }
      all_small = all_small && (ta * tb < EPA_SCALE_THRESHOLD);
    }
  }
  flags[(size_t) blockIdx.x * n + site] = all_small ? 1 : 0;
}

template <int S, int R>
__global__ void __launch_bounds__(128)
lookup_build_kernel(const DevModel * __restrict__ m, DevTree tree, int n, int n_pad, int K,
                    const EdgeDev * __restrict__ edges, const double * __restrict__ pmats_half,
                    const double * __restrict__ coltab, double * __restrict__ lookup,
                    const uint8_t * __restrict__ ti_flags = nullptr)
{
  extern __shared__ double smem[];
  double * P = smem;                    // [R][S][S]
  double * M = smem + R * S * S;        // [R][K][S]
  const EdgeDev e = edges[blockIdx.x];
  stage_doubles(P, pmats_half + (size_t) blockIdx.x * R * S * S, R * S * S);
  stage_doubles(M, coltab, R * K * S);
  __syncthreads();
  const int site = blockIdx.y * blockDim.x + threadIdx.x;
  if (site >= n) return;

  const double * D = tree.clv + e.distal * tree.clv_stride + (size_t) site * (R * S);
  const double * X = tree.clv + e.proximal * tree.clv_stride + (size_t) site * (R * S);
  double inner[R][S];
  bool all_small = true;
  #pragma unroll
  for (int r = 0; r < R; ++r)
  {
    double dv[S], xv[S];
    load_vec<S>(D + r * S, dv);
    load_vec<S>(X + r * S, xv);
    #pragma unroll
    for (int i = 0; i < S; ++i)
    {
      double ta = 0.0, tb = 0.0;
      #pragma unroll
      for (int j = 0; j < S; ++j)
      {
        ta += P[(r * S + i) * S + j] * dv[j];
        tb += P[(r * S + i) * S + j] * xv[j];
      }
      inner[r][i] = ta * tb;
      all_small = all_small && (inner[r][i] < EPA_SCALE_THRESHOLD);
    }
  }
  uint32_t sc;
  if (tree.sr > 1)
  {
    // per-rate scalers (as in lookup_build_site_kernel): rate weights 2^(-256 d) relative to the site's minimum count
    // (ti_flags: the whole-site rescalings of a tip edge's inner CLV and where the reference counts them)
    const uint8_t * tf = (ti_flags && e.distal < tree.n_tips) ? ti_flags + (size_t) blockIdx.x * n : nullptr;
    uint32_t kr[R], kmin = 0xffffffffu;
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      kr[r] = tree.scaler[((size_t) e.distal * n + site) * R + r] + tree.scaler[((size_t) e.proximal * n + site) * R + r];
      if (tf && site * R + r < n) kr[r] += tf[site * R + r];
      kmin = min(kmin, kr[r]);
    }
    const double whole = (tf && tf[site]) ? EPA_SCALE_FACTOR : 1.0;
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      const double f = rate_scale_factor(min(kr[r] - kmin, EPA_RATE_MAXDIFF)) * whole;
      #pragma unroll
      for (int i = 0; i < S; ++i) inner[r][i] *= f;
    }
    sc = kmin;
  }
  else
  {
    sc = tree.scaler[(size_t) e.distal * n + site] + tree.scaler[(size_t) e.proximal * n + site];
    if (all_small)
    {
      sc += 1;
      #pragma unroll
      for (int r = 0; r < R; ++r)
        #pragma unroll
        for (int i = 0; i < S; ++i) inner[r][i] *= EPA_SCALE_FACTOR;
    }
  }
  const double inv = tree.inv ? __ldg(tree.inv + site) : 0.0;
  double wr[R];
  #pragma unroll
  for (int r = 0; r < R; ++r) wr[r] = m->weights[r];

  double * out = lookup + ((size_t) blockIdx.x * n_pad + site) * K;
  for (int c = 0; c < K; ++c)
  {
    double v = 0.0;
    if (m->colmask[c] != 0)
    {
      double terma = 0.0;
      #pragma unroll
      for (int r = 0; r < R; ++r)
      {
        double tr = 0.0;
        #pragma unroll
        for (int i = 0; i < S; ++i) tr += inner[r][i] * M[(r * K + c) * S + i];
        terma += tr * wr[r];
      }
      v = site_loglk(terma, sc, inv);
    }
    out[c] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// DNA fast path of the lookup build (S = 4, R = 4, K = 16): four lanes per site, lane = rate.
// Every lane loads one 32-byte sector of each CLV (a warp covers 1 KB contiguous per CLV), the
// rate sum and the scaling decision run over 2 xor-shuffles, and each lane finishes 4 of the 16
// columns (log + 32-byte store; the four lanes of a site write one 128-byte line).
// grid = (n_edges, ceil(n / LOOKUP_DNA_SITES_PER_BLOCK)), block = 256 threads = 64 sites per trip.
// ---------------------------------------------------------------------------------------------
constexpr int LOOKUP_DNA_SITES_PER_BLOCK = 256;

__global__ void __launch_bounds__(256, 3)
lookup_build_dna_kernel(const DevModel * __restrict__ m, DevTree tree, int n, int n_pad,
                        const EdgeDev * __restrict__ edges, const double * __restrict__ pmats_half,
                        const double * __restrict__ coltab, double * __restrict__ lookup)
{
  constexpr int S = 4, R = 4, K = 16;
  __shared__ double P[R * S * S];
  // column table transposed to [c][i/2][r] pairs: the four rate lanes of a site read four
  // consecutive 16-byte words (conflict-free), the eight site groups of a warp broadcast
  __shared__ double2 M2[K * 2 * R];
  const EdgeDev e = edges[blockIdx.x];
  stage_doubles(P, pmats_half + (size_t) blockIdx.x * R * S * S, R * S * S);
  for (int idx = threadIdx.x; idx < K * 2 * R; idx += blockDim.x)
  {
    const int c = idx / (2 * R), ih = (idx / R) & 1, rr = idx % R;
    M2[idx] = make_double2(__ldg(coltab + (rr * K + c) * S + 2 * ih), __ldg(coltab + (rr * K + c) * S + 2 * ih + 1));
  }
  __syncthreads();
  const int r = threadIdx.x & 3;
  // this lane's rate block of P(len/2) stays in registers for all its sites
  double p[S * S];
  #pragma unroll
  for (int k = 0; k < S * S; ++k) p[k] = P[r * S * S + k];
  const double w = m->weights[r];
  const double * Dn = tree.clv + e.distal * tree.clv_stride;
  const double * Xn = tree.clv + e.proximal * tree.clv_stride;
  const uint32_t * sDn = tree.scaler + (size_t) e.distal * n;
  const uint32_t * sXn = tree.scaler + (size_t) e.proximal * n;

  const int site0 = blockIdx.y * LOOKUP_DNA_SITES_PER_BLOCK;
  #pragma unroll 1
  for (int it = 0; it < LOOKUP_DNA_SITES_PER_BLOCK / 64; ++it)
  {
    const int site = site0 + it * 64 + (threadIdx.x >> 2);
    if (site0 + it * 64 >= n) break;                      // block-uniform
    const bool active = site < n;
    const int s = active ? site : n - 1;

    double dv[S], xv[S], inner[S];
    load_vec<S>(Dn + (size_t) s * (R * S) + r * S, dv);
    load_vec<S>(Xn + (size_t) s * (R * S) + r * S, xv);
    uint32_t sc = __ldg(sDn + s) + __ldg(sXn + s);
    bool small = true;
    #pragma unroll
    for (int i = 0; i < S; ++i)
    {
      const double ta = p[i * S] * dv[0] + p[i * S + 1] * dv[1] + p[i * S + 2] * dv[2] + p[i * S + 3] * dv[3];
      const double tb = p[i * S] * xv[0] + p[i * S + 1] * xv[1] + p[i * S + 2] * xv[2] + p[i * S + 3] * xv[3];
      inner[i] = ta * tb;
      small = small && (inner[i] < EPA_SCALE_THRESHOLD);
    }
    // all 16 entries of the site below the threshold? (lanes 4k..4k+3 hold one site)
    const unsigned ballot = __ballot_sync(0xffffffffu, small);
    const unsigned grp = (ballot >> ((threadIdx.x & 31) & ~3)) & 0xfu;
    if (grp == 0xfu)
    {
      sc += 1;
      #pragma unroll
      for (int i = 0; i < S; ++i) inner[i] *= EPA_SCALE_FACTOR;
    }
    const double inv = tree.inv ? __ldg(tree.inv + s) : 0.0;

    // per-rate contribution of every column, then sum over the four rate lanes (fixed order)
    double mine[4];
    #pragma unroll
    for (int c = 0; c < K; ++c)
    {
      const double2 m01 = M2[(c * 2) * R + r], m23 = M2[(c * 2 + 1) * R + r];
      double tr = inner[0] * m01.x + inner[1] * m01.y + inner[2] * m23.x + inner[3] * m23.y;
      tr *= w;
      tr += __shfl_xor_sync(0xffffffffu, tr, 1);
      tr += __shfl_xor_sync(0xffffffffu, tr, 2);
      if ((c >> 2) == r) mine[c & 3] = tr;      // lane r finishes columns 4r..4r+3
    }
    double res[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k)
      res[k] = (r == 0 && k == 0) ? 0.0 : site_loglk(mine[k], sc, inv);   // column 0 = zero column
    if (active)
      store_vec<4>(lookup + ((size_t) blockIdx.x * n_pad + site) * K + r * 4, res);
  }
}

// ---------------------------------------------------------------------------------------------
// Log-likelihood of the reference tree evaluated across one edge (inspection / tests):
// per-block partial sums, summed on the host in block order.
// ---------------------------------------------------------------------------------------------
template <int S, int R>
__global__ void __launch_bounds__(128)
edge_logl_kernel(const DevModel * __restrict__ m, DevTree tree, int n, EdgeDev e,
                 const double * __restrict__ pmat, double * __restrict__ partial)
{
  extern __shared__ double smem[];
  double * P = smem;
  __shared__ double red[128];
  stage_doubles(P, pmat, R * S * S);
  __syncthreads();
  const int site = blockIdx.x * blockDim.x + threadIdx.x;
  double lk = 0.0;
  if (site < n)
  {
    const double * A = tree.clv + e.distal * tree.clv_stride + (size_t) site * (R * S);
    const double * B = tree.clv + e.proximal * tree.clv_stride + (size_t) site * (R * S);
    double terma = 0.0;
    uint32_t kr[R], kmin = 0xffffffffu;
    if (tree.sr > 1)
    {
      #pragma unroll
      for (int r = 0; r < R; ++r)
      {
        kr[r] = tree.scaler[((size_t) e.distal * n + site) * R + r] + tree.scaler[((size_t) e.proximal * n + site) * R + r];
        kmin = min(kmin, kr[r]);
      }
    }
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      double av[S], bv[S];
      load_vec<S>(A + r * S, av);
      load_vec<S>(B + r * S, bv);
      double tr = 0.0;
      #pragma unroll
      for (int i = 0; i < S; ++i)
      {
        double tb = 0.0;
        #pragma unroll
        for (int j = 0; j < S; ++j) tb += P[(r * S + i) * S + j] * bv[j];
        tr += av[i] * m->freqs[i] * tb;
      }
      if (tree.sr > 1) tr *= rate_scale_factor(min(kr[r] - kmin, EPA_RATE_MAXDIFF));
      terma += tr * m->weights[r];
    }
    const uint32_t sc = tree.sr > 1 ? kmin
                                    : tree.scaler[(size_t) e.distal * n + site] + tree.scaler[(size_t) e.proximal * n + site];
    lk = site_loglk(terma, sc, tree.inv ? __ldg(tree.inv + site) : 0.0);
  }
  red[threadIdx.x] = lk;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1)
  {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

}  // namespace epa
