// kernels_blo_aa.cuh - HOT LOOP B for amino acids on the fp64 tensor-core path (DMMA.884) with the
// sumtable in tensor memory.
//
// Same algorithm and reference citations as kernels_blo.cuh / kernels_blo_generic.cuh (Tiny_Tree::place with
// branch-length optimisation, src/tree/Tiny_Tree.cpp:131-218, src/core/pll/optimize.cpp:60-286, libpll sumtable /
// derivatives / partials / likelihood kernels LP/core_derivatives.c:321-858, LP/core_partials.c:202-352,612-766,
// LP/core_likelihood.c:351-578, PM/optimize/opt_algorithms.c:86-262). With 20 states the contractions of a CLV pass
// are real matrix products - per site and rate category 20 x 20 for each child, 20 x 20 for the eigen rotation(s) -
// and the derivative sums are a [sites] x [R * 20] by [R * 20] x 3 product. They run as mma.sync.m8n8k4.f64:
// measured on B200 (tools/ubench/dmma_issue.cu) DMMA.884 delivers the DFMA pipe's 64 FMA/clk/SM from two warps per
// scheduler at 1/16 of the issue slots, where the DFMA formulation of kernels_blo_generic.cuh reaches 28 %.
//
// ONE CTA OF 8 WARPS PER PAIR (AA_WARPS; 12 and 16 measured no faster); a warp task = 8 consecutive window sites (the M rows of an MMA tile), all rate
// categories in turn. Fragment layouts (checked by tools/ubench/dmma_layout.cu), g = lane / 4, q = lane % 4:
//   A (8 x 4):  A[g][q]          B (4 x 8):  B[q][g]          C (8 x 8):  C[g][2q], C[g][2q + 1]
//  * child products  T[s][i] = sum_j CLV[s][j] P[i][j]: A fragments straight from the site-blocked CLV copy (five
//    k-steps of 4 states), B fragments from the transition matrix kept in shared memory IN FRAGMENT ORDER
//    [n-tile][k-step][lane] (three 8-wide n-tiles: 20 states padded to 24 with zero rows), conflict-free;
//  * the element-wise products (inner CLV, tip factor, left x right) happen in the C layout;
//  * eigen rotation  right[s][jj] = sum_i V[jj][i] in[s][i]: the C-layout registers of `in` are used AS A fragments -
//    lane q holds i = 8t + 2q + {0, 1}, which is a valid k-step (t, slot) under a permuted order of i, and the
//    constant B table (V) is stored under the same permutation: no shuffles, six k-steps instead of five;
//  * sumtable: the C-layout result of a (site group, rate) is 6 doubles per lane = 12 columns of the lane's own
//    tensor-memory row, at a column given by the site group (aa_tm_col): windows up to 320 sites;
//  * derivative sums: plain DFMA on the stored C-layout rows (each lane weighs its 6 components of every rate with
//    the decay factors e, lambda e, lambda^2 e, the quad adds up): as an MMA the [sites] x [R * 24] by [R * 24] x 3
//    product would use 3 of 8 output columns, 2.7 x the pipe time of the DFMA form. The stationary component
//    (jj = 0) is one more entry with (w_r, 0, 0).
// Longer windows, --raxml-blo and other state counts keep the kernel of kernels_blo_generic.cuh.
#pragma once
#include "kernels_blo_generic.cuh"

namespace epa {

#ifndef AA_WARPS
#define AA_WARPS 8                        /* warps per CTA (a multiple of 4: every warp keeps its 32-lane quarter) */
#endif
constexpr int AA_THREADS = AA_WARPS * 32;
constexpr int AA_SP = 24;                 // 20 states padded to three 8-wide MMA tiles
constexpr int AA_GROUPS = 40;             // site groups of 8 whose rows fit the tensor memory (aa_tm_col)
constexpr int AA_MAX_WINDOW = 8 * AA_GROUPS;                             // 320 sites
constexpr int AA_TASKS = (AA_GROUPS + AA_WARPS - 1) / AA_WARPS;          // site groups per warp, at most

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ void tc_st4(uint32_t taddr, const uint32_t (&v)[4])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

template <int R>
struct AaSmem {
  static constexpr int S = 20;
  static constexpr int PB_MAT = R * 3 * 5 * 32;        // one transition matrix in fragment order [r][t][kk][lane]
  static constexpr int PB = 0;                         // [3: distal, proximal, pendant][PB_MAT]
  static constexpr int TV = PB + 3 * PB_MAT;           // [R][MAX_CODES][AA_SP] tip vectors of the pendant matrix
  static constexpr int TL = TV + R * MAX_CODES * AA_SP;   // [MAX_CODES][AA_SP] sum_{k in mask} pi_k Vinv[k][jj]
  static constexpr int VB = TL + MAX_CODES * AA_SP;    // [3][6][32] V as B fragments under the permuted i order
  static constexpr int LB = VB + 3 * 6 * 32;           // [3][5][32] pi Vinv as B fragments
  static constexpr int DG = LB + 3 * 5 * 32;           // [R][3: e, lambda e, lambda^2 e][AA_SP] decay factors (per evaluation)
  static constexpr int FR = DG + R * 3 * AA_SP;        // [AA_SP] stationary frequencies, padded
  static constexpr int EX = FR + AA_SP;                // [R][S] expm1 / exp scratch
  static constexpr int VA = EX + R * S;                // [3][5][32] Vinv as A fragments (matrix build)
  static constexpr int VN = VA + 3 * 5 * 32;           // [3][5][32] V as B fragments in natural order (matrix build)
  static constexpr int RED = VN + 3 * 5 * 32;          // [2][AA_WARPS] block-reduction scratch
  static constexpr int TOTAL = RED + 2 * AA_WARPS;
};

// deterministic block-wide sum of two values (fixed order over the warps)
__device__ __forceinline__ void aa_block_sum2(double & a, double & b, double * red)
{
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5;
  __syncthreads();                                     // previous readers of `red` are done
  if ((threadIdx.x & 31) == 0) { red[warp] = a; red[AA_WARPS + warp] = b; }
  __syncthreads();
  double sa = 0.0, sb = 0.0;
  #pragma unroll
  for (int w = 0; w < AA_WARPS; ++w) { sa += red[w]; sb += red[AA_WARPS + w]; }
  a = sa; b = sb;
}

// element (i, j) of rate r inside one fragment-ordered matrix: n-tile i / 8, k-step j / 4, lane (i % 8) * 4 + j % 4
__device__ __forceinline__ int aa_pb_index(int r, int i, int j)
{
  return ((r * 3 + (i >> 3)) * 5 + (j >> 2)) * 32 + (i & 7) * 4 + (j & 3);
}

// state index that lane q's C-layout slot (t, slot) holds = k-step kk' = 2 t + slot of the permuted order
__device__ __forceinline__ int aa_perm(int kk, int q) { return 8 * (kk >> 1) + 2 * q + (kk & 1); }

// P[r][i][j] = delta_ij + sum_k Vinv[i][k] expm1(lambda_k rate_r t) V[k][j]   (LP/core_pmatrix.c:185-249), written in
// fragment order. One 8 x 8 tile of one rate per warp task: five MMA k-steps with A = Vinv[i][k] expm1_k (the constant
// fragment scaled by this lane's k) and B = V[k][j].
template <int R>
__device__ __forceinline__ void aa_pmatrix(double * sm, double t, int which)
{
  using L = AaSmem<R>;
  constexpr int S = 20;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  __syncthreads();                                     // readers of EX and of the matrix being replaced are done
  for (int idx = threadIdx.x; idx < R * S; idx += AA_THREADS)
    sm[L::EX + idx] = expm1(c_model.eigenvals[idx % S] * c_model.rates[idx / S] * t);
  __syncthreads();
  double * P = sm + L::PB + which * L::PB_MAT;
  #pragma unroll 1
  for (int task = warp; task < R * 9; task += AA_WARPS)
  {
    const int r = task / 9, ti = (task % 9) / 3, tj = task % 3;
    double c[2] = {0.0, 0.0};
    #pragma unroll
    for (int kk = 0; kk < 5; ++kk)
      dmma884(c, sm[L::VA + (ti * 5 + kk) * 32 + lane] * sm[L::EX + r * S + 4 * kk + q], sm[L::VN + (tj * 5 + kk) * 32 + lane]);
    const int i = 8 * ti + g, j = 8 * tj + 2 * q;
    if (i < S && j < S)
    {
      P[aa_pb_index(r, i, j)] = c[0] + (i == j ? 1.0 : 0.0);
      P[aa_pb_index(r, i, j + 1)] = c[1] + (i == j + 1 ? 1.0 : 0.0);
    }
  }
  __syncthreads();
}

// tv[r][code][i] = sum_{j in mask(code)} Ppendant[r][i][j]
template <int R>
__device__ __forceinline__ void aa_tipvec(double * sm)
{
  using L = AaSmem<R>;
  constexpr int S = 20;
  const double * P = sm + L::PB + 2 * L::PB_MAT;
  const int ncodes = c_model.ncodes;
  for (int idx = threadIdx.x; idx < R * ncodes * S; idx += AA_THREADS)
  {
    const int r = idx / (ncodes * S), c = (idx / S) % ncodes, i = idx % S;
    uint32_t mask = c_model.code2mask[c];
    double acc = 0.0;
    while (mask)                                       // ascending states, as the full loop would add them
    {
      const int j = __ffs((int) mask) - 1;
      acc += P[aa_pb_index(r, i, j)];
      mask &= mask - 1u;
    }
    sm[L::TV + (r * MAX_CODES + c) * AA_SP + i] = acc;
  }
  __syncthreads();
}

// Tensor-memory columns of (site group, rate): group grp belongs to the 32-lane quarter grp % 4 - the quarter of every
// warp that can be given the group (warp % 4 == grp % 4) - and takes 12 R columns at slot grp / 4: 40 groups fill 480 of
// the 512 columns of each quarter.
__device__ __forceinline__ uint32_t aa_tm_col(int grp, int r, int R) { return (uint32_t) ((grp >> 2) * R + r) * 12u; }

// the 6 doubles of one (task, rate) row of this lane <-> 12 tensor-memory columns
__device__ __forceinline__ void aa_tm_store(uint32_t taddr, const double (&st)[3][2])
{
  uint32_t a[8], b[4];
  #pragma unroll
  for (int k = 0; k < 4; ++k) { a[2 * k] = (uint32_t) __double2loint(st[k >> 1][k & 1]); a[2 * k + 1] = (uint32_t) __double2hiint(st[k >> 1][k & 1]); }
  #pragma unroll
  for (int k = 0; k < 2; ++k) { b[2 * k] = (uint32_t) __double2loint(st[2][k]); b[2 * k + 1] = (uint32_t) __double2hiint(st[2][k]); }
  tc_st8(taddr, a);
  tc_st4(taddr + 8, b);
}

// Pass A: inner CLV toward the new tip from (D, X); returns the window log-likelihood (new tip | inner) and leaves
// the pendant sumtable in tensor memory. Pass B (DISTAL): inner CLV toward the distal node from (tip, X); leaves
// the distal sumtable (D vs inner).
template <int R, bool DISTAL>
__device__ __forceinline__ double aa_pass(double * sm, uint32_t tm, const double * __restrict__ DT, const double * __restrict__ XT,
                                          const uint32_t * __restrict__ sD, const uint32_t * __restrict__ sX,
                                          const uint8_t * __restrict__ qc, int begin, int w, const double * __restrict__ inv_w)
{
  using L = AaSmem<R>;
  constexpr int S = 20;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const int n_groups = (w + 7) >> 3;
  const double * PBd = sm + L::PB, * PBx = sm + L::PB + L::PB_MAT;
  double acc = 0.0;
  // A fragments of (site group, rate) come from L2 (each element is read once per pass): the loads of the NEXT
  // (group, rate) are issued before the arithmetic of the current one
  auto frag_offset = [&](int grp) -> size_t
  {
    const int s0 = grp * 8 + g;
    return clvt_off_c<S>(begin + (s0 < w ? s0 : w - 1), R * S);
  };
  double nD[5], nX[5];
  if (warp < n_groups)
  {
    const size_t off0 = frag_offset(warp);
    #pragma unroll
    for (int kk = 0; kk < 5; ++kk)
    {
      const size_t o = off0 + (size_t) (4 * kk + q) * 32;
      nX[kk] = __ldg(XT + o);
      nD[kk] = __ldg(DT + o);
    }
  }
  #pragma unroll 1
  for (int grp = warp; grp < n_groups; grp += AA_WARPS)
  {
    const int s0 = grp * 8 + g;
    const bool act = s0 < w;
    const int s = act ? s0 : w - 1;                           // rows beyond the window recompute the last site
    const size_t off = frag_offset(grp);
    const int code = qc[s] & (MAX_CODES - 1);
    double term = 0.0;
    #pragma unroll 1
    for (int r = 0; r < R; ++r)
    {
      double aD[5], aX[5];
      #pragma unroll
      for (int kk = 0; kk < 5; ++kk) { aD[kk] = nD[kk]; aX[kk] = nX[kk]; }
      {
        const bool more_r = r + 1 < R;
        const int ngrp = more_r ? grp : grp + AA_WARPS;
        if (ngrp < n_groups)
        {
          const size_t noff = (more_r ? off : frag_offset(ngrp)) + (size_t) ((more_r ? r + 1 : 0) * S) * 32;
          #pragma unroll
          for (int kk = 0; kk < 5; ++kk)
          {
            const size_t o = noff + (size_t) (4 * kk + q) * 32;
            nX[kk] = __ldg(XT + o);
            nD[kk] = __ldg(DT + o);
          }
        }
      }
      double in[3][2];
      const double2 * tv2 = reinterpret_cast<const double2 *>(sm + L::TV + (r * MAX_CODES + code) * AA_SP + 2 * q);
      if constexpr (!DISTAL)
      {
        double ta[3][2] = {}, tb[3][2] = {};
        #pragma unroll
        for (int t = 0; t < 3; ++t)
          #pragma unroll
          for (int kk = 0; kk < 5; ++kk)
          {
            dmma884(ta[t], aD[kk], PBd[((r * 3 + t) * 5 + kk) * 32 + lane]);
            dmma884(tb[t], aX[kk], PBx[((r * 3 + t) * 5 + kk) * 32 + lane]);
          }
        const double2 * fr2 = reinterpret_cast<const double2 *>(sm + L::FR + 2 * q);
        double part = 0.0;
        #pragma unroll
        for (int t = 0; t < 3; ++t)
        {
          in[t][0] = ta[t][0] * tb[t][0]; in[t][1] = ta[t][1] * tb[t][1];
          const double2 tv = tv2[4 * t], fr = fr2[4 * t];
          part += (in[t][0] * fr.x) * tv.x;
          part += (in[t][1] * fr.y) * tv.y;
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        term += part * c_model.weights[r];
      }
      else
      {
        double tb[3][2] = {};
        #pragma unroll
        for (int t = 0; t < 3; ++t)
          #pragma unroll
          for (int kk = 0; kk < 5; ++kk) dmma884(tb[t], aX[kk], PBx[((r * 3 + t) * 5 + kk) * 32 + lane]);
        #pragma unroll
        for (int t = 0; t < 3; ++t)
        {
          const double2 tv = tv2[4 * t];
          in[t][0] = tv.x * tb[t][0]; in[t][1] = tv.y * tb[t][1];
        }
      }
      // eigen rotation of the inner CLV: C-layout registers as A fragments of the permuted order
      double right[3][2] = {};
      #pragma unroll
      for (int t = 0; t < 3; ++t)
        #pragma unroll
        for (int kk = 0; kk < 6; ++kk) dmma884(right[t], in[kk >> 1][kk & 1], sm[L::VB + (t * 6 + kk) * 32 + lane]);
      double st[3][2];
      if constexpr (!DISTAL)
      {
        const double2 * tl2 = reinterpret_cast<const double2 *>(sm + L::TL + code * AA_SP + 2 * q);
        #pragma unroll
        for (int t = 0; t < 3; ++t) { const double2 tl = tl2[4 * t]; st[t][0] = tl.x * right[t][0]; st[t][1] = tl.y * right[t][1]; }
      }
      else
      {
        double left[3][2] = {};
        #pragma unroll
        for (int t = 0; t < 3; ++t)
          #pragma unroll
          for (int kk = 0; kk < 5; ++kk) dmma884(left[t], aD[kk], sm[L::LB + (t * 5 + kk) * 32 + lane]);
        #pragma unroll
        for (int t = 0; t < 3; ++t) { st[t][0] = left[t][0] * right[t][0]; st[t][1] = left[t][1] * right[t][1]; }
      }
      aa_tm_store(tm + aa_tm_col(grp, r, R), st);
    }
    if constexpr (!DISTAL)
    {
      if (act && q == 0)
      {
        const double inv = inv_w ? __ldg(inv_w + s) : 0.0;
        const uint32_t scal = __ldg(sD + s) + __ldg(sX + s);
        acc += site_loglk(term, scal, inv);
      }
    }
  }
  tc_wait_st();
  if constexpr (!DISTAL)
  {
    double unused = 0.0;
    aa_block_sum2(acc, unused, sm + AaSmem<R>::RED);
  }
  return acc;
}

// Derivative sums over the window (LP/core_derivatives.c:643-858). The rows stay in the C layout they were stored
// in: lane q of a site's quad holds 6 of the 24 (padded) eigen-components of every rate. Each lane multiplies its
// components with their decay factors (DFMA: as an MMA the product would use 3 of 8 output columns), the quad adds
// up, lane q = 0 of the quad finishes the site. Rates outer (18 table values per rate in registers), tasks inner.
template <int R>
__device__ __forceinline__ void aa_derivatives(double * sm, uint32_t tm, int w, double t, const double * __restrict__ inv_w,
                                               double & f, double & df)
{
  using L = AaSmem<R>;
  constexpr int S = 20;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  __syncthreads();
  // DG[r][m][jj] for the padded 24 components: e, lambda e, lambda^2 e with e = exp(lambda_jj rate_r t) w_r; the
  // stationary component carries (w_r, 0, 0), the pads are zero
  for (int k = threadIdx.x; k < R * AA_SP; k += AA_THREADS)
  {
    const int r = k / AA_SP, jj = k % AA_SP;
    double e = 0.0, lk = 0.0;
    if (jj < S)
    {
      lk = jj == 0 ? 0.0 : c_model.eigenvals[jj] * c_model.rates[r];
      e = (jj == 0 ? 1.0 : exp(lk * t)) * c_model.weights[r];
    }
    sm[L::DG + (r * 3 + 0) * AA_SP + jj] = e;
    sm[L::DG + (r * 3 + 1) * AA_SP + jj] = lk * e;
    sm[L::DG + (r * 3 + 2) * AA_SP + jj] = lk * lk * e;
  }
  __syncthreads();
  const int n_groups = (w + 7) >> 3;
  const int n_tasks = warp < n_groups ? (n_groups - warp + AA_THREADS / 32 - 1) / (AA_THREADS / 32) : 0;      // warp-uniform
  double c0[AA_TASKS], c1[AA_TASKS], c2[AA_TASKS];
  #pragma unroll
  for (int k = 0; k < AA_TASKS; ++k) { c0[k] = 0.0; c1[k] = 0.0; c2[k] = 0.0; }
  #pragma unroll 1
  for (int r = 0; r < R; ++r)
  {
    // this lane's components: state 8 t + 2 q + slot for (t, slot) = (kk / 2, kk % 2)
    double d0[6], d1[6], d2[6];
    #pragma unroll
    for (int t3 = 0; t3 < 3; ++t3)
    {
      const double2 x0 = *reinterpret_cast<const double2 *>(sm + L::DG + (r * 3 + 0) * AA_SP + 8 * t3 + 2 * q);
      const double2 x1 = *reinterpret_cast<const double2 *>(sm + L::DG + (r * 3 + 1) * AA_SP + 8 * t3 + 2 * q);
      const double2 x2 = *reinterpret_cast<const double2 *>(sm + L::DG + (r * 3 + 2) * AA_SP + 8 * t3 + 2 * q);
      d0[2 * t3] = x0.x; d0[2 * t3 + 1] = x0.y; d1[2 * t3] = x1.x; d1[2 * t3 + 1] = x1.y; d2[2 * t3] = x2.x; d2[2 * t3 + 1] = x2.y;
    }
    // all rows of this rate first, one wait, then the arithmetic
    uint32_t a[AA_TASKS][8], b[AA_TASKS][4];
    #pragma unroll
    for (int task = 0; task < AA_TASKS; ++task)
      if (task < n_tasks)
      {
        const uint32_t ta = tm + aa_tm_col(warp + task * (AA_THREADS / 32), r, R);
        tc_ld8(ta, a[task]);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b[task][0]), "=r"(b[task][1]), "=r"(b[task][2]), "=r"(b[task][3]) : "r"(ta + 8u) : "memory");
      }
    tc_wait_ld();
    #pragma unroll
    for (int task = 0; task < AA_TASKS; ++task)
      if (task < n_tasks)
      {
        #pragma unroll
        for (int kk = 0; kk < 6; ++kk)
        {
          const double x = kk < 4 ? __hiloint2double((int) a[task][2 * kk + 1], (int) a[task][2 * kk])
                                  : __hiloint2double((int) b[task][2 * (kk - 4) + 1], (int) b[task][2 * (kk - 4)]);
          c0[task] += x * d0[kk]; c1[task] += x * d1[kk]; c2[task] += x * d2[kk];
        }
      }
  }
  double a1 = 0.0, a2 = 0.0;
  #pragma unroll
  for (int task = 0; task < AA_TASKS; ++task)
  {
    if (task < n_tasks)
    {
      double x0 = c0[task], x1 = c1[task], x2 = c2[task];
      x0 += __shfl_xor_sync(0xffffffffu, x0, 1); x1 += __shfl_xor_sync(0xffffffffu, x1, 1); x2 += __shfl_xor_sync(0xffffffffu, x2, 1);
      x0 += __shfl_xor_sync(0xffffffffu, x0, 2); x1 += __shfl_xor_sync(0xffffffffu, x1, 2); x2 += __shfl_xor_sync(0xffffffffu, x2, 2);
      const int s = (warp + task * (AA_THREADS / 32)) * 8 + g;
      if (q == 0 && s < w)
      {
        const double l0 = x0 + (inv_w ? __ldg(inv_w + s) : 0.0);
        const double inv = 1.0 / l0;
        const double g1 = -x1 * inv;
        a1 += g1;
        a2 += g1 * g1 - x2 * inv;
      }
    }
  }
  aa_block_sum2(a1, a2, sm + L::RED);
  f = a1; df = a2;
}

// bounded Newton-Raphson, PM/optimize/opt_algorithms.c:133-262; returns 0.0 on failure
template <int R>
__device__ __forceinline__ double aa_newton(double * sm, uint32_t tm, int w, const double * __restrict__ inv_w,
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
    aa_derivatives<R>(sm, tm, w, x, inv_w, f, df);
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

template <int R>
__global__ void __launch_bounds__(AA_THREADS, 1)
blo_aa_kernel(BloArgs a, const double * __restrict__ clvT, size_t t_stride)
{
  using L = AaSmem<R>;
  constexpr int S = 20;
  extern __shared__ __align__(16) double sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t tmem_slot;
  const int ncodes = c_model.ncodes;
  // constant tables: V / Vinv copies, padded frequencies, tip factors, B fragments of V and pi Vinv; zeroed matrix pads
  for (int i = threadIdx.x; i < L::TOTAL; i += AA_THREADS) sm[i] = 0.0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 3 * 5 * 32; idx += AA_THREADS)
  {
    const int t = idx / (5 * 32), kk = (idx / 32) % 5, ln = idx & 31;
    const int row = 8 * t + (ln >> 2), k = 4 * kk + (ln & 3);
    sm[L::VA + idx] = row < S ? c_model.inv_eigenvecs[row * S + k] : 0.0;        // A[i][k] = Vinv[i][k]
    sm[L::VN + idx] = row < S ? c_model.eigenvecs[k * S + row] : 0.0;            // B[k][j] = V[k][j], j = row
  }
  for (int i = threadIdx.x; i < S; i += AA_THREADS) sm[L::FR + i] = c_model.freqs[i];
  for (int i = threadIdx.x; i < ncodes * S; i += AA_THREADS)
  {
    const int c = i / S, j = i % S;
    const uint32_t mask = c_model.code2mask[c];
    double acc = 0.0;
    for (int k = 0; k < S; ++k)
      if ((mask >> k) & 1u) acc += c_model.pivinv[k * S + j];
    sm[L::TL + c * AA_SP + j] = acc;
  }
  for (int idx = threadIdx.x; idx < 3 * 6 * 32; idx += AA_THREADS)
  {
    const int t = idx / (6 * 32), kk = (idx / 32) % 6, ln = idx & 31;
    const int jj = 8 * t + (ln >> 2), i = aa_perm(kk, ln & 3);
    sm[L::VB + idx] = (jj < S && i < S) ? c_model.eigenvecs[jj * S + i] : 0.0;
  }
  for (int idx = threadIdx.x; idx < 3 * 5 * 32; idx += AA_THREADS)
  {
    const int t = idx / (5 * 32), kk = (idx / 32) % 5, ln = idx & 31;
    const int jj = 8 * t + (ln >> 2), k = 4 * kk + (ln & 3);
    sm[L::LB + idx] = jj < S ? c_model.pivinv[k * S + jj] : 0.0;
  }
  if (threadIdx.x < 32)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // this thread's tensor-memory row: its lane of the warp's quarter (columns by site group, aa_tm_col)
  const uint32_t tm = tmem_slot + ((uint32_t) (((threadIdx.x >> 5) & 3) * 32) << 16);

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
    if (w <= 0 || w > AA_MAX_WINDOW)
    {
      if (threadIdx.x == 0) a.out[pid] = BloResult{NAN, NAN, NAN};
      continue;
    }
    const int n = a.n;
    const double * DT = clvT + (size_t) ed.distal * t_stride;
    const double * XT = clvT + (size_t) ed.proximal * t_stride;
    const uint32_t * sD = a.tree.scaler + (size_t) ed.distal * n + begin;
    const uint32_t * sX = a.tree.scaler + (size_t) ed.proximal * n + begin;
    const uint8_t * qc = a.codes + (size_t) q * n + begin;
    const double * inv_w = a.tree.inv ? a.tree.inv + begin : nullptr;      // +I: pll_util.cpp:413-414

    // optimize_branch_triplet / opt_branch_lengths_pplacer as half rounds (see kernels_blo.cuh)
    const double orig = ed.length;
    double len[3] = {orig / 2.0, orig / 2.0, EPA_DEFAULT_PENDANT};
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
          aa_pmatrix<R>(sm, len[mi], mi);
          if (mi == 2) aa_tipvec<R>(sm);
        }
      rebuild = 0u;
      double xmin, xmax, xguess;
      if (!distal_phase)
      {
        const double new_logl = -aa_pass<R, false>(sm, tm, DT, XT, sD, sX, qc, begin, w, inv_w);
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
        (void) aa_pass<R, true>(sm, tm, DT, XT, sD, sX, qc, begin, w, inv_w);
        xmin = fmin(EPA_MIN_BRLEN / 2.0, original_length / 2.0);
        xmax = original_length - xmin / 10.0;
        xguess = len[0];
        if (xguess < xmin || xguess > xmax) xguess = original_length / 2.0;
      }
      const double xres = aa_newton<R>(sm, tm, w, inv_w, xmin, xguess, xmax, xmin / 10.0);
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
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_slot) : "memory");
}

template <int R>
inline cudaError_t launch_blo_aa(int sm_count, BloArgs & a, cudaStream_t stream, const double * clvT, size_t t_stride)
{
  const size_t smem = (size_t) AaSmem<R>::TOTAL * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(blo_aa_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned) std::min<uint64_t>((uint64_t) sm_count, a.n_pairs);
  blo_aa_kernel<R><<<grid, AA_THREADS, smem, stream>>>(a, clvT, t_stride);
  return cudaGetLastError();
}

}  // namespace epa
