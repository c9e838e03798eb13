// kernels_blo_site.cuh - HOT LOOP B, lane = site formulation of the DNA branch-length optimisation.
//
// Same reference behaviour as kernels_blo.cuh (Tiny_Tree::place src/tree/Tiny_Tree.cpp:131-218,
// opt_branch_lengths_pplacer src/core/pll/optimize.cpp:60-248, Newton-Raphson
// PM/optimize/opt_algorithms.c:86-262, sumtable/derivatives LP/core_derivatives.c:321-858, CLV update
// LP/core_partials.c:202-352,612-766, edge log-likelihood LP/core_likelihood.c:351-578), one warp per
// (query, edge) pair, but EVERY phase maps one lane to one site:
//   * the CLV passes read a site-blocked copy of the reference CLVs, clvT[node][site/32][r*4+k][32],
//     so that the 32 lanes of a warp (32 consecutive sites, any alignment) read 256 contiguous bytes
//     per component: coalesced, no lane owns less than a whole site, hence no cross-lane rate sums,
//     no ballots for the scaling test, one logarithm per site;
//   * the sumtable row of a site is written and read by the same lane (shared memory, odd row
//     length = conflict-free), so passes and Newton iterations need no warp synchronisation;
//   * transition matrices are read as warp-uniform 128-bit shared-memory broadcasts; the per-mask
//     tip vectors use a skewed row (4R + 2 doubles) that keeps the A/C/G/T/N rows in disjoint banks;
//   * the 3R decay factors of a Newton evaluation are computed by 3R lanes and broadcast through
//     shared memory (no shuffles).
#pragma once
#include "kernels_blo.cuh"
#include "kernels_preplace_mma.cuh"     // tcgen05.ld helpers

namespace epa {

constexpr int CLVT_BLOCK = 32;      // sites per block of the site-blocked CLV copy

__host__ __device__ inline size_t clvt_node_stride(int n, int R)
{
  return (size_t) ((n + CLVT_BLOCK - 1) / CLVT_BLOCK) * (size_t) (R * 4) * CLVT_BLOCK;
}

// clv[node][site][c] -> clvT[node][site/32][c][site%32]; pad sites read as zero. C = R * 4.
// grid = (site blocks, nodes), block = 256 threads, dynamic smem = 32 * (C + 1) doubles.
__global__ void __launch_bounds__(256)
clv_site_block_kernel(const double * __restrict__ clv, size_t clv_stride, int n, int C,
                      double * __restrict__ clvT, size_t t_stride)
{
  extern __shared__ double tile[];                 // [32][C + 1]
  const int blk = blockIdx.x;
  const size_t node = blockIdx.y;
  const double * src = clv + node * clv_stride + (size_t) blk * CLVT_BLOCK * C;
  double * dst = clvT + node * t_stride + (size_t) blk * CLVT_BLOCK * C;
  const int total = CLVT_BLOCK * C;
  const int valid = min(CLVT_BLOCK, n - blk * CLVT_BLOCK) * C;
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    tile[(i / C) * (C + 1) + (i % C)] = i < valid ? src[i] : 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    dst[i] = tile[(i % CLVT_BLOCK) * (C + 1) + (i / CLVT_BLOCK)];
}

template <int R>
struct SiteWarpSmem {
  static constexpr int C = 4 * R;
  static constexpr int P_D = 0;                     // [r][i][j]
  static constexpr int P_P = 16 * R;
  static constexpr int P_E = 32 * R;
  static constexpr int TVS = C + 2;                 // tip-vector row: [r][i] + 2 doubles of skew
  static constexpr int TV = 48 * R;                 // [16 mask positions][TVS]
  static constexpr int EX = TV + 16 * TVS;          // [3][3R] decay tables / [4R] expm1 scratch
  static constexpr int EXN = (9 * R > 4 * R ? 9 * R : 4 * R);
  static constexpr int SUM = (EX + EXN + 1) & ~1;   // [w][blo_row(R)]
  __host__ __device__ static constexpr size_t doubles(int wcap)
  {
    return (((size_t) SUM + (size_t) wcap * blo_row(R)) + 1) & ~(size_t) 1;
  }
};

struct SiteCtaSmem {
  double V[16], Vinv[16];
  __align__(16) double tipleft[64];     // [tv_pos(mask)][j] = sum_{k in mask} pi_k Vinv[k][j]
  __align__(16) double tippi[64];       // [tv_pos(mask)][k] = pi_k if k in mask else 0
  unsigned long long q_next, q_end;
  int q_lock;
};

// Where a warp keeps its sumtable: shared-memory rows, global planes (GS kernels), or - for the
// first SITE_TMEM_WARPS warps of a CTA - tensor memory. TMEM is idle in this kernel, every thread
// of a warp can reach its own TMEM lane with tcgen05.ld/st (32x32b shape), and a sumtable row is
// only ever touched by the lane that owns the site: 32 columns (128 bytes) per row and lane,
// 8 rows per warp, two warps per 32-lane quarter. This lifts the shared-memory limit on the
// number of pairs in flight per SM.
constexpr int SITE_TMEM_WARPS = 8;
#ifndef SITE_MAX_WARPS
#define SITE_MAX_WARPS 12          /* warps per CTA: 12 x 32 threads x 168 registers fill the register file */
#endif
constexpr int SITE_TMEM_ROWS = 8;        // rows (trips of 32 sites) of a TMEM-backed warp when 8 warps share the 512 columns:
                                         // windows <= 256; windows <= 512 give 16 rows to 4 warps
struct SumRef {
  double * base;        // global planes (GS)
  int gstride;
  uint32_t taddr;       // tensor-memory address of the warp's first row
  int tm;
  uint32_t saddr;       // shared-memory address of this LANE's row of trip 0
};

// Shared-memory sumtable rows: 1 + 3R doubles padded to an even count whose half is odd, so that the
// lane = site accesses are 128-bit and conflict-free (a quarter-warp's eight rows start in eight
// different 16-byte bank groups). Every storage kind keeps a row for all 32 lanes of every trip: a
// lane beyond the window stores (1, 0, ..., 0), whose derivative terms are exact zeros, so the
// Newton sweeps need no validity predicates.
__host__ __device__ constexpr int site_row_pad(int NK)        // NK = decaying components per site
{
  int p = (1 + NK + 1) & ~1;
  if (((p / 2) & 1) == 0) p += 2;
  return p;
}
__host__ __device__ constexpr uint32_t site_trip_bytes(int NK) { return 32u * (uint32_t) site_row_pad(NK) * 8u; }

__device__ __forceinline__ void lds2(uint32_t addr, double & a, double & b)
{
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts2(uint32_t addr, double a, double b)
{
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(a), "d"(b) : "memory");
}

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                  "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                  "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct BloSiteArgs {
  BloArgs b;
  int bugcompat;                        // per-rate scalers: reproduce the reference's window offset
  const double * clvT;                  // site-blocked CLV copy
  size_t t_stride;                      // doubles per node in clvT
  double * gscratch;                    // global sumtable scratch [warp][blo_row][wpad] (GS variant)
  int wpad;                             // padded window capacity of the global scratch
  int n_tmem_warps;                     // leading warps of a CTA that keep their sumtable in tensor memory
  int tmem_cols;                        // TMEM columns per such warp: 256 (8 rows, 8 warps) or 512 (16 rows, 4 warps)
  // first-round tables (NULL = not available): every pair of an edge starts from the same three
  // lengths, so the inner CLV of the first pass, rotated into the eigenbasis, is per-edge data
  const double * gT;                    // [edge] site-blocked V * inner(orig/2, orig/2), scaled like the CLV update
  size_t g_stride;                      // doubles per edge in gT
  const double * lookup;                // preplacement tables [edge][n_pad][16]: the first-pass site log-likelihoods
  int n_pad;
};

// gT[e][s/32][r*4+j][s%32] = sum_k V[j][k] * inner[s][r][k], inner = (P(orig/2) D) * (P(orig/2) X), multiplied by
// 2^256 where the CLV update would rescale (LP/core_partials.c:690-766). grid = (edges, ceil(n / 128)).
template <int R>
__global__ void __launch_bounds__(128)
blo_first_table_kernel(DevTree tree, int n, const EdgeDev * __restrict__ edges, const double * __restrict__ pm,
                       double * __restrict__ gT, size_t g_stride)
{
  __shared__ double P[R * 16];
  const uint32_t e = blockIdx.x;
  for (int i = threadIdx.x; i < R * 16; i += blockDim.x) P[i] = pm[(size_t) e * R * 16 + i];
  __syncthreads();
  const int s = blockIdx.y * 128 + threadIdx.x;
  if (s >= n) return;
  const EdgeDev ed = edges[e];
  const double * D = tree.clv + ed.distal * tree.clv_stride + (size_t) s * R * 4;
  const double * X = tree.clv + ed.proximal * tree.clv_stride + (size_t) s * R * 4;
  double in[4 * R];
  bool small = true;
  #pragma unroll
  for (int r = 0; r < R; ++r)
  {
    double dv[4], xv[4];
    load_vec<4>(D + r * 4, dv);
    load_vec<4>(X + r * 4, xv);
    #pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const double * p = P + r * 16 + i * 4;
      const double ta = p[0] * dv[0] + p[1] * dv[1] + p[2] * dv[2] + p[3] * dv[3];
      const double tb = p[0] * xv[0] + p[1] * xv[1] + p[2] * xv[2] + p[3] * xv[3];
      in[r * 4 + i] = ta * tb;
      small = small && (in[r * 4 + i] < EPA_SCALE_THRESHOLD);
    }
  }
  if (small)
  {
    #pragma unroll
    for (int c = 0; c < 4 * R; ++c) in[c] *= EPA_SCALE_FACTOR;
  }
  double * out = gT + (size_t) e * g_stride + (size_t) (s >> 5) * (R * 4 * CLVT_BLOCK) + (s & 31);
  #pragma unroll
  for (int r = 0; r < R; ++r)
    #pragma unroll
    for (int j = 0; j < 4; ++j)
      out[(size_t) (r * 4 + j) * CLVT_BLOCK] = c_model.eigenvecs[j * 4] * in[r * 4] + c_model.eigenvecs[j * 4 + 1] * in[r * 4 + 1]
                                             + c_model.eigenvecs[j * 4 + 2] * in[r * 4 + 2] + c_model.eigenvecs[j * 4 + 3] * in[r * 4 + 3];
}

// P[r][i][j] = delta_ij + sum_k Vinv[i][k] expm1(lambda_k rate_r t) V[k][j]   (LP/core_pmatrix.c:185-249)
template <int R>
__device__ __forceinline__ void site_pmatrix(const SiteCtaSmem & cs, double t, double * P, double * ex, int lane)
{
  __syncwarp();
  for (int idx = lane; idx < R * 4; idx += 32)
    ex[idx] = expm1(c_model.eigenvals[idx & 3] * c_model.rates[idx >> 2] * t);
  __syncwarp();
  for (int idx = lane; idx < R * 16; idx += 32)
  {
    const int r = idx >> 4, i = (idx >> 2) & 3, j = idx & 3;
    double acc = (i == j) ? 1.0 : 0.0;
    #pragma unroll
    for (int k = 0; k < 4; ++k) acc += (cs.Vinv[i * 4 + k] * ex[r * 4 + k]) * cs.V[k * 4 + j];
    P[idx] = acc;
  }
  __syncwarp();
}

// tv[pos(mask)][r*4+i] = sum_{j in mask} P[r][i][j]
template <int R>
__device__ __forceinline__ void site_tipvec(const double * P, double * tv, int lane)
{
  for (int idx = lane; idx < R * 64; idx += 32)
  {
    const int mask = idx / (4 * R), ri = idx % (4 * R);
    double acc = 0.0;
    #pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((mask >> j) & 1) acc += P[ri * 4 + j];
    tv[tv_pos(mask) * SiteWarpSmem<R>::TVS + ri] = acc;
  }
  __syncwarp();
}

template <int N>
__device__ __forceinline__ void lds_vec(const double * p, double (&v)[N])
{
  const double2 * p2 = reinterpret_cast<const double2 *>(p);
  #pragma unroll
  for (int i = 0; i < N / 2; ++i) { const double2 t = p2[i]; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
}

// element offset of (window-relative site s, component 0) inside a node of the site-blocked copy
template <int R>
__device__ __forceinline__ size_t clvt_offset(int abs_site)
{
  return (size_t) (abs_site >> 5) * (size_t) (R * 4 * CLVT_BLOCK) + (size_t) (abs_site & 31);
}

// ---------------------------------------------------------------------------------------------
// DNA lookup build, lane = site over the site-blocked CLV copy (rows 2-3 of the scope table:
// Tiny_Tree ctor src/tree/Tiny_Tree.cpp:84-128 + precompute_sites_static :18-46):
//   inner[r][i] = (P(len/2) D)[r][i] * (P(len/2) X)[r][i]          (+ per-site scaling)
//   lookup[e][s][c] = log( sum_r w_r sum_i inner[r][i] M[r][c][i] ) + scaler * log(2^-256), column 0 = zero
// A warp reads 256 contiguous bytes per CLV component, keeps the whole site in registers (no
// shuffles), reads the transition matrix and the column table as warp-uniform 128-bit
// shared-memory broadcasts, and every lane writes its site's 128-byte table row.
// grid = (n_edges, ceil(n / 128)), block = 128.
// ---------------------------------------------------------------------------------------------
// column table of the lookup build in constant memory, [c][r][i] (copied device-to-device before the
// launch): the 240 table reads of a site become constant-bank operands of the FMAs
__constant__ double c_coltab[16 * 4 * MAX_RATES / 2];      // R <= 4

// Natural logarithm of a positive normal double through a 128-entry table: x = 2^e m, m in [1, 2),
// c = tab[top 7 mantissa bits].x ~ 1 / m, r = m c - 1 exactly rounded once (|r| <= 2^-8),
// log x = e ln 2 - log c + log1p(r) with a degree-7 Taylor polynomial (truncation 2^-59 relative).
// Absolute error ~1e-16 (|log x| + 1): the table sums of the preplacement are insensitive to it.
// The CUDA log() is ~90 instructions with its special-case handling, and 15 of them per site made
// the lookup build issue bound; zero, subnormal, negative and non-finite arguments still take it.
__device__ __forceinline__ bool table_log_ok(double x)      // positive normal
{
  return (unsigned) (__double2hiint(x) - 0x00100000) < 0x7fe00000u;
}
__device__ __forceinline__ double table_log_fast(double x, const double2 * __restrict__ tab);
__device__ __forceinline__ double table_log(double x, const double2 * __restrict__ tab)
{
  if (!table_log_ok(x)) return log(x);
  return table_log_fast(x, tab);
}
__device__ __forceinline__ double table_log_fast(double x, const double2 * __restrict__ tab)
{
  const int hi = __double2hiint(x);
  const double2 t = tab[(hi >> 13) & 127];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, t.x, -1.0);
  double p = 1.0 / 7.0;
  p = fma(p, r, -1.0 / 6.0);
  p = fma(p, r, 1.0 / 5.0);
  p = fma(p, r, -1.0 / 4.0);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double l1p = fma(p * r, r, r);
  const double e = (double) ((hi >> 20) - 1023);
  // ln 2 split so that e * hi part is exact for |e| < 2^11
  return fma(e, 0x1.62e42fefa38p-1, t.y) + fma(e, 0x1.ef35793c7673p-45, l1p);
}

// site_loglk (common.cuh) with the table logarithm
__device__ __forceinline__ double site_loglk_tab(double term, uint32_t sc, double inv, const double2 * __restrict__ tab)
{
  if (inv > 0.0) return table_log(sc ? term * rate_scale_factor(min(sc, EPA_RATE_MAXDIFF)) + inv : term + inv, tab);
  return table_log(term, tab) + (sc ? (double) sc * EPA_LOG_SCALE_THRESHOLD : 0.0);
}

// (c_i, -log c_i), c_i = 1 / (1 + (i + 1/2) / 128): written once per device at context creation
__device__ double2 g_logtab[128];

// shared memory of one warp of the lookup build: its two CLV columns (32 sites x C components each),
// later overwritten by its 32 x 16 result rows (stride 17)
template <int R>
__host__ __device__ constexpr int lookup_warp_doubles() { return (2 * 4 * R * 32 > 32 * 17 ? 2 * 4 * R * 32 : 32 * 17 + 32) & ~31; }

// A block takes LOOKUP_ITER consecutive 128-site groups of its edge: the transition matrix, the logarithm table
// and the barrier set-up are paid once per block.
#ifndef LOOKUP_ITER
#define LOOKUP_ITER 2
#endif
template <int R>
__global__ void __launch_bounds__(128, 6)
lookup_build_site_kernel(const DevModel * __restrict__ m, const double * __restrict__ clvT, size_t t_stride,
                         const uint32_t * __restrict__ scaler, int sr, const double * __restrict__ inv_lk, int n, int n_pad,
                         const EdgeDev * __restrict__ edges, const double * __restrict__ pmats_half,
                         double * __restrict__ lookup, const uint8_t * __restrict__ tipmask = nullptr, uint32_t n_tips = 0)
{
  constexpr int K = 16, C = 4 * R, WD = lookup_warp_doubles<R>();
  extern __shared__ __align__(128) double lk_stage[];   // [4 warps][WD]
  __shared__ __align__(16) double P[R * 16];
  __shared__ double2 logtab[128];                    // lanes index it with their own mantissa bits
  __shared__ uint64_t bar[4];
  const EdgeDev e = edges[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double * stage = lk_stage + warp * WD;
  // every warp owns one barrier: set up by its lane 0, no block-wide synchronisation
  if (lane == 0) { mbar_init(&bar[warp], 1); mbar_fence_init(); }
  __syncwarp();
  // The two 32-site CLV columns of the warp (C x 256 bytes each, contiguous in the site-blocked copy)
  // arrive as two bulk copies: no registers are tied up while they are in flight, six blocks stay
  // resident per SM and their load and arithmetic phases overlap.
  // A tip edge (the tip is always distal) reads the tip's n state masks, as the reference reads its tipchars,
  // instead of streaming a 0/1 CLV: 32 bytes per warp in place of 4 KB x R.
  const bool tip_d = tipmask != nullptr && e.distal < n_tips;          // block-uniform
  auto issue = [&](int site0)
  {
    // (the site-blocked copy is padded to whole 32-site blocks)
    if (site0 < n && lane == 0)
    {
      constexpr uint32_t bytes = C * 32 * sizeof(double);
      const size_t boff = (size_t) (site0 >> 5) * (size_t) (C * CLVT_BLOCK);
      mbar_expect_tx(&bar[warp], tip_d ? bytes : 2 * bytes);
      if (!tip_d) bulk_g2s(stage, clvT + (size_t) e.distal * t_stride + boff, bytes, &bar[warp]);
      bulk_g2s(stage + C * 32, clvT + (size_t) e.proximal * t_stride + boff, bytes, &bar[warp]);
    }
  };
  const int first0 = blockIdx.y * (LOOKUP_ITER * 128) + warp * 32;
  issue(first0);
  for (int i = threadIdx.x; i < R * 16; i += blockDim.x) P[i] = __ldg(pmats_half + (size_t) blockIdx.x * R * 16 + i);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) logtab[i] = g_logtab[i];
  __syncthreads();                                   // (no block-wide barrier below)
  #pragma unroll 1
  for (int it = 0; it < LOOKUP_ITER; ++it)
  {
    const int site0 = first0 + it * 128;
    if (site0 >= n) break;
    if (it > 0)
    {
      // the slice held result rows written and read through the generic proxy: order them before the bulk copies
      fence_proxy_async_smem();
      __syncwarp();
      issue(site0);
    }
    const int site = site0 + lane;
    const int s = site < n ? site : n - 1;
    const uint32_t dmask = tip_d ? __ldg(tipmask + (size_t) e.distal * n + s) : 0u;
    uint32_t kr[R];
    if (sr == 1)
      kr[0] = __ldg(scaler + (size_t) e.distal * n + s) + __ldg(scaler + (size_t) e.proximal * n + s);
    else
    {
      #pragma unroll
      for (int r = 0; r < R; ++r)
        kr[r] = __ldg(scaler + ((size_t) e.distal * n + s) * R + r) + __ldg(scaler + ((size_t) e.proximal * n + s) * R + r);
    }
    const double inv = inv_lk ? __ldg(inv_lk + s) : 0.0;
    mbar_wait(&bar[warp], (uint32_t) (it & 1));
    uint32_t sc = 0;
    double in[C];
    bool small = true;
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      double dv[4], xv[4];
      #pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        dv[k] = tip_d ? (((dmask >> k) & 1u) ? 1.0 : 0.0) : stage[(r * 4 + k) * 32 + lane];
        xv[k] = stage[(C + r * 4 + k) * 32 + lane];
      }
      #pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        double p[4];
        lds_vec<4>(P + r * 16 + i * 4, p);
        const double ta = p[0] * dv[0] + p[1] * dv[1] + p[2] * dv[2] + p[3] * dv[3];
        const double tb = p[0] * xv[0] + p[1] * xv[1] + p[2] * xv[2] + p[3] * xv[3];
        in[r * 4 + i] = ta * tb;
        small = small && (in[r * 4 + i] < EPA_SCALE_THRESHOLD);
      }
    }
    __syncwarp();                                      // every lane has consumed its columns: the slice is reused for the result rows
    if (sr == 1)
    {
      sc = kr[0];
      if (small)
      {
        sc += 1;
        #pragma unroll
        for (int c = 0; c < C; ++c) in[c] *= EPA_SCALE_FACTOR;
      }
    }
    else
    {
      // per-rate scalers: a rate whose count is d above the site's minimum weighs 2^(-256 d). The inner
      // CLV is not rescaled here: its own per-rate rescaling would only move factors of 2^256 between
      // the values and these counts.
      uint32_t kmin = 0xffffffffu;
      #pragma unroll
      for (int r = 0; r < R; ++r) kmin = min(kmin, kr[r]);
      #pragma unroll
      for (int r = 0; r < R; ++r)
      {
        const double f = rate_scale_factor(min(kr[r] - kmin, EPA_RATE_MAXDIFF));
        #pragma unroll
        for (int i = 0; i < 4; ++i) in[r * 4 + i] *= f;
      }
      sc = kmin;
    }
    // The likelihood of a state set is the sum of the likelihoods of its states: four single-state terms
    // (columns 1, 2, 4, 8 of the column table), eleven sums, fifteen logarithms.
    double single[4];
    #pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const int c = 1 << j;
      double term = 0.0;
      #pragma unroll
      for (int r = 0; r < R; ++r)
      {
        const double tr = in[r * 4] * c_coltab[c * C + r * 4] + in[r * 4 + 1] * c_coltab[c * C + r * 4 + 1]
                        + in[r * 4 + 2] * c_coltab[c * C + r * 4 + 2] + in[r * 4 + 3] * c_coltab[c * C + r * 4 + 3];
        term += tr * c_model.weights[r];
      }
      single[j] = term;
    }
    double * trow = stage + lane * 17;
    trow[0] = 0.0;                                       // column 0 = zero column
    // site_loglk_tab for the fifteen columns, with ONE test for the table logarithm's fast path (all arguments
    // positive and normal) instead of a branch per column
    double x[K];
    bool ok = true;
    const double undo = (inv > 0.0 && sc) ? rate_scale_factor(min(sc, EPA_RATE_MAXDIFF)) : 1.0;
    #pragma unroll
    for (int c = 1; c < K; ++c)
    {
      double term = 0.0;
      bool any = false;
      #pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((c >> j) & 1) { term = any ? term + single[j] : single[j]; any = true; }
      x[c] = inv > 0.0 ? (sc ? term * undo + inv : term + inv) : term;
      ok = ok && table_log_ok(x[c]);
    }
    const double add = (!(inv > 0.0) && sc) ? (double) sc * EPA_LOG_SCALE_THRESHOLD : 0.0;
    if (ok)
    {
      #pragma unroll
      for (int c = 1; c < K; ++c) trow[c] = table_log_fast(x[c], logtab) + add;
    }
    else
    {
      #pragma unroll
      for (int c = 1; c < K; ++c) trow[c] = table_log(x[c], logtab) + add;
    }
    __syncwarp();
    // the warp's 32 table rows are 4 KB contiguous in global memory: coalesced 256-byte stores
    const int n_valid = min(32, n - site0) * K;
    double * out = lookup + ((size_t) blockIdx.x * n_pad + site0) * K;
    #pragma unroll
    for (int k = 0; k < K; ++k)
    {
      const int idx = k * 32 + lane;
      if (idx < n_valid) out[idx] = stage[(idx >> 4) * 17 + (idx & 15)];
    }
  }
}

// 1/x for positive normal x (a site likelihood): hardware seed (relative error e <= 2^-23) times
// 1 + e + e^2, no slow-path branch, so that the compiler can interleave the reciprocals of several
// sites. Truncation e^3 < 2^-69: within 2 ulp.
__device__ __forceinline__ double fast_rcp(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}

// sums two values over the warp with one butterfly: after the first exchange the low half-warp
// carries the first value, the high half-warp the second (fixed order: every lane gets the same bits)
__device__ __forceinline__ void warp_sum2(double & a, double & b, int lane)
{
  const bool hi = lane >= 16;
  const double send = hi ? a : b, keep = hi ? b : a;
  double v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  #pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  a = __shfl_sync(0xffffffffu, v, 0);
  b = __shfl_sync(0xffffffffu, v, 16);
}

// derivative sums over the window, lane = site (LP/core_derivatives.c:643-858)
// G = number of distinct non-zero eigenvalues (3, or 2 / 1 when the context found equal ones and
// permuted them to the front: {1,2},{3} / {1,2,3}): components that decay alike share one entry.
template <int R, int G, bool GS>
__device__ __forceinline__ void site_derivatives(const SumRef & sr, double * ex, int w, double t,
                                                 int lane, double & f, double & df)
{
  constexpr int NK = G * R;
  if (lane < NK)
  {
    const int g = lane % G;
    const double lk = c_model.eigenvals[G == 3 ? 1 + g : (G == 2 ? 1 + 2 * g : 1)] * c_model.rates[lane / G];
    const double e = exp(lk * t) * c_model.weights[lane / G];
    const double e1 = lk * e;
    ex[lane] = e; ex[NK + lane] = e1; ex[2 * NK + lane] = lk * e1;
  }
  __syncwarp();
  double d0[NK], d1[NK], d2[NK];
  if constexpr ((NK % 2) == 0)
  {
    lds_vec<NK>(ex, d0); lds_vec<NK>(ex + NK, d1); lds_vec<NK>(ex + 2 * NK, d2);
  }
  else
  {
    #pragma unroll
    for (int k = 0; k < NK; ++k) { d0[k] = ex[k]; d1[k] = ex[NK + k]; d2[k] = ex[2 * NK + k]; }
  }
  // Trips of 32 sites are taken two at a time (sites s and s + 32 of a lane): six independent FMA
  // chains and two branch-free reciprocals in flight; an odd last trip runs alone. Rows of lanes
  // beyond the window hold (1, 0, ..., 0) and add exact zeros.
  double a1 = 0.0, a2 = 0.0;
  const int trips = (w + 31) >> 5;
  auto load_row = [&](int tr, uint32_t ra, double (&x)[NK + 2])
  {
    if constexpr (GS)
    {
      const double * g = sr.base + tr * 32 + lane;
      #pragma unroll
      for (int k = 0; k <= NK; ++k) x[k] = g[(size_t) k * sr.gstride];
    }
    else
    {
      #pragma unroll
      for (int k = 0; k < (NK + 2) / 2; ++k) lds2(ra + 16u * k, x[2 * k], x[2 * k + 1]);
    }
  };
  auto unpack = [](const uint32_t (&v)[32], double (&x)[NK + 2])
  {
    #pragma unroll
    for (int k = 0; k <= NK; ++k) x[k] = __hiloint2double((int) v[2 * k + 1], (int) v[2 * k]);
  };
  constexpr uint32_t TB = site_trip_bytes(NK);
  uint32_t ra = sr.saddr;
  int tr = 0;
  #pragma unroll 1
  for (; tr + 2 <= trips; tr += 2, ra += 2 * TB)
  {
    double xa[NK + 2], xb[NK + 2];
    if (!GS && sr.tm)
    {
      uint32_t va[32], vb[32];
      tc_ld32(sr.taddr + tr * 32, va);
      tc_ld32(sr.taddr + tr * 32 + 32, vb);
      tc_wait_ld();
      unpack(va, xa); unpack(vb, xb);
    }
    else
    {
      load_row(tr, ra, xa);
      load_row(tr + 1, ra + TB, xb);
    }
    double c0a = xa[0], c1a = 0.0, c2a = 0.0, c0b = xb[0], c1b = 0.0, c2b = 0.0;
    #pragma unroll
    for (int k = 0; k < NK; ++k)
    {
      c0a += xa[k + 1] * d0[k]; c1a += xa[k + 1] * d1[k]; c2a += xa[k + 1] * d2[k];
      c0b += xb[k + 1] * d0[k]; c1b += xb[k + 1] * d1[k]; c2b += xb[k + 1] * d2[k];
    }
    const double ia = fast_rcp(c0a), ib = fast_rcp(c0b);
    const double g1a = -c1a * ia, g1b = -c1b * ib;
    a1 += g1a; a2 += g1a * g1a - c2a * ia;
    a1 += g1b; a2 += g1b * g1b - c2b * ib;
  }
  if (tr < trips)
  {
    double xa[NK + 2];
    if (!GS && sr.tm)
    {
      uint32_t va[32];
      tc_ld32(sr.taddr + tr * 32, va);
      tc_wait_ld();
      unpack(va, xa);
    }
    else
      load_row(tr, ra, xa);
    double c0a = xa[0], c1a = 0.0, c2a = 0.0;
    #pragma unroll
    for (int k = 0; k < NK; ++k) { c0a += xa[k + 1] * d0[k]; c1a += xa[k + 1] * d1[k]; c2a += xa[k + 1] * d2[k]; }
    const double ia = fast_rcp(c0a);
    const double g1a = -c1a * ia;
    a1 += g1a; a2 += g1a * g1a - c2a * ia;
  }
  warp_sum2(a1, a2, lane);
  f = a1;
  df = a2;
  __syncwarp();            // the decay tables in `ex` are rewritten by the next evaluation / matrix build
}

// bounded Newton-Raphson, PM/optimize/opt_algorithms.c:133-262; returns 0.0 on failure
template <int R, int G, bool GS>
__device__ __forceinline__ double site_newton(const SumRef & sr, double * ex, int w, int lane,
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
    site_derivatives<R, G, GS>(sr, ex, w, x, lane, f, df);
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

// the three decaying components (j = 1, 2, 3) of one rate category -> G sumtable entries
template <int R, int G>
__device__ __forceinline__ void site_merge(double (&st)[G * R], int r, double v1, double v2, double v3)
{
  if constexpr (G == 3) { st[r * 3] = v1; st[r * 3 + 1] = v2; st[r * 3 + 2] = v3; }
  else if constexpr (G == 2) { st[r * 2] = v1 + v2; st[r * 2 + 1] = v3; }
  else st[r] = (v1 + v2) + v3;
}

// stores one finished sumtable row: st[r][j], j = 0 is the stationary component, into the slot of
// (trip, this lane). All 32 lanes store: a lane beyond the window keeps (1, 0, ..., 0) in its slot.
template <int R, int G, bool GS>
__device__ __forceinline__ void site_store_row(const SumRef & sr, int trip, int lane, bool act, double base,
                                               const double (&st)[G * R])
{
  constexpr int NK = G * R;
  const double b0 = act ? base : 1.0;
  if constexpr (GS)
  {
    double * g = sr.base + trip * 32 + lane;
    g[0] = b0;
    #pragma unroll
    for (int k = 0; k < NK; ++k) g[(size_t) (k + 1) * sr.gstride] = act ? st[k] : 0.0;
  }
  else
  {
    if (sr.tm)
    {
      uint32_t v[32];
      v[0] = (uint32_t) __double2loint(b0); v[1] = (uint32_t) __double2hiint(b0);
      #pragma unroll
      for (int k = 0; k < NK; ++k)
      {
        v[2 * k + 2] = (uint32_t) __double2loint(act ? st[k] : 0.0);
        v[2 * k + 3] = (uint32_t) __double2hiint(act ? st[k] : 0.0);
      }
      #pragma unroll
      for (int k = 2 * NK + 2; k < 32; ++k) v[k] = 0u;
      tc_st32(sr.taddr + trip * 32, v);
    }
    else
    {
      const uint32_t ra = sr.saddr + (uint32_t) trip * site_trip_bytes(NK);
      double v[NK + 2];
      v[0] = b0;
      #pragma unroll
      for (int k = 0; k < NK; ++k) v[k + 1] = act ? st[k] : 0.0;
      v[NK + 1] = 0.0;
      #pragma unroll
      for (int k = 0; k < (NK + 2) / 2; ++k) sts2(ra + 16u * k, v[2 * k], v[2 * k + 1]);
    }
  }
}

// The CLV updates of the reference rescale a site by 2^256 when all of its entries drop below
// 2^-256 (LP/core_partials.c:690-766). Inside the tiny tree that only protects later products from
// underflow: the derivative sums are ratios of per-site sums (scale free) and the site
// log-likelihood adds back exactly what the scaling took out. The entries here are products of
// two scaled CLV values and transition probabilities, far above the double range's lower end, so
// the passes below skip the rescaling; results agree with the rescaled arithmetic to rounding.

// Pass A: inner CLV toward the new tip from (D, X); returns the edge log-likelihood new_tip | inner
// over the window and leaves the pendant sumtable (inner vs tip) in `sum`.
// Per-rate scaler weights of one window site (PR kernels): counts of the two edge CLVs per rate,
// read where the reference reads them. shift_partition_focus advances the scale buffers by `begin`
// ELEMENTS even though a site has R of them (src/core/pll/pll_util.cpp:405-408), so the thorough
// phase of the reference sees entry begin + s*R + r; with bugcompat = 0 the entry of the site itself.
template <int R>
__device__ __forceinline__ uint32_t rate_weights(const uint32_t * __restrict__ sDn, const uint32_t * __restrict__ sXn,
                                                 int begin, int s, int bugcompat, double (&f)[R])
{
  const size_t base = bugcompat ? (size_t) begin + (size_t) s * R : ((size_t) begin + s) * R;
  uint32_t kr[R], kmin = 0xffffffffu;
  #pragma unroll
  for (int r = 0; r < R; ++r) { kr[r] = __ldg(sDn + base + r) + __ldg(sXn + base + r); kmin = min(kmin, kr[r]); }
  #pragma unroll
  for (int r = 0; r < R; ++r) f[r] = rate_scale_factor(min(kr[r] - kmin, EPA_RATE_MAXDIFF));
  return kmin;
}

// +I models (INV kernels): `inv_w` is the window's slice of the per-site invariant terms. The term
// joins the t-independent sumtable entry and the site likelihood (LP/core_derivatives.c:676-687,
// LP/core_likelihood.c:524-556). The reference adds it to sums built from a CLV its update may have
// rescaled by 2^256, so an invariant site does take that rescaling here (per-site scalers).
template <int R>
__device__ __forceinline__ uint32_t site_rescale_inv(double (&in)[4 * R])
{
  bool small = true;
  #pragma unroll
  for (int c = 0; c < 4 * R; ++c) small = small && (in[c] < EPA_SCALE_THRESHOLD);
  if (!small) return 0u;
  #pragma unroll
  for (int c = 0; c < 4 * R; ++c) in[c] *= EPA_SCALE_FACTOR;
  return 1u;
}

template <int R, int G, bool GS, bool PR, bool INV>
__device__ __forceinline__ double site_pass_tip(const SiteCtaSmem & cs, const double * ws, const SumRef & sr,
                                             const double * __restrict__ DT, const double * __restrict__ XT,
                                             const uint32_t * __restrict__ sD, const uint32_t * __restrict__ sX,
                                             const uint8_t * __restrict__ qc, int begin, int w, int lane, int bugcompat,
                                             const double * __restrict__ inv_w)
{
  using L = SiteWarpSmem<R>;
  // The window log-likelihood is a sum of per-site logarithms. A lane multiplies the mantissas of
  // its sites' likelihoods (each in [0.5, 1)) and adds their exponents as integers, so that one
  // logarithm per 16 sites of a lane replaces one per site.
  double acc = 0.0, prod = 1.0;
  int esum = 0, ssum = 0, since = 0;
  const int trips = (w + 31) >> 5;
  #pragma unroll 1
  for (int tr = 0; tr < trips; ++tr)
  {
    const bool act = lane + 32 * tr < w;
    const int s = act ? lane + 32 * tr : w - 1;     // lanes beyond the window recompute the last site
    const size_t off = clvt_offset<R>(begin + s);
    const double * dp = DT + off;
    const double * xp = XT + off;
    double dv[4 * R], xv[4 * R];
    #pragma unroll
    for (int c = 0; c < 4 * R; ++c) { dv[c] = __ldg(dp + (size_t) c * CLVT_BLOCK); xv[c] = __ldg(xp + (size_t) c * CLVT_BLOCK); }
    const int mask = qc[s] & 15;
    const int pos = tv_pos(mask);
    uint32_t scal;
    double rf[R];
    if constexpr (PR) scal = rate_weights<R>(sD, sX, begin, s, bugcompat, rf);
    else scal = __ldg(sD + s) + __ldg(sX + s);
    double in[4 * R];
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      #pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        double pd[4], pp[4];
        lds_vec<4>(ws + L::P_D + r * 16 + i * 4, pd);
        lds_vec<4>(ws + L::P_P + r * 16 + i * 4, pp);
        const double ta = pd[0] * dv[r * 4] + pd[1] * dv[r * 4 + 1] + pd[2] * dv[r * 4 + 2] + pd[3] * dv[r * 4 + 3];
        const double tb = pp[0] * xv[r * 4] + pp[1] * xv[r * 4 + 1] + pp[2] * xv[r * 4 + 2] + pp[3] * xv[r * 4 + 3];
        in[r * 4 + i] = ta * tb;
        if constexpr (PR) in[r * 4 + i] *= rf[r];
      }
    }
    double inv = 0.0;
    if constexpr (INV)
    {
      inv = __ldg(inv_w + s);
      if constexpr (!PR)
        if (inv > 0.0) scal += site_rescale_inv<R>(in);
    }
    double tl[4];
    lds_vec<4>(cs.tipleft + pos * 4, tl);
    const double * tvp = ws + L::TV + pos * L::TVS;
    double term = 0.0, base = 0.0;
    double st[G * R];
    double tpi[4];
    if constexpr (G == 1) lds_vec<4>(cs.tippi + pos * 4, tpi);
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      double tp[4];
      lds_vec<4>(tvp + r * 4, tp);
      const double tr = (in[r * 4] * c_model.freqs[0]) * tp[0] + (in[r * 4 + 1] * c_model.freqs[1]) * tp[1]
                      + (in[r * 4 + 2] * c_model.freqs[2]) * tp[2] + (in[r * 4 + 3] * c_model.freqs[3]) * tp[3];
      term += tr * c_model.weights[r];
      // pendant sumtable: tip side takes pi*Vinv (tipleft), inner side takes V
      if constexpr (G == 1)
      {
        // one decaying group: sum_{j>0} tl[j] (V in)[j] = sum_{k in mask} pi_k in[k] - tl[0] (V in)[0], because
        // sum_j pi_k Vinv[k][j] V[j][i] = pi_k delta_ki
        const double tot = tpi[0] * in[r * 4] + tpi[1] * in[r * 4 + 1] + tpi[2] * in[r * 4 + 2] + tpi[3] * in[r * 4 + 3];
        const double right0 = c_model.eigenvecs[0] * in[r * 4] + c_model.eigenvecs[1] * in[r * 4 + 1]
                            + c_model.eigenvecs[2] * in[r * 4 + 2] + c_model.eigenvecs[3] * in[r * 4 + 3];
        const double b = tl[0] * right0;
        base += b * c_model.weights[r];
        st[r] = tot - b;
      }
      else
      {
        double v[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const double right = c_model.eigenvecs[j * 4] * in[r * 4] + c_model.eigenvecs[j * 4 + 1] * in[r * 4 + 1]
                             + c_model.eigenvecs[j * 4 + 2] * in[r * 4 + 2] + c_model.eigenvecs[j * 4 + 3] * in[r * 4 + 3];
          v[j] = tl[j] * right;
        }
        base += v[0] * c_model.weights[r];
        site_merge<R, G>(st, r, v[1], v[2], v[3]);
      }
    }
    if constexpr (INV)
    {
      base += inv;
      if (inv > 0.0)
      {
        if (scal) term *= rate_scale_factor(min(scal, EPA_RATE_MAXDIFF));
        term += inv;
        scal = 0;
      }
    }
    site_store_row<R, G, GS>(sr, tr, lane, act, base, st);
    if (act)
    {
      ssum += (int) scal;
      const int hi = __double2hiint(term);
      const int ef = (hi >> 20) & 0x7ff;
      if (hi > 0 && ef != 0 && ef != 0x7ff)
      {
        prod *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(term));
        esum += ef - 1022;
      }
      else
        acc += log(term);                     // zero, denormal, negative or non-finite: plain path
      if (++since == 16)
      {
        acc += log(prod); prod = 1.0; since = 0;
      }
    }
  }
  if (!GS && sr.tm) tc_wait_st();
  acc += log(prod) + (double) esum * 0.693147180559945309417 + (double) ssum * EPA_LOG_SCALE_THRESHOLD;
  return warp_sum(acc);
}

// Pass A of the FIRST half round, from the per-edge tables: the pendant sumtable is the tip factor
// times the stored eigen-rotated inner CLV, the window log-likelihood is the sum of the
// preplacement table entries (the same three lengths, the same tiny tree).
template <int R, int G, bool GS, bool INV>
__device__ __forceinline__ double site_pass_first(const SiteCtaSmem & cs, const SumRef & sr,
                                                  const double * __restrict__ GT, const double * __restrict__ lk,
                                                  const uint8_t * __restrict__ qc, int begin, int w, int lane,
                                                  const double * __restrict__ inv_w)
{
  double acc = 0.0;
  const int trips = (w + 31) >> 5;
  #pragma unroll 1
  for (int tr = 0; tr < trips; ++tr)
  {
    const bool act = lane + 32 * tr < w;
    const int s = act ? lane + 32 * tr : w - 1;
    const double * gp = GT + clvt_offset<R>(begin + s);
    double gv[4 * R];
    #pragma unroll
    for (int c = 0; c < 4 * R; ++c) gv[c] = __ldg(gp + (size_t) c * CLVT_BLOCK);
    const int mask = qc[s] & 15;
    if (act) acc += __ldg(lk + (size_t) (begin + s) * 16 + mask);
    double tl[4];
    lds_vec<4>(cs.tipleft + tv_pos(mask) * 4, tl);
    double base = 0.0;
    double st[G * R];
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      base += (tl[0] * gv[r * 4]) * c_model.weights[r];
      site_merge<R, G>(st, r, tl[1] * gv[r * 4 + 1], tl[2] * gv[r * 4 + 2], tl[3] * gv[r * 4 + 3]);
    }
    if constexpr (INV) base += __ldg(inv_w + s);
    site_store_row<R, G, GS>(sr, tr, lane, act, base, st);
  }
  if (!GS && sr.tm) tc_wait_st();
  return warp_sum(acc);
}

// Pass B: inner CLV toward the distal node from (T, X); leaves the distal sumtable (D vs inner).
// `pm_x` is the offset of X's transition matrix in the warp's slice. (--raxml-blo also optimises the
// proximal edge on its own: the same pass with the two nodes swapped and D's matrix.)
template <int R, int G, bool GS, bool PR, bool INV>
__device__ __forceinline__ void site_pass_distal(const double * ws, const SumRef & sr,
                                              const double * __restrict__ DT, const double * __restrict__ XT,
                                              const uint32_t * __restrict__ sD, const uint32_t * __restrict__ sX,
                                              const uint8_t * __restrict__ qc, int begin, int w, int lane, int bugcompat,
                                              const double * __restrict__ inv_w, int pm_x = SiteWarpSmem<R>::P_P)
{
  using L = SiteWarpSmem<R>;
  const int trips = (w + 31) >> 5;
  #pragma unroll 1
  for (int tr = 0; tr < trips; ++tr)
  {
    const bool act = lane + 32 * tr < w;
    const int s = act ? lane + 32 * tr : w - 1;
    const size_t off = clvt_offset<R>(begin + s);
    const double * dp = DT + off;
    const double * xp = XT + off;
    double dv[4 * R], xv[4 * R];
    #pragma unroll
    for (int c = 0; c < 4 * R; ++c) { dv[c] = __ldg(dp + (size_t) c * CLVT_BLOCK); xv[c] = __ldg(xp + (size_t) c * CLVT_BLOCK); }
    const int pos = tv_pos(qc[s] & 15);
    const double * tvp = ws + L::TV + pos * L::TVS;
    double rf[R];
    if constexpr (PR) (void) rate_weights<R>(sD, sX, begin, s, bugcompat, rf);
    double in[4 * R];
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      double tp[4];
      lds_vec<4>(tvp + r * 4, tp);
      #pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        double pp[4];
        lds_vec<4>(ws + pm_x + r * 16 + i * 4, pp);
        const double tb = pp[0] * xv[r * 4] + pp[1] * xv[r * 4 + 1] + pp[2] * xv[r * 4 + 2] + pp[3] * xv[r * 4 + 3];
        in[r * 4 + i] = tp[i] * tb;
        if constexpr (PR) in[r * 4 + i] *= rf[r];
      }
    }
    double inv = 0.0;
    if constexpr (INV)
    {
      inv = __ldg(inv_w + s);
      if constexpr (!PR)
        if (inv > 0.0) (void) site_rescale_inv<R>(in);
    }
    double base = 0.0;
    double st[G * R];
    #pragma unroll
    for (int r = 0; r < R; ++r)
    {
      if constexpr (G == 1)
      {
        // sum over all four eigen-components = sum_k pi_k D[k] in[k]; minus the stationary one
        const double tot = (dv[r * 4] * c_model.freqs[0]) * in[r * 4] + (dv[r * 4 + 1] * c_model.freqs[1]) * in[r * 4 + 1]
                         + (dv[r * 4 + 2] * c_model.freqs[2]) * in[r * 4 + 2] + (dv[r * 4 + 3] * c_model.freqs[3]) * in[r * 4 + 3];
        const double left0 = dv[r * 4] * c_model.pivinv[0] + dv[r * 4 + 1] * c_model.pivinv[4]
                           + dv[r * 4 + 2] * c_model.pivinv[8] + dv[r * 4 + 3] * c_model.pivinv[12];
        const double right0 = c_model.eigenvecs[0] * in[r * 4] + c_model.eigenvecs[1] * in[r * 4 + 1]
                            + c_model.eigenvecs[2] * in[r * 4 + 2] + c_model.eigenvecs[3] * in[r * 4 + 3];
        const double b = left0 * right0;
        base += b * c_model.weights[r];
        st[r] = tot - b;
      }
      else
      {
        double v[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const double left = dv[r * 4] * c_model.pivinv[j] + dv[r * 4 + 1] * c_model.pivinv[4 + j]
                            + dv[r * 4 + 2] * c_model.pivinv[8 + j] + dv[r * 4 + 3] * c_model.pivinv[12 + j];
          const double right = c_model.eigenvecs[j * 4] * in[r * 4] + c_model.eigenvecs[j * 4 + 1] * in[r * 4 + 1]
                             + c_model.eigenvecs[j * 4 + 2] * in[r * 4 + 2] + c_model.eigenvecs[j * 4 + 3] * in[r * 4 + 3];
          v[j] = left * right;
        }
        base += v[0] * c_model.weights[r];
        site_merge<R, G>(st, r, v[1], v[2], v[3]);
      }
    }
    if constexpr (INV) base += inv;
    site_store_row<R, G, GS>(sr, tr, lane, act, base, st);
  }
  if (!GS && sr.tm) tc_wait_st();
}

// (q_next, q_end are only touched under the q_lock spin lock; compute-sanitizer's racecheck does not
// model the lock and reports them)
__device__ __forceinline__ unsigned long long site_next_item(SiteCtaSmem & cs, unsigned long long * counter)
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
// PR = per-rate scalers (the scaler pointers then address the node's [n][R] block, not the window)
// INV = +I model (a.tree.inv holds the per-site invariant terms)
// RAXML = --raxml-blo (optimize.cpp:274-278): pendant, distal and proximal edge optimised one after the
// other, each unconstrained in [1e-4, 100], then the pendant edge once more from the tip's side
// (pllmod_opt_optimize_branch_lengths_local with radius 1, PM/optimize/pll_optimize.c:778-1097)
template <int R, bool GS, bool PR = false, bool INV = false, bool RAXML = false, int G = 3>
__global__ void __launch_bounds__(SITE_MAX_WARPS * 32, 1)
blo_site_kernel(BloSiteArgs sa)
{
  using L = SiteWarpSmem<R>;
  const BloArgs & a = sa.b;
  extern __shared__ __align__(16) double smem_d[];
  __shared__ SiteCtaSmem cs;
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
    cs.tipleft[tv_pos(mask) * 4 + j] = acc;
    cs.tippi[tv_pos(mask) * 4 + j] = ((mask >> j) & 1) ? c_model.freqs[j] : 0.0;
  }
  if (threadIdx.x == 0) { cs.q_next = 0; cs.q_end = 0; cs.q_lock = 0; }
  __syncthreads();

  // shared memory: the fixed part (matrices, tip vectors, decay tables) of every warp, then the
  // sumtable rows of the warps that are not TMEM-backed
  __shared__ uint32_t tmem_slot;
  const int n_tm = GS ? 0 : sa.n_tmem_warps;
  if (n_tm > 0 && warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (n_tm > 0)
  {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  constexpr size_t FIX = (size_t) L::SUM;
  double * ws = smem_d + (size_t) warp * FIX;
  SumRef sr;
  sr.tm = warp < n_tm ? 1 : 0;
  sr.taddr = n_tm > 0 ? tmem_slot + ((uint32_t) ((warp & 3) * 32) << 16) + (uint32_t) (warp >> 2) * (uint32_t) sa.tmem_cols : 0u;
  sr.gstride = GS ? sa.wpad : 0;
  sr.base = GS ? sa.gscratch + ((size_t) blockIdx.x * n_warps + warp) * (size_t) sa.wpad * blo_row(R) : nullptr;
  // shared-memory rows: [warp - n_tm][trips * 32][site_row_pad(G * R)]
  const int wcap32 = (a.wcap + 31) & ~31;
  sr.saddr = (GS || sr.tm) ? 0u
           : smem_u32(smem_d + (size_t) n_warps * FIX + ((size_t) (warp - n_tm) * wcap32 + lane) * site_row_pad(G * R));
  double * ex = ws + L::EX;

  for (;;)
  {
    unsigned long long item = 0;
    if (lane == 0) item = site_next_item(cs, a.counter);
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
    if (w <= 0 || (!GS && w > a.wcap) || (GS && w > sa.wpad))
    {
      if (lane == 0) a.out[pid] = BloResult{NAN, NAN, NAN};
      continue;
    }
    const int n = a.n;
    const double * DT = sa.clvT + (size_t) ed.distal * sa.t_stride;
    const double * XT = sa.clvT + (size_t) ed.proximal * sa.t_stride;
    const uint32_t * sD = PR ? a.tree.scaler + (size_t) ed.distal * n * R : a.tree.scaler + (size_t) ed.distal * n + begin;
    const uint32_t * sX = PR ? a.tree.scaler + (size_t) ed.proximal * n * R : a.tree.scaler + (size_t) ed.proximal * n + begin;
    const uint8_t * qc = a.codes + (size_t) q * n + begin;
    const double * inv_w = INV ? a.tree.inv + begin : nullptr;         // pll_util.cpp:413-414

    const double orig = ed.length;
    double len[3] = {orig / 2.0, orig / 2.0, EPA_DEFAULT_PENDANT};      // distal, proximal, pendant
    if constexpr (RAXML)
    {
      #pragma unroll 1
      for (int mi = 0; mi < 3; ++mi) site_pmatrix<R>(cs, len[mi], ws + mi * (R * 16), ex, lane);
      site_tipvec<R>(ws + L::P_E, ws + L::TV, lane);
      auto pass_tip = [&]() -> double
      {
        return site_pass_tip<R, G, GS, PR, INV>(cs, ws, sr, DT, XT, sD, sX, qc, begin, w, lane, sa.bugcompat, inv_w);
      };
      // one recomp_iterative step (pll_optimize.c:799-833) on the edge whose sumtable is current:
      // Newton, new length, matrix rebuilt if the length moved by more than 1e-10
      auto edge = [&](int mi) -> bool
      {
        double xguess = len[mi];
        if (xguess < EPA_MIN_BRLEN || xguess > EPA_MAX_BRLEN) xguess = EPA_DEFAULT_BRLEN;
        bool failed;
        const double xres = newton_old([&](double x, double & f, double & df) { site_derivatives<R, G, GS>(sr, ex, w, x, lane, f, df); },
                                       EPA_MIN_BRLEN, xguess, EPA_MAX_BRLEN, EPA_MIN_BRLEN / 10.0, failed);
        if (failed) return false;
        const bool moved = fabs(xres - len[mi]) > 1e-10;
        len[mi] = xres;
        if (moved)
        {
          site_pmatrix<R>(cs, xres, ws + mi * (R * 16), ex, lane);
          if (mi == 2) site_tipvec<R>(ws + L::P_E, ws + L::TV, lane);
        }
        return true;
      };
      double loglikelihood = (sa.gT)
          ? site_pass_first<R, G, GS, INV>(cs, sr, sa.gT + (size_t) e * sa.g_stride, sa.lookup + (size_t) e * sa.n_pad * 16,
                                        qc, begin, w, lane, inv_w)
          : pass_tip();
      int iters = EPA_SMOOTHINGS;
      bool ok = true;
      while (iters)
      {
        if (!(ok = edge(2))) break;                                                    // pendant
        site_pass_distal<R, G, GS, PR, INV>(ws, sr, DT, XT, sD, sX, qc, begin, w, lane, sa.bugcompat, inv_w, L::P_P);
        if (!(ok = edge(0))) break;                                                    // distal
        site_pass_distal<R, G, GS, PR, INV>(ws, sr, XT, DT, sX, sD, qc, begin, w, lane, sa.bugcompat, inv_w, L::P_D);
        if (!(ok = edge(1))) break;                                                    // proximal
        double new_logl = pass_tip();                                                  // inner CLV back toward the tip
        const double pend_before = len[2];
        if (!(ok = edge(2))) break;                                                    // pendant, from the tip's side
        if (fabs(len[2] - pend_before) > 1e-10) new_logl = pass_tip();                 // score with the new matrix
        if (new_logl - loglikelihood > new_logl * 1e-13)
        {
          --iters;
          if (fabs(new_logl - loglikelihood) < EPA_BLO_EPSILON) iters = 0;
          loglikelihood = new_logl;
        }
        else
        {
          loglikelihood = new_logl;          // the older optimiser keeps a worse score and stops (:1084-1088)
          break;
        }
      }
      if (lane == 0)
      {
        BloResult res;
        res.logl = ok ? loglikelihood : 0.0;                  // a failed Newton call: PLL_FAILURE = 0 comes back
        res.distal = (orig / (len[0] + len[1])) * len[0];     // Tiny_Tree.cpp:183-185
        res.pendant = len[2];
        a.out[pid] = res;
      }
      __syncwarp();
      continue;
    }
    // optimize_branch_triplet / opt_branch_lengths_pplacer as half rounds (see kernels_blo.cuh)
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
          site_pmatrix<R>(cs, t, ws + mi * (R * 16), ex, lane);
          if (mi == 2) site_tipvec<R>(ws + L::P_E, ws + L::TV, lane);
        }
      rebuild = 0u;
      double xmin, xmax, xguess;
      if (!distal_phase)
      {
        double new_logl;
        if (first && sa.gT)
          new_logl = -site_pass_first<R, G, GS, INV>(cs, sr, sa.gT + (size_t) e * sa.g_stride,
                                                  sa.lookup + (size_t) e * sa.n_pad * 16, qc, begin, w, lane, inv_w);
        else
          new_logl = -site_pass_tip<R, G, GS, PR, INV>(cs, ws, sr, DT, XT, sD, sX, qc, begin, w, lane, sa.bugcompat, inv_w);
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
        site_pass_distal<R, G, GS, PR, INV>(ws, sr, DT, XT, sD, sX, qc, begin, w, lane, sa.bugcompat, inv_w);
        xmin = fmin(EPA_MIN_BRLEN / 2.0, original_length / 2.0);
        xmax = original_length - xmin / 10.0;
        xguess = len[0];
        if (xguess < xmin || xguess > xmax) xguess = original_length / 2.0;
      }
      const double xres = site_newton<R, G, GS>(sr, ex, w, lane, xmin, xguess, xmax, xmin / 10.0);
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
  if (n_tm > 0)
  {
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_slot) : "memory");
  }
}

}  // namespace epa
