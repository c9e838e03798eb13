// kernels_collect.cuh - output stage of a chunk: LWR over the evaluated candidates, sort, filter,
// pack into fixed-stride placement records.
//
// Reference behaviour restated (paths relative to /root/reference):
//   compute_and_set_lwr               src/set_manipulators.cpp:43-69
//   discard_by_support_threshold      src/set_manipulators.cpp:131-163
//   discard_by_accumulated_threshold  src/set_manipulators.cpp:90-113,165-190
//   Placement record                  src/sample/Placement.hpp:49-53
#pragma once
#include "common.cuh"
#include "kernels_blo.cuh"
#include "kernels_preplace.cuh"

namespace epa {

struct PlacementRec { uint64_t branch_id; double likelihood, lwr, pendant_length, distal_length; };

struct CollectArgs {
  const BloResult * res;          // [pair id]
  const uint32_t * pair_e;        // NULL = implicit all-pairs mode (pair id = q*n_edges + e)
  const uint32_t * off;           // [nq] first pair of each query (explicit mode)
  const uint32_t * cnt;           // [nq] number of pairs of each query (explicit mode)
  uint32_t nq, n_edges;
  int acc_mode;                   // 0: min-LWR filter, 1: accumulated-LWR filter
  double thresh;
  uint32_t fmin, fmax;            // fmax = record stride (> 0)
  PlacementRec * out;             // [nq][fmax]
  uint32_t * out_cnt;             // [nq]
  int * err;                      // err[0] = 3 when a likelihood is not finite, err[1] = query + 1
};

// One warp per query. Candidates are extracted best-first (log-likelihood descending, ties by the
// lower edge index - LWR is monotone in the log-likelihood) until the filter rule says stop.
__global__ void __launch_bounds__(256)
collect_kernel(CollectArgs a)
{
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= a.nq) return;
  const uint32_t base = a.pair_e ? a.off[q] : q * a.n_edges;
  const int C = (int) (a.pair_e ? a.cnt[q] : a.n_edges);
  const BloResult * r = a.res + base;

  double mx = -INFINITY;
  bool bad = false;
  for (int k = lane; k < C; k += 32)
  {
    const double l = r[k].logl;
    bad = bad || !isfinite(l);
    mx = fmax(mx, l);
  }
  mx = warp_max(mx);
  bad = __any_sync(0xffffffffu, bad);
  if (bad && lane == 0) { if (atomicCAS(&a.err[0], 0, 3) == 0) a.err[1] = (int) q + 1; }
  double tot = 0.0;
  for (int k = lane; k < C; k += 32) tot += exp(r[k].logl - mx);
  tot = warp_sum(tot);

  const int limit = min(C, (int) max(a.fmax, a.fmin));
  double pv = INFINITY; int pi = -1;          // previous pick in (logl desc, edge asc) order
  double acc = 0.0;
  int kept = 0, above = 0, summed = 0;
  bool acc_open = true;
  PlacementRec * out = a.out + (size_t) q * a.fmax;
  for (int round = 0; round < limit; ++round)
  {
    double bv = -INFINITY; int bi = INT_MAX, bk = -1;
    for (int k = lane; k < C; k += 32)
    {
      const double v = r[k].logl;
      const int e = a.pair_e ? (int) a.pair_e[base + k] : k;
      if (ranks_before(pv, pi, v, e) && ranks_before(v, e, bv, bi)) { bv = v; bi = e; bk = k; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
      if (ranks_before(ov, oi, bv, bi)) { bv = ov; bi = oi; bk = ok; }
    }
    if (bk < 0) break;
    const double lwr = exp(bv - mx) / tot;
    bool take;
    if (a.acc_mode)
    {
      // until_accumulated_reached (src/set_manipulators.cpp:90-113): add while summed < max and sum < thresh; the
      // iterator is then advanced to begin + min - 1, i.e. max(summed, min - 1) entries are kept (none for thresh = 0
      // and min = 1)
      if (acc_open && (uint32_t) summed < a.fmax && acc < a.thresh) { acc += lwr; ++summed; take = true; }
      else { acc_open = false; take = (uint32_t) kept + 1 < a.fmin; }
    }
    else
    {
      // support threshold: everything with lwr > thresh (at most max), but at least min
      const bool is_above = lwr > a.thresh;
      if (is_above) ++above;
      take = (is_above && (a.fmax == 0 || (uint32_t) above <= a.fmax)) || (!is_above && (uint32_t) kept < a.fmin);
    }
    if (!take) break;
    if (lane == 0 && (uint32_t) kept < a.fmax)
    {
      PlacementRec rec;
      rec.branch_id = (uint64_t) bi;
      rec.likelihood = bv;
      rec.lwr = lwr;
      rec.pendant_length = r[bk].pendant;
      rec.distal_length = r[bk].distal;
      out[kept] = rec;
    }
    ++kept;
    pv = bv; pi = bi;
  }
  if (lane == 0)
  {
    const uint32_t n_out = (uint32_t) min(kept, (int) a.fmax);
    a.out_cnt[q] = n_out;
    for (uint32_t k = n_out; k < a.fmax; ++k) out[k] = PlacementRec{0, 0.0, 0.0, 0.0, 0.0};   // unused slots
  }
}

}  // namespace epa
