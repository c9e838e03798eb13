// kernels_preplace.cuh - query ingest, preplacement (HOT LOOP A) and candidate selection.
//
// Reference behaviour restated (paths relative to /root/reference):
//   valid range            src/util/Range.hpp:34-49  (only '-' trims)
//   preplacement score     src/core/Lookup_Store.hpp:110-141, src/core/place.cpp:41-95
//   LWR + candidate choice src/set_manipulators.cpp:43-69,90-113, src/core/heuristics.hpp:40-64
#pragma once
#include "common.cuh"
#include "kernels_preplace_mma.cuh"     // RowSummary (fused selection)

namespace epa {

// ---------------------------------------------------------------------------------------------
// small generic helpers
// ---------------------------------------------------------------------------------------------
__global__ void histogram_kernel(const int * __restrict__ keys, uint32_t count, uint32_t * __restrict__ hist)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) atomicAdd(&hist[keys[i]], 1u);
}

// exclusive scan of `count` uint32 by ONE block of 1024 threads; out may alias in; total optional
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const uint32_t * in, uint32_t * out, uint32_t count, uint64_t * total)
{
  // every warp owns one contiguous segment and walks it 32 elements at a time (coalesced)
  __shared__ uint64_t wsum[32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t seg = (((count + 31u) / 32u) + 31u) & ~31u;
  const uint32_t lo = min(count, warp * seg), hi = min(count, lo + seg);
  uint64_t s = 0;
  for (uint32_t i = lo + lane; i < hi; i += 32) s += in[i];
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) wsum[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint64_t run = 0;
    for (int w = 0; w < 32; ++w) { const uint64_t v = wsum[w]; wsum[w] = run; run += v; }
    if (total) *total = run;
  }
  __syncthreads();
  uint64_t run = wsum[warp];
  for (uint32_t base = lo; base < hi; base += 32)
  {
    const uint32_t i = base + lane;
    const uint32_t v = i < hi ? in[i] : 0u;
    uint32_t x = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (uint32_t) o) x += y;
    }
    if (i < hi) out[i] = (uint32_t) (run + x - v);
    run += __shfl_sync(0xffffffffu, x, 31);
  }
}

// Multi-block exclusive scan for long arrays (the single block above takes 50-170 us on 10^5 elements):
// block sums -> single-block scan of the sums -> per-block scan with its offset. SCAN_ITEMS elements
// per block (256 threads x 16 consecutive elements); out may alias in; total optional.
constexpr uint32_t SCAN_ITEMS = 4096;

__global__ void __launch_bounds__(256)
scan_block_sums_kernel(const uint32_t * __restrict__ in, uint32_t count, uint32_t * __restrict__ sums)
{
  __shared__ uint32_t ws[8];
  const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 16;
  uint32_t s = 0;
  #pragma unroll
  for (int k = 0; k < 16; ++k) s += base + k < count ? in[base + k] : 0u;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    sums[blockIdx.x] = t;
  }
}

// sums[] already holds the exclusive offsets of the blocks (low 32 bits are enough for the entries; the
// grand total is kept in 64 bits by the sums scan)
__global__ void __launch_bounds__(256)
scan_apply_kernel(const uint32_t * in, uint32_t * out, uint32_t count, const uint32_t * __restrict__ offsets)
{
  __shared__ uint32_t ws[8];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 16;
  uint32_t v[16], s = 0;
  #pragma unroll
  for (int k = 0; k < 16; ++k) { v[k] = base + k < count ? in[base + k] : 0u; s += v[k]; }
  uint32_t x = s;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t) o) x += y;
  }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  uint32_t run = offsets[blockIdx.x] + (x - s);
  for (uint32_t w = 0; w < warp; ++w) run += ws[w];
  #pragma unroll
  for (int k = 0; k < 16; ++k)
  {
    if (base + k < count) out[base + k] = run;
    run += v[k];
  }
}

// perm[offset[key] + k] = i  (order inside one key is arbitrary)
__global__ void scatter_by_key_kernel(const int * __restrict__ keys, uint32_t count,
                                      uint32_t * __restrict__ cursor, uint32_t * __restrict__ perm)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) perm[atomicAdd(&cursor[keys[i]], 1u)] = i;
}

// ---------------------------------------------------------------------------------------------
// ASCII -> codes, valid range. One warp per query.
// err[0] = status (0 ok, 1 invalid character, 2 all-gap query), err[1] = first offending query + 1,
// err[3] = longest valid range of the chunk, err[4] = number of "simple" queries.
// A DNA query is simple when it holds A, C, G, T, fully ambiguous characters and at most amb_cap other
// ambiguity codes (amb[q] = their number): those take the tensor-core (or pair-table) preplacement kernel.
// sortkey = begin for simple queries, n + 1 + begin otherwise. err[8] = simple queries with amb[q] > 0.
// ---------------------------------------------------------------------------------------------
// V = bytes per lane and step (8 when the alignment width and both buffers allow 64-bit accesses, else 1):
// a warp then moves 256 bytes per load instead of 32.
template <int V>
__global__ void __launch_bounds__(256)
encode_queries_kernel(const DevModel * __restrict__ m, const uint8_t * __restrict__ raw, uint32_t nq, int n,
                      int premask, uint8_t * __restrict__ codes, int * __restrict__ begin,
                      int * __restrict__ span, int * __restrict__ sortkey, int * __restrict__ err, int amb_cap,
                      uint8_t * __restrict__ amb)
{
  __shared__ uint8_t a2c[256];
  __shared__ int s_simple, s_maxw, s_ambig;  // per-block sums: one global atomic each instead of one per query
  a2c[threadIdx.x] = m->ascii2code[threadIdx.x];
  if (threadIdx.x == 0) { s_simple = 0; s_maxw = 0; s_ambig = 0; }
  __syncthreads();
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const bool valid = q < nq;
  const uint8_t * row = raw + (size_t) (valid ? q : 0) * n;
  uint8_t * crow = codes + (size_t) (valid ? q : 0) * n;
  int lo = n, hi = -1;
  bool bad = false;
  const bool dna = (m->S == 4);
  int n_amb = 0;                            // characters other than A, C, G, T and the fully ambiguous ones
  for (int s0 = lane * V; valid && s0 < n; s0 += 32 * V)
  {
    uint8_t chv[V], cv[V];
    if constexpr (V == 8)
    {
      const uint2 w = __ldg(reinterpret_cast<const uint2 *>(row + s0));
      #pragma unroll
      for (int k = 0; k < 4; ++k) { chv[k] = (uint8_t) (w.x >> (8 * k)); chv[4 + k] = (uint8_t) (w.y >> (8 * k)); }
    }
    else
      chv[0] = row[s0];
    #pragma unroll
    for (int k = 0; k < V; ++k)
    {
      const int s = s0 + k;
      const uint8_t ch = chv[k];
      const uint8_t c = a2c[ch];
      cv[k] = c;
      bad = bad || (c == 255);
      n_amb += ((0x8116u >> (c & 15)) & 1u) ? 0 : 1;        // masks 1, 2, 4, 8, 15 are the plain ones
      if (ch != '-') { lo = min(lo, s); hi = max(hi, s); }
    }
    if constexpr (V == 8)
    {
      uint2 o;
      o.x = (uint32_t) cv[0] | ((uint32_t) cv[1] << 8) | ((uint32_t) cv[2] << 16) | ((uint32_t) cv[3] << 24);
      o.y = (uint32_t) cv[4] | ((uint32_t) cv[5] << 8) | ((uint32_t) cv[6] << 16) | ((uint32_t) cv[7] << 24);
      *reinterpret_cast<uint2 *>(crow + s0) = o;
    }
    else
      crow[s0] = cv[0];
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  bad = __any_sync(0xffffffffu, bad);
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_amb += __shfl_xor_sync(0xffffffffu, n_amb, o);
  const bool simple = dna && n_amb <= amb_cap;
  if (valid && lane == 0)
  {
    amb[q] = (uint8_t) (simple ? n_amb : 0);
    if (simple && n_amb) atomicAdd(&s_ambig, 1);
    int b = 0, w = n;
    if (premask) { b = (hi < 0) ? 0 : lo; w = (hi < 0) ? 0 : hi - lo + 1; }
    begin[q] = b;
    span[q] = w;
    sortkey[q] = simple ? b : n + 1 + b;
    if (simple) atomicAdd(&s_simple, 1);
    atomicMax(&s_maxw, w);
    if (bad) { if (atomicCAS(&err[0], 0, 1) == 0) err[1] = (int) q + 1; }
    else if (hi < 0) { if (atomicCAS(&err[0], 0, 2) == 0) err[1] = (int) q + 1; }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    if (s_simple) atomicAdd(&err[4], s_simple);
    if (s_ambig) atomicAdd(&err[8], s_ambig);
    atomicMax(&err[3], s_maxw);
  }
}

// Queries with a few ambiguity codes other than N (R, Y, K, M, S, W, B, D, H, V; src/core/Lookup_Store.hpp:33-68) after
// the tensor-core preplacement, which scored those sites as fully ambiguous: one warp per query adds
// lookup[e][s][code] - lookup[e][s][N] for every such site s of its window and every edge e, and withdraws the row
// maximum the tensor-core epilogue recorded (the selection then finds it itself).
constexpr int AMBIG_CAP = 16;
__global__ void __launch_bounds__(256)
preplace_ambig_fix_kernel(const double * __restrict__ lookup, int n_pad, int n_edges, const uint8_t * __restrict__ codes, int n,
                          const int * __restrict__ begin, const int * __restrict__ span, const uint8_t * __restrict__ amb,
                          const uint32_t * __restrict__ perm, uint32_t n_simple, double * __restrict__ pre, size_t pre_stride,
                          double * __restrict__ qmax)
{
  const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= n_simple) return;
  const uint32_t q = perm[slot];
  if (amb[q] == 0) return;
  const uint8_t * crow = codes + (size_t) q * n;
  const int b = begin[q], w = span[q];
  double * row = pre + (size_t) q * pre_stride;
  for (int base = 0; base < w; base += 32)
  {
    const int s = b + base + lane;
    const int c = base + lane < w ? (crow[s] & 15) : 15;
    unsigned todo = __ballot_sync(0xffffffffu, !((0x8116u >> c) & 1u));
    while (todo)
    {
      const int src = __ffs((int) todo) - 1;
      todo &= todo - 1u;
      const int site = b + base + src;
      const int code = __shfl_sync(0xffffffffu, c, src);
      for (int e = lane; e < n_edges; e += 32)
      {
        const double * lk = lookup + ((size_t) e * n_pad + site) * 16;
        row[e] += __ldg(lk + code) - __ldg(lk + 15);
      }
    }
  }
  if (lane == 0) qmax[q] = NAN;
}

// ---------------------------------------------------------------------------------------------
// Site range covered by each tile of TQ (begin-sorted) queries. One warp per tile.
// range[t] = (lo rounded down to `align`, hi); maxw = max over tiles of the aligned width.
// ---------------------------------------------------------------------------------------------
__global__ void tile_range_kernel(const uint32_t * __restrict__ perm, const int * __restrict__ begin,
                                  const int * __restrict__ span, uint32_t nq, int tq, uint32_t n_tiles, int align,
                                  int2 * __restrict__ range, int * __restrict__ maxw)
{
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= n_tiles) return;
  int lo = INT_MAX, hi = 0;
  for (uint32_t k = lane; k < (uint32_t) tq; k += 32)
  {
    const uint32_t idx = t * tq + k;
    if (idx < nq)
    {
      const uint32_t q = perm[idx];
      const int b = begin[q], w = span[q];
      if (w > 0) { lo = min(lo, b); hi = max(hi, b + w); }
    }
  }
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0)
  {
    if (hi == 0) lo = 0;
    lo &= ~(align - 1);
    range[t] = make_int2(lo, hi);
    atomicMax(maxw, ((hi - lo) + align - 1) & ~(align - 1));
  }
}

// ---------------------------------------------------------------------------------------------
// HOT LOOP A. One CTA = one tile of TQ begin-sorted queries (thread = query) x ALL edges.
//   pre[q][b] = sum_{s in [begin_q, begin_q+span_q)} lookup[b][s][col(q_s)]
// The tile's column indices are staged once in shared memory (4 sites per 32-bit word, already
// multiplied by 8; out-of-range sites point at the zero column), then the CTA streams the
// [lo,hi) slice of every edge's table through an NS-deep TMA (cp.async.bulk) ring.
// Four accumulators keyed by the ABSOLUTE site index mod 4 make the result independent of how
// the tiles were cut. Tiles wider than `wc` sites run in several chunks (read-modify-write).
// lookup layout: [edge][n_pad][K] doubles, n_pad = n rounded up to 4, pad rows zero.
// dynamic smem: NS*wc*K*8 (stages, 128-byte aligned) + (wc/4)*TQ*4 (codes) + NS*8 (barriers)
// ---------------------------------------------------------------------------------------------
template <int K, int TQ, int NS>
__global__ void __launch_bounds__(TQ)
preplace_kernel(const DevModel * __restrict__ m, const double * __restrict__ lookup, int n, int n_pad,
                uint32_t n_edges, const uint8_t * __restrict__ codes, const int * __restrict__ begin,
                const int * __restrict__ span, const uint32_t * __restrict__ perm, uint32_t nq,
                const int2 * __restrict__ range, int wc, double * __restrict__ pre, size_t pre_stride)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  double * stage = reinterpret_cast<double *>(smem_raw);                         // [NS][wc*K]
  uint32_t * cw = reinterpret_cast<uint32_t *>(smem_raw + (size_t) NS * wc * K * 8);   // [wc/4][TQ]
  uint64_t * bars = reinterpret_cast<uint64_t *>(cw + (size_t) (wc / 4) * TQ);
  __shared__ uint8_t c2c[MAX_CODES];
  __shared__ uint32_t tq_q[TQ];
  __shared__ int tq_b[TQ], tq_e[TQ];

  const int tid = threadIdx.x;
  if (tid < MAX_CODES) c2c[tid] = m->code2col[tid];
  const uint32_t slot = blockIdx.x * TQ + tid;
  const bool valid = slot < nq;
  const uint32_t q = valid ? perm[slot] : 0;
  {
    const int b = valid ? begin[q] : 0, w = valid ? span[q] : 0;
    tq_q[tid] = q; tq_b[tid] = b; tq_e[tid] = b + w;
  }
  if (tid == 0)
  {
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int2 rg = range[blockIdx.x];
  const int lo = rg.x, hi = rg.y;
  uint32_t it = 0;                                  // running stage counter (ring position + parity)
  double * out = pre + (size_t) q * pre_stride;

  for (int c0 = lo; c0 < hi; c0 += wc)
  {
    const int w4 = (min(wc, hi - c0) + 3) >> 2;     // words (4 sites each) in this chunk
    const uint32_t bytes = (uint32_t) w4 * 4u * K * 8u;
    // ---- stage the tile's column indices: word (j, t) covers sites c0+4j .. c0+4j+3 of query t
    for (int idx = tid; idx < w4 * TQ; idx += TQ)
    {
      const int t = idx / w4, j = idx - t * w4;
      const int b = tq_b[t], e = tq_e[t];
      const uint8_t * crow = codes + (size_t) tq_q[t] * n;
      uint32_t word = 0;
      #pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        const int s = c0 + 4 * j + k;
        uint32_t col8 = (K == 16 ? 0u : 24u) * 8u;      // zero column: 0 for DNA, 24 for amino acids
        if (s >= b && s < e) col8 = (uint32_t) c2c[crow[s] & (MAX_CODES - 1)] * 8u;
        word |= col8 << (8 * k);
      }
      cw[(size_t) j * TQ + t] = word;
    }
    __syncthreads();
    const bool first = (c0 == lo);
    const double * src0 = lookup + (size_t) c0 * K;
    const size_t edge_stride = (size_t) n_pad * K;
    if (tid == 0)
    {
      for (uint32_t p = 0; p < (uint32_t) (NS - 1) && p < n_edges; ++p)
      {
        const uint32_t st = (it + p) % NS;
        mbar_expect_tx(&bars[st], bytes);
        bulk_g2s(stage + (size_t) st * wc * K, src0 + p * edge_stride, bytes, &bars[st]);
      }
    }
    for (uint32_t b = 0; b < n_edges; ++b, ++it)
    {
      if (tid == 0 && b + NS - 1 < n_edges)
      {
        const uint32_t st = (it + NS - 1) % NS;
        mbar_expect_tx(&bars[st], bytes);
        bulk_g2s(stage + (size_t) st * wc * K, src0 + (size_t) (b + NS - 1) * edge_stride, bytes, &bars[st]);
      }
      const uint32_t st = it % NS;
      mbar_wait(&bars[st], (it / NS) & 1u);
      const uint8_t * T = reinterpret_cast<const uint8_t *>(stage + (size_t) st * wc * K);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      #pragma unroll 2
      for (int j = 0; j < w4; ++j)
      {
        const uint32_t word = cw[(size_t) j * TQ + tid];
        const uint8_t * row = T + (size_t) j * (4 * K * 8);
        a0 += *reinterpret_cast<const double *>(row + (word & 0xffu));
        a1 += *reinterpret_cast<const double *>(row + K * 8 + ((word >> 8) & 0xffu));
        a2 += *reinterpret_cast<const double *>(row + 2 * K * 8 + ((word >> 16) & 0xffu));
        a3 += *reinterpret_cast<const double *>(row + 3 * K * 8 + (word >> 24));
      }
      const double sum = (a0 + a1) + (a2 + a3);
      if (valid)
      {
        if (first) out[b] = sum;
        else out[b] += sum;
      }
      __syncthreads();                              // stage st may be refilled from now on
    }
  }
}

// ---------------------------------------------------------------------------------------------
// DNA pair tables. A warp-wide 64-bit shared-memory load costs two wavefronts however many lanes
// share an address, so HOT LOOP A is bound by the NUMBER of lookups. For queries made of A, C, G, T
// and fully ambiguous characters only (practically all reads) two adjacent sites are scored with
// ONE lookup into a table of pair sums:
//   T2[edge][p][idx(c1, c2)] = lookup[edge][2p][col(c1)] + lookup[edge][2p+1][col(c2)]
// with character classes 0 = outside the query's range (zero column), 1..4 = A, C, G, T, 5 = fully
// ambiguous. The 16 unambiguous combinations take idx 0..15 (one 128-byte bank window: conflict
// free), the 20 combinations with a range end or a gap follow; a row is padded to 40 doubles.
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_ROW = 40;

__host__ __device__ __forceinline__ int pair_index(int c1, int c2)
{
  if (c1 >= 1 && c1 <= 4 && c2 >= 1 && c2 <= 4) return (c1 - 1) * 4 + (c2 - 1);
  // remaining 20 combinations: enumerate (c1, c2) with c1 or c2 in {0, 5}
  const int e1 = (c1 == 0) ? 0 : (c1 == 5 ? 1 : -1);
  const int e2 = (c2 == 0) ? 0 : (c2 == 5 ? 1 : -1);
  if (e1 >= 0) return 16 + e1 * 6 + c2;           // 16..27: c1 in {0,5}, c2 in 0..5
  return 28 + e2 * 4 + (c1 - 1);                  // 28..35: c1 in 1..4, c2 in {0,5}
}

// one thread per (edge, pair, idx)
__global__ void pairtab_build_kernel(const double * __restrict__ lookup, int n_pad, uint32_t n_edges,
                                     double * __restrict__ pairtab)
{
  const size_t total = (size_t) n_edges * (n_pad / 2) * PAIR_ROW;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int idx = (int) (t % PAIR_ROW);
  const size_t ep = t / PAIR_ROW;                  // edge * (n_pad/2) + pair
  int c1 = -1, c2 = -1;
  for (int a = 0; a < 6 && c1 < 0; ++a)
    for (int b = 0; b < 6; ++b)
      if (pair_index(a, b) == idx) { c1 = a; c2 = b; break; }
  double v = 0.0;
  if (c1 >= 0)
  {
    const int col[6] = {0, 1, 2, 4, 8, 15};
    const double * r0 = lookup + (ep * 2) * 16;
    v = r0[col[c1]] + r0[16 + col[c2]];
  }
  pairtab[t] = v;
}

// Same structure as preplace_kernel; one shared-memory lookup per two sites.
// dynamic smem: NS * (wc/2) * PAIR_ROW * 8 (stages) + (wc/8) * TQ * 4 (pair indices, 4 per word) + NS * 8
template <int TQ, int NS>
__global__ void __launch_bounds__(TQ)
preplace_pair_kernel(const double * __restrict__ pairtab, int n, int n_pad, uint32_t n_edges,
                     const uint8_t * __restrict__ codes, const int * __restrict__ begin,
                     const int * __restrict__ span, const uint32_t * __restrict__ perm, uint32_t nq,
                     const int2 * __restrict__ range, int wc, double * __restrict__ pre, size_t pre_stride)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const size_t stage_doubles = (size_t) (wc / 2) * PAIR_ROW;
  double * stage = reinterpret_cast<double *>(smem_raw);                              // [NS][wc/2][PAIR_ROW]
  uint32_t * cw = reinterpret_cast<uint32_t *>(smem_raw + NS * stage_doubles * 8);    // [wc/8][TQ]
  uint64_t * bars = reinterpret_cast<uint64_t *>(cw + (size_t) (wc / 8) * TQ);
  __shared__ uint32_t tq_q[TQ];
  __shared__ int tq_b[TQ], tq_e[TQ];

  const int tid = threadIdx.x;
  const uint32_t slot = blockIdx.x * TQ + tid;
  const bool valid = slot < nq;
  const uint32_t q = valid ? perm[slot] : 0;
  {
    const int b = valid ? begin[q] : 0, w = valid ? span[q] : 0;
    tq_q[tid] = q; tq_b[tid] = b; tq_e[tid] = b + w;
  }
  if (tid == 0)
  {
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int2 rg = range[blockIdx.x];
  const int lo = rg.x, hi = rg.y;                  // lo is a multiple of 8
  uint32_t it = 0;
  double * out = pre + (size_t) q * pre_stride;
  const size_t edge_stride = (size_t) (n_pad / 2) * PAIR_ROW;

  for (int c0 = lo; c0 < hi; c0 += wc)
  {
    const int w8 = (min(wc, hi - c0) + 7) >> 3;    // words (4 pairs = 8 sites each) in this chunk
    // rows past the padded alignment end do not exist: clamp the copy, the indices there are 0 -> idx(0,0)
    const int pairs_avail = n_pad / 2 - c0 / 2;
    const int pairs = min(w8 * 4, pairs_avail);
    const uint32_t bytes = (uint32_t) pairs * PAIR_ROW * 8u;
    for (int idx = tid; idx < w8 * TQ; idx += TQ)
    {
      const int t = idx / w8, j = idx - t * w8;
      const int b = tq_b[t], e = tq_e[t];
      const uint8_t * crow = codes + (size_t) tq_q[t] * n;
      uint32_t word = 0;
      #pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        const int s = c0 + 8 * j + 2 * k;
        int c1 = 0, c2 = 0;
        if (s >= b && s < e) { const int m = crow[s] & 15; c1 = m == 15 ? 5 : (m == 8 ? 4 : (m == 4 ? 3 : m)); }
        if (s + 1 >= b && s + 1 < e) { const int m = crow[s + 1] & 15; c2 = m == 15 ? 5 : (m == 8 ? 4 : (m == 4 ? 3 : m)); }
        // a pair beyond the copied rows must not be read: both classes are 0 there by construction
        word |= (uint32_t) pair_index(c1, c2) << (8 * k);
      }
      cw[(size_t) j * TQ + t] = word;
    }
    __syncthreads();
    const bool first = (c0 == lo);
    const double * src0 = pairtab + (size_t) (c0 / 2) * PAIR_ROW;
    if (tid == 0)
    {
      for (uint32_t p = 0; p < (uint32_t) (NS - 1) && p < n_edges; ++p)
      {
        const uint32_t st = (it + p) % NS;
        mbar_expect_tx(&bars[st], bytes);
        bulk_g2s(stage + st * stage_doubles, src0 + p * edge_stride, bytes, &bars[st]);
      }
    }
    for (uint32_t b = 0; b < n_edges; ++b, ++it)
    {
      if (tid == 0 && b + NS - 1 < n_edges)
      {
        const uint32_t st = (it + NS - 1) % NS;
        mbar_expect_tx(&bars[st], bytes);
        bulk_g2s(stage + st * stage_doubles, src0 + (size_t) (b + NS - 1) * edge_stride, bytes, &bars[st]);
      }
      const uint32_t st = it % NS;
      mbar_wait(&bars[st], (it / NS) & 1u);
      const double * T = stage + st * stage_doubles;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      // words past the copied rows hold idx(0,0) and read row 0 of the stage instead (value 0.0)
      const int w8_safe = pairs / 4;
      #pragma unroll 2
      for (int j = 0; j < w8_safe; ++j)
      {
        const uint32_t word = cw[(size_t) j * TQ + tid];
        const double * row = T + (size_t) j * (4 * PAIR_ROW);
        a0 += row[word & 0xffu];
        a1 += row[PAIR_ROW + ((word >> 8) & 0xffu)];
        a2 += row[2 * PAIR_ROW + ((word >> 16) & 0xffu)];
        a3 += row[3 * PAIR_ROW + (word >> 24)];
      }
      for (int pp = w8_safe * 4; pp < pairs; ++pp)       // tail pairs of a clamped copy
      {
        const uint32_t word = cw[(size_t) (pp >> 2) * TQ + tid];
        const double v = T[(size_t) pp * PAIR_ROW + ((word >> (8 * (pp & 3))) & 0xffu)];
        // same accumulator as in the unrolled loop (keyed by the absolute pair index mod 4)
        if ((pp & 3) == 0) a0 += v; else if ((pp & 3) == 1) a1 += v; else if ((pp & 3) == 2) a2 += v; else a3 += v;
      }
      const double sum = (a0 + a1) + (a2 + a3);
      if (valid)
      {
        if (first) out[b] = sum;
        else out[b] += sum;
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Candidate selection, dynamic heuristic (accumulated LWR threshold). One warp per query.
// Order: LWR descending, ties by lower edge index. Elements are taken while the running LWR sum
// is still below the threshold (the element that crosses it is included).
// Pass 1 (count): cnt[q], cut_v[q], cut_i[q] = the last element taken.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ranks_before(double v, int i, double pv, int pi)
{
  // (v,i) strictly before (pv,pi) in (value desc, index asc) order
  return (v > pv) || (v == pv && i < pi);
}

constexpr int SEL_CAP = 16;          // candidates staged per query by the count pass
constexpr int SEL_LIST = 64;         // near-best entries a warp keeps while it scans a row
#define SEL_CUT 40.0                 /* entries below max - 40 carry less than 4.3e-18 of the weight */

// Fast path of the dynamic heuristic: ONE scan of the row computes the LWR normaliser and keeps the
// entries within SEL_CUT of the best in shared memory (in ascending edge order); the best-first
// accumulation then runs on that short list. It is exact as long as the threshold is crossed inside
// the list, which the caller guarantees by only enabling it when 1 - thresh exceeds the weight
// that can hide below the cut; otherwise, and for the other heuristics, every extraction scans the
// row again (slow path). qmax (optional, NaN = absent) is the row maximum from the preplacement kernel.
__global__ void __launch_bounds__(256)
select_count_kernel(const double * __restrict__ pre, size_t pre_stride, int n_edges, uint32_t nq, int mode,
                    double thresh, int fast_ok, const double * __restrict__ qmax, uint32_t * __restrict__ cnt,
                    double * __restrict__ cut_v, int * __restrict__ cut_i, uint32_t * __restrict__ cand,
                    const uint32_t * __restrict__ qlist = nullptr)
{
  // mode 0: dynamic  - accumulated LWR threshold (until_accumulated_reached, set_manipulators.cpp:90-113)
  // mode 1: fixed    - the best ceil(thresh * edges) (until_top_percent, set_manipulators.cpp:82-88)
  // mode 2: baseball - everything within 3 log-likelihood units of the best, plus min(40 - hits, 6)
  //                    more (baseball_heuristic, src/core/heuristics.hpp:74-117)
  __shared__ double lv[8][SEL_LIST];
  __shared__ int li[8][SEL_LIST];
  // qlist: the warps work through the listed queries (the ones the fused epilogue selection left over)
  const uint32_t wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wq >= nq) return;
  const uint32_t q = qlist ? qlist[wq] : wq;
  const double * row = pre + (size_t) q * pre_stride;
  double mx = qmax ? qmax[q] : NAN;
  if (mx != mx)
  {
    mx = -INFINITY;
    for (int b = lane; b < n_edges; b += 32) mx = fmax(mx, row[b]);
    mx = warp_max(mx);
  }
  const bool fast = fast_ok && mode == 0;
  double tot = 0.0;
  int nl = 0;
  // The row is a pure stream (16 KB per query): eight 256-byte loads of the warp stay in flight while the
  // previous eight are processed (registers as the prefetch buffer; the loads of a plain unrolled loop
  // were issued too late and the kernel sat at a third of the HBM bandwidth waiting for them).
  constexpr int PF = 8;
  double buf[PF];
  #pragma unroll
  for (int i = 0; i < PF; ++i) { const int b = 32 * i + lane; buf[i] = b < n_edges ? __ldg(row + b) : -INFINITY; }
  #pragma unroll 1
  for (int base = 0; base < n_edges; base += 32 * PF)
  {
    double nxt[PF];
    #pragma unroll
    for (int i = 0; i < PF; ++i) { const int b = base + 32 * (PF + i) + lane; nxt[i] = b < n_edges ? __ldg(row + b) : -INFINITY; }
    #pragma unroll
    for (int i = 0; i < PF; ++i)
    {
      const int b = base + 32 * i + lane;
      const double v = buf[i];             // -inf beyond the row: adds nothing, is never near
      // exp(-60) = 8.8e-27: thousands of such terms cannot reach half an ulp of a sum that holds exp(0)
      if (v - mx > -60.0) tot += exp(v - mx);
      if (fast)
      {
        const bool near = v > mx - SEL_CUT;
        const unsigned m = __ballot_sync(0xffffffffu, near);
        if (near)
        {
          const int pos = nl + __popc(m & ((1u << lane) - 1u));
          if (pos < SEL_LIST) { lv[wib][pos] = v; li[wib][pos] = b; }
        }
        nl += __popc(m);
      }
    }
    #pragma unroll
    for (int i = 0; i < PF; ++i) buf[i] = nxt[i];
  }
  tot = warp_sum(tot);
  __syncwarp();

  if (fast && nl <= SEL_LIST)
  {
    double v0 = lane < nl ? lv[wib][lane] : -INFINITY, v1 = lane + 32 < nl ? lv[wib][lane + 32] : -INFINITY;
    const int i0 = lane < nl ? li[wib][lane] : INT_MAX, i1 = lane + 32 < nl ? li[wib][lane + 32] : INT_MAX;
    bool t0 = false, t1 = false;
    double acc = 0.0, pv = INFINITY; int pi = -1;
    uint32_t c = 0;
    bool exhausted = false;
    while (acc < thresh)
    {
      double bv = v0; int bi = i0;
      if (t0 || bi == INT_MAX) { bv = -INFINITY; bi = INT_MAX; }
      if (!t1 && i1 != INT_MAX && ranks_before(v1, i1, bv, bi)) { bv = v1; bi = i1; }
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ranks_before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (bi == INT_MAX) { exhausted = true; break; }
      if (bi == i0) t0 = true;
      if (bi == i1) t1 = true;
      acc += exp(bv - mx) / tot;
      pv = bv; pi = bi;
      ++c;
    }
    if (!exhausted || nl >= n_edges)
    {
      const unsigned m0 = __ballot_sync(0xffffffffu, t0), m1 = __ballot_sync(0xffffffffu, t1);
      uint32_t * cq = cand + (size_t) q * SEL_CAP;
      if (c <= SEL_CAP)
      {
        if (t0) cq[__popc(m0 & ((1u << lane) - 1u))] = (uint32_t) i0;
        if (t1) cq[__popc(m0) + __popc(m1 & ((1u << lane) - 1u))] = (uint32_t) i1;
      }
      else if (lane == 0) cq[0] = 0xffffffffu;
      if (lane == 0) { cnt[q] = c; cut_v[q] = pv; cut_i[q] = pi; }
      return;
    }
  }

  double pv = INFINITY; int pi = -1;
  double acc = 0.0;
  uint32_t c = 0;
  uint32_t target = (uint32_t) n_edges;            // modes 1, 2: number of candidates to take
  if (mode == 1) target = min((uint32_t) n_edges, (uint32_t) ceil(thresh * (double) n_edges));
  bool counting_hits = (mode == 2);
  while (c < (uint32_t) n_edges && (mode == 0 ? acc < thresh : c < target))
  {
    double bv = -INFINITY; int bi = INT_MAX;
    for (int b = lane; b < n_edges; b += 32)
    {
      const double v = row[b];
      if (ranks_before(pv, pi, v, b) && ranks_before(v, b, bv, bi)) { bv = v; bi = b; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ranks_before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (bi == INT_MAX) break;                       // nothing left (NaNs)
    if (counting_hits && bv < mx - 3.0)
    {
      // first element outside the strike box: c hits so far, add up to 6 more (at most 40 in all)
      const uint32_t hits = c;
      const uint32_t extra = min((uint32_t) 40u - hits, 6u);      // size_t arithmetic of the reference (wraps)
      target = min((uint32_t) n_edges, hits + extra);
      counting_hits = false;
      if (c >= target) break;
    }
    acc += exp(bv - mx) / tot;
    pv = bv; pi = bi;
    ++c;
  }
  if (lane == 0) { cnt[q] = c; cut_v[q] = pv; cut_i[q] = pi; cand[(size_t) q * SEL_CAP] = 0xffffffffu; }
}

// Selection from the epilogue summaries of the fused tensor-core kernel (RowSummary, kernels_preplace_mma.cuh):
// one thread per query merges its two half-row summaries and accumulates best-first exactly like the
// fast path above. A selection that would need more than SUM_K candidates, or a tile the fused kernel
// skipped (m = NaN), goes to over_list: those queries take the unfused kernels.
__global__ void __launch_bounds__(256)
select_finish_kernel(const RowSummary * __restrict__ summary, const uint32_t * __restrict__ perm, uint32_t n_fused,
                     double thresh, uint32_t * __restrict__ cnt, uint32_t * __restrict__ cand,
                     uint32_t * __restrict__ over_list, uint32_t * __restrict__ n_over)
{
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_fused) return;
  const uint32_t q = perm[slot];
  const RowSummary a = summary[2 * (size_t) q], b = summary[2 * (size_t) q + 1];
  bool over = (a.m != a.m) || (b.m != b.m);
  uint32_t chosen[SUM_K];
  uint32_t c = 0;
  if (!over)
  {
    const double m = fmax(a.m, b.m);
    const double tot = (a.S > 0.0 ? a.S * exp(a.m - m) : 0.0) + (b.S > 0.0 ? b.S * exp(b.m - m) : 0.0);
    int ia = 0, ib = 0;
    double acc = 0.0;
    while (acc < thresh)
    {
      if (c == SUM_K) { over = true; break; }          // the next best is not guaranteed to be in the lists
      const bool ha = ia < SUM_K && a.e[ia] != 0xffffffffu, hb = ib < SUM_K && b.e[ib] != 0xffffffffu;
      if (!ha && !hb) { over = true; break; }          // (trees with fewer than 2 x SUM_K edges)
      bool take_a = ha;
      if (ha && hb) take_a = ranks_before(a.v[ia], (int) a.e[ia], b.v[ib], (int) b.e[ib]);
      const double v = take_a ? a.v[ia] : b.v[ib];
      chosen[c++] = take_a ? a.e[ia] : b.e[ib];
      if (take_a) ++ia; else ++ib;
      acc += exp(v - m) / tot;
    }
  }
  if (over)
  {
    cnt[q] = 0;
    cand[(size_t) q * SEL_CAP] = 0xffffffffu;
    over_list[atomicAdd(n_over, 1u)] = q;
    return;
  }
  // staged candidates: ascending edge order
  #pragma unroll
  for (int i = 1; i < SUM_K; ++i)
    for (int j = i; j > 0; --j)
      if ((uint32_t) j < c && chosen[j] < chosen[j - 1]) { const uint32_t t = chosen[j]; chosen[j] = chosen[j - 1]; chosen[j - 1] = t; }
  for (uint32_t i = 0; i < c; ++i) cand[(size_t) q * SEL_CAP + i] = chosen[i];
  cnt[q] = c;
}

// Pass 2 (fill): pair list in query-major order, edges ascending inside a query. Queries whose
// candidates were staged by the count pass are copied, the others scan their row again.
__global__ void __launch_bounds__(256)
select_fill_kernel(const double * __restrict__ pre, size_t pre_stride, int n_edges, uint32_t nq,
                   const uint32_t * __restrict__ off, const uint32_t * __restrict__ cnt,
                   const uint32_t * __restrict__ cand, const double * __restrict__ cut_v,
                   const int * __restrict__ cut_i, const int * __restrict__ begin, int window_bin,
                   uint32_t nbins, uint32_t * __restrict__ pair_q,
                   uint32_t * __restrict__ pair_e, uint32_t * __restrict__ key_hist)
{
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const uint32_t wbin = (uint32_t) (begin[q] / window_bin);
  uint32_t o = off[q];
  const uint32_t c = cnt[q];
  if (c <= SEL_CAP && cand[(size_t) q * SEL_CAP] != 0xffffffffu)
  {
    if ((uint32_t) lane < c)
    {
      const uint32_t b = cand[(size_t) q * SEL_CAP + lane];
      pair_q[o + lane] = q;
      pair_e[o + lane] = b;
      atomicAdd(&key_hist[(size_t) b * nbins + wbin], 1u);
    }
    return;
  }
  const double * row = pre + (size_t) q * pre_stride;
  const double cv = cut_v[q]; const int ci = cut_i[q];
  for (int base = 0; base < n_edges; base += 32)
  {
    const int b = base + lane;
    bool sel = false;
    if (b < n_edges)
    {
      const double v = row[b];
      sel = (v > cv) || (v == cv && b <= ci);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, sel);
    if (sel)
    {
      const uint32_t p = o + __popc(mask & ((1u << lane) - 1u));
      pair_q[p] = q;
      pair_e[p] = (uint32_t) b;
      atomicAdd(&key_hist[(size_t) b * nbins + wbin], 1u);
    }
    o += __popc(mask);
  }
}

// work[offset(edge, window bin) + k] = pair id: edge-major, window-sorted work list so that the
// warps of a CTA (which take consecutive items) share CLV windows in L1/L2
__global__ void work_scatter_kernel(const uint32_t * __restrict__ pair_q, const uint32_t * __restrict__ pair_e,
                                    uint32_t n_pairs, const int * __restrict__ begin, int window_bin, uint32_t nbins,
                                    uint32_t * __restrict__ cursor, uint32_t * __restrict__ work)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pairs)
  {
    const size_t key = (size_t) pair_e[p] * nbins + (uint32_t) (begin[pair_q[p]] / window_bin);
    work[atomicAdd(&cursor[key], 1u)] = p;
  }
}

}  // namespace epa
