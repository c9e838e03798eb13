// kernels_preplace_mma.cuh - HOT LOOP A on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference behaviour (paths relative to /root/reference): the preplacement score
//   pre[q][b] = sum_{s in [begin_q, end_q)} lookup[b][s][col(q_s)]
// of Lookup_Store::sum_precomputed_sitelk (src/core/Lookup_Store.hpp:110-141) for every query and
// every branch (src/core/place.cpp:41-95).
//
// Formulation. For DNA queries that only hold A, C, G, T and fully ambiguous characters write
//   lookup[b][s][c] = lookup[b][s][N] - delta[b][s][c],   delta >= 0, delta[.][.][N] = 0
// (N = fully ambiguous; its likelihood is the sum over all states, so it bounds every other column).
// Then
//   pre[q][b] = (PN[b][end_q] - PN[b][begin_q]) - sum_s OneHot[q][(s, c)] * delta[b][s][c]
// with PN the per-branch prefix sums of the N column. The second term is a dense product of a 0/1
// matrix (queries x (site, state)) with the delta table. delta is stored as a 48-bit fixed-point
// number (38 fraction bits, resolution 3.6e-12) split into six unsigned 8-bit digits; the product
// runs as an EXACT integer GEMM  u8 x u8 -> s32  (tcgen05.mma.kind::i8): a 0/1 row picks at most
// one digit per site, so a digit sum over a window stays far below 2^31, and the six digit sums
// are recombined in 64-bit integer arithmetic. The only rounding is the quantisation of delta.
//
// Mapping. One persistent CTA per SM, tiles of 128 begin-sorted queries (UMMA M = 128):
//   * all warps build the tile's one-hot operand A once in shared memory (K = 4 bytes per site,
//     K-major, no swizzle: 16-byte K chunks of 8-row core matrices);
//   * warp 0 (one lane) streams the table, pre-arranged in global memory in exactly the shared
//     memory operand layout, through a 4-stage cp.async.bulk (TMA) + mbarrier ring: one stage is
//     32 sites x 32 branches x 6 digits;
//   * warp 1 (one lane) issues tcgen05.mma 128 x 192 x 32 into one of two TMEM accumulators
//     (192 columns each) and signals with tcgen05.commit;
//   * warps 2-5 drain the other accumulator with tcgen05.ld (thread = query row), recombine the
//     digits, add the prefix-sum term and write 32 scores per row.
#pragma once
#include "common.cuh"

namespace epa {

constexpr int MMA_TQ = 128;                       // queries per tile (UMMA M)
constexpr int MMA_EB = 32;                        // branches per accumulator block
constexpr int MMA_P = 6;                          // 8-bit digits per table entry
constexpr int MMA_N = MMA_EB * MMA_P;             // 192 accumulator columns (UMMA N)
constexpr int MMA_KC_STAGE = 8;                   // 16-byte K chunks (4 sites each) per table stage
constexpr int MMA_STAGES = 4;
constexpr int MMA_KC_MAX = 56;                    // A tile: at most 224 sites
constexpr int MMA_FRAC = 38;                      // fraction bits of the fixed-point delta
constexpr int MMA_THREADS = 320;                  // producer warp, MMA warp, 2 x 4 epilogue warps
constexpr uint32_t MMA_A_CHUNK_BYTES = MMA_TQ * 16;                       // 2048
constexpr uint32_t MMA_B_CHUNK_BYTES = MMA_N * 16;                        // 3072
constexpr uint32_t MMA_B_STAGE_BYTES = MMA_KC_STAGE * MMA_B_CHUNK_BYTES;  // 24576
constexpr size_t MMA_SMEM_BYTES = (size_t) MMA_KC_MAX * MMA_A_CHUNK_BYTES + (size_t) MMA_STAGES * MMA_B_STAGE_BYTES + 256;

__host__ __device__ inline int mma_kc_total(int n) { return (((n + 3) / 4) + 2 * MMA_KC_STAGE + 7) & ~7; }

// ---- table construction ---------------------------------------------------------------------
// btab[eb][kc][j * 6 + p][(s % 4) * 4 + c]: digit p of delta[eb * 32 + j][4 kc + s%4][c], c = A, C, G, T.
// One thread per (edge, K chunk): 96 contiguous bytes. flag[0] is raised when a delta does not fit.
__global__ void __launch_bounds__(256)
mma_table_kernel(const double * __restrict__ lookup, int n, int n_pad, uint32_t n_edges, int kc_total,
                 uint8_t * __restrict__ btab, int * __restrict__ flag)
{
  const int kc_used = (n + 3) / 4;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t) n_edges * kc_used) return;
  const uint32_t e = (uint32_t) (t % n_edges);
  const int kc = (int) (t / n_edges);
  uint32_t w[MMA_P][4] = {};
  bool bad = false;
  #pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    const int s = kc * 4 + i;
    if (s >= n) continue;
    const double * row = lookup + ((size_t) e * n_pad + s) * 16;
    const double ln = row[15];
    #pragma unroll
    for (int c = 0; c < 4; ++c)
    {
      const double d = (ln - row[1 << c]) * (double) (1ull << MMA_FRAC);
      // negated test catches NaN
      if (!(d >= -0.5 && d < 281474976710655.0)) { bad = true; continue; }
      const unsigned long long q = (unsigned long long) llrint(fmax(d, 0.0));
      #pragma unroll
      for (int p = 0; p < MMA_P; ++p) w[p][i] |= (uint32_t) ((q >> (8 * p)) & 0xffull) << (8 * c);
    }
  }
  if (bad) atomicExch(flag, 1);
  const uint32_t eb = e / MMA_EB, j = e % MMA_EB;
  uint4 * dst = reinterpret_cast<uint4 *>(btab + (((size_t) eb * kc_total + kc) * MMA_N + (size_t) j * MMA_P) * 16);
  #pragma unroll
  for (int p = 0; p < MMA_P; ++p) dst[p] = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

// pn[s][e] = sum_{s' < s} lookup[e][s'][N], s = 0..n, rows of e_pad doubles (site-major: the 16
// branches an epilogue thread needs for one site are 128 contiguous bytes). One warp per edge
// (sequential over 32-site groups, fixed order).
__global__ void __launch_bounds__(256)
mma_prefix_kernel(const double * __restrict__ lookup, int n, int n_pad, uint32_t n_edges, uint32_t e_pad,
                  double * __restrict__ pn)
{
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= n_edges) return;
  double * out = pn + e;
  double run = 0.0;
  if (lane == 0) out[0] = 0.0;
  for (int base = 0; base < n; base += 32)
  {
    const int s = base + lane;
    double x = s < n ? lookup[((size_t) e * n_pad + s) * 16 + 15] : 0.0;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const double y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (s < n) out[(size_t) (s + 1) * e_pad] = run + x;
    run += __shfl_sync(0xffffffffu, x, 31);
  }
}

// ---- PTX helpers ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t * bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], u8 x u8 -> s32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices; LBO = byte distance between the
// two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups
// (cute/arch/mma_sm100_desc.hpp, UMMA::SmemDescriptor; version 1 = Blackwell)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
  return (uint64_t) ((addr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// 32 lanes x 32 columns of 32-bit accumulators: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 256-bit store: one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v4(double * p, double a, double b, double c, double d)
{
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

struct PreMmaArgs {
  const uint8_t * btab;        // [n_eb][kc_total][192][16]
  const double * pn;           // [n + 1][n_eb * 32] prefix sums of the N column, site-major
  int kc_total, n;
  uint32_t n_edges, n_eb;
  const uint8_t * codes;       // [nq][n] state masks
  const int * begin;
  const int * span;
  const uint32_t * perm;       // begin-sorted queries of this launch
  uint32_t nq;
  const int2 * range;          // per tile: (lo rounded down to 4, hi)
  uint32_t n_tiles;
  double * pre;
  size_t pre_stride;
  double * qmax;               // [query] row maximum (read by the candidate selection)
  struct RowSummary * summary; // FUSED kernel: [query][2 halves], instead of pre / qmax
};

// Candidate selection fused into the epilogue (dynamic heuristic): an epilogue thread sees the scores of
// its query on half of the edges, in ascending edge order, and keeps what the selection needs of them -
// the running maximum m, S = sum exp(score - m) over everything within 60 of m (rescaled when m rises),
// and the best SUM_K scores with their edges (value descending, lower edge first among equals). The
// [query][edge] score matrix is never written. The global k-th best score, k <= SUM_K, is among the two
// halves' lists, so a selection that ends within SUM_K candidates is exact; longer ones are flagged and
// take the unfused kernels (select_finish_kernel, kernels_preplace.cuh).
constexpr int SUM_K = 8;
struct RowSummary {
  double m, S;
  double v[SUM_K];
  uint32_t e[SUM_K];
};

// instruction descriptor (UMMA::InstrDescriptor): D = s32, A = B = unsigned 8 bit, both K-major
__host__ __device__ constexpr uint32_t mma_idesc_i8()
{
  return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t) (MMA_N >> 3) << 17) | ((uint32_t) (MMA_TQ >> 4) << 24);
}

template <bool FUSED>
__global__ void __launch_bounds__(MMA_THREADS, 1)
preplace_mma_kernel(PreMmaArgs a)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t * sA = smem_raw;                                              // [kc][128][16]
  uint8_t * sB = smem_raw + (size_t) MMA_KC_MAX * MMA_A_CHUNK_BYTES;    // [stage][8][192][16]
  uint64_t * bars = reinterpret_cast<uint64_t *>(sB + (size_t) MMA_STAGES * MMA_B_STAGE_BYTES);
  uint64_t * full = bars, * empty = bars + MMA_STAGES, * tfull = bars + 2 * MMA_STAGES, * tempty = tfull + 2;
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  __shared__ uint32_t row_q[MMA_TQ];
  __shared__ int row_b[MMA_TQ], row_e[MMA_TQ];
  __shared__ double row_max[2][MMA_TQ];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
  {
    for (int s = 0; s < MMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
    mbar_fence_init();
  }
  if (warp == 1)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  uint32_t stage_it = 0;        // running table-stage counter (producer and MMA warp advance alike)
  uint32_t blk_it = 0;          // running accumulator-block counter (MMA warp and epilogue alike)

  for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
  {
    const int2 rg = a.range[tile];
    const int nkc_all = (rg.y - rg.x + 3) >> 2;                          // K chunks that hold real sites
    const int n_pass = max(1, (nkc_all + MMA_KC_MAX - 1) / MMA_KC_MAX);  // windows wider than the A tile: K is split
    if (FUSED && n_pass > 1)
    {
      // the fused epilogue needs final scores: such a tile is flagged (m = NaN) and redone by the unfused kernel
      if (tid < MMA_TQ)
      {
        const uint32_t slot = tile * MMA_TQ + tid;
        if (slot < a.nq)
        {
          const uint32_t q = a.perm[slot];
          a.summary[2 * (size_t) q].m = NAN;
          a.summary[2 * (size_t) q + 1].m = NAN;
        }
      }
      continue;
    }
    if (tid < MMA_TQ)
    {
      const uint32_t slot = tile * MMA_TQ + tid;
      const bool valid = slot < a.nq;
      const uint32_t q = valid ? a.perm[slot] : 0u;
      const int b = valid ? a.begin[q] : 0, w = valid ? a.span[q] : 0;
      row_q[tid] = valid ? q : 0xffffffffu; row_b[tid] = b; row_e[tid] = b + w;
    }
    __syncthreads();
   for (int pass = 0; pass < n_pass; ++pass)
   {
    // pass p covers sites [lo, lo + 4 nkc): later passes add to the scores the earlier ones wrote
    const int lo = rg.x + pass * MMA_KC_MAX * 4;
    const int nkc = min(MMA_KC_MAX, nkc_all - pass * MMA_KC_MAX);
    const int n_ks = (nkc + MMA_KC_STAGE - 1) / MMA_KC_STAGE;            // table stages per block
    const int kc0 = lo >> 2;
    const bool first_pass = pass == 0, last_pass = pass == n_pass - 1;
    // ---- one-hot operand: chunk kc of row r = 4 sites x 4 state bytes
    for (int idx = tid; idx < n_ks * MMA_KC_STAGE * MMA_TQ; idx += MMA_THREADS)
    {
      const int r = idx & (MMA_TQ - 1), kc = idx >> 7;
      const int b = row_b[r], e = row_e[r];
      uint32_t w4[4] = {0u, 0u, 0u, 0u};
      if (row_q[r] != 0xffffffffu && kc < nkc)
      {
        const uint8_t * crow = a.codes + (size_t) row_q[r] * a.n;
        #pragma unroll
        for (int i = 0; i < 4; ++i)
        {
          const int s = lo + 4 * kc + i;
          if (s >= b && s < e)
          {
            const int m = crow[s] & 15;
            // masks 1, 2, 4, 8 -> byte 0..3; anything else (fully ambiguous) selects nothing
            if (m == 1) w4[i] = 1u; else if (m == 2) w4[i] = 1u << 8; else if (m == 4) w4[i] = 1u << 16; else if (m == 8) w4[i] = 1u << 24;
          }
        }
      }
      *reinterpret_cast<uint4 *>(sA + (size_t) kc * MMA_A_CHUNK_BYTES + (size_t) r * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
    fence_proxy_async_smem();
    __syncthreads();

    if (warp == 0)
    {
      // ===== table producer =====
      if (lane == 0)
      {
        uint32_t it = stage_it;
        for (uint32_t eb = 0; eb < a.n_eb; ++eb)
        {
          const uint8_t * src = a.btab + ((size_t) eb * a.kc_total + kc0) * MMA_B_CHUNK_BYTES;
          for (int ks = 0; ks < n_ks; ++ks, ++it)
          {
            const uint32_t st = it % MMA_STAGES, use = it / MMA_STAGES;
            mbar_wait(&empty[st], (use & 1u) ^ 1u);
            mbar_expect_tx(&full[st], MMA_B_STAGE_BYTES);
            bulk_g2s(sB + (size_t) st * MMA_B_STAGE_BYTES, src + (size_t) ks * MMA_B_STAGE_BYTES, MMA_B_STAGE_BYTES, &full[st]);
          }
        }
      }
    }
    else if (warp == 1)
    {
      // ===== MMA issuer =====
      if (lane == 0)
      {
        uint32_t it = stage_it, bi = blk_it;
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        for (uint32_t eb = 0; eb < a.n_eb; ++eb, ++bi)
        {
          const uint32_t acc = bi & 1u, use_acc = bi >> 1;
          mbar_wait(&tempty[acc], (use_acc & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * 256u;
          for (int ks = 0; ks < n_ks; ++ks, ++it)
          {
            const uint32_t st = it % MMA_STAGES, use = it / MMA_STAGES;
            mbar_wait(&full[st], use & 1u);
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < MMA_KC_STAGE / 2; ++k)
            {
              const uint64_t da = tc_smem_desc(a_base + (uint32_t) (ks * MMA_KC_STAGE + 2 * k) * MMA_A_CHUNK_BYTES, MMA_A_CHUNK_BYTES, 128);
              const uint64_t db = tc_smem_desc(b_base + st * MMA_B_STAGE_BYTES + (uint32_t) (2 * k) * MMA_B_CHUNK_BYTES, MMA_B_CHUNK_BYTES, 128);
              tc_mma_i8(tmem_d, da, db, mma_idesc_i8(), (ks | k) ? 1u : 0u);
            }
            tc_commit(&empty[st]);            // the stage may be refilled once these MMAs have read it
          }
          tc_commit(&tfull[acc]);             // accumulator complete
        }
      }
    }
    else
    {
      // ===== epilogue: thread = query row; warps 2-5 take branches 0-15 of a block, warps 6-9
      // branches 16-31 =====
      const int quarter = warp & 3;                       // TMEM lanes this warp may read
      const int half = (warp - 2) >> 2;
      const int r = quarter * 32 + lane;
      const uint32_t q = row_q[r];
      const size_t e_pad = (size_t) a.n_eb * MMA_EB;
      const double2 * __restrict__ pnb = reinterpret_cast<const double2 *>(a.pn + (size_t) row_b[r] * e_pad + half * 16);
      const double2 * __restrict__ pne = reinterpret_cast<const double2 *>(a.pn + (size_t) row_e[r] * e_pad + half * 16);
      double * out = a.pre + (size_t) (q == 0xffffffffu ? 0u : q) * a.pre_stride + half * 16;
      uint32_t bi = blk_it;
      double rmax = -INFINITY;     // only the last pass sees final scores
      // fused selection state of this (query, half)
      double fm = -INFINITY, fS = 0.0;
      double fv[SUM_K];
      uint32_t fe[SUM_K];
      #pragma unroll
      for (int k = 0; k < SUM_K; ++k) { fv[k] = -INFINITY; fe[k] = 0xffffffffu; }
      // A block is drained in two steps of 8 branches (48 accumulator columns). The prefix-sum rows
      // of a step are requested one step ahead into one of two register buffers (raw values: the
      // subtraction waits until they are used, so the loads stay in flight behind the tensor work).
      struct PnBuf { double2 hi[4], lo[4]; };
      auto load_pn = [&](PnBuf & b, uint32_t eb, int sub)
      {
        if (first_pass)
        {
          #pragma unroll
          for (int j = 0; j < 4; ++j)
          {
            b.hi[j] = __ldg(pne + (size_t) eb * 16 + sub * 4 + j);
            b.lo[j] = __ldg(pnb + (size_t) eb * 16 + sub * 4 + j);
          }
        }
        else
        {
          // read-modify-write of this thread's own scores of the previous pass
          const double2 * prev = reinterpret_cast<const double2 *>(out + (size_t) eb * MMA_EB + sub * 8);
          #pragma unroll
          for (int j = 0; j < 4; ++j)
          {
            b.hi[j] = q != 0xffffffffu ? prev[j] : make_double2(0.0, 0.0);
            b.lo[j] = make_double2(0.0, 0.0);
          }
        }
      };
      auto finish = [&](const PnBuf & b, const uint32_t (&v0)[32], const uint32_t (&v1)[16], uint32_t e0, double * dst)
      {
        double res[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j)
        {
          unsigned long long sum = 0;
          #pragma unroll
          for (int p = 0; p < MMA_P; ++p)
          {
            const int c = j * MMA_P + p;
            const uint32_t d = c < 32 ? v0[c] : v1[c - 32];
            sum += (unsigned long long) d << (8 * p);
          }
          const double base = (j & 1) == 0 ? b.hi[j / 2].x - b.lo[j / 2].x : b.hi[j / 2].y - b.lo[j / 2].y;
          res[j] = base - (double) sum * (1.0 / (double) (1ull << MMA_FRAC));
        }
        if constexpr (FUSED)
        {
          // edges beyond the tree score -inf: they add nothing and never enter the list
          double lm = -INFINITY;
          #pragma unroll
          for (int j = 0; j < 8; ++j)
          {
            if (e0 + j >= a.n_edges) res[j] = -INFINITY;
            lm = fmax(lm, res[j]);
          }
          if (lm > fm)
          {
            fS = fm == -INFINITY ? 0.0 : fS * exp(fm - lm);
            fm = lm;
          }
          // most groups lie far below the running maximum and below the list: one test skips them
          if (lm - fm > -60.0)
          {
            #pragma unroll
            for (int j = 0; j < 8; ++j)
            {
              const double d = res[j] - fm;
              // a term within 30 of the final maximum is within 30 of the running one: those get the double
              // exponential; the rest weigh less than 1e-13 of the sum, where single precision is exact enough
              if (d > -30.0) fS += exp(d);
              else if (d > -60.0) fS += (double) __expf((float) d);
            }
          }
          if (lm > fv[SUM_K - 1])
          {
            #pragma unroll
            for (int j = 0; j < 8; ++j)
            {
              if (res[j] > fv[SUM_K - 1])
              {
                // bubble up behind every entry that is >= the new one (edges arrive ascending: lower edge first among equals)
                fv[SUM_K - 1] = res[j]; fe[SUM_K - 1] = e0 + j;
                #pragma unroll
                for (int k = SUM_K - 1; k > 0; --k)
                  if (fv[k] > fv[k - 1])
                  {
                    const double tv = fv[k]; fv[k] = fv[k - 1]; fv[k - 1] = tv;
                    const uint32_t te = fe[k]; fe[k] = fe[k - 1]; fe[k - 1] = te;
                  }
              }
            }
          }
        }
        else if (q != 0xffffffffu)
        {
          #pragma unroll
          for (int j = 0; j < 8; j += 4)
          {
            if (e0 + j + 3 < a.n_edges)
            {
              st_global_v4(dst + j, res[j], res[j + 1], res[j + 2], res[j + 3]);
              rmax = fmax(fmax(rmax, fmax(res[j], res[j + 1])), fmax(res[j + 2], res[j + 3]));
            }
            else
            {
              #pragma unroll
              for (int k = 0; k < 4; ++k)
                if (e0 + j + k < a.n_edges) { dst[j + k] = res[j + k]; rmax = fmax(rmax, res[j + k]); }
            }
          }
        }
      };
      auto drain = [&](const PnBuf & b, uint32_t taddr, uint32_t e0, double * dst)
      {
        uint32_t v0[32], v1[16];
        tc_ld32(taddr, v0);
        tc_ld16(taddr + 32, v1);
        tc_wait_ld();
        finish(b, v0, v1, e0, dst);
      };
      PnBuf bufA, bufB;
      load_pn(bufA, 0, 0);
      for (uint32_t eb = 0; eb < a.n_eb; ++eb, ++bi)
      {
        const uint32_t acc = bi & 1u, use_acc = bi >> 1;
        const uint32_t taddr = tmem_base + ((uint32_t) (quarter * 32) << 16) + acc * 256u + half * 96;
        const uint32_t e0 = eb * MMA_EB + half * 16;
        load_pn(bufB, eb, 1);
        mbar_wait(&tfull[acc], use_acc & 1u);
        tc_fence_after();
        drain(bufA, taddr, e0, out + (size_t) eb * MMA_EB);
        if (eb + 1 < a.n_eb) load_pn(bufA, eb + 1, 0);
        {
          // second step: after its accumulator columns are in registers the block is handed back
          uint32_t v0[32], v1[16];
          tc_ld32(taddr + 48, v0);
          tc_ld16(taddr + 80, v1);
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          finish(bufB, v0, v1, e0 + 8, out + (size_t) eb * MMA_EB + 8);
        }
      }
      if constexpr (FUSED)
      {
        if (q != 0xffffffffu)
        {
          RowSummary & o = a.summary[2 * (size_t) q + half];
          o.m = fm; o.S = fS;
          #pragma unroll
          for (int k = 0; k < SUM_K; ++k) { o.v[k] = fv[k]; o.e[k] = fe[k]; }
        }
      }
      else
        row_max[half][r] = rmax;
    }
    stage_it += a.n_eb * (uint32_t) n_ks;
    blk_it += a.n_eb;
    __syncthreads();           // the epilogue of the last block implies every MMA of the pass is done
    if (!FUSED && last_pass && tid < MMA_TQ && row_q[tid] != 0xffffffffu) a.qmax[row_q[tid]] = fmax(row_max[0][tid], row_max[1][tid]);
   }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

}  // namespace epa
