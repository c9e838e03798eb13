// common.cuh - shared device-side definitions of libepa_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace epa {

constexpr int MAX_STATES = 20;
constexpr int MAX_RATES = 8;
constexpr int MAX_CODES = 32;                 // query character codes (DNA: 16 masks, AA: 25 codes)

// 2^256 / 2^-256: libpll's scaling factor / threshold (libpll pll.h:96-97)
#define EPA_SCALE_FACTOR 0x1p256
#define EPA_SCALE_THRESHOLD 0x1p-256
// log(2^-256)
#define EPA_LOG_SCALE_THRESHOLD (-177.445678223345993274)

// Model tables. One copy lives in global memory per context; kernels read the few entries they
// need through the read-only path (uniform addresses broadcast) or stage them in shared memory.
struct DevModel {
  int S, R, n, K;                                 // states, rate cats, sites, lookup columns (internal)
  int ncodes;                                     // number of query character codes
  int per_rate, bugcompat;
  int ngroups;                                    // distinct non-zero eigenvalues (DNA: 3, 2 or 1; see epa_ctx_create)
  double eigenvals[MAX_STATES];
  double eigenvecs[MAX_STATES * MAX_STATES];      // V    [j*S+k]
  double inv_eigenvecs[MAX_STATES * MAX_STATES];  // Vinv [k*S+j]
  double pivinv[MAX_STATES * MAX_STATES];         // freqs[k] * Vinv[k*S+j]  (sumtable, left side)
  double freqs[MAX_STATES];
  double rates[MAX_RATES];
  double weights[MAX_RATES];
  uint32_t code2mask[MAX_CODES];                  // query code -> state mask (thorough path)
  uint32_t colmask[MAX_CODES];                    // lookup column -> state mask (0 = zero column)
  uint8_t code2col[MAX_CODES];                    // query code -> lookup column (preplacement)
  uint8_t ascii2code[256];                        // 255 = invalid character
};

// Node id -> storage: ids 0..n_tips-1 are tips (their 0/1 CLVs are materialised once), ids
// >= n_tips are the directional CLV slots of the inner nodes.
struct DevTree {
  double * clv;                                   // [n_nodes][n][R][S]
  uint32_t * scaler;                              // [n_nodes][n][sr]: sr = 1 (per-site scaling) or R (per-rate scalers)
  size_t clv_stride;                              // n*R*S
  uint32_t n_tips;
  uint32_t n_nodes;
  uint32_t sr;                                    // scaler entries per site
  const double * inv;                             // +I models: [n] invariant-site term of a site likelihood,
                                                  // pinv * freq[state] * sum of rate weights where all tips share
                                                  // one state, else 0 (LP/models.c:651-760); NULL when pinv = 0
};

// libpll caps the per-rate scaler difference of a site (PLL_SCALE_RATE_MAXDIFF, LP/pll.h:104)
constexpr uint32_t EPA_RATE_MAXDIFF = 4;
// 2^(-256 d), d = 0..4 (2^-1024 is subnormal): the weight of a rate whose scaler count is d above
// the site's minimum (LP/core_likelihood.c:474-491,518-521)
__device__ __forceinline__ double rate_scale_factor(uint32_t d)
{
  const int hi = d == 0 ? 0x3ff00000 : (d == 1 ? 0x2ff00000 : (d == 2 ? 0x1ff00000 : (d == 3 ? 0x0ff00000 : 0x00040000)));
  return __hiloint2double(hi, 0);
}

// Logarithm of a site likelihood under a +I model (LP/core_likelihood.c:524-556). `term` already
// carries the factor (1 - pinv) through the rate weights, `inv` is the site's invariant term (0 for a
// variable site or a model without +I), `sc` its scaler count: where the invariant term is present
// the scaling is undone on the variable term (capped like the per-rate differences) instead of being
// added to the logarithm.
__device__ __forceinline__ double site_loglk(double term, uint32_t sc, double inv)
{
  if (inv > 0.0) return log(sc ? term * rate_scale_factor(min(sc, EPA_RATE_MAXDIFF)) + inv : term + inv);
  return log(term) + (sc ? (double) sc * EPA_LOG_SCALE_THRESHOLD : 0.0);
}

struct ClvOpDev {
  uint32_t parent, left, right;
  uint32_t tip_tip;                               // 1 = both children are tips: never rescale; 2 = exactly one child is a tip
  uint32_t lmat, rmat;                            // indices into the pmatrix array of the launch
};

struct EdgeDev {
  uint32_t distal, proximal;                      // node ids (a tip is always distal)
  double length;
};

__device__ __forceinline__ double warp_sum(double v)
{
  // fixed-order butterfly: every lane ends with the same bits
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v)
{
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int S>
__device__ __forceinline__ void load_vec(const double * __restrict__ p, double (&v)[S])
{
  static_assert(S % 2 == 0, "even state count expected");
  const double2 * p2 = reinterpret_cast<const double2 *>(p);
  #pragma unroll
  for (int i = 0; i < S / 2; ++i) { double2 t = __ldg(p2 + i); v[2 * i] = t.x; v[2 * i + 1] = t.y; }
}

template <int S>
__device__ __forceinline__ void store_vec(double * __restrict__ p, const double (&v)[S])
{
  double2 * p2 = reinterpret_cast<double2 *>(p);
  #pragma unroll
  for (int i = 0; i < S / 2; ++i) p2[i] = make_double2(v[2 * i], v[2 * i + 1]);
}

// ---- device memory through a per-process block cache (defined in epa_b200.cu) -----------------
// Allocating and freeing several GB per context makes cudaMalloc / cudaFree stall for hundreds of milliseconds now and
// then; blocks of destroyed contexts are kept per device (up to EPA_B200_DEVICE_POOL_MB, default 16384; 0 = off) and
// handed to the next context that asks for a similar size. dev_free waits for the device like cudaFree does.
cudaError_t dev_alloc_raw(void ** p, size_t bytes);
void dev_free(void * p);
template <class T> inline cudaError_t dev_alloc(T ** p, size_t bytes) { return dev_alloc_raw(reinterpret_cast<void **>(p), bytes); }

// ---- mbarrier / bulk-copy (TMA 1D) helpers, sm_90+ PTX -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void * p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

}  // namespace epa
