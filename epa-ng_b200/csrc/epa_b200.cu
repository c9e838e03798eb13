// epa_b200.cu - context management and the C ABI of libepa_b200.so (see include/epa_b200.h).
// Everything that computes runs in the hand-written sm_100a kernels of kernels_*.cuh; there is
// no CPU fallback: without a CUDA device every entry point fails with EPA_ERR_CUDA.
#include "../../include/epa_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_clv.cuh"
#include "kernels_preplace.cuh"
#include "kernels_preplace_mma.cuh"
#include "kernels_blo.cuh"
#include "kernels_blo_site.cuh"
#include "kernels_blo_generic.cuh"
#include "kernels_blo_aa.cuh"
#include "kernels_collect.cuh"

using namespace epa;

static_assert(sizeof(epa_placement) == sizeof(PlacementRec), "record layout");
static_assert(sizeof(epa_placement) == 40, "Placement is 40 bytes");


// ---- device block cache (common.cuh) -----------------------------------------------------------
namespace epa {
namespace devpool_detail {
struct DevPool {
  std::mutex m;
  struct Block { void * p; size_t cap; int dev; };
  std::vector<Block> idle;
  std::vector<Block> live;
  size_t idle_bytes = 0;
  size_t limit = 0;
  bool init = false;
  void setup()
  {
    if (init) return;
    init = true;
    const char * v = getenv("EPA_B200_DEVICE_POOL_MB");
    limit = (size_t) (v ? std::max(0, atoi(v)) : 16384) << 20;
  }
};
DevPool g_devpool;
}  // namespace devpool_detail
using devpool_detail::DevPool;
using devpool_detail::g_devpool;

cudaError_t dev_alloc_raw(void ** p, size_t bytes)
{
  *p = nullptr;
  if (bytes == 0) bytes = 1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cudaGetLastError();
  // sizes in 2 MB steps above 2 MB (the driver's own granularity): the same buffer of the next context fits
  const size_t cap = bytes > (2u << 20) ? (bytes + (2u << 20) - 1) & ~(size_t) ((2u << 20) - 1) : bytes;
  {
    std::lock_guard<std::mutex> lk(g_devpool.m);
    g_devpool.setup();
    size_t best = g_devpool.idle.size();
    for (size_t i = 0; i < g_devpool.idle.size(); ++i)
    {
      const auto & b = g_devpool.idle[i];
      if (b.dev == dev && b.cap >= cap && b.cap <= cap + cap / 4 + (1u << 20) &&
          (best == g_devpool.idle.size() || b.cap < g_devpool.idle[best].cap)) best = i;
    }
    if (best != g_devpool.idle.size())
    {
      DevPool::Block b = g_devpool.idle[best];
      g_devpool.idle.erase(g_devpool.idle.begin() + (long) best);
      g_devpool.idle_bytes -= b.cap;
      g_devpool.live.push_back(b);
      *p = b.p;
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, cap);
  if (e != cudaSuccess)
  {
    // out of memory: give the cached blocks of this device back and try once more
    (void) cudaGetLastError();
    std::vector<void *> drop;
    {
      std::lock_guard<std::mutex> lk(g_devpool.m);
      for (size_t i = g_devpool.idle.size(); i-- > 0;)
        if (g_devpool.idle[i].dev == dev)
        {
          drop.push_back(g_devpool.idle[i].p);
          g_devpool.idle_bytes -= g_devpool.idle[i].cap;
          g_devpool.idle.erase(g_devpool.idle.begin() + (long) i);
        }
    }
    for (void * q : drop) (void) cudaFree(q);
    e = cudaMalloc(p, cap);
    if (e != cudaSuccess) return e;
  }
  std::lock_guard<std::mutex> lk(g_devpool.m);
  g_devpool.live.push_back({*p, cap, dev});
  return cudaSuccess;
}

void dev_free(void * p)
{
  if (!p) return;
  DevPool::Block b{p, 0, -1};
  {
    std::lock_guard<std::mutex> lk(g_devpool.m);
    for (size_t i = 0; i < g_devpool.live.size(); ++i)
      if (g_devpool.live[i].p == p) { b = g_devpool.live[i]; g_devpool.live.erase(g_devpool.live.begin() + (long) i); break; }
  }
  if (b.dev < 0) { (void) cudaFree(p); return; }            // not ours
  int cur = 0;
  (void) cudaGetDevice(&cur);
  if (cur != b.dev) (void) cudaSetDevice(b.dev);
  (void) cudaDeviceSynchronize();                            // nothing in flight still uses the block (cudaFree semantics)
  bool keep = false;
  {
    std::lock_guard<std::mutex> lk(g_devpool.m);
    if (g_devpool.idle_bytes + b.cap <= g_devpool.limit) { g_devpool.idle.push_back(b); g_devpool.idle_bytes += b.cap; keep = true; }
  }
  if (!keep) (void) cudaFree(p);
  if (cur != b.dev) (void) cudaSetDevice(cur);
}
}  // namespace epa

extern "C" void epa_device_pool_trim(void)
{
  std::vector<epa::devpool_detail::DevPool::Block> drop;
  {
    std::lock_guard<std::mutex> lk(epa::devpool_detail::g_devpool.m);
    drop.swap(epa::devpool_detail::g_devpool.idle);
    epa::devpool_detail::g_devpool.idle_bytes = 0;
  }
  int cur = 0;
  (void) cudaGetDevice(&cur);
  for (auto & b : drop) { (void) cudaSetDevice(b.dev); (void) cudaFree(b.p); }
  (void) cudaSetDevice(cur);
}

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void * p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) { dev_free(p); p = nullptr; cap = 0; }
    // grow with some slack so that slightly larger chunks do not reallocate
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = dev_alloc(&p, want);
    if (e != cudaSuccess) { (void) cudaGetLastError(); e = dev_alloc(&p, bytes); }
    if (e == cudaSuccess) cap = (want > bytes && p) ? want : bytes;
    return e;
  }
  void release() { if (p) dev_free(p); p = nullptr; cap = 0; }
  template <typename T> T * as() const { return static_cast<T *>(p); }
};

enum Stage { ST_NONE = 0, ST_QUERIES = 1, ST_PREPLACED = 2, ST_SELECTED = 3, ST_PLACED = 4 };
constexpr int kPreplaceTQ = 256;        // queries per tile of the per-site preplacement kernel (= CTA size)
constexpr int kPairTQ = 256;            // queries per tile of the DNA pair-table kernel
constexpr int kWindowBin = 4;           // work list is sorted by (edge, window start / kWindowBin)

}  // namespace

struct epa_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int sm_count = 0;
  size_t smem_optin = 0;

  DevModel hm;
  DevModel * d_model = nullptr;
  int S = 0, R = 0, n = 0, n_pad = 0, K = 0;
  uint32_t n_tips = 0, n_slots = 0, n_nodes = 0, n_edges = 0;
  DevTree tree{};
  std::vector<EdgeDev> h_edges;
  EdgeDev * d_edges = nullptr;
  double * d_lookup = nullptr;
  double * d_pairtab = nullptr;    // DNA pair-sum tables [edge][n_pad/2][PAIR_ROW] (fallback kernel, lazy)
  bool pairtab_ready = false;
  uint8_t * d_btab = nullptr;      // DNA: fixed-point digit table of the tensor-core preplacement
  double * d_pn = nullptr;         // DNA: prefix sums of the fully-ambiguous lookup column [edge][n + 1]
  bool mma_ok = false;             // every table entry fits the fixed-point format
  double * d_clvT = nullptr;       // DNA: site-blocked CLV copy read by the lane = site BLO kernel
  uint8_t * d_tipmask = nullptr;   // DNA: [tip][site] 4-bit state masks (the lookup build reads them instead of a tip's 0/1 CLV)
  double * d_gT = nullptr;         // DNA: per-edge first-round tables of the BLO kernel (site-blocked)
  bool clvT_ready = false;         // covers d_clvT and d_gT
  bool clvs_ready = false, lookup_ready = false;
  std::vector<uint8_t> slot_filled;

  // chunk state
  int stage = ST_NONE;
  uint32_t nq = 0;
  int max_span = 0;
  uint32_t n_simple = 0;           // queries that take the pair-table preplacement kernel (sorted first)
  uint32_t n_ambig = 0;            // of those: queries with a few other ambiguity codes (fixed up after the tensor-core kernel)
  bool implicit_pairs = false;
  uint64_t n_pairs = 0;
  size_t pre_stride = 0;
  DevBuf raw, codes, begin, span, sortkey, perm, hist, range, pre, cnt, cutv, cuti, off, pair_q, pair_e,
         edge_hist, edge_off, work, res, out_rec, out_cnt, scratch, tmp, qmax, cand, scan_sums, summary, over_list, range2, amb;
  // candidate selection fused into the tensor-core preplacement (dynamic heuristic): announced by
  // epa_hint_selection before epa_preplace; the [query][edge] score matrix is then not written
  bool sel_hint = false;
  int sel_hint_mode = 0;
  double sel_hint_thresh = 0.0;
  bool fused = false;              // the last epa_preplace ran the fused kernel on perm[0, fused_n)
  uint32_t fused_n = 0;
  // host -> device prefetch of the NEXT chunk on a second stream (epa_hint_next_chunk)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr;
  DevBuf raw_next;
  // device -> host copies of the placement records also run on the copy stream, from two
  // alternating device buffers, so that they can overlap the next chunk (epa_set_deferred_results)
  DevBuf out_rec2, out_cnt2;
  cudaEvent_t ev_collect = nullptr, ev_d2h[2] = {nullptr, nullptr};
  int out_flip = 0;
  bool defer_results = false;
  const char * hint_ptr = nullptr; uint32_t hint_n = 0;         // announced next chunk
  const char * staged_ptr = nullptr; uint32_t staged_n = 0;     // chunk whose copy into raw_next is in flight
  // developer switches (environment, read once at context creation): each one routes a stage back
  // to its previous kernel so that two paths can be compared on the same inputs
  struct Switches {
    bool old_lookup = false;   // EPA_B200_OLD_LOOKUP: 4-lanes-per-site lookup build
    bool no_mma = false;       // EPA_B200_NO_MMA: shared-memory preplacement kernels instead of tcgen05
    bool no_first = false;     // EPA_B200_NO_FIRST: full first CLV pass instead of the per-edge tables
    bool no_tmem = false;      // EPA_B200_NO_TMEM: sumtables in shared memory only
    bool old_aa = false;       // EPA_B200_OLD_AA: (site, rate)-per-thread amino-acid passes
    bool lookup_tip_clv = false;   // EPA_B200_LOOKUP_TIP_CLV: the lookup build streams a tip's 0/1 CLV instead of its masks (A/B)
    bool aa_dfma = false;      // EPA_B200_AA_DFMA: the DFMA amino-acid kernel (kernels_blo_generic.cuh) instead of the DMMA one
    int gs_below = 6;          // EPA_B200_BLO_GS_BELOW: resident warps below which the global-scratch variant runs
    int gs_warps = 8;          // EPA_B200_GS_WARPS: warps per CTA of the global-scratch variant
    int site_warps = 0;        // EPA_B200_SITE_WARPS: cap on the warps per CTA of the lane = site kernel (0 = none)
    bool fused_select = false; // EPA_B200_FUSED_SELECT: candidate selection in the tensor-core epilogue (measured slower: the
                               // rare per-query events - a new list entry, a term near the running maximum - happen in
                               // some lane of almost every epilogue step; off unless requested)
  } sw;
  int * d_flags = nullptr;              // [0..1] error, [2] max tile width, [3] max span
  unsigned long long * d_counter = nullptr;
  uint64_t * d_total = nullptr;

  cudaEvent_t ev[6] = {};
  float ms[5] = {0, 0, 0, 0, 0};
  uint64_t launches = 0;
  std::string err;
};

namespace {

std::mutex g_const_mutex;
epa_ctx * g_const_owner[64] = {};

int fail(epa_ctx * ctx, int code, const char * fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
    {                                                                                              \
      (void) cudaGetLastError();                                                                   \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? EPA_ERR_NOMEM : EPA_ERR_CUDA,             \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
    }                                                                                              \
  } while (0)

#define LAUNCHED(ctx) do { (ctx)->launches++; CU(cudaGetLastError()); } while (0)

// dispatch over the compiled (states, rate categories) combinations
#define EPA_DISPATCH_SR(ctx, BODY)                                                                 \
  do {                                                                                             \
    const int S__ = (ctx)->S, R__ = (ctx)->R;                                                      \
    if (S__ == 4 && R__ == 1) { constexpr int S_ = 4, R_ = 1; BODY; }                              \
    else if (S__ == 4 && R__ == 2) { constexpr int S_ = 4, R_ = 2; BODY; }                         \
    else if (S__ == 4 && R__ == 4) { constexpr int S_ = 4, R_ = 4; BODY; }                         \
    else if (S__ == 4 && R__ == 8) { constexpr int S_ = 4, R_ = 8; BODY; }                         \
    else if (S__ == 20 && R__ == 1) { constexpr int S_ = 20, R_ = 1; BODY; }                       \
    else if (S__ == 20 && R__ == 4) { constexpr int S_ = 20, R_ = 4; BODY; }                       \
    else return fail(ctx, EPA_ERR_ARG, "unsupported states/rate_cats combination %d/%d", S__, R__);\
  } while (0)

bool supported_sr(int S, int R)
{
  return (S == 4 && (R == 1 || R == 2 || R == 4 || R == 8)) || (S == 20 && (R == 1 || R == 4));
}

int bind_constants(epa_ctx * ctx)
{
  std::lock_guard<std::mutex> lock(g_const_mutex);
  if (g_const_owner[ctx->device] == ctx) return EPA_OK;
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyToSymbol(c_model, &ctx->hm, sizeof(DevModel)));
  g_const_owner[ctx->device] = ctx;
  return EPA_OK;
}

int set_device(epa_ctx * ctx)
{
  CU(cudaSetDevice(ctx->device));
  return EPA_OK;
}

// ---- character tables ------------------------------------------------------------------------
// DNA: code = libpll state mask (A=1,C=2,G=4,T=8; libpll maps.c:46-64), lookup column = mask,
//      column 0 is the zero column.
// AA:  codes 0..19 = ARNDCQEGHILKMFPSTWYV singles, 20 = B, 21 = Z, 22 = J, 23 = fully ambiguous
//      ('-', '?', '*', '.'), 24 = X. Lookup columns 0..23 follow the codes, column 24 is the zero
//      column; X scores on the column of N in preplacement (reference quirk,
//      src/core/Lookup_Store.hpp:63-66) but is fully ambiguous in the thorough phase.
void build_char_tables(DevModel & m)
{
  memset(m.ascii2code, 255, sizeof m.ascii2code);
  memset(m.code2mask, 0, sizeof m.code2mask);
  memset(m.colmask, 0, sizeof m.colmask);
  memset(m.code2col, 0, sizeof m.code2col);
  auto both = [&](char c, uint8_t code) {
    m.ascii2code[(unsigned char) c] = code;
    if (c >= 'A' && c <= 'Z') m.ascii2code[(unsigned char) (c - 'A' + 'a')] = code;
  };
  if (m.S == 4)
  {
    m.K = 16; m.ncodes = 16;
    const char * chars = "ABCDGHKMNORSTUVWXY-.?";
    const int masks[] = {1, 14, 2, 13, 4, 11, 12, 3, 15, 15, 5, 6, 8, 8, 7, 9, 15, 10, 15, 15, 15};
    for (int i = 0; chars[i]; ++i) both(chars[i], (uint8_t) masks[i]);
    for (int c = 0; c < 16; ++c) { m.code2mask[c] = c; m.code2col[c] = (uint8_t) c; m.colmask[c] = c; }
  }
  else
  {
    m.K = 26; m.ncodes = 25;
    const char * order = "ARNDCQEGHILKMFPSTWYV";
    for (int i = 0; i < 20; ++i) { both(order[i], (uint8_t) i); m.code2mask[i] = 1u << i; }
    auto bit = [&](char c) { return 1u << (uint32_t) (strchr(order, c) - order); };
    both('B', 20); m.code2mask[20] = bit('N') | bit('D');
    both('Z', 21); m.code2mask[21] = bit('Q') | bit('E');
    both('J', 22); m.code2mask[22] = bit('I') | bit('L');
    for (const char * p = "-?*."; *p; ++p) both(*p, 23);
    m.code2mask[23] = 0xfffffu;
    both('X', 24); m.code2mask[24] = 0xfffffu;
    for (int c = 0; c < 24; ++c) { m.code2col[c] = (uint8_t) c; m.colmask[c] = m.code2mask[c]; }
    m.code2col[24] = (uint8_t) (strchr(order, 'N') - order);
    m.colmask[24] = 0; m.colmask[25] = 0;       // zero column + padding
  }
}

}  // namespace

// ==============================================================================================
//  context
// ==============================================================================================
extern "C" void epa_options_default(epa_options * o)
{
  if (!o) return;
  o->prescoring = 1;
  o->heuristic = 0;
  o->prescoring_threshold = 0.99999;
  o->premasking = 1;
  o->sliding_blo = 1;
  o->filter_acc_lwr = 0;
  o->support_threshold = 0.01;
  o->filter_min = 1;
  o->filter_max = 7;
}

extern "C" const char * epa_last_error(const epa_ctx * ctx)
{
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" uint64_t epa_launch_count(const epa_ctx * ctx) { return ctx ? ctx->launches : 0; }

extern "C" void epa_ctx_destroy(epa_ctx * ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_const_owner[ctx->device] == ctx) g_const_owner[ctx->device] = nullptr;
  }
  DevBuf * bufs[] = {&ctx->raw, &ctx->codes, &ctx->begin, &ctx->span, &ctx->sortkey, &ctx->perm, &ctx->hist, &ctx->range,
                     &ctx->pre, &ctx->cnt, &ctx->cutv, &ctx->cuti, &ctx->off, &ctx->pair_q, &ctx->pair_e,
                     &ctx->edge_hist, &ctx->edge_off, &ctx->work, &ctx->res, &ctx->out_rec, &ctx->out_cnt,
                     &ctx->scratch, &ctx->tmp, &ctx->qmax, &ctx->cand};
  for (DevBuf * b : bufs) b->release();
  ctx->raw_next.release(); ctx->out_rec2.release(); ctx->out_cnt2.release();
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->ev_collect) cudaEventDestroy(ctx->ev_collect);
  for (auto & e : ctx->ev_d2h) if (e) cudaEventDestroy(e);
  dev_free(ctx->d_model); dev_free(ctx->tree.clv); dev_free(ctx->tree.scaler); dev_free(ctx->d_edges);
  dev_free(const_cast<double *>(ctx->tree.inv));
  dev_free(ctx->d_lookup); dev_free(ctx->d_pairtab); dev_free(ctx->d_clvT); dev_free(ctx->d_tipmask); dev_free(ctx->d_gT); dev_free(ctx->d_btab); dev_free(ctx->d_pn); dev_free(ctx->d_flags); dev_free(ctx->d_counter); dev_free(ctx->d_total);
  for (auto & e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int epa_ctx_create(epa_ctx ** out, int device, const epa_model_desc * model,
                              uint32_t n_tips, const uint32_t * tip_masks, uint32_t n_clv_slots,
                              const epa_edge_desc * edges, uint32_t n_edges)
{
  epa_ctx * ctx = nullptr;      // errors before the context exists go to the thread-local message
  if (!out || !model || !tip_masks || !edges) return fail(ctx, EPA_ERR_ARG, "null argument");
  *out = nullptr;
  if (!supported_sr((int) model->states, (int) model->rate_cats))
    return fail(ctx, EPA_ERR_ARG, "unsupported states/rate_cats combination %u/%u", model->states, model->rate_cats);
  if (!(model->pinv >= 0.0 && model->pinv < 1.0))        // LP/models.c:510-518
    return fail(ctx, EPA_ERR_ARG, "Invalid proportion of invariant sites (%f)", model->pinv);
  if (model->sites == 0 || n_tips < 3 || n_edges == 0) return fail(ctx, EPA_ERR_ARG, "empty tree or alignment");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    (void) cudaGetLastError();
    return fail(ctx, EPA_ERR_CUDA, "no CUDA device available: libepa_b200 has no CPU path");
  }
  if (device < 0 || device >= ndev || device >= 64) return fail(ctx, EPA_ERR_ARG, "invalid device %d", device);
  for (uint32_t i = 0; i < n_edges; ++i)
  {
    const uint32_t lim = n_tips + n_clv_slots;
    if (edges[i].distal >= lim || edges[i].proximal >= lim)
      return fail(ctx, EPA_ERR_ARG, "edge %u references node out of range", i);
    if (edges[i].proximal < n_tips && edges[i].distal >= n_tips)
      return fail(ctx, EPA_ERR_ARG, "edge %u: a tip must be the distal side", i);
  }

  ctx = new epa_ctx();
  ctx->device = device;
  auto bail = [&](int code) { std::string msg = ctx->err; epa_ctx_destroy(ctx); g_create_error = msg; return code; };
#define CUC(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
    {                                                                                              \
      (void) cudaGetLastError();                                                                   \
      fail(ctx, EPA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                     \
      return bail(e_ == cudaErrorMemoryAllocation ? EPA_ERR_NOMEM : EPA_ERR_CUDA);                 \
    }                                                                                              \
  } while (0)

  CUC(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUC(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
  {
    fail(ctx, EPA_ERR_CUDA, "device %d is sm_%d%d; libepa_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return bail(EPA_ERR_CUDA);
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (auto & e : ctx->ev) CUC(cudaEventCreate(&e));

  ctx->sw.old_lookup = getenv("EPA_B200_OLD_LOOKUP") != nullptr;
  ctx->sw.no_mma = getenv("EPA_B200_NO_MMA") != nullptr;
  ctx->sw.no_first = getenv("EPA_B200_NO_FIRST") != nullptr;
  ctx->sw.no_tmem = getenv("EPA_B200_NO_TMEM") != nullptr;
  ctx->sw.old_aa = getenv("EPA_B200_OLD_AA") != nullptr;
  ctx->sw.aa_dfma = getenv("EPA_B200_AA_DFMA") != nullptr;
  ctx->sw.lookup_tip_clv = getenv("EPA_B200_LOOKUP_TIP_CLV") != nullptr;
  if (const char * v = getenv("EPA_B200_BLO_GS_BELOW")) ctx->sw.gs_below = atoi(v);
  if (const char * v = getenv("EPA_B200_GS_WARPS")) ctx->sw.gs_warps = std::max(1, std::min(12, atoi(v)));
  if (const char * v = getenv("EPA_B200_SITE_WARPS")) ctx->sw.site_warps = atoi(v);
  ctx->sw.fused_select = getenv("EPA_B200_FUSED_SELECT") != nullptr;
  DevModel & m = ctx->hm;
  memset(&m, 0, sizeof m);
  const int S = (int) model->states, R = (int) model->rate_cats;
  m.S = S; m.R = R; m.n = (int) model->sites;
  m.per_rate = (model->flags & EPA_FLAG_RATE_SCALERS) ? 1 : 0;
  m.bugcompat = (model->flags & EPA_FLAG_BUGCOMPAT_FOCUS) ? 1 : 0;
  for (int i = 0; i < S; ++i) { m.eigenvals[i] = model->eigenvals[i]; m.freqs[i] = model->freqs[i]; }
  for (int i = 0; i < S * S; ++i) { m.eigenvecs[i] = model->eigenvecs[i]; m.inv_eigenvecs[i] = model->inv_eigenvecs[i]; }
  {
    // Put the stationary eigenpair (eigenvalue 0 of a reversible model) first: a permutation of the
    // eigenpairs leaves P(t) unchanged, and the thorough kernel folds that component into one plane.
    int z = 0;
    double scale = 0.0;
    for (int i = 0; i < S; ++i) { if (m.eigenvals[i] > m.eigenvals[z]) z = i; scale = std::max(scale, std::fabs(m.eigenvals[i])); }
    if (std::fabs(m.eigenvals[z]) > 1e-9 * std::max(scale, 1e-300))
    {
      fail(ctx, EPA_ERR_ARG, "the rate matrix has no zero eigenvalue (largest is %g): not a reversible model", m.eigenvals[z]);
      return bail(EPA_ERR_ARG);
    }
    if (z != 0)
    {
      std::swap(m.eigenvals[0], m.eigenvals[z]);
      for (int k = 0; k < S; ++k)
      {
        std::swap(m.eigenvecs[0 * S + k], m.eigenvecs[z * S + k]);             // rows of V
        std::swap(m.inv_eigenvecs[k * S + 0], m.inv_eigenvecs[k * S + z]);     // columns of Vinv
      }
    }
  }
  // Equal non-zero eigenvalues (JC69, F81: all three; K80, ...: two): their sumtable components decay
  // alike, so the thorough DNA kernel keeps one merged entry per group. Equal ones are permuted to
  // indices 1, 2 (a permutation of the eigenpairs leaves P(t) unchanged).
  m.ngroups = S - 1;
  if (S == 4 && !getenv("EPA_B200_NO_EIGEN_GROUPS"))
  {
    auto swap_pair = [&](int a, int b)
    {
      std::swap(m.eigenvals[a], m.eigenvals[b]);
      for (int k = 0; k < S; ++k)
      {
        std::swap(m.eigenvecs[a * S + k], m.eigenvecs[b * S + k]);
        std::swap(m.inv_eigenvecs[k * S + a], m.inv_eigenvecs[k * S + b]);
      }
    };
    double scale = 0.0;
    for (int i = 1; i < 4; ++i) scale = std::max(scale, std::fabs(m.eigenvals[i]));
    const double tol = 1e-13 * scale;
    auto eq = [&](int a, int b) { return std::fabs(m.eigenvals[a] - m.eigenvals[b]) <= tol; };
    if (eq(1, 2) && eq(1, 3) && eq(2, 3)) m.ngroups = 1;
    else if (eq(1, 2)) m.ngroups = 2;
    else if (eq(1, 3)) { swap_pair(2, 3); m.ngroups = 2; }
    else if (eq(2, 3)) { swap_pair(1, 3); m.ngroups = 2; }
  }
  for (int i = 0; i < S * S; ++i) m.pivinv[i] = m.freqs[i / S] * m.inv_eigenvecs[i];
  // +I (LP/core_pmatrix.c:209-220, LP/core_derivatives.c:757-772, LP/core_likelihood.c:524-537): the
  // rates are stretched by 1 / (1 - pinv) wherever a branch length meets them, and the variable part
  // of every site likelihood carries (1 - pinv) - folded into the rate weights here
  const double pinv = model->pinv;
  for (int r = 0; r < R; ++r)
  {
    m.rates[r] = pinv > 0.0 ? model->rates[r] / (1.0 - pinv) : model->rates[r];
    m.weights[r] = pinv > 0.0 ? model->rate_weights[r] * (1.0 - pinv) : model->rate_weights[r];
  }
  build_char_tables(m);
  ctx->S = S; ctx->R = R; ctx->n = m.n; ctx->n_pad = (m.n + 3) & ~3; ctx->K = m.K;
  ctx->n_tips = n_tips; ctx->n_slots = n_clv_slots; ctx->n_nodes = n_tips + n_clv_slots; ctx->n_edges = n_edges;
  ctx->slot_filled.assign(ctx->n_nodes, 0);
  for (uint32_t i = 0; i < n_tips; ++i) ctx->slot_filled[i] = 1;

  const uint32_t all = S == 4 ? 0xfu : 0xfffffu;
  const size_t tip_sites = (size_t) n_tips * m.n;
  for (size_t i = 0; i < tip_sites; ++i)
    if (tip_masks[i] == 0 || (tip_masks[i] & ~all))
    {
      fail(ctx, EPA_ERR_ARG, "invalid tip state mask at tip %zu site %zu", i / m.n, i % m.n);
      return bail(EPA_ERR_ARG);
    }

  ctx->tree.inv = nullptr;
  if (pinv > 0.0)
  {
    // pll_update_invariant_sites (LP/models.c:651-760): a site is invariant when the AND of all tip
    // masks leaves exactly one state; its term of the site likelihood is pinv * freq[state], summed
    // over the rate categories with their weights (LP/core_likelihood.c:529-533)
    std::vector<double> inv((size_t) m.n, 0.0);
    for (int s = 0; s < m.n; ++s)
    {
      uint32_t st = all;
      for (uint32_t t = 0; t < n_tips; ++t) st &= tip_masks[(size_t) t * m.n + s];
      if (st != 0 && (st & (st - 1)) == 0)
      {
        const double f = model->freqs[__builtin_ctz(st)];
        double acc = 0.0;
        for (int r = 0; r < R; ++r) acc += model->rate_weights[r] * f * pinv;
        inv[s] = acc;
      }
    }
    double * d_inv = nullptr;
    CUC(dev_alloc(&d_inv, inv.size() * sizeof(double)));
    ctx->tree.inv = d_inv;
    CUC(cudaMemcpy(d_inv, inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  CUC(dev_alloc(&ctx->d_model, sizeof(DevModel)));
  CUC(cudaMemcpy(ctx->d_model, &m, sizeof(DevModel), cudaMemcpyHostToDevice));
  ctx->tree.clv_stride = (size_t) m.n * R * S;
  ctx->tree.n_tips = n_tips; ctx->tree.n_nodes = ctx->n_nodes;
  CUC(dev_alloc(&ctx->tree.clv, ctx->tree.clv_stride * ctx->n_nodes * sizeof(double)));
  ctx->tree.sr = m.per_rate ? (uint32_t) R : 1u;
  CUC(dev_alloc(&ctx->tree.scaler, (size_t) m.n * ctx->tree.sr * ctx->n_nodes * sizeof(uint32_t)));
  CUC(cudaMemset(ctx->tree.scaler, 0, (size_t) m.n * ctx->tree.sr * ctx->n_nodes * sizeof(uint32_t)));
  ctx->h_edges.resize(n_edges);
  for (uint32_t i = 0; i < n_edges; ++i) ctx->h_edges[i] = EdgeDev{edges[i].distal, edges[i].proximal, edges[i].length};
  CUC(dev_alloc(&ctx->d_edges, n_edges * sizeof(EdgeDev)));
  CUC(cudaMemcpy(ctx->d_edges, ctx->h_edges.data(), n_edges * sizeof(EdgeDev), cudaMemcpyHostToDevice));
  CUC(dev_alloc(&ctx->d_flags, 16 * sizeof(int)));
  CUC(cudaMemset(ctx->d_flags, 0, 16 * sizeof(int)));
  CUC(dev_alloc(&ctx->d_counter, sizeof(unsigned long long)));
  CUC(dev_alloc(&ctx->d_total, 2 * sizeof(uint64_t)));

  // tips -> 0/1 CLVs
  {
    uint32_t * d_masks = nullptr;
    CUC(dev_alloc(&d_masks, tip_sites * sizeof(uint32_t)));
    CUC(cudaMemcpy(d_masks, tip_masks, tip_sites * sizeof(uint32_t), cudaMemcpyHostToDevice));
    const unsigned blocks = (unsigned) ((tip_sites + 255) / 256);
    if (S == 4) tip_expand_kernel<4><<<blocks, 256, 0, ctx->stream>>>(ctx->d_model, ctx->tree, d_masks, tip_sites);
    else tip_expand_kernel<20><<<blocks, 256, 0, ctx->stream>>>(ctx->d_model, ctx->tree, d_masks, tip_sites);
    ctx->launches++;
    CUC(cudaGetLastError());
    CUC(cudaStreamSynchronize(ctx->stream));
    dev_free(d_masks);
    if (S == 4)
    {
      std::vector<uint8_t> m8(tip_sites);
      for (size_t i = 0; i < tip_sites; ++i) m8[i] = (uint8_t) (tip_masks[i] & 15u);
      CUC(dev_alloc(&ctx->d_tipmask, tip_sites));
      CUC(cudaMemcpy(ctx->d_tipmask, m8.data(), tip_sites, cudaMemcpyHostToDevice));
    }
  }
  {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    cudaDeviceSynchronize();
    CUC(cudaMemcpyToSymbol(c_model, &ctx->hm, sizeof(DevModel)));
    g_const_owner[device] = ctx;
    // table of the lookup build's logarithm (kernels_blo_site.cuh: table_log), the same for every context
    double2 logtab[128];
    for (int i = 0; i < 128; ++i)
    {
      const double c = 1.0 / (1.0 + ((double) i + 0.5) * (1.0 / 128.0));
      logtab[i] = make_double2(c, -std::log(c));
    }
    CUC(cudaMemcpyToSymbol(g_logtab, logtab, sizeof logtab));
  }
#undef CUC
  *out = ctx;
  return EPA_OK;
}

// ==============================================================================================
//  reference CLVs
// ==============================================================================================
static int launch_pmatrices(epa_ctx * ctx, const double * d_lengths, double * d_out, uint32_t count)
{
  if (ctx->S == 4) pmatrix_kernel<4><<<count, 128, 0, ctx->stream>>>(ctx->d_model, d_lengths, d_out);
  else pmatrix_kernel<20><<<count, 128, 0, ctx->stream>>>(ctx->d_model, d_lengths, d_out);
  LAUNCHED(ctx);
  return EPA_OK;
}

extern "C" int epa_compute_clvs(epa_ctx * ctx, const epa_clv_op * ops, uint32_t n_ops)
{
  if (!ctx) return EPA_ERR_ARG;
  if (!ops && n_ops) return fail(ctx, EPA_ERR_ARG, "null ops");
  if (int rc = set_device(ctx)) return rc;
  const uint32_t N = ctx->n_nodes, T = ctx->n_tips;
  // dependency depth of every op
  std::vector<int> producer(N, -1);
  for (uint32_t i = 0; i < n_ops; ++i)
  {
    const epa_clv_op & o = ops[i];
    if (o.parent < T || o.parent >= N || o.left >= N || o.right >= N)
      return fail(ctx, EPA_ERR_ARG, "op %u references node out of range", i);
    if (producer[o.parent] != -1) return fail(ctx, EPA_ERR_ARG, "CLV slot %u computed twice", o.parent - T);
    producer[o.parent] = (int) i;
  }
  std::vector<int> level(N, -1);
  for (uint32_t i = 0; i < N; ++i)
    if (ctx->slot_filled[i] && producer[i] == -1) level[i] = 0;
  // iterative DFS
  std::vector<uint32_t> stack;
  for (uint32_t i = 0; i < n_ops; ++i)
  {
    stack.push_back(ops[i].parent);
    while (!stack.empty())
    {
      const uint32_t node = stack.back();
      if (level[node] >= 0) { stack.pop_back(); continue; }
      if (producer[node] < 0) return fail(ctx, EPA_ERR_STATE, "CLV slot %u is needed but never computed or uploaded", node - T);
      const epa_clv_op & o = ops[producer[node]];
      if (level[node] == -2 && (level[o.left] < 0 || level[o.right] < 0))
        return fail(ctx, EPA_ERR_ARG, "cyclic CLV dependencies at slot %u", node - T);
      if (level[o.left] >= 0 && level[o.right] >= 0)
      {
        level[node] = 1 + std::max(level[o.left], level[o.right]);
        stack.pop_back();
      }
      else
      {
        level[node] = -2;      // visiting
        if (level[o.left] == -2 || level[o.right] == -2)
          return fail(ctx, EPA_ERR_ARG, "cyclic CLV dependencies at slot %u", node - T);
        if (level[o.left] < 0) stack.push_back(o.left);
        if (level[o.right] < 0) stack.push_back(o.right);
      }
    }
  }
  int max_level = 0;
  for (uint32_t i = 0; i < n_ops; ++i) max_level = std::max(max_level, level[ops[i].parent]);
  std::vector<std::vector<uint32_t>> by_level(max_level + 1);
  for (uint32_t i = 0; i < n_ops; ++i) by_level[level[ops[i].parent]].push_back(i);

  std::vector<ClvOpDev> hops;
  std::vector<double> lengths;
  hops.reserve(n_ops); lengths.reserve(2 * (size_t) n_ops);
  std::vector<std::pair<uint32_t, uint32_t>> ranges;
  for (int l = 1; l <= max_level; ++l)
  {
    const uint32_t start = (uint32_t) hops.size();
    for (uint32_t i : by_level[l])
    {
      const epa_clv_op & o = ops[i];
      ClvOpDev d;
      d.parent = o.parent; d.left = o.left; d.right = o.right;
      d.tip_tip = (o.left < T && o.right < T) ? 1u : ((o.left < T || o.right < T) ? 2u : 0u);
      d.lmat = (uint32_t) lengths.size(); lengths.push_back(o.left_length);
      d.rmat = (uint32_t) lengths.size(); lengths.push_back(o.right_length);
      hops.push_back(d);
    }
    ranges.emplace_back(start, (uint32_t) hops.size() - start);
  }
  ctx->clvT_ready = false;
  if (hops.empty()) { ctx->clvs_ready = true; return EPA_OK; }

  const size_t pm = (size_t) ctx->R * ctx->S * ctx->S;
  ClvOpDev * d_ops = nullptr; double * d_len = nullptr; double * d_pm = nullptr;
  CU(dev_alloc(&d_ops, hops.size() * sizeof(ClvOpDev)));
  CU(dev_alloc(&d_len, lengths.size() * sizeof(double)));
  CU(dev_alloc(&d_pm, lengths.size() * pm * sizeof(double)));
  CU(cudaMemcpyAsync(d_ops, hops.data(), hops.size() * sizeof(ClvOpDev), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_len, lengths.data(), lengths.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = launch_pmatrices(ctx, d_len, d_pm, (uint32_t) lengths.size())) return rc;
  const size_t smem = 2 * pm * sizeof(double);
  if (ctx->tree.sr > 1 && ctx->S != 4)
  {
    // tip-inner updates of the generic per-rate path add their counts atomically (kernels_clv.cuh)
    const size_t sn = (size_t) ctx->n * ctx->tree.sr;
    for (auto & o : hops)
      if (o.tip_tip == 2) CU(cudaMemsetAsync(ctx->tree.scaler + (size_t) o.parent * sn, 0, sn * sizeof(uint32_t), ctx->stream));
  }
  for (auto & rg : ranges)
  {
    if (!rg.second) continue;
    dim3 grid(rg.second, (ctx->n + 127) / 128);
    EPA_DISPATCH_SR(ctx, (clv_update_kernel<S_, R_><<<grid, 128, smem, ctx->stream>>>(ctx->tree, ctx->n, d_ops + rg.first, d_pm)));
    LAUNCHED(ctx);
  }
  CU(cudaStreamSynchronize(ctx->stream));
  dev_free(d_ops); dev_free(d_len); dev_free(d_pm);
  for (auto & o : hops) ctx->slot_filled[o.parent] = 1;
  ctx->clvs_ready = true;
  ctx->lookup_ready = false;
  return EPA_OK;
}

extern "C" int epa_upload_clvs(epa_ctx * ctx, const epa_host_clv * clvs, uint32_t n_clvs)
{
  if (!ctx) return EPA_ERR_ARG;
  if (!clvs && n_clvs) return fail(ctx, EPA_ERR_ARG, "null clvs");
  if (int rc = set_device(ctx)) return rc;
  for (uint32_t i = 0; i < n_clvs; ++i)
  {
    const epa_host_clv & c = clvs[i];
    if (c.slot < ctx->n_tips || c.slot >= ctx->n_nodes || !c.clv)
      return fail(ctx, EPA_ERR_ARG, "clv %u: invalid slot %u", i, c.slot);
    CU(cudaMemcpyAsync(ctx->tree.clv + c.slot * ctx->tree.clv_stride, c.clv, ctx->tree.clv_stride * sizeof(double),
                       cudaMemcpyHostToDevice, ctx->stream));
    const size_t sn = (size_t) ctx->n * ctx->tree.sr;      // scaler entries per node
    if (c.scaler)
      CU(cudaMemcpyAsync(ctx->tree.scaler + (size_t) c.slot * sn, c.scaler, sn * sizeof(uint32_t),
                         cudaMemcpyHostToDevice, ctx->stream));
    else
      CU(cudaMemsetAsync(ctx->tree.scaler + (size_t) c.slot * sn, 0, sn * sizeof(uint32_t), ctx->stream));
    ctx->slot_filled[c.slot] = 1;
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->clvs_ready = true;
  ctx->lookup_ready = false;
  ctx->clvT_ready = false;
  return EPA_OK;
}

extern "C" int epa_get_clv(epa_ctx * ctx, uint32_t slot, double * clv, uint32_t * scaler)
{
  if (!ctx) return EPA_ERR_ARG;
  if (slot >= ctx->n_nodes) return fail(ctx, EPA_ERR_ARG, "node %u out of range", slot);
  if (int rc = set_device(ctx)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  if (clv) CU(cudaMemcpy(clv, ctx->tree.clv + slot * ctx->tree.clv_stride, ctx->tree.clv_stride * sizeof(double), cudaMemcpyDeviceToHost));
  const size_t sn = (size_t) ctx->n * ctx->tree.sr;
  if (scaler) CU(cudaMemcpy(scaler, ctx->tree.scaler + (size_t) slot * sn, sn * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return EPA_OK;
}

static int check_edges_ready(epa_ctx * ctx)
{
  for (const EdgeDev & e : ctx->h_edges)
    if (!ctx->slot_filled[e.distal] || !ctx->slot_filled[e.proximal])
      return fail(ctx, EPA_ERR_STATE, "edge CLVs have not been computed or uploaded");
  return EPA_OK;
}

extern "C" int epa_edge_loglikelihood(epa_ctx * ctx, uint32_t edge, double * logl)
{
  if (!ctx || !logl) return EPA_ERR_ARG;
  if (edge >= ctx->n_edges) return fail(ctx, EPA_ERR_ARG, "edge %u out of range", edge);
  if (int rc = set_device(ctx)) return rc;
  if (int rc = check_edges_ready(ctx)) return rc;
  const size_t pm = (size_t) ctx->R * ctx->S * ctx->S;
  const int blocks = (ctx->n + 127) / 128;
  CU(ctx->tmp.ensure((pm + 1 + blocks) * sizeof(double)));
  double * d_len = ctx->tmp.as<double>(), * d_pm = d_len + 1, * d_part = d_pm + pm;
  const EdgeDev e = ctx->h_edges[edge];
  CU(cudaMemcpyAsync(d_len, &e.length, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = launch_pmatrices(ctx, d_len, d_pm, 1)) return rc;
  EPA_DISPATCH_SR(ctx, (edge_logl_kernel<S_, R_><<<blocks, 128, pm * sizeof(double), ctx->stream>>>(ctx->d_model, ctx->tree, ctx->n, e, d_pm, d_part)));
  LAUNCHED(ctx);
  std::vector<double> part(blocks);
  CU(cudaMemcpyAsync(part.data(), d_part, blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  double s = 0;
  for (double v : part) s += v;
  *logl = s;
  return EPA_OK;
}

// ==============================================================================================
//  lookup tables
// ==============================================================================================
namespace { int ensure_clvT(epa_ctx * ctx); }

extern "C" int epa_build_lookup(epa_ctx * ctx)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  if (int rc = check_edges_ready(ctx)) return rc;
  const int S = ctx->S, R = ctx->R, K = ctx->K, n = ctx->n;
  if (S == 4 && (R == 1 || R == 2 || R == 4))
  {
    // the DNA lookup kernel reads the site-blocked CLV copy (uses ctx->tmp: must come first)
    if (int rc = bind_constants(ctx)) return rc;
    if (int rc = ensure_clvT(ctx)) return rc;
  }
  const size_t pm = (size_t) R * S * S;
  const uint32_t B = ctx->n_edges;
  const size_t lookup_doubles = (size_t) B * ctx->n_pad * K;
  if (!ctx->d_lookup) CU(dev_alloc(&ctx->d_lookup, lookup_doubles * sizeof(double)));
  if (ctx->n_pad != n)      // pad rows must read as zero; every real row is written by the kernel
    CU(cudaMemsetAsync(ctx->d_lookup, 0, lookup_doubles * sizeof(double), ctx->stream));

  std::vector<double> lengths(B + 1);
  for (uint32_t i = 0; i < B; ++i) lengths[i] = ctx->h_edges[i].length / 2.0;
  lengths[B] = EPA_DEFAULT_PENDANT;
  const size_t coltab = (size_t) R * K * S;
  CU(ctx->tmp.ensure(((B + 1) * (pm + 1) + coltab) * sizeof(double)));
  double * d_len = ctx->tmp.as<double>(), * d_pm = d_len + (B + 1), * d_col = d_pm + (B + 1) * pm;
  CU(cudaMemcpyAsync(d_len, lengths.data(), (B + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (int rc = launch_pmatrices(ctx, d_len, d_pm, B + 1)) return rc;
  if (S == 4) lookup_coltable_kernel<4><<<1, 256, 0, ctx->stream>>>(ctx->d_model, d_pm + (size_t) B * pm, d_col);
  else lookup_coltable_kernel<20><<<1, 256, 0, ctx->stream>>>(ctx->d_model, d_pm + (size_t) B * pm, d_col);
  LAUNCHED(ctx);
  if (S == 4 && (R == 1 || R == 2 || R == 4) && (ctx->tree.sr > 1 || !ctx->sw.old_lookup))
  {
    // lane = site kernel over the site-blocked CLV copy; its column table goes to constant memory
    // as [c][r][i] (the mutex covers copy + launch: the symbol is shared by the contexts of a process)
    const size_t t_stride = clvt_node_stride(n, R);
    dim3 grid(B, (n + 128 * LOOKUP_ITER - 1) / (128 * LOOKUP_ITER));
    std::vector<double> hcol((size_t) R * K * 4), hperm((size_t) K * R * 4);
    CU(cudaMemcpyAsync(hcol.data(), d_col, hcol.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < K; ++c)
        for (int i = 0; i < 4; ++i) hperm[((size_t) c * R + r) * 4 + i] = hcol[((size_t) r * K + c) * 4 + i];
    {
      std::lock_guard<std::mutex> lock(g_const_mutex);
      CU(cudaMemcpyToSymbol(c_coltab, hperm.data(), hperm.size() * sizeof(double)));
      CU(cudaEventRecord(ctx->ev[0], ctx->stream));
      const uint8_t * tipmask = ctx->sw.lookup_tip_clv ? nullptr : ctx->d_tipmask;
      switch (R)
      {
        case 1: lookup_build_site_kernel<1><<<grid, 128, 4 * lookup_warp_doubles<1>() * sizeof(double), ctx->stream>>>(ctx->d_model, ctx->d_clvT, t_stride, ctx->tree.scaler, (int) ctx->tree.sr, ctx->tree.inv, n, ctx->n_pad, ctx->d_edges, d_pm, ctx->d_lookup, tipmask, ctx->n_tips); break;
        case 2: lookup_build_site_kernel<2><<<grid, 128, 4 * lookup_warp_doubles<2>() * sizeof(double), ctx->stream>>>(ctx->d_model, ctx->d_clvT, t_stride, ctx->tree.scaler, (int) ctx->tree.sr, ctx->tree.inv, n, ctx->n_pad, ctx->d_edges, d_pm, ctx->d_lookup, tipmask, ctx->n_tips); break;
        default: lookup_build_site_kernel<4><<<grid, 128, 4 * lookup_warp_doubles<4>() * sizeof(double), ctx->stream>>>(ctx->d_model, ctx->d_clvT, t_stride, ctx->tree.scaler, (int) ctx->tree.sr, ctx->tree.inv, n, ctx->n_pad, ctx->d_edges, d_pm, ctx->d_lookup, tipmask, ctx->n_tips); break;
      }
      LAUNCHED(ctx);
      CU(cudaStreamSynchronize(ctx->stream));
    }
  }
  else if (S == 4 && R == 4)
  {
    dim3 grid(B, (n + LOOKUP_DNA_SITES_PER_BLOCK - 1) / LOOKUP_DNA_SITES_PER_BLOCK);
    lookup_build_dna_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_model, ctx->tree, n, ctx->n_pad, ctx->d_edges, d_pm, d_col, ctx->d_lookup);
    LAUNCHED(ctx);
  }
  else
  {
    const size_t smem = (pm + coltab) * sizeof(double);
    dim3 grid(B, (n + 127) / 128);
    uint8_t * d_tiflags = nullptr;
    if (S != 4 && ctx->tree.sr > 1)
    {
      // amino acids under per-rate scalers: whole-site rescalings of the tip edges' inner CLVs (kernels_clv.cuh)
      CU(dev_alloc(&d_tiflags, (size_t) B * n));
      EPA_DISPATCH_SR(ctx, (lookup_ti_flags_kernel<S_, R_><<<grid, 128, pm * sizeof(double), ctx->stream>>>(ctx->tree, n, ctx->d_edges, d_pm, d_tiflags)));
      LAUNCHED(ctx);
    }
    EPA_DISPATCH_SR(ctx, (lookup_build_kernel<S_, R_><<<grid, 128, smem, ctx->stream>>>(ctx->d_model, ctx->tree, n, ctx->n_pad, K, ctx->d_edges, d_pm, d_col, ctx->d_lookup, d_tiflags)));
    LAUNCHED(ctx);
    if (d_tiflags)
    {
      CU(cudaStreamSynchronize(ctx->stream));
      dev_free(d_tiflags);
    }
  }
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (S == 4)
  {
    ctx->pairtab_ready = false;       // pair-sum tables of the fallback kernel are built on first use
    // fixed-point digit table + prefix sums of the tensor-core preplacement (kernels_preplace_mma.cuh)
    const int kc_total = mma_kc_total(n);
    const uint32_t n_eb = (B + MMA_EB - 1) / MMA_EB;
    const size_t btab_bytes = (size_t) n_eb * kc_total * MMA_B_CHUNK_BYTES;
    if (!ctx->d_btab) CU(dev_alloc(&ctx->d_btab, btab_bytes));
    const uint32_t e_pad = n_eb * MMA_EB;
    if (!ctx->d_pn) CU(dev_alloc(&ctx->d_pn, (size_t) e_pad * (n + 1) * sizeof(double)));
    CU(cudaMemsetAsync(ctx->d_pn, 0, (size_t) e_pad * (n + 1) * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_btab, 0, btab_bytes, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_flags + 6, 0, sizeof(int), ctx->stream));
    const size_t tthreads = (size_t) B * ((n + 3) / 4);
    mma_table_kernel<<<(unsigned) ((tthreads + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_lookup, n, ctx->n_pad, B, kc_total, ctx->d_btab, ctx->d_flags + 6);
    LAUNCHED(ctx);
    mma_prefix_kernel<<<(B + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_lookup, n, ctx->n_pad, B, e_pad, ctx->d_pn);
    LAUNCHED(ctx);
    int bad = 1;
    CU(cudaMemcpyAsync(&bad, ctx->d_flags + 6, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->mma_ok = (bad == 0) && !ctx->sw.no_mma;
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->lookup_ready = true;
  return EPA_OK;
}

extern "C" int epa_last_lookup_ms(epa_ctx * ctx, float * ms)
{
  if (!ctx || !ms) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  CU(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
  return EPA_OK;
}

extern "C" int epa_get_lookup(epa_ctx * ctx, uint32_t edge, double * out)
{
  if (!ctx || !out) return EPA_ERR_ARG;
  if (!ctx->lookup_ready) return fail(ctx, EPA_ERR_STATE, "epa_build_lookup has not run");
  if (edge >= ctx->n_edges) return fail(ctx, EPA_ERR_ARG, "edge %u out of range", edge);
  if (int rc = set_device(ctx)) return rc;
  const int K = ctx->K, n = ctx->n;
  std::vector<double> raw((size_t) n * K);
  CU(cudaMemcpy(raw.data(), ctx->d_lookup + (size_t) edge * ctx->n_pad * K, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
  // reference column order: NT_MAP / AA_MAP (src/util/maps.hpp:9-28)
  const char * map = ctx->S == 4 ? "-TGKCYSBAWRDMHVN" : "ACDEFGHIKLMNPQRSTVWY-XBZ";
  const int KR = (int) strlen(map);
  for (int c = 0; c < KR; ++c)
  {
    int code = ctx->hm.ascii2code[(unsigned char) map[c]];
    // the reference's own X column is a genuine fully-ambiguous column (only the char->column
    // map redirects X to N), so report the fully ambiguous column for it
    if (ctx->S == 20 && map[c] == 'X') code = 23;
    const int col = (ctx->S == 20 && map[c] == 'X') ? 23 : ctx->hm.code2col[code];
    for (int s = 0; s < n; ++s) out[(size_t) s * KR + c] = raw[(size_t) s * K + col];
  }
  return EPA_OK;
}

// ==============================================================================================
//  chunk pipeline
// ==============================================================================================
static int read_flags(epa_ctx * ctx, int flags[16])
{
  CU(cudaMemcpyAsync(flags, ctx->d_flags, 16 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return EPA_OK;
}

static int ensure_copy_stream(epa_ctx * ctx)
{
  if (ctx->copy_stream) return EPA_OK;
  CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_collect, cudaEventDisableTiming));
  for (auto & e : ctx->ev_d2h) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return EPA_OK;
}

// exclusive scan on the context's stream: one block for short arrays, three launches for long ones
static int launch_exclusive_scan(epa_ctx * ctx, const uint32_t * in, uint32_t * out, uint32_t count, uint64_t * total)
{
  const uint32_t nb = (count + SCAN_ITEMS - 1) / SCAN_ITEMS;
  if (count <= 2 * SCAN_ITEMS || nb > 65536)
  {
    exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(in, out, count, total);
    LAUNCHED(ctx);
    return EPA_OK;
  }
  CU(ctx->scan_sums.ensure(nb * sizeof(uint32_t)));
  uint32_t * sums = ctx->scan_sums.as<uint32_t>();
  scan_block_sums_kernel<<<nb, 256, 0, ctx->stream>>>(in, count, sums);
  LAUNCHED(ctx);
  exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(sums, sums, nb, total);
  LAUNCHED(ctx);
  scan_apply_kernel<<<nb, 256, 0, ctx->stream>>>(in, out, count, sums);
  LAUNCHED(ctx);
  return EPA_OK;
}

extern "C" int epa_upload_queries(epa_ctx * ctx, const char * seqs, uint32_t n_queries, int premasking)
{
  if (!ctx) return EPA_ERR_ARG;
  if (!seqs && n_queries) return fail(ctx, EPA_ERR_ARG, "null seqs");
  if (int rc = set_device(ctx)) return rc;
  ctx->stage = ST_NONE;
  ctx->nq = n_queries;
  if (n_queries == 0) { ctx->stage = ST_QUERIES; ctx->max_span = 0; return EPA_OK; }
  const size_t bytes = (size_t) n_queries * ctx->n;
  CU(ctx->codes.ensure(bytes));
  CU(ctx->begin.ensure(n_queries * sizeof(int)));
  CU(ctx->span.ensure(n_queries * sizeof(int)));
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (ctx->staged_ptr == seqs && ctx->staged_n == n_queries)
  {
    // this chunk was prefetched while the previous one was being placed
    std::swap(ctx->raw, ctx->raw_next);
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
  }
  else
  {
    if (ctx->staged_ptr) CU(cudaStreamSynchronize(ctx->copy_stream));     // stale prefetch: let it land first
    CU(ctx->raw.ensure(bytes));
    CU(cudaMemcpyAsync(ctx->raw.p, seqs, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->staged_ptr = nullptr; ctx->staged_n = 0;
  const int rc = epa_encode_queries_dev(ctx, nullptr, n_queries, premasking);
  if (rc == EPA_OK && ctx->hint_ptr && ctx->hint_n)
  {
    // start the copy of the announced next chunk; it overlaps the placement of this one
    if (int rc2 = ensure_copy_stream(ctx)) return rc2;
    const size_t nbytes = (size_t) ctx->hint_n * ctx->n;
    CU(ctx->raw_next.ensure(nbytes));
    CU(cudaMemcpyAsync(ctx->raw_next.p, ctx->hint_ptr, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    ctx->staged_ptr = ctx->hint_ptr; ctx->staged_n = ctx->hint_n;
  }
  ctx->hint_ptr = nullptr; ctx->hint_n = 0;
  return rc;
}

extern "C" int epa_set_deferred_results(epa_ctx * ctx, int on)
{
  if (!ctx) return EPA_ERR_ARG;
  ctx->defer_results = on != 0;
  return EPA_OK;
}

extern "C" int epa_wait_results(epa_ctx * ctx)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
  return EPA_OK;
}

// Deferred results: blocks until the records of the chunk BEFORE the most recent epa_collect /
// epa_place_chunk have landed in host memory (that copy overlapped the most recent chunk's kernels).
extern "C" int epa_wait_older_results(epa_ctx * ctx)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  if (ctx->copy_stream) CU(cudaEventSynchronize(ctx->ev_d2h[ctx->out_flip]));
  return EPA_OK;
}

extern "C" int epa_hint_next_chunk(epa_ctx * ctx, const char * next_seqs, uint32_t next_n_queries)
{
  if (!ctx) return EPA_ERR_ARG;
  ctx->hint_ptr = next_seqs; ctx->hint_n = next_seqs ? next_n_queries : 0;
  return EPA_OK;
}

// Device-resident variant: `seqs_dev` already lives in HBM (NULL = the context's own staging
// buffer filled by epa_upload_queries).
extern "C" int epa_encode_queries_dev(epa_ctx * ctx, const char * seqs_dev, uint32_t n_queries, int premasking)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  ctx->stage = ST_NONE;
  ctx->nq = n_queries;
  if (n_queries == 0) { ctx->stage = ST_QUERIES; ctx->max_span = 0; return EPA_OK; }
  const size_t bytes = (size_t) n_queries * ctx->n;
  if (seqs_dev)
  {
    CU(ctx->codes.ensure(bytes));
    CU(ctx->begin.ensure(n_queries * sizeof(int)));
    CU(ctx->span.ensure(n_queries * sizeof(int)));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  }
  const uint8_t * src = seqs_dev ? reinterpret_cast<const uint8_t *>(seqs_dev) : ctx->raw.as<uint8_t>();
  CU(cudaMemsetAsync(ctx->d_flags, 0, 16 * sizeof(int), ctx->stream));
  const unsigned blocks = (n_queries + 7) / 8;
  CU(ctx->sortkey.ensure(n_queries * sizeof(int)));
  CU(ctx->amb.ensure(n_queries * sizeof(uint8_t)));
  // queries with a few other ambiguity codes (R, Y, K, M, ...) still take the tensor-core preplacement: those sites
  // score as fully ambiguous there and preplace_ambig_fix_kernel adds the difference afterwards
  const int amb_cap = (ctx->mma_ok && !ctx->sw.no_mma && !ctx->sw.fused_select) ? AMBIG_CAP : 0;
  // 64-bit accesses when every row of both buffers is 8-byte aligned
  if (ctx->n % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0 && (reinterpret_cast<uintptr_t>(ctx->codes.as<uint8_t>()) & 7) == 0)
    encode_queries_kernel<8><<<blocks, 256, 0, ctx->stream>>>(ctx->d_model, src, n_queries, ctx->n, premasking ? 1 : 0,
                                                         ctx->codes.as<uint8_t>(), ctx->begin.as<int>(),
                                                         ctx->span.as<int>(), ctx->sortkey.as<int>(), ctx->d_flags, amb_cap, ctx->amb.as<uint8_t>());
  else
    encode_queries_kernel<1><<<blocks, 256, 0, ctx->stream>>>(ctx->d_model, src, n_queries, ctx->n, premasking ? 1 : 0,
                                                         ctx->codes.as<uint8_t>(), ctx->begin.as<int>(),
                                                         ctx->span.as<int>(), ctx->sortkey.as<int>(), ctx->d_flags, amb_cap, ctx->amb.as<uint8_t>());
  LAUNCHED(ctx);
  // counting sort of the queries by (class, window start): simple queries first. The tiles of the
  // preplacement kernels and the all-pairs work order of the thorough kernel follow this order.
  {
    const uint32_t nq = n_queries;
    const uint32_t nkeys = 2u * (uint32_t) (ctx->n + 1);
    CU(ctx->perm.ensure(nq * sizeof(uint32_t)));
    CU(ctx->hist.ensure((size_t) (nkeys + 1) * sizeof(uint32_t)));
    CU(cudaMemsetAsync(ctx->hist.p, 0, (size_t) (nkeys + 1) * sizeof(uint32_t), ctx->stream));
    histogram_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->sortkey.as<int>(), nq, ctx->hist.as<uint32_t>());
    LAUNCHED(ctx);
    if (int rc = launch_exclusive_scan(ctx, ctx->hist.as<uint32_t>(), ctx->hist.as<uint32_t>(), nkeys, nullptr)) return rc;
    scatter_by_key_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->sortkey.as<int>(), nq, ctx->hist.as<uint32_t>(), ctx->perm.as<uint32_t>());
    LAUNCHED(ctx);
  }
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  int flags[16];
  if (int rc = read_flags(ctx, flags)) return rc;
  if (flags[0] == 1) return fail(ctx, EPA_ERR_QUERY, "query %d contains a character that is not valid for this data type", flags[1] - 1);
  if (flags[0] == 2) return fail(ctx, EPA_ERR_QUERY, "query %d consists entirely of gaps", flags[1] - 1);
  ctx->max_span = flags[3];
  ctx->n_simple = (uint32_t) flags[4];
  ctx->n_ambig = (uint32_t) flags[8];
  ctx->stage = ST_QUERIES;
  return EPA_OK;
}

namespace {
// generic kernel over perm[first, first + count)
template <int K>
int launch_preplace(epa_ctx * ctx, uint32_t first, uint32_t count, const int2 * range, int maxw)
{
  constexpr int TQ = kPreplaceTQ, NS = 2;
  const uint32_t n_tiles = (count + TQ - 1) / TQ;
  const size_t per_site = (size_t) NS * K * 8 + TQ;              // stage bytes + code bytes per site
  const size_t budget = ctx->smem_optin - 4096;
  int wc = maxw;
  if ((size_t) wc * per_site > budget) wc = (int) (budget / per_site) & ~3;
  const size_t smem = (size_t) wc * per_site + NS * 8;
  CU(cudaFuncSetAttribute(preplace_kernel<K, TQ, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  preplace_kernel<K, TQ, NS><<<n_tiles, TQ, smem, ctx->stream>>>(
      ctx->d_model, ctx->d_lookup, ctx->n, ctx->n_pad, ctx->n_edges, ctx->codes.as<uint8_t>(), ctx->begin.as<int>(),
      ctx->span.as<int>(), ctx->perm.as<uint32_t>() + first, count, range, wc, ctx->pre.as<double>(), ctx->pre_stride);
  LAUNCHED(ctx);
  return EPA_OK;
}

// DNA pair-table kernel over perm[0, count)
int launch_preplace_pair(epa_ctx * ctx, uint32_t count, const int2 * range, int maxw)
{
  constexpr int TQ = kPairTQ, NS = 2;
  if (!ctx->pairtab_ready)
  {
    const size_t pair_doubles = (size_t) ctx->n_edges * (ctx->n_pad / 2) * PAIR_ROW;
    if (!ctx->d_pairtab) CU(dev_alloc(&ctx->d_pairtab, pair_doubles * sizeof(double)));
    pairtab_build_kernel<<<(unsigned) ((pair_doubles + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_lookup, ctx->n_pad, ctx->n_edges, ctx->d_pairtab);
    LAUNCHED(ctx);
    ctx->pairtab_ready = true;
  }
  const uint32_t n_tiles = (count + TQ - 1) / TQ;
  // per 8 sites: 4 pair rows per stage + one index word per query
  const size_t per_word = (size_t) NS * 4 * PAIR_ROW * 8 + (size_t) TQ * 4;
  const size_t budget = ctx->smem_optin - 8192;
  int wc = maxw;                                                  // multiple of 8
  if ((size_t) (wc / 8) * per_word > budget) wc = (int) (budget / per_word) * 8;
  const size_t smem = (size_t) (wc / 8) * per_word + NS * 8;
  CU(cudaFuncSetAttribute(preplace_pair_kernel<TQ, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  preplace_pair_kernel<TQ, NS><<<n_tiles, TQ, smem, ctx->stream>>>(
      ctx->d_pairtab, ctx->n, ctx->n_pad, ctx->n_edges, ctx->codes.as<uint8_t>(), ctx->begin.as<int>(),
      ctx->span.as<int>(), ctx->perm.as<uint32_t>(), count, range, wc, ctx->pre.as<double>(), ctx->pre_stride);
  LAUNCHED(ctx);
  return EPA_OK;
}

// DNA tensor-core kernel over perm[0, count): tiles of MMA_TQ queries, persistent CTAs. fused = the
// epilogue keeps the selection summaries instead of writing scores (kernels_preplace_mma.cuh).
int launch_preplace_mma(epa_ctx * ctx, const uint32_t * perm, uint32_t count, const int2 * range, bool fused)
{
  PreMmaArgs a{};
  a.btab = ctx->d_btab; a.pn = ctx->d_pn; a.kc_total = mma_kc_total(ctx->n); a.n = ctx->n;
  a.n_edges = ctx->n_edges; a.n_eb = (ctx->n_edges + MMA_EB - 1) / MMA_EB;
  a.codes = ctx->codes.as<uint8_t>(); a.begin = ctx->begin.as<int>(); a.span = ctx->span.as<int>();
  a.perm = perm; a.nq = count; a.range = range;
  a.n_tiles = (count + MMA_TQ - 1) / MMA_TQ;
  a.pre = fused ? nullptr : ctx->pre.as<double>(); a.pre_stride = ctx->pre_stride; a.qmax = fused ? nullptr : ctx->qmax.as<double>();
  a.summary = fused ? ctx->summary.as<RowSummary>() : nullptr;
  const unsigned grid = (unsigned) std::min<uint32_t>((uint32_t) ctx->sm_count, a.n_tiles);
  if (fused)
  {
    CU(cudaFuncSetAttribute(preplace_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) MMA_SMEM_BYTES));
    preplace_mma_kernel<true><<<grid, MMA_THREADS, MMA_SMEM_BYTES, ctx->stream>>>(a);
  }
  else
  {
    CU(cudaFuncSetAttribute(preplace_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) MMA_SMEM_BYTES));
    preplace_mma_kernel<false><<<grid, MMA_THREADS, MMA_SMEM_BYTES, ctx->stream>>>(a);
  }
  LAUNCHED(ctx);
  return EPA_OK;
}

// the [query][edge] score matrix and the row maxima (not needed while every query takes the fused kernel)
int ensure_prescores(epa_ctx * ctx)
{
  const uint32_t nq = ctx->nq;
  const bool fresh = (size_t) nq * ctx->pre_stride * sizeof(double) > ctx->pre.cap || nq * sizeof(double) > ctx->qmax.cap;
  CU(ctx->pre.ensure((size_t) nq * ctx->pre_stride * sizeof(double)));
  CU(ctx->qmax.ensure(nq * sizeof(double)));
  (void) fresh;
  return EPA_OK;
}
}  // namespace

extern "C" int epa_preplace(epa_ctx * ctx)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  if (!ctx->lookup_ready) return fail(ctx, EPA_ERR_STATE, "epa_build_lookup has not run");
  if (ctx->stage < ST_QUERIES) return fail(ctx, EPA_ERR_STATE, "no queries uploaded");
  const uint32_t nq = ctx->nq;
  ctx->pre_stride = (ctx->n_edges + 3u) & ~3u;
  ctx->fused = false; ctx->fused_n = 0;
  if (nq == 0) { ctx->stage = ST_PREPLACED; return EPA_OK; }
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  // simple DNA queries (sorted first) take the tensor-core kernel (the pair-table kernel when the
  // digit table could not be built); the rest take the per-site kernel
  const uint32_t nA = ctx->S == 4 ? ctx->n_simple : 0u, nB = nq - nA;
  const uint32_t tilesM = (nA + MMA_TQ - 1) / MMA_TQ;
  const uint32_t tilesA = (nA + kPairTQ - 1) / kPairTQ, tilesB = (nB + kPreplaceTQ - 1) / kPreplaceTQ;
  CU(ctx->range.ensure((size_t) (std::max(tilesA, tilesM) + tilesB) * sizeof(int2)));
  int2 * rangeB = ctx->range.as<int2>() + std::max(tilesA, tilesM);
  const bool use_mma = ctx->mma_ok && nA > 0;
  // fused selection: dynamic heuristic announced, single-scan selection exact (see epa_select), tensor-core path
  const bool fuse = use_mma && ctx->sel_hint && ctx->sel_hint_mode == 0 && ctx->sw.fused_select &&
                    (1.0 - ctx->sel_hint_thresh) > 4.0 * (double) ctx->n_edges * std::exp(-SEL_CUT);
  ctx->sel_hint = false;                       // a hint covers one epa_preplace
  if (!fuse || nB)
  {
    if (int rc = ensure_prescores(ctx)) return rc;
    CU(cudaMemsetAsync(ctx->qmax.p, 0xff, nq * sizeof(double), ctx->stream));      // NaN = no row maximum recorded
  }
  if (fuse) CU(ctx->summary.ensure((size_t) nq * 2 * sizeof(RowSummary)));
  int flags[16];
  if (use_mma)
  {
    tile_range_kernel<<<(tilesM + 7) / 8, 256, 0, ctx->stream>>>(ctx->perm.as<uint32_t>(), ctx->begin.as<int>(), ctx->span.as<int>(),
                                                              nA, MMA_TQ, tilesM, 4, ctx->range.as<int2>(), ctx->d_flags + 2);
    LAUNCHED(ctx);
  }
  CU(cudaMemsetAsync(ctx->d_flags + 5, 0, sizeof(int), ctx->stream));
  if (nA && !use_mma)
  {
    CU(cudaMemsetAsync(ctx->d_flags + 2, 0, sizeof(int), ctx->stream));
    tile_range_kernel<<<(tilesA + 7) / 8, 256, 0, ctx->stream>>>(ctx->perm.as<uint32_t>(), ctx->begin.as<int>(), ctx->span.as<int>(),
                                                              nA, kPairTQ, tilesA, 8, ctx->range.as<int2>(), ctx->d_flags + 2);
    LAUNCHED(ctx);
  }
  if (nB)
  {
    tile_range_kernel<<<(tilesB + 7) / 8, 256, 0, ctx->stream>>>(ctx->perm.as<uint32_t>() + nA, ctx->begin.as<int>(), ctx->span.as<int>(),
                                                              nB, kPreplaceTQ, tilesB, 4, rangeB, ctx->d_flags + 5);
    LAUNCHED(ctx);
  }
  if ((nA && !use_mma) || nB)
    if (int rc = read_flags(ctx, flags)) return rc;
  if (nA && use_mma)
  {
    if (int rc = launch_preplace_mma(ctx, ctx->perm.as<uint32_t>(), nA, ctx->range.as<int2>(), fuse)) return rc;
    ctx->fused = fuse; ctx->fused_n = fuse ? nA : 0;
    if (ctx->n_ambig && !fuse)
    {
      preplace_ambig_fix_kernel<<<(nA + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_lookup, ctx->n_pad, (int) ctx->n_edges, ctx->codes.as<uint8_t>(),
                                                                     ctx->n, ctx->begin.as<int>(), ctx->span.as<int>(), ctx->amb.as<uint8_t>(),
                                                                     ctx->perm.as<uint32_t>(), nA, ctx->pre.as<double>(), ctx->pre_stride,
                                                                     ctx->qmax.as<double>());
      LAUNCHED(ctx);
    }
  }
  else if (nA)
    if (int rc = launch_preplace_pair(ctx, nA, ctx->range.as<int2>(), std::max(8, flags[2]))) return rc;
  if (nB)
  {
    const int maxw = std::max(4, flags[5]);
    const int rc = (ctx->K == 16) ? launch_preplace<16>(ctx, nA, nB, rangeB, maxw)
                                  : launch_preplace<26>(ctx, nA, nB, rangeB, maxw);
    if (rc) return rc;
  }
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->stage = ST_PREPLACED;
  return EPA_OK;
}

// Announces the options of the epa_select that will follow the next epa_preplace: with the dynamic heuristic
// the tensor-core kernel then selects in its epilogue and never writes the score matrix.
extern "C" int epa_hint_selection(epa_ctx * ctx, const epa_options * opts)
{
  if (!ctx) return EPA_ERR_ARG;
  ctx->sel_hint = opts != nullptr && opts->prescoring != 0;
  if (ctx->sel_hint) { ctx->sel_hint_mode = opts->heuristic; ctx->sel_hint_thresh = opts->prescoring_threshold; }
  return EPA_OK;
}

extern "C" int epa_get_prescores(epa_ctx * ctx, double * out)
{
  if (!ctx || !out) return EPA_ERR_ARG;
  if (ctx->stage < ST_PREPLACED) return fail(ctx, EPA_ERR_STATE, "epa_preplace has not run");
  if (ctx->fused) return fail(ctx, EPA_ERR_STATE, "the scores were not materialised: epa_preplace ran with a selection hint");
  if (int rc = set_device(ctx)) return rc;
  if (ctx->nq == 0) return EPA_OK;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy2D(out, (size_t) ctx->n_edges * sizeof(double), ctx->pre.p, ctx->pre_stride * sizeof(double),
                  (size_t) ctx->n_edges * sizeof(double), ctx->nq, cudaMemcpyDeviceToHost));
  return EPA_OK;
}

extern "C" int epa_select(epa_ctx * ctx, const epa_options * opts, uint64_t * n_pairs)
{
  if (!ctx || !opts) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  const uint32_t nq = ctx->nq, B = ctx->n_edges;
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (!opts->prescoring)
  {
    if (ctx->stage < ST_QUERIES) return fail(ctx, EPA_ERR_STATE, "no queries uploaded");
    if ((uint64_t) nq * B > 0xffffffffull) return fail(ctx, EPA_ERR_ARG, "chunk too large: %u queries x %u edges exceeds 2^32 pairs", nq, B);
    ctx->implicit_pairs = true;
    ctx->n_pairs = (uint64_t) nq * B;
  }
  else
  {
    if (ctx->stage < ST_PREPLACED) return fail(ctx, EPA_ERR_STATE, "epa_preplace has not run");
    if (opts->heuristic < 0 || opts->heuristic > 2) return fail(ctx, EPA_ERR_ARG, "unknown heuristic %d", opts->heuristic);
    if (opts->heuristic != 2 && !(opts->prescoring_threshold >= 0.0 && opts->prescoring_threshold <= 1.0))
      return fail(ctx, EPA_ERR_ARG, "prescoring threshold outside [0,1]");
    ctx->implicit_pairs = false;
    ctx->n_pairs = 0;
    if (nq)
    {
      CU(ctx->cnt.ensure(nq * sizeof(uint32_t)));
      CU(ctx->off.ensure(nq * sizeof(uint32_t)));
      CU(ctx->cutv.ensure(nq * sizeof(double)));
      CU(ctx->cuti.ensure(nq * sizeof(int)));
      CU(ctx->cand.ensure((size_t) nq * SEL_CAP * sizeof(uint32_t)));
      // single-scan selection is exact while the weight hidden below the cut cannot reach 1 - threshold
      const int fast_ok = (1.0 - opts->prescoring_threshold) > 4.0 * (double) B * std::exp(-SEL_CUT) ? 1 : 0;
      // work list order: edge-major, then window start in bins of kWindowBin sites
      const uint32_t nbins = (uint32_t) (ctx->n / kWindowBin) + 1;
      const size_t nkeys = (size_t) B * nbins;
      if (nkeys > 0x7fffffffull) return fail(ctx, EPA_ERR_ARG, "tree x alignment too large for the work-list sort");
      CU(ctx->edge_hist.ensure((nkeys + 1) * sizeof(uint32_t)));
      const unsigned blocks = (nq + 7) / 8;
      if (ctx->fused && (opts->heuristic != ctx->sel_hint_mode || opts->prescoring_threshold != ctx->sel_hint_thresh))
      {
        // other options than announced: materialise the scores after all
        ctx->sel_hint = false;
        if (int rc = epa_preplace(ctx)) return rc;
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
      }
      if (ctx->fused)
      {
        // 1. queries of the fused kernel: selection from the epilogue summaries
        const uint32_t nA = ctx->fused_n;
        CU(ctx->over_list.ensure((size_t) nA * sizeof(uint32_t)));
        uint32_t * d_nover = reinterpret_cast<uint32_t *>(ctx->d_flags + 6);
        CU(cudaMemsetAsync(d_nover, 0, sizeof(uint32_t), ctx->stream));
        select_finish_kernel<<<(nA + 255) / 256, 256, 0, ctx->stream>>>(ctx->summary.as<RowSummary>(), ctx->perm.as<uint32_t>(), nA,
                                                                        opts->prescoring_threshold, ctx->cnt.as<uint32_t>(),
                                                                        ctx->cand.as<uint32_t>(), ctx->over_list.as<uint32_t>(), d_nover);
        LAUNCHED(ctx);
        uint32_t n_over = 0;
        CU(cudaMemcpyAsync(&n_over, d_nover, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (getenv("EPA_B200_DEBUG_FUSED")) fprintf(stderr, "[fused] %u of %u queries left over\n", n_over, nA);
        // 2. the ones it left over (long candidate lists, multi-pass tiles): unfused kernels on that list
        if (n_over)
        {
          if (int rc = ensure_prescores(ctx)) return rc;
          const uint32_t tiles = (n_over + MMA_TQ - 1) / MMA_TQ;
          CU(ctx->range2.ensure((size_t) tiles * sizeof(int2)));
          CU(cudaMemsetAsync(ctx->d_flags + 7, 0, sizeof(int), ctx->stream));
          tile_range_kernel<<<(tiles + 7) / 8, 256, 0, ctx->stream>>>(ctx->over_list.as<uint32_t>(), ctx->begin.as<int>(), ctx->span.as<int>(),
                                                                    n_over, MMA_TQ, tiles, 4, ctx->range2.as<int2>(), ctx->d_flags + 7);
          LAUNCHED(ctx);
          if (int rc = launch_preplace_mma(ctx, ctx->over_list.as<uint32_t>(), n_over, ctx->range2.as<int2>(), false)) return rc;
          select_count_kernel<<<(n_over + 7) / 8, 256, 0, ctx->stream>>>(ctx->pre.as<double>(), ctx->pre_stride, (int) B, n_over,
                                                                         opts->heuristic, opts->prescoring_threshold, fast_ok,
                                                                         ctx->qmax.as<double>(), ctx->cnt.as<uint32_t>(),
                                                                         ctx->cutv.as<double>(), ctx->cuti.as<int>(), ctx->cand.as<uint32_t>(),
                                                                         ctx->over_list.as<uint32_t>());
          LAUNCHED(ctx);
        }
        // 3. queries that never took the tensor-core kernel (other ambiguity codes): their rows exist
        if (nq > nA)
        {
          select_count_kernel<<<(nq - nA + 7) / 8, 256, 0, ctx->stream>>>(ctx->pre.as<double>(), ctx->pre_stride, (int) B, nq - nA,
                                                                          opts->heuristic, opts->prescoring_threshold, fast_ok,
                                                                          ctx->qmax.as<double>(), ctx->cnt.as<uint32_t>(),
                                                                          ctx->cutv.as<double>(), ctx->cuti.as<int>(), ctx->cand.as<uint32_t>(),
                                                                          ctx->perm.as<uint32_t>() + nA);
          LAUNCHED(ctx);
        }
      }
      else
      {
        select_count_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->pre.as<double>(), ctx->pre_stride, (int) B, nq,
                                                             opts->heuristic, opts->prescoring_threshold, fast_ok,
                                                             ctx->qmax.as<double>(), ctx->cnt.as<uint32_t>(),
                                                             ctx->cutv.as<double>(), ctx->cuti.as<int>(), ctx->cand.as<uint32_t>());
        LAUNCHED(ctx);
      }
      if (int rc = launch_exclusive_scan(ctx, ctx->cnt.as<uint32_t>(), ctx->off.as<uint32_t>(), nq, ctx->d_total)) return rc;
      uint64_t total = 0;
      CU(cudaMemcpyAsync(&total, ctx->d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      if (total > 0xffffffffull) return fail(ctx, EPA_ERR_ARG, "chunk too large: %llu candidate pairs", (unsigned long long) total);
      ctx->n_pairs = total;
      CU(ctx->pair_q.ensure(total * sizeof(uint32_t)));
      CU(ctx->pair_e.ensure(total * sizeof(uint32_t)));
      CU(ctx->work.ensure(total * sizeof(uint32_t)));
      CU(cudaMemsetAsync(ctx->edge_hist.p, 0, (nkeys + 1) * sizeof(uint32_t), ctx->stream));
      select_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->pre.as<double>(), ctx->pre_stride, (int) B, nq,
                                                          ctx->off.as<uint32_t>(), ctx->cnt.as<uint32_t>(), ctx->cand.as<uint32_t>(),
                                                          ctx->cutv.as<double>(), ctx->cuti.as<int>(),
                                                          ctx->begin.as<int>(), kWindowBin, nbins,
                                                          ctx->pair_q.as<uint32_t>(), ctx->pair_e.as<uint32_t>(),
                                                          ctx->edge_hist.as<uint32_t>());
      LAUNCHED(ctx);
      if (int rc = launch_exclusive_scan(ctx, ctx->edge_hist.as<uint32_t>(), ctx->edge_hist.as<uint32_t>(), (uint32_t) nkeys, nullptr)) return rc;
      if (total)
      {
        work_scatter_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->pair_q.as<uint32_t>(), ctx->pair_e.as<uint32_t>(),
                                                                                    (uint32_t) total, ctx->begin.as<int>(), kWindowBin, nbins,
                                                                                    ctx->edge_hist.as<uint32_t>(), ctx->work.as<uint32_t>());
        LAUNCHED(ctx);
      }
    }
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  if (n_pairs) *n_pairs = ctx->n_pairs;
  ctx->stage = ST_SELECTED;
  return EPA_OK;
}

namespace {
// site-blocked CLV copy for the lane = site kernel (derived data, rebuilt when the CLVs change)
int ensure_clvT(epa_ctx * ctx)
{
  if (ctx->clvT_ready) return EPA_OK;
  const int C = ctx->R * ctx->S;
  const size_t t_stride = (size_t) ((ctx->n + CLVT_BLOCK - 1) / CLVT_BLOCK) * (size_t) C * CLVT_BLOCK;
  if (!ctx->d_clvT) CU(dev_alloc(&ctx->d_clvT, (size_t) ctx->n_nodes * t_stride * sizeof(double)));
  dim3 grid((ctx->n + CLVT_BLOCK - 1) / CLVT_BLOCK, ctx->n_nodes);
  clv_site_block_kernel<<<grid, 256, (size_t) CLVT_BLOCK * (C + 1) * sizeof(double), ctx->stream>>>(
      ctx->tree.clv, ctx->tree.clv_stride, ctx->n, C, ctx->d_clvT, t_stride);
  LAUNCHED(ctx);
  if (ctx->S != 4)
  {
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->clvT_ready = true;
    return EPA_OK;
  }
  // first-round tables: transition matrices of orig/2 per edge, then V * inner per (edge, site)
  const uint32_t B = ctx->n_edges;
  const size_t pm = (size_t) ctx->R * 16;
  std::vector<double> lengths(B);
  for (uint32_t i = 0; i < B; ++i) lengths[i] = ctx->h_edges[i].length / 2.0;
  CU(ctx->tmp.ensure((size_t) B * (pm + 1) * sizeof(double)));
  double * d_len = ctx->tmp.as<double>(), * d_pm = d_len + B;
  CU(cudaMemcpyAsync(d_len, lengths.data(), B * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = launch_pmatrices(ctx, d_len, d_pm, B)) return rc;
  if (!ctx->d_gT && dev_alloc(&ctx->d_gT, (size_t) B * t_stride * sizeof(double)) != cudaSuccess)
  {
    (void) cudaGetLastError();
    ctx->d_gT = nullptr;                         // optional table: the kernel then runs the full first pass
  }
  if (ctx->d_gT)
  {
    dim3 g2(B, (ctx->n + 127) / 128);
    switch (ctx->R)
    {
      case 1: blo_first_table_kernel<1><<<g2, 128, 0, ctx->stream>>>(ctx->tree, ctx->n, ctx->d_edges, d_pm, ctx->d_gT, t_stride); break;
      case 2: blo_first_table_kernel<2><<<g2, 128, 0, ctx->stream>>>(ctx->tree, ctx->n, ctx->d_edges, d_pm, ctx->d_gT, t_stride); break;
      default: blo_first_table_kernel<4><<<g2, 128, 0, ctx->stream>>>(ctx->tree, ctx->n, ctx->d_edges, d_pm, ctx->d_gT, t_stride); break;
    }
    LAUNCHED(ctx);
  }
  CU(cudaStreamSynchronize(ctx->stream));        // lengths[] is host memory
  ctx->clvT_ready = true;
  return EPA_OK;
}

// one instantiation of the lane = site kernel: GS = sumtable in global scratch, pr = per-rate
// scalers, inv = +I model
template <int R, bool GS, bool PR, bool INV>
int launch_site_kernel(epa_ctx * ctx, const BloSiteArgs & sa, unsigned grid, int warps, size_t smem)
{
#ifndef EPA_DEV_MIN
  if (sa.b.raxml)
  {
    CU(cudaFuncSetAttribute(blo_site_kernel<R, GS, PR, INV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    blo_site_kernel<R, GS, PR, INV, true><<<grid, warps * 32, smem, ctx->stream>>>(sa);
    return EPA_OK;
  }
#endif
  // models with equal non-zero eigenvalues (JC, F81, K80, every model at the reference's default rates, ...):
  // merged sumtable components (per-site and per-rate scalers; the +I and --raxml-blo variants keep three)
  if constexpr (!INV)
  {
    if (ctx->hm.ngroups == 1)
    {
      CU(cudaFuncSetAttribute(blo_site_kernel<R, GS, PR, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      blo_site_kernel<R, GS, PR, false, false, 1><<<grid, warps * 32, smem, ctx->stream>>>(sa);
      return EPA_OK;
    }
    if (ctx->hm.ngroups == 2)
    {
      CU(cudaFuncSetAttribute(blo_site_kernel<R, GS, PR, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      blo_site_kernel<R, GS, PR, false, false, 2><<<grid, warps * 32, smem, ctx->stream>>>(sa);
      return EPA_OK;
    }
  }
  CU(cudaFuncSetAttribute(blo_site_kernel<R, GS, PR, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  blo_site_kernel<R, GS, PR, INV><<<grid, warps * 32, smem, ctx->stream>>>(sa);
  return EPA_OK;
}

template <int R, bool GS>
int launch_site_variant(epa_ctx * ctx, const BloSiteArgs & sa, unsigned grid, int warps, size_t smem, bool pr, bool inv)
{
#ifdef EPA_DEV_MIN            /* developer builds: the default variant only (compiles in seconds) */
  return launch_site_kernel<R, GS, false, false>(ctx, sa, grid, warps, smem);
#else
  if (pr) return inv ? launch_site_kernel<R, GS, true, true>(ctx, sa, grid, warps, smem)
                     : launch_site_kernel<R, GS, true, false>(ctx, sa, grid, warps, smem);
  return inv ? launch_site_kernel<R, GS, false, true>(ctx, sa, grid, warps, smem)
             : launch_site_kernel<R, GS, false, false>(ctx, sa, grid, warps, smem);
#endif
}

// R = 1, 2, 4: lane = site kernel (kernels_blo_site.cuh)
template <int R>
int launch_blo_site(epa_ctx * ctx, BloArgs & a)
{
  if (int rc = ensure_clvT(ctx)) return rc;
  BloSiteArgs sa{};
  sa.clvT = ctx->d_clvT; sa.t_stride = clvt_node_stride(ctx->n, ctx->R);
  sa.bugcompat = ctx->hm.bugcompat;
  const bool pr = ctx->tree.sr > 1, inv = ctx->tree.inv != nullptr;
  if (ctx->d_gT && ctx->lookup_ready && !pr && !ctx->sw.no_first)
  {
    sa.gT = ctx->d_gT; sa.g_stride = sa.t_stride; sa.lookup = ctx->d_lookup; sa.n_pad = ctx->n_pad;
  }
  const int wmax = std::max(1, ctx->max_span);
  // up to 12 warps per CTA (register file): the first 8 keep their sumtable in tensor memory when
  // the windows fit 8 rows of 32 sites, the others in shared memory
  const size_t fix = (size_t) SiteWarpSmem<R>::SUM * sizeof(double);
  // eigenvalue groups only in the default variant (launch_site_kernel)
  const int G = (inv || a.raxml) ? 3 : ctx->hm.ngroups;
  const size_t rows = (size_t) ((wmax + 31) & ~31) * site_row_pad(G * R) * sizeof(double);
  const size_t budget = ctx->smem_optin - 2048;
  const int max_warps = SITE_MAX_WARPS;
  const bool tm_ok = !ctx->sw.no_tmem;
  int n_tm = !tm_ok ? 0 : (wmax <= SITE_TMEM_ROWS * 32 ? SITE_TMEM_WARPS : (wmax <= SITE_TMEM_ROWS * 64 ? SITE_TMEM_WARPS / 2 : 0));
  if (ctx->sw.site_warps > 0) n_tm = std::min(n_tm, ctx->sw.site_warps);
  sa.tmem_cols = wmax <= SITE_TMEM_ROWS * 32 ? 256 : 512;
  int n_sm = 0;
  while (n_sm < (n_tm ? max_warps - n_tm : 9) && (size_t) (n_tm + n_sm + 1) * fix + (size_t) (n_sm + 1) * rows <= budget) ++n_sm;
  if (ctx->sw.site_warps > 0) n_sm = std::max(0, std::min(n_sm, ctx->sw.site_warps - n_tm));
  int warps = n_tm + n_sm;
  // few resident warps (long windows) lose to the global-scratch variant with 8 warps per SM (measured on
  // 450..1000-site windows, tools/bench_window.py)
  const int gs_below = ctx->sw.gs_below;
  if (warps >= gs_below)
  {
    a.wcap = wmax;
    sa.b = a;
    sa.n_tmem_warps = n_tm;
    const size_t smem = (size_t) warps * fix + (size_t) n_sm * rows;
    uint64_t grid = (uint64_t) ctx->sm_count;
    grid = std::min<uint64_t>(grid, (a.n_pairs + warps - 1) / warps);
    if (int rc = launch_site_variant<R, false>(ctx, sa, (unsigned) grid, warps, smem, pr, inv)) return rc;
  }
  else
  {
    warps = ctx->sw.gs_warps;
    const unsigned grid = (unsigned) std::min<uint64_t>((uint64_t) ctx->sm_count, (a.n_pairs + warps - 1) / warps);
    sa.wpad = (wmax + 31) & ~31;
    CU(ctx->scratch.ensure((size_t) grid * warps * sa.wpad * blo_row(R) * sizeof(double)));
    sa.gscratch = ctx->scratch.as<double>();
    a.wcap = 0;
    sa.b = a;
    const size_t smem = SiteWarpSmem<R>::doubles(0) * sizeof(double) * warps;
    if (int rc = launch_site_variant<R, true>(ctx, sa, grid, warps, smem, pr, inv)) return rc;
  }
  LAUNCHED(ctx);
  return EPA_OK;
}

template <int R, bool GS, bool RAXML, bool PR>
int launch_dna_kernel(epa_ctx * ctx, const BloArgs & a, unsigned grid, int warps, size_t smem)
{
  CU(cudaFuncSetAttribute(blo_dna_kernel<R, GS, RAXML, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  blo_dna_kernel<R, GS, RAXML, PR><<<grid, warps * 32, smem, ctx->stream>>>(a);
  return EPA_OK;
}

template <int R, bool GS>
int launch_dna_variant(epa_ctx * ctx, const BloArgs & a, unsigned grid, int warps, size_t smem)
{
  const bool pr = ctx->tree.sr > 1;
  if (a.raxml) return pr ? launch_dna_kernel<R, GS, true, true>(ctx, a, grid, warps, smem) : launch_dna_kernel<R, GS, true, false>(ctx, a, grid, warps, smem);
  return pr ? launch_dna_kernel<R, GS, false, true>(ctx, a, grid, warps, smem) : launch_dna_kernel<R, GS, false, false>(ctx, a, grid, warps, smem);
}

template <int R>
int launch_blo_dna(epa_ctx * ctx, BloArgs & a)
{
  const int wmax = std::max(1, ctx->max_span);
  const size_t per_warp = BloWarpSmem<R>::doubles(wmax) * sizeof(double);
  const size_t budget = ctx->smem_optin - 2048;
  int warps = (int) std::min<size_t>(8, budget / per_warp);      // __launch_bounds__(256, 1)
  if (warps >= 1)
  {
    a.wcap = wmax;
    const size_t smem = per_warp * warps;
    int ctas_per_sm = (int) std::max<size_t>(1, std::min<size_t>(4, (ctx->smem_optin + 1024) / (smem + 2048)));
    // keep every SM busy but do not launch far more warps than there are pairs
    uint64_t grid = (uint64_t) ctx->sm_count * ctas_per_sm;
    grid = std::min<uint64_t>(grid, (a.n_pairs + warps - 1) / warps);
    if (int rc = launch_dna_variant<R, false>(ctx, a, (unsigned) grid, warps, smem)) return rc;
  }
  else
  {
    warps = 8;
    const unsigned grid = (unsigned) std::min<uint64_t>((uint64_t) ctx->sm_count * 2, (a.n_pairs + warps - 1) / warps);
    CU(ctx->scratch.ensure((size_t) grid * warps * ctx->n * blo_row(R) * sizeof(double)));
    a.scratch = ctx->scratch.as<double>();
    a.wcap = 0;
    const size_t smem = BloWarpSmem<R>::doubles(0) * sizeof(double) * warps;
    if (int rc = launch_dna_variant<R, true>(ctx, a, grid, warps, smem)) return rc;
  }
  LAUNCHED(ctx);
  return EPA_OK;
}
}  // namespace

extern "C" int epa_place_pairs(epa_ctx * ctx, const epa_options * opts)
{
  if (!ctx || !opts) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  if (ctx->stage < ST_SELECTED) return fail(ctx, EPA_ERR_STATE, "epa_select has not run");
  if (int rc = check_edges_ready(ctx)) return rc;
  if (int rc = bind_constants(ctx)) return rc;
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  if (ctx->n_pairs)
  {
    CU(ctx->res.ensure(ctx->n_pairs * sizeof(BloResult)));
    CU(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), ctx->stream));
    BloArgs a{};
    a.tree = ctx->tree; a.n = ctx->n; a.edges = ctx->d_edges; a.codes = ctx->codes.as<uint8_t>();
    a.begin = ctx->begin.as<int>(); a.span = ctx->span.as<int>();
    a.work = ctx->implicit_pairs ? nullptr : ctx->work.as<uint32_t>();
    a.pair_q = ctx->implicit_pairs ? nullptr : ctx->pair_q.as<uint32_t>();
    a.pair_e = ctx->implicit_pairs ? nullptr : ctx->pair_e.as<uint32_t>();
    a.perm = ctx->perm.as<uint32_t>();
    a.n_pairs = (uint32_t) ctx->n_pairs; a.nq = ctx->nq; a.n_edges = ctx->n_edges;
    a.counter = ctx->d_counter; a.out = ctx->res.as<BloResult>(); a.scratch = nullptr; a.wcap = 0;
    a.raxml = opts->sliding_blo ? 0 : 1;
    a.bugcompat = ctx->hm.bugcompat;
    int rc;
    if (ctx->S == 4)
    {
      switch (ctx->R)
      {
#ifndef EPA_DEV_MIN
        case 1: rc = launch_blo_site<1>(ctx, a); break;
        case 2: rc = launch_blo_site<2>(ctx, a); break;
        case 8: rc = launch_blo_dna<8>(ctx, a); break;
#endif
        case 4: rc = launch_blo_site<4>(ctx, a); break;
        default: return fail(ctx, EPA_ERR_ARG, "unsupported rate category count %d", ctx->R);
      }
    }
    else
    {
      const bool pr = ctx->tree.sr > 1;
      const bool site = !ctx->sw.old_aa || pr;
      if (site)
        if (int rc2 = ensure_clvT(ctx)) return rc2;
      const size_t aa_stride = (size_t) ((ctx->n + CLVT_BLOCK - 1) / CLVT_BLOCK) * (size_t) (ctx->R * ctx->S) * CLVT_BLOCK;
      // 20 states, windows up to 320 sites, pplacer-style BLO: fp64 tensor-core kernel with the sumtable in tensor memory
      const bool mma = site && ctx->S == 20 && (ctx->R == 4 || ctx->R == 1) && !a.raxml && ctx->max_span <= AA_MAX_WINDOW &&
                       !ctx->sw.aa_dfma && !ctx->sw.no_tmem && !pr;
      if (getenv("EPA_B200_DEBUG_AA")) fprintf(stderr, "[aa] mma=%d site=%d S=%d R=%d raxml=%d max_span=%d\n", (int) mma, (int) site, ctx->S, ctx->R, a.raxml, ctx->max_span);
      cudaError_t ce;
      if (mma)
        ce = ctx->R == 4 ? launch_blo_aa<4>(ctx->sm_count, a, ctx->stream, ctx->d_clvT, aa_stride)
                         : launch_blo_aa<1>(ctx->sm_count, a, ctx->stream, ctx->d_clvT, aa_stride);
      else
        ce = launch_blo_generic(ctx->S, ctx->R, ctx->sm_count, ctx->smem_optin, ctx->max_span, ctx->d_model, a, &ctx->scratch.p,
                                &ctx->scratch.cap, ctx->stream, site ? ctx->d_clvT : nullptr, aa_stride, ctx->sw.no_tmem ? 0 : 1,
                                pr ? 1 : 0, ctx->hm.bugcompat);
      rc = ce == cudaSuccess ? EPA_OK : fail(ctx, EPA_ERR_CUDA, "amino-acid BLO launch failed: %s", cudaGetErrorString(ce));
    }
    if (rc) return rc;
    if (ctx->S != 4) LAUNCHED(ctx);
  }
  CU(cudaEventRecord(ctx->ev[4], ctx->stream));
  ctx->stage = ST_PLACED;
  return EPA_OK;
}

static int collect_impl(epa_ctx * ctx, const epa_options * opts, epa_placement * out, uint32_t * out_counts, bool to_device)
{
  if (!ctx || !opts) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  if (ctx->stage < ST_PLACED) return fail(ctx, EPA_ERR_STATE, "epa_place_pairs has not run");
  if (opts->filter_max == 0) return fail(ctx, EPA_ERR_ARG, "filter_max = 0 (unlimited) is not supported: it is the record stride");
  if (opts->filter_min < 1) return fail(ctx, EPA_ERR_ARG, "Filter min cannot be smaller than 1!");
  if (!(opts->support_threshold >= 0.0 && opts->support_threshold <= 1.0))
    return fail(ctx, EPA_ERR_ARG, "thresh is not a valid likelihood weight ratio (outside of [0,1])");
  if (opts->filter_acc_lwr && opts->filter_min > opts->filter_max) return fail(ctx, EPA_ERR_ARG, "Filter min cannot be smaller than max!");
  const uint32_t nq = ctx->nq;
  if (nq == 0) return EPA_OK;
  if (to_device && (!out || !out_counts)) return fail(ctx, EPA_ERR_ARG, "null device output");
  DevBuf & rec_buf = ctx->out_flip ? ctx->out_rec2 : ctx->out_rec;
  DevBuf & cnt_buf = ctx->out_flip ? ctx->out_cnt2 : ctx->out_cnt;
  if (!to_device)
  {
    if (int rc = ensure_copy_stream(ctx)) return rc;
    // the copy that last read this buffer pair (two chunks ago) must have finished
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[ctx->out_flip], 0));
    if ((size_t) nq * opts->filter_max * sizeof(PlacementRec) > rec_buf.cap || nq * sizeof(uint32_t) > cnt_buf.cap)
      CU(cudaStreamSynchronize(ctx->copy_stream));          // growing frees the old buffer
    CU(rec_buf.ensure((size_t) nq * opts->filter_max * sizeof(PlacementRec)));
    CU(cnt_buf.ensure(nq * sizeof(uint32_t)));
  }
  CU(cudaMemsetAsync(ctx->d_flags, 0, 2 * sizeof(int), ctx->stream));
  CollectArgs a{};
  a.res = ctx->res.as<BloResult>();
  a.pair_e = ctx->implicit_pairs ? nullptr : ctx->pair_e.as<uint32_t>();
  a.off = ctx->off.as<uint32_t>(); a.cnt = ctx->cnt.as<uint32_t>();
  a.nq = nq; a.n_edges = ctx->n_edges;
  a.acc_mode = opts->filter_acc_lwr ? 1 : 0;
  a.thresh = opts->support_threshold; a.fmin = opts->filter_min; a.fmax = opts->filter_max;
  a.out = to_device ? reinterpret_cast<PlacementRec *>(out) : rec_buf.as<PlacementRec>();
  a.out_cnt = to_device ? out_counts : cnt_buf.as<uint32_t>();
  a.err = ctx->d_flags;
  collect_kernel<<<(nq + 7) / 8, 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
  CU(cudaEventRecord(ctx->ev[5], ctx->stream));
  if (!to_device && (out || out_counts))
  {
    CU(cudaEventRecord(ctx->ev_collect, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_collect, 0));
    if (out) CU(cudaMemcpyAsync(out, rec_buf.p, (size_t) nq * opts->filter_max * sizeof(PlacementRec), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (out_counts) CU(cudaMemcpyAsync(out_counts, cnt_buf.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_d2h[ctx->out_flip], ctx->copy_stream));
    ctx->out_flip ^= 1;
    if (!ctx->defer_results) CU(cudaStreamSynchronize(ctx->copy_stream));
  }
  int flags[16];
  if (int rc = read_flags(ctx, flags)) return rc;
  for (int i = 0; i < 5; ++i) (void) cudaEventElapsedTime(&ctx->ms[i], ctx->ev[i], ctx->ev[i + 1]);
  (void) cudaGetLastError();
  if (flags[0] == 3) return fail(ctx, EPA_ERR_QUERY, "query %d: likelihood is not finite", flags[1] - 1);
  return EPA_OK;
}

extern "C" int epa_collect(epa_ctx * ctx, const epa_options * opts, epa_placement * out, uint32_t * out_counts)
{
  return collect_impl(ctx, opts, out, out_counts, false);
}

extern "C" int epa_collect_dev(epa_ctx * ctx, const epa_options * opts, epa_placement * out_dev, uint32_t * out_counts_dev)
{
  return collect_impl(ctx, opts, out_dev, out_counts_dev, true);
}

extern "C" int epa_ctx_set_stream(epa_ctx * ctx, void * cuda_stream)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  return EPA_OK;
}

extern "C" int epa_place_chunk(epa_ctx * ctx, const char * seqs, uint32_t n_queries, const epa_options * opts,
                               epa_placement * out, uint32_t * out_counts)
{
  if (!ctx || !opts) return EPA_ERR_ARG;
  if (!ctx->lookup_ready && opts->prescoring)
    if (int rc = epa_build_lookup(ctx)) return rc;
  if (int rc = epa_upload_queries(ctx, seqs, n_queries, opts->premasking)) return rc;
  if (opts->prescoring)
  {
    epa_hint_selection(ctx, opts);
    if (int rc = epa_preplace(ctx)) return rc;
  }
  if (int rc = epa_select(ctx, opts, nullptr)) return rc;
  if (int rc = epa_place_pairs(ctx, opts)) return rc;
  return epa_collect(ctx, opts, out, out_counts);
}

extern "C" int epa_get_pairs(epa_ctx * ctx, uint32_t * query_ids, uint32_t * edge_ids, epa_placement * raw, uint64_t capacity)
{
  if (!ctx) return EPA_ERR_ARG;
  if (ctx->stage < ST_SELECTED) return fail(ctx, EPA_ERR_STATE, "epa_select has not run");
  if (capacity < ctx->n_pairs) return fail(ctx, EPA_ERR_ARG, "capacity %llu < %llu pairs", (unsigned long long) capacity, (unsigned long long) ctx->n_pairs);
  if (int rc = set_device(ctx)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  const uint64_t np = ctx->n_pairs;
  std::vector<uint32_t> q(np), e(np);
  if (ctx->implicit_pairs)
    for (uint64_t p = 0; p < np; ++p) { q[p] = (uint32_t) (p / ctx->n_edges); e[p] = (uint32_t) (p % ctx->n_edges); }
  else if (np)
  {
    CU(cudaMemcpy(q.data(), ctx->pair_q.p, np * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(e.data(), ctx->pair_e.p, np * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  if (query_ids) memcpy(query_ids, q.data(), np * sizeof(uint32_t));
  if (edge_ids) memcpy(edge_ids, e.data(), np * sizeof(uint32_t));
  if (raw)
  {
    if (ctx->stage < ST_PLACED) return fail(ctx, EPA_ERR_STATE, "epa_place_pairs has not run");
    std::vector<BloResult> r(np);
    if (np) CU(cudaMemcpy(r.data(), ctx->res.p, np * sizeof(BloResult), cudaMemcpyDeviceToHost));
    for (uint64_t p = 0; p < np; ++p)
    {
      raw[p].branch_id = e[p];
      raw[p].likelihood = r[p].logl;
      raw[p].lwr = 0.0;
      raw[p].pendant_length = r[p].pendant;
      raw[p].distal_length = r[p].distal;
    }
  }
  return EPA_OK;
}

extern "C" int epa_last_timings(epa_ctx * ctx, float ms[5])
{
  if (!ctx || !ms) return EPA_ERR_ARG;
  for (int i = 0; i < 5; ++i) ms[i] = ctx->ms[i];
  return EPA_OK;
}

extern "C" int epa_num_pairs(epa_ctx * ctx, uint64_t * n_pairs)
{
  if (!ctx || !n_pairs) return EPA_ERR_ARG;
  *n_pairs = ctx->n_pairs;
  return EPA_OK;
}

// ---- fp64 roofline denominator: dependent-chain-free DFMA stream, 16 warps per SM ----------------
namespace {
__global__ void __launch_bounds__(512)
fp64_peak_kernel(double * out, int iters, double x, double y)
{
  double a[8];
  #pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  #pragma unroll 1
  for (int it = 0; it < iters; ++it)
  {
    #pragma unroll
    for (int rep = 0; rep < 8; ++rep)
      #pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0.0;
  #pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int epa_measure_fp64_peak(int device, double * tflops)
{
  if (!tflops) return EPA_ERR_ARG;
  *tflops = 0.0;
  if (cudaSetDevice(device) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_CUDA; }
  const int blocks = prop.multiProcessorCount * 2, threads = 512, iters = 20000;
  double * out = nullptr;
  if (dev_alloc(&out, (size_t) blocks * threads * sizeof(double)) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_NOMEM; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 0.0f;
  for (int rep = 0; rep < 4; ++rep)                 // first repetition warms up
  {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.0f || ms < best)) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  dev_free(out);
  if (cudaGetLastError() != cudaSuccess || best <= 0.0f) return EPA_ERR_CUDA;
  *tflops = 2.0 * (double) blocks * threads * iters * 64.0 / (best * 1e-3) / 1e12;
  return EPA_OK;
}

extern "C" int epa_pinned_alloc(void ** ptr, size_t bytes)
{
  if (!ptr) return EPA_ERR_ARG;
  *ptr = nullptr;
  if (cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess)
  {
    (void) cudaGetLastError();
    return EPA_ERR_NOMEM;
  }
  return EPA_OK;
}

extern "C" void epa_pinned_free(void * ptr)
{
  if (ptr) (void) cudaFreeHost(ptr);
}

// ---- peer memory (multi-GPU, one process per GPU) ----------------------------------------------------------
// The gather of the placement records without a collective: the owning rank allocates the buffer of all shards and
// exports it (CUDA IPC), the other ranks of the box map it and hand their slice to epa_collect_dev - the collect
// kernel then writes its records straight into the owner's memory over NVLink.
extern "C" int epa_peer_alloc(int device, size_t bytes, void ** dptr, unsigned char * handle64)
{
  epa_ctx * none = nullptr;
  if (!dptr || !handle64) return fail(none, EPA_ERR_ARG, "null argument");
  *dptr = nullptr;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (cudaSetDevice(device) != cudaSuccess) { (void) cudaGetLastError(); return fail(none, EPA_ERR_CUDA, "cannot select device %d", device); }
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);        // its own allocation: the handle covers exactly this buffer
  if (e != cudaSuccess) { (void) cudaGetLastError(); return fail(none, EPA_ERR_NOMEM, "peer buffer of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *dptr);
  if (e != cudaSuccess)
  {
    (void) cudaGetLastError();
    (void) cudaFree(*dptr); *dptr = nullptr;
    return fail(none, EPA_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle64, &h, 64);
  return EPA_OK;
}

extern "C" int epa_peer_open(int device, const unsigned char * handle64, void ** dptr)
{
  epa_ctx * none = nullptr;
  if (!dptr || !handle64) return fail(none, EPA_ERR_ARG, "null argument");
  *dptr = nullptr;
  if (cudaSetDevice(device) != cudaSuccess) { (void) cudaGetLastError(); return fail(none, EPA_ERR_CUDA, "cannot select device %d", device); }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  const cudaError_t e = cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { (void) cudaGetLastError(); *dptr = nullptr; return fail(none, EPA_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
  return EPA_OK;
}

extern "C" int epa_peer_close(int device, void * dptr)
{
  if (!dptr) return EPA_OK;
  if (cudaSetDevice(device) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_CUDA; }
  if (cudaIpcCloseMemHandle(dptr) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_CUDA; }
  return EPA_OK;
}

extern "C" int epa_peer_free(int device, void * dptr)
{
  if (!dptr) return EPA_OK;
  if (cudaSetDevice(device) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_CUDA; }
  (void) cudaDeviceSynchronize();
  if (cudaFree(dptr) != cudaSuccess) { (void) cudaGetLastError(); return EPA_ERR_CUDA; }
  return EPA_OK;
}

extern "C" int epa_synchronize(epa_ctx * ctx)
{
  if (!ctx) return EPA_ERR_ARG;
  if (int rc = set_device(ctx)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return EPA_OK;
}
