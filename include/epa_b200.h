/*
 * epa_b200.h - C ABI of libepa_b200.so: the B200-native replacement for EPA-ng's per-query
 * placement hot path.
 *
 * EPA-ng has no plugin API; the seam this library replaces is the pair of chunk loops
 *   place()          /root/reference/src/core/place.cpp:41-95    (preplacement, all branches)
 *   place_thorough() /root/reference/src/core/place.cpp:97-171   (branch-length optimisation)
 * and everything they call below: Tiny_Tree (src/tree/Tiny_Tree.cpp:48-218), Lookup_Store
 * (src/core/Lookup_Store.hpp:73-141), optimize_branch_triplet (src/core/pll/optimize.cpp:60-286)
 * and the libpll kernels listed in SURVEY.md 3.5. Candidate selection (src/core/heuristics.hpp:40-64)
 * and the output filter (src/set_manipulators.cpp:43-204) sit between/after the two loops and are
 * part of the same device pipeline.
 *
 * Conventions
 *   - plain C types only; every pointer is a HOST pointer unless the name ends in _dev
 *   - all functions return 0 on success, a negative epa_status otherwise; epa_last_error() gives
 *     the message (the reference throws std::runtime_error, e.g. Tiny_Tree.cpp:145-156,209-212)
 *   - calls on one epa_ctx must be serialised by the caller (one host thread per GPU); contexts on DIFFERENT
 *     devices may be driven concurrently from different threads (epa_run_files_multi does), contexts on the SAME
 *     device must not: the model tables live in one __constant__ symbol per device, bound by the context that
 *     launches (switching contexts on one thread is fine)
 *   - the library never falls back to a CPU path: without a CUDA device every compute entry
 *     point fails with EPA_ERR_CUDA
 *
 * Layouts (identical to libpll so that reference buffers can be handed over unchanged):
 *   CLV      double[sites][rate_cats][states]      libpll core_likelihood.c:1419-1459
 *   scaler   uint32[sites] or uint32[sites][rate_cats] with EPA_FLAG_RATE_SCALERS
 *   eigen    eigenvecs / inv_eigenvecs double[states*states] row-major, libpll models.c:394-404
 *   tip      uint32 state mask per site (DNA: A=1,C=2,G=4,T=8; AA: bit i = i-th state of ARNDCQEGHILKMFPSTWYV)
 */
#ifndef EPA_B200_H
#define EPA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct epa_ctx epa_ctx;

typedef enum {
  EPA_OK = 0,
  EPA_ERR_ARG = -1,        /* invalid argument / unsupported configuration */
  EPA_ERR_CUDA = -2,       /* CUDA runtime failure or no device */
  EPA_ERR_NOMEM = -3,
  EPA_ERR_STATE = -4,      /* call sequence violated (e.g. preplace before build_lookup) */
  EPA_ERR_QUERY = -5,      /* invalid query: bad character, all-gap sequence, -inf likelihood */
} epa_status;

enum {
  EPA_FLAG_RATE_SCALERS = 1u << 0,     /* PLL_ATTRIB_RATE_SCALERS (src/io/file_io.cpp:211-214). With 20 states the counters
                                          are placed as libpll's generic tip-inner update places them
                                          (LP/core_partials.c:461-506: whole-site rescaling, entry [site index]);
                                          uploaded CLVs are expected to carry the reference's own counters */
  EPA_FLAG_BUGCOMPAT_FOCUS = 1u << 1,  /* reproduce shift_partition_focus' per-rate scaler offset
                                          (src/core/pll/pll_util.cpp:405-408) */
};

/* raxml::Model + pll_partition_t model tables (src/core/raxml/Model.hpp, libpll pll.h:209-260) */
typedef struct {
  uint32_t states;                /* 4 or 20 */
  uint32_t rate_cats;             /* 1..8 */
  uint32_t sites;                 /* alignment width after premasking */
  uint32_t flags;                 /* EPA_FLAG_* */
  const double * eigenvals;       /* [states] */
  const double * eigenvecs;       /* [states*states] */
  const double * inv_eigenvecs;   /* [states*states] */
  const double * freqs;           /* [states] */
  const double * rates;           /* [rate_cats] */
  const double * rate_weights;    /* [rate_cats]; the weights of the placement likelihoods. (The reference's tiny
                                     partition keeps the default 1/rate_cats even for +R{..}{weights} models,
                                     src/tree/tiny_util.cpp:110-111: pass those to reproduce its placements.) */
  double pinv;                    /* proportion of invariant sites in [0, 1) (+IU{p}); the invariant sites are
                                     derived from the tip masks (libpll models.c:495-760) */
} epa_model_desc;

/* One reference-tree edge as Tiny_Tree sees it (src/tree/Tiny_Tree.cpp:48-76): the two CLVs
 * looking away from the edge. Node ids: 0..n_tips-1 = tips, n_tips.. = directional CLV slots.
 * If one side is a tip it must be `distal` (Tiny_Tree.cpp:64-74). */
typedef struct {
  uint32_t distal;
  uint32_t proximal;
  double length;
} epa_edge_desc;

/* One Felsenstein pruning step of the reference-tree precompute
 * (src/core/pll/epa_pll_util.cpp:62-107 -> pll_update_partials): CLV slot `parent` is computed
 * from children `left`/`right` (tip id or CLV slot) across branches of the given lengths.
 * Operations may be listed in any order; the library schedules them by dependency depth. */
typedef struct {
  uint32_t parent;
  uint32_t left;
  uint32_t right;
  uint32_t reserved;
  double left_length;
  double right_length;
} epa_clv_op;

/* Host-owned CLV handed over as is (drop-in for Tree::get_clv, src/tree/Tree.cpp:80-117) */
typedef struct {
  uint32_t slot;                  /* CLV slot id (>= n_tips) */
  uint32_t reserved;
  const double * clv;             /* [sites][rate_cats][states] */
  const uint32_t * scaler;        /* may be NULL (= all zero) */
} epa_host_clv;

/* = Placement (src/sample/Placement.hpp:49-53), 40 bytes */
typedef struct {
  uint64_t branch_id;
  double likelihood;
  double lwr;
  double pendant_length;
  double distal_length;
} epa_placement;

/* Options on the path (src/util/Options.hpp:5-35) */
typedef struct {
  int32_t prescoring;             /* 1: preplacement + heuristic (default); 0: --no-heur */
  int32_t heuristic;              /* 0 dynamic (accumulated LWR, -g), 1 fixed fraction (-G), 2 baseball */
  double prescoring_threshold;    /* default 0.99999 */
  int32_t premasking;             /* 1: restrict each query to [first non-gap, last non-gap] */
  int32_t sliding_blo;            /* 1: pplacer-style BLO (default); 0: --raxml-blo (optimize.cpp:274-278) */
  int32_t filter_acc_lwr;         /* 0: min-LWR filter (default), 1: accumulated-LWR filter */
  double support_threshold;       /* default 0.01 */
  uint32_t filter_min;            /* default 1 */
  uint32_t filter_max;            /* default 7; also the record stride of the output */
} epa_options;

/* -------------------------------------------------------------------------------------------- */
/* context                                                                                      */
/* -------------------------------------------------------------------------------------------- */

void epa_options_default(epa_options * opts);

/* Creates a context on CUDA device `device`, uploads the model and the tip state masks
 * (tip_masks[n_tips][sites]), reserves n_clv_slots directional CLVs and registers the edges.
 * Mirrors Tree::Tree + make_partition (src/tree/Tree.cpp:16-56, src/io/file_io.cpp:205-235). */
int epa_ctx_create(epa_ctx ** ctx, int device, const epa_model_desc * model,
                   uint32_t n_tips, const uint32_t * tip_masks,
                   uint32_t n_clv_slots,
                   const epa_edge_desc * edges, uint32_t n_edges);

/* Fills CLV slots on the device by Felsenstein pruning (replaces precompute_clvs). */
int epa_compute_clvs(epa_ctx * ctx, const epa_clv_op * ops, uint32_t n_ops);

/* Alternative to epa_compute_clvs: copies CLVs computed by the caller (e.g. by libpll). */
int epa_upload_clvs(epa_ctx * ctx, const epa_host_clv * clvs, uint32_t n_clvs);

/* Per-edge lookup tables: Tiny_Tree ctor + precompute_sites_static + Lookup_Store::init_branch
 * (src/tree/Tiny_Tree.cpp:18-46,84-128). */
int epa_build_lookup(epa_ctx * ctx);

/* Whole hot path for one chunk of queries (src/core/place.cpp:207-246 minus file I/O):
 * seqs = n_queries rows of `sites` ASCII characters (already premasked column-wise, any case).
 * out  = n_queries * opts->filter_max records, row q holds out_counts[q] placements sorted by
 *        descending LWR (compute_and_set_lwr + filter, place.cpp:238-239). */
int epa_place_chunk(epa_ctx * ctx, const char * seqs, uint32_t n_queries,
                    const epa_options * opts, epa_placement * out, uint32_t * out_counts);

/* -------------------------------------------------------------------------------------------- */
/* staged interface (what epa_place_chunk runs, exposed for tests and device-resident timing)    */
/* -------------------------------------------------------------------------------------------- */

/* H2D copy of the chunk + encoding + valid-range detection (src/util/Range.hpp:34-49). */
int epa_upload_queries(epa_ctx * ctx, const char * seqs, uint32_t n_queries, int premasking);
/* Announces the chunk that will follow the next epa_upload_queries / epa_place_chunk call: its
 * host-to-device copy then runs on a second stream while that chunk is being placed (the
 * reference prefetches the next chunk with std::async, src/seq/MSA_Stream.cpp:79-85). The host
 * buffer must stay valid and unchanged until it has been uploaded. NULL cancels the hint. */
int epa_hint_next_chunk(epa_ctx * ctx, const char * next_seqs, uint32_t next_n_queries);
/* Deferred results: with on != 0, epa_place_chunk / epa_collect return once the records are
 * computed and their device-to-host copy has been QUEUED on the copy stream; the host buffers may
 * only be read after epa_wait_results (the reference writes its jplace chunks asynchronously too,
 * src/io/jplace_writer.hpp:59-65). Default: off (results are in the host buffers on return). */
int epa_set_deferred_results(epa_ctx * ctx, int on);
int epa_wait_results(epa_ctx * ctx);
/* With deferred results on: waits only for the records of the chunk before the most recent one (their copy
 * ran under the most recent chunk's kernels), so a caller can hand chunk k - 1 on while k's copy is in flight. */
int epa_wait_older_results(epa_ctx * ctx);
/* Same for a chunk that already lives in device memory (seqs_dev = DEVICE pointer to
 * n_queries * sites ASCII bytes): used for device-resident timing and by callers that stage
 * the query file in HBM themselves. */
int epa_encode_queries_dev(epa_ctx * ctx, const char * seqs_dev, uint32_t n_queries, int premasking);
/* HOT LOOP A: pre[q][b] for all queries x all edges (Lookup_Store::sum_precomputed_sitelk). */
int epa_preplace(epa_ctx * ctx);
/* Optional, before epa_preplace: the options the following epa_select will be called with. In a context created
 * with EPA_B200_FUSED_SELECT set in the environment, the tensor-core preplacement then selects the candidates of the
 * dynamic heuristic (-g) in its epilogue and never writes the [query][edge] score matrix (epa_get_prescores is
 * refused for that chunk). The hint covers one epa_preplace; an epa_select with other options re-runs the
 * preplacement unfused. Measured slower than the two-kernel path on B200, hence opt-in;
 * without the switch the hint is accepted and ignored. */
int epa_hint_selection(epa_ctx * ctx, const epa_options * opts);
/* Candidate selection -> (query, edge) work list; n_pairs receives its size. With
 * opts->prescoring == 0 the list is all queries x all edges. */
int epa_select(epa_ctx * ctx, const epa_options * opts, uint64_t * n_pairs);
/* HOT LOOP B: branch-length optimisation of every pair in the work list. */
int epa_place_pairs(epa_ctx * ctx, const epa_options * opts);
/* LWR over the evaluated candidates, filter, pack; D2H of the records. */
int epa_collect(epa_ctx * ctx, const epa_options * opts, epa_placement * out, uint32_t * out_counts);

/* Same, but the records stay in device memory: out_dev[n_queries][filter_max] and
 * counts_dev[n_queries] are DEVICE pointers owned by the caller (e.g. the send buffer of the
 * NCCL gather that collects the shards of a multi-GPU run). */
int epa_collect_dev(epa_ctx * ctx, const epa_options * opts, epa_placement * out_dev, uint32_t * counts_dev);

/* Makes the context issue all its work on the caller's CUDA stream (a cudaStream_t; NULL = the
 * legacy default stream) so that the caller's events and collectives order against it. */
int epa_ctx_set_stream(epa_ctx * ctx, void * cuda_stream);

/* -------------------------------------------------------------------------------------------- */
/* inspection (parity tests)                                                                    */
/* -------------------------------------------------------------------------------------------- */

/* CLV slot / scaler as stored on the device. */
int epa_get_clv(epa_ctx * ctx, uint32_t slot, double * clv, uint32_t * scaler);
/* Lookup table of one edge in the reference's column order (NT_MAP / AA_MAP,
 * src/util/maps.hpp:9-28): out[sites][16 or 24]. */
int epa_get_lookup(epa_ctx * ctx, uint32_t edge, double * out);
/* pre[q][b] of the current chunk: out[n_queries][n_edges]. */
int epa_get_prescores(epa_ctx * ctx, double * out);
/* Work list of the current chunk (query index, edge index) and the raw BLO result per pair. */
int epa_get_pairs(epa_ctx * ctx, uint32_t * query_ids, uint32_t * edge_ids, epa_placement * raw,
                  uint64_t capacity);
/* Reference-tree log-likelihood evaluated across one edge (Tree::ref_tree_logl, Tree.cpp:119-131). */
int epa_edge_loglikelihood(epa_ctx * ctx, uint32_t edge, double * logl);

/* Device-time of the stages of the last chunk in milliseconds (CUDA events on the ctx stream):
 * [0] upload+encode [1] preplace [2] select [3] thorough [4] collect. */
int epa_last_timings(epa_ctx * ctx, float ms[5]);
/* Device time of the last epa_build_lookup (pmatrices + column table + lookup kernel), ms. */
int epa_last_lookup_ms(epa_ctx * ctx, float * ms);
/* Size of the current work list. */
int epa_num_pairs(epa_ctx * ctx, uint64_t * n_pairs);
/* Blocks until all work queued on the context's stream has finished. */
int epa_synchronize(epa_ctx * ctx);
/* Measured fp64 FMA throughput of the device (TFLOP/s, 2 flops per DFMA lane): an unrolled stream of
 * independent DFMAs on every SM. The roofline denominator of the fp64-bound thorough kernel. */
int epa_measure_fp64_peak(int device, double * tflops);
/* The library keeps the device memory blocks of destroyed contexts (per device, up to EPA_B200_DEVICE_POOL_MB MB in the
 * environment, default 16384, 0 = off) for the next context - allocating and freeing several GB per context otherwise
 * stalls for hundreds of milliseconds now and then. This call returns the cached blocks to the driver. */
void epa_device_pool_trim(void);
/* Page-locked host memory for query rows and result records (full PCIe speed, asynchronous copies). */
int epa_pinned_alloc(void ** ptr, size_t bytes);
void epa_pinned_free(void * ptr);
/* Peer memory for the gather of the records on a multi-GPU box with one process per GPU - the counterpart of the
 * reference's MPI gather of the samples to rank 0 (src/net/epa_mpi_util.hpp, src/io/jplace_writer.hpp:92-132) without a
 * collective: the owning rank allocates the buffer of all shards (epa_peer_alloc: device memory + a 64-byte CUDA IPC handle
 * to send to the other ranks by any means), they map it (epa_peer_open) and pass their slice of it to epa_collect_dev:
 * the collect kernel writes the records straight into the owner's memory over NVLink. A barrier between the ranks, after
 * their streams are synchronised, completes the gather. */
int epa_peer_alloc(int device, size_t bytes, void ** dptr, unsigned char * handle64);
int epa_peer_open(int device, const unsigned char * handle64, void ** dptr);
int epa_peer_close(int device, void * dptr);
int epa_peer_free(int device, void * dptr);
/* Number of kernel launches issued on behalf of the ctx since creation. */
uint64_t epa_launch_count(const epa_ctx * ctx);

const char * epa_last_error(const epa_ctx * ctx);   /* ctx may be NULL: error of a failed create */
void epa_ctx_destroy(epa_ctx * ctx);

#ifdef __cplusplus
}
#endif
#endif /* EPA_B200_H */
