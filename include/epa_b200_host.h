/*
 * epa_b200_host.h - C ABI of the C++ host layer in libepa_b200.so: everything the reference does
 * around the hot path to get from files to a jplace, re-typed (not accelerated) so that the
 * library is usable end to end:
 *
 *   tree parsing + edge numbering   /root/reference/src/io/file_io.cpp:120-192, src/core/pll/pll_util.cpp:182-352
 *   model string                    src/core/raxml/Model.cpp:123-560 (subset, see csrc/host/model.hpp)
 *   MSA reading + pre-masking       src/seq/MSA_Info.hpp:22-111, src/main.cpp:470-494
 *   Tree(...) construction          src/tree/Tree.cpp:16-56 (CLVs are computed ON THE DEVICE)
 *   chunk loop                      src/core/place.cpp:173-251 (simple_mpi)
 *   jplace output                   src/io/jplace_util.cpp:20-86
 *
 * A session owns one epa_ctx (include/epa_b200.h); all compute goes through that C ABI.
 * All functions return 0 or a negative epa_status; epa_host_last_error() gives the message of
 * the last failure on the calling thread.
 */
#ifndef EPA_B200_HOST_H
#define EPA_B200_HOST_H

#include "epa_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct epa_session epa_session;

/* Builds the reference state on `device`: parses the newick text, matches the n_taxa rows of the
 * reference alignment (ref_rows[n_taxa][sites], ASCII, already column-masked) to the tips by
 * name, parses the model string, computes all directional CLVs and the lookup tables. */
int epa_session_open(epa_session ** session, const char * newick, uint32_t n_taxa,
                     const char * const * names, const char * ref_rows, uint32_t sites,
                     const char * model, int device);

/* The same with the per-rate scaler policy of THIS session (the reference's --rate-scalers): rate_scalers 0 = off,
 * 1 = on, 2 = auto (on above 2000 tips), -1 = the process-wide policy of epa_host_set_rate_scalers; bugcompat_focus
 * 1 / 0 / -1 likewise (the reference's scaler window offset of the thorough phase). */
int epa_session_open_ex(epa_session ** session, const char * newick, uint32_t n_taxa,
                        const char * const * names, const char * ref_rows, uint32_t sites,
                        const char * model, int device, int rate_scalers, int bugcompat_focus);

/* Places n_queries rows (query_rows[n_queries][sites], ASCII, HOST memory; pinned memory gives
 * full PCIe speed) in chunks of chunk_size queries (0 = default): the reference's chunk loop.
 * out[n_queries][opts->filter_max], counts[n_queries] as in epa_place_chunk. */
int epa_session_place(epa_session * session, const char * query_rows, uint64_t n_queries,
                      const epa_options * opts, uint32_t chunk_size, epa_placement * out,
                      uint32_t * counts);

epa_ctx * epa_session_ctx(epa_session * session);
uint32_t epa_session_num_edges(const epa_session * session);
uint32_t epa_session_num_tips(const epa_session * session);
uint32_t epa_session_sites(const epa_session * session);
/* Rooted reference trees are unrooted for the computation (build_tree_from_file,
 * src/io/file_io.cpp:129-173); by default placements and the jplace tree are reported on the ROOTED
 * tree (--preserve-rooting on; rtree_mapper, src/core/pll/rtree_mapper.hpp:38-61). */
int epa_session_set_preserve_rooting(epa_session * session, int on);
int epa_session_is_rooted(const epa_session * session);
/* jplace "tree" string: newick with {edge_num} annotations. */
const char * epa_session_numbered_newick(epa_session * session, int precision);
/* Reference-tree log-likelihood evaluated across edge 0 (Tree::ref_tree_logl). */
int epa_session_tree_logl(epa_session * session, double * logl);
void epa_session_close(epa_session * session);

/* Whole run, files to jplace: what the reference's main() does for
 *   epa-ng -t tree -s ref_msa -q query -m model -w outdir [options]
 * Writes <outdir>/epa_result.jplace and <outdir>/epa_info.log. */
/* Per-rate scaler policy of the sessions opened afterwards (the reference's --rate-scalers):
 * mode 0 = off, 1 = on, 2 = auto (default: on for more than 2000 tips, src/tree/Tree_Numbers.hpp:11,
 * src/io/file_io.cpp:211-214). bugcompat != 0 (default) reproduces the reference's scaler window
 * offset in the thorough phase (shift_partition_focus, src/core/pll/pll_util.cpp:405-408). */
int epa_host_set_rate_scalers(int mode, int bugcompat);

int epa_run_files(const char * tree_file, const char * ref_msa_file, const char * query_file,
                  const char * model, const char * outdir, const epa_options * opts,
                  uint32_t chunk_size, int precision, int device, const char * invocation);

int epa_run_files_ex(const char * tree_file, const char * ref_msa_file, const char * query_file,
                     const char * model, const char * outdir, const epa_options * opts,
                     uint32_t chunk_size, int precision, int device, const char * invocation,
                     int preserve_rooting);

/* The same run as a pipeline over several GPUs of the box (devices[n_devices], CUDA ordinals): one host
 * thread and one reference state per GPU, query chunks handed out in file order, one reader stage
 * (memory-mapped file, decoded by host threads into pinned memory) and one writer stage shared by all
 * - what the reference does with MPI ranks over query blocks (src/net/epa_mpi_util.cpp:10-30) and its
 * async chunk reader / jplace writer (src/seq/MSA_Stream.cpp:79-85, src/io/jplace_writer.hpp:58-132).
 * host_threads <= 0 = all hardware threads. The jplace does not depend on n_devices or host_threads.
 * stats (may be NULL) receives wall-clock figures of the run. */
typedef struct {
  uint64_t n_queries;
  double seconds_total;      /* whole call */
  double seconds_index;      /* tree + reference MSA read, query file indexed (first pass) */
  double seconds_setup;      /* reference state on the device(s): context, CLVs, lookup tables */
  double seconds_place;      /* pipeline: decode -> place -> format/write, from first chunk to closed file */
  double busy_read;          /* time the reader stage spent decoding */
  double busy_write;         /* time the writer stage spent formatting and writing */
  double busy_device_max;    /* longest time a device thread spent inside epa_session_place */
} epa_run_stats;

int epa_run_files_multi(const char * tree_file, const char * ref_msa_file, const char * query_file,
                        const char * model, const char * outdir, const epa_options * opts,
                        uint32_t chunk_size, int precision, const int * devices, uint32_t n_devices,
                        const char * invocation, int preserve_rooting, int host_threads,
                        epa_run_stats * stats);

/* epa_run_files* keep their page-locked staging blocks (up to (3 n_devices + 2) x chunk_size x (width + 284) bytes)
 * for the next run in the same process; this gives them back to the system. */
void epa_host_release_pinned_pool(void);

/* Formats placement records as a jplace document (src/io/jplace_util.cpp:20-86) into `path`. */
int epa_write_jplace(const char * path, const char * numbered_newick, const char * invocation,
                     const char * const * query_names, uint64_t n_queries, const epa_placement * recs,
                     const uint32_t * counts, uint32_t stride, int precision);

/* ---- device-free inspection of the host logic (CPU tests) ---------------------------------- */
/* Parses the tree; writes the numbered newick (NUL-terminated, truncated to cap) and the counts. */
int epa_host_parse_tree(const char * newick, int precision, char * out_newick, size_t cap,
                        uint32_t * n_tips, uint32_t * n_edges);
/* Reads an aligned sequence file (FASTA, or the reference's bfast: src/io/Binary_Fasta.hpp) into
 * upper-case rows [n][sites]; labels receives the names separated by '\n'. Pass rows = NULL to query
 * the sizes only. */
int epa_host_read_alignment(const char * path, uint32_t * n_sequences, uint32_t * sites, char * rows, size_t rows_cap,
                            char * labels, size_t labels_cap);
/* The reader of the files -> jplace pipeline: memory-mapped file, record index and all-gap column mask
 * built by `threads` host threads, rows decoded in parallel (same results as epa_host_read_alignment).
 * gap_mask_out (may be NULL) receives `sites` bytes, 1 = every sequence has one of "NOX.-?" there. */
int epa_host_read_alignment_mt(const char * path, int threads, int want_mask, uint32_t * n_sequences, uint32_t * sites,
                               char * rows, size_t rows_cap, char * labels, size_t labels_cap, uint8_t * gap_mask_out);
/* printf("%.*f") digit for digit without printf (the jplace writer's number formatting); returns the
 * length, out must hold 400 bytes. */
int epa_host_format_fixed(double value, int precision, char * out, size_t cap);
/* -c/--bfast of the reference (Binary_Fasta::fasta_to_bfast, src/io/Binary_Fasta.hpp:214-246, src/main.cpp:284-288):
 * converts an aligned DNA FASTA file to <out_dir>/<file name>.bfast; out_path (may be NULL) receives the path. */
int epa_host_fasta_to_bfast(const char * fasta_path, const char * out_dir, char * out_path, size_t cap);
/* +F / +FC models: base frequencies counted on the reference MSA's tip state masks [n_tips][sites]
 * (compute_and_set_empirical_frequencies, src/core/pll/optimize.cpp:457-472); freqs / eigenvals (optional) hold `states` doubles.
 * epa_session_open does this itself; the entry point exists for callers that build their own epa_model_desc. */
int epa_host_empirical_frequencies(const char * model, const uint32_t * tip_masks, uint32_t n_tips, uint32_t sites,
                                   double * freqs, double * eigenvals);
/* -m <file> of the reference (src/main.cpp:433-436 -> src/util/parse_model.hpp): the model string of a RAxML 8 info
 * file, a raxml-ng .bestModel file or an IQ-TREE report. epa_run_files does this itself when its model argument
 * names an existing file. */
int epa_host_model_from_file(const char * path, char * out, size_t cap);
/* Rooted input only: translates (edge, distal length) pairs of the unrooted working tree to the
 * rooted tree, in place; writes the numbered newick of the working tree when out_newick != NULL. */
int epa_host_map_rooted(const char * newick, uint32_t * edges, double * distal, uint32_t count,
                        char * out_unrooted_newick, size_t cap);
/* Pruning schedule and edge list that epa_session_open hands to the device API; tip_labels
 * receives the tip names separated by '\n' in tip-id order. Capacities are in elements. */
int epa_host_tree_schedule(const char * newick, uint32_t * n_slots, epa_clv_op * ops, uint32_t ops_cap,
                           uint32_t * n_ops, epa_edge_desc * edges, uint32_t edges_cap, uint32_t * n_edges,
                           char * tip_labels, size_t labels_cap);
/* Parses the model string: arrays must hold 20 / 8 / 400 doubles. */
int epa_host_parse_model(const char * model, uint32_t * states, uint32_t * rate_cats, double * rates,
                         double * weights, double * freqs, double * eigenvals, double * eigenvecs,
                         double * inv_eigenvecs);

const char * epa_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* EPA_B200_HOST_H */
