// session.cpp - host layer: reference state construction, chunk loop, jplace, whole-run driver,
// and the C ABI of include/epa_b200_host.h. All computation is delegated to the epa_* device API.
#include "../../../include/epa_b200_host.h"

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "fastio.hpp"
#include "model.hpp"
#include "seqio.hpp"
#include "tree.hpp"
#include "session_internal.hpp"

using namespace epa_host;

namespace {
thread_local std::string g_host_error;

int host_fail(int code, const std::string & msg)
{
  g_host_error = msg;
  return code;
}

constexpr uint32_t kDefaultChunk = 131072;
}  // namespace

namespace epa_host {
int host_fail_msg(int code, const std::string & msg) { return host_fail(code, msg); }
}  // namespace epa_host

extern "C" const char * epa_host_last_error(void) { return g_host_error.c_str(); }

// DNA / protein state masks of the reference alignment (libpll maps.c:46-111)
static uint32_t tip_mask(int states, uint8_t ch)
{
  const char c = (char) std::toupper(ch);
  if (states == 4)
  {
    switch (c)
    {
      case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8;
      case 'M': return 3; case 'R': return 5; case 'W': return 9; case 'S': return 6; case 'Y': return 10;
      case 'K': return 12; case 'V': return 7; case 'H': return 11; case 'D': return 13; case 'B': return 14;
      case 'N': case 'O': case 'X': case '-': case '.': case '?': return 15;
      default: return 0;
    }
  }
  static const char * order = "ARNDCQEGHILKMFPSTWYV";
  if (const char * p = c ? std::strchr(order, c) : nullptr) return 1u << (uint32_t) (p - order);
  auto bit = [&](char x) { return 1u << (uint32_t) (std::strchr(order, x) - order); };
  switch (c)
  {
    case 'B': return bit('N') | bit('D');
    case 'Z': return bit('Q') | bit('E');
    case 'J': return bit('I') | bit('L');
    case 'X': case '*': case '-': case '.': case '?': return 0xfffffu;
    default: return 0;
  }
}

// per-rate scaler policy of the sessions opened next (process-wide, like a command-line option):
// 0 = off, 1 = on, 2 = auto = on for trees with more than 2000 tips (src/tree/Tree_Numbers.hpp:11,
// src/io/file_io.cpp:211-214). bugcompat = reproduce the reference's scaler window offset
// (SURVEY 8a quirk 4); 0 reads the scalers of the site itself.
static int g_rate_scalers = 2, g_rate_bugcompat = 1;
extern "C" int epa_host_set_rate_scalers(int mode, int bugcompat)
{
  if (mode < 0 || mode > 2) return host_fail(EPA_ERR_ARG, "rate scaler mode must be 0 (off), 1 (on) or 2 (auto)");
  g_rate_scalers = mode; g_rate_bugcompat = bugcompat ? 1 : 0;
  return EPA_OK;
}

extern "C" int epa_session_open(epa_session ** out, const char * newick, uint32_t n_taxa, const char * const * names,
                                const char * ref_rows, uint32_t sites, const char * model_desc, int device)
{
  return epa_session_open_ex(out, newick, n_taxa, names, ref_rows, sites, model_desc, device, -1, -1);
}

extern "C" int epa_session_open_ex(epa_session ** out, const char * newick, uint32_t n_taxa, const char * const * names,
                                   const char * ref_rows, uint32_t sites, const char * model_desc, int device,
                                   int rate_scalers, int bugcompat_focus)
{
  if (!out || !newick || !names || !ref_rows || !model_desc) return host_fail(EPA_ERR_ARG, "null argument");
  if (rate_scalers < -1 || rate_scalers > 2) return host_fail(EPA_ERR_ARG, "rate scaler mode must be 0 (off), 1 (on), 2 (auto) or -1 (process policy)");
  const int rate_mode = rate_scalers < 0 ? g_rate_scalers : rate_scalers;
  const int rate_bug = bugcompat_focus < 0 ? g_rate_bugcompat : (bugcompat_focus ? 1 : 0);
  *out = nullptr;
  try
  {
    std::unique_ptr<epa_session> s(new epa_session());
    s->tree = Tree::parse(newick);
    s->model = Model::parse(model_desc);
    s->sites = sites;
    const size_t T = s->tree.num_tips();
    if (T > n_taxa)
      return host_fail(EPA_ERR_ARG, "tree has " + std::to_string(T) + " tips but the reference MSA has only " + std::to_string(n_taxa) + " sequences");
    std::unordered_map<std::string, uint32_t> by_name;
    for (uint32_t i = 0; i < n_taxa; ++i) by_name.emplace(names[i], i);
    std::vector<uint32_t> masks(T * (size_t) sites);
    for (size_t t = 0; t < T; ++t)
    {
      const std::string & label = s->tree.nodes[s->tree.tip_node[t]].label;
      auto it = by_name.find(label);
      if (it == by_name.end()) return host_fail(EPA_ERR_ARG, "Sequence with header '" + label + "' does not appear in the tree / MSA");
      const uint8_t * row = reinterpret_cast<const uint8_t *>(ref_rows) + (size_t) it->second * sites;
      for (uint32_t k = 0; k < sites; ++k)
      {
        const uint32_t m = tip_mask(s->model.states, row[k]);
        if (!m) return host_fail(EPA_ERR_ARG, "invalid character '" + std::string(1, (char) row[k]) + "' in reference sequence " + label);
        masks[t * sites + k] = m;
      }
    }
    if (s->model.empirical_freqs) s->model.set_empirical_freqs(masks.data(), T, sites);     // epa_pll_util.cpp:55-57
    const Tree::Schedule sch = s->tree.schedule();
    epa_model_desc md{};
    md.states = (uint32_t) s->model.states;
    md.rate_cats = (uint32_t) s->model.rate_cats;
    md.sites = sites;
    md.flags = 0;
    {
      const bool want = rate_mode == 1 || (rate_mode == 2 && T > 2000);
      if (want) md.flags |= EPA_FLAG_RATE_SCALERS | (rate_bug ? EPA_FLAG_BUGCOMPAT_FOCUS : 0u);
    }
    md.eigenvals = s->model.eigenvals.data();
    md.eigenvecs = s->model.eigenvecs.data();
    md.inv_eigenvecs = s->model.inv_eigenvecs.data();
    md.freqs = s->model.freqs.data();
    md.rates = s->model.rates.data();
    // Reference quirk: the tiny partition aliases the reference partition's category RATES but not its rate
    // WEIGHTS (src/tree/tiny_util.cpp:110-111): every placement likelihood uses pll_partition_create's default
    // weights 1/R, whatever +R{rates}{weights} says (the rates are still normalised with the user's weights).
    const std::vector<double> tiny_weights((size_t) s->model.rate_cats, 1.0 / s->model.rate_cats);
    md.rate_weights = tiny_weights.data();
    md.pinv = s->model.pinv;
    int rc = epa_ctx_create(&s->ctx, device, &md, (uint32_t) T, masks.data(), sch.n_slots, sch.edges.data(), (uint32_t) sch.edges.size());
    if (rc) return host_fail(rc, epa_last_error(nullptr));
    rc = epa_compute_clvs(s->ctx, sch.ops.data(), (uint32_t) sch.ops.size());
    if (rc) return host_fail(rc, epa_last_error(s->ctx));
    rc = epa_build_lookup(s->ctx);
    if (rc) return host_fail(rc, epa_last_error(s->ctx));
    *out = s.release();
    return EPA_OK;
  }
  catch (const std::exception & e)
  {
    return host_fail(EPA_ERR_ARG, e.what());
  }
}

extern "C" void epa_session_close(epa_session * s) { delete s; }
extern "C" epa_ctx * epa_session_ctx(epa_session * s) { return s ? s->ctx : nullptr; }
extern "C" uint32_t epa_session_num_edges(const epa_session * s) { return s ? (uint32_t) s->tree.num_edges() : 0; }
extern "C" uint32_t epa_session_num_tips(const epa_session * s) { return s ? (uint32_t) s->tree.num_tips() : 0; }
extern "C" uint32_t epa_session_sites(const epa_session * s) { return s ? s->sites : 0; }

extern "C" const char * epa_session_numbered_newick(epa_session * s, int precision)
{
  if (!s) return "";
  if (s->newick_precision != precision)
  {
    s->newick_cache = s->tree.numbered_newick(precision, s->preserve_rooting);
    s->newick_precision = precision;
  }
  return s->newick_cache.c_str();
}

extern "C" int epa_session_tree_logl(epa_session * s, double * logl)
{
  if (!s || !logl) return host_fail(EPA_ERR_ARG, "null argument");
  const int rc = epa_edge_loglikelihood(s->ctx, 0, logl);
  return rc ? host_fail(rc, epa_last_error(s->ctx)) : EPA_OK;
}

extern "C" int epa_session_place(epa_session * s, const char * query_rows, uint64_t n_queries, const epa_options * opts,
                                 uint32_t chunk_size, epa_placement * out, uint32_t * counts)
{
  if (!s || !opts || (!query_rows && n_queries)) return host_fail(EPA_ERR_ARG, "null argument");
  if ((out == nullptr) != (counts == nullptr)) return host_fail(EPA_ERR_ARG, "out and counts must both be given or both be NULL");
  if (chunk_size == 0) chunk_size = kDefaultChunk;
  if (!opts->prescoring)
  {
    // all-pairs mode: keep a chunk below 2^32 pairs and a few GB of results
    const uint64_t max_q = std::max<uint64_t>(1, (1ull << 28) / std::max<uint32_t>(1, epa_session_num_edges(s)));
    chunk_size = (uint32_t) std::min<uint64_t>(chunk_size, max_q);
  }
  // the record copies of a chunk overlap the next chunk unless the rooted-tree mapping must touch
  // them right away
  const bool defer = !(s->tree.mapper.active && s->preserve_rooting);
  struct DeferGuard {
    epa_ctx * c; bool on;
    DeferGuard(epa_ctx * c_, bool on_) : c(c_), on(on_) { if (on) epa_set_deferred_results(c, 1); }
    ~DeferGuard() { if (on) { epa_wait_results(c); epa_set_deferred_results(c, 0); } }
  } guard(s->ctx, defer);
  // Chunk schedule: a short first chunk (its upload cannot overlap anything) and a short last one
  // (neither can the copy of its records), full chunks in between. Results do not depend on the cut.
  std::vector<uint32_t> sizes;
  {
    uint64_t left = n_queries;
    const uint32_t small = std::min<uint32_t>(chunk_size, std::max<uint32_t>(4096u, chunk_size / 8));    // never above the cap
    if (left > (uint64_t) chunk_size + 2 * (uint64_t) small)
    {
      sizes.push_back(small); left -= small;
      while (left > (uint64_t) chunk_size + small) { sizes.push_back(chunk_size); left -= chunk_size; }
      if (left > small) { sizes.push_back((uint32_t) (left - small)); left = small; }
      sizes.push_back((uint32_t) left);
    }
    else
      while (left) { const uint32_t c = (uint32_t) std::min<uint64_t>(chunk_size, left); sizes.push_back(c); left -= c; }
  }
  uint64_t done = 0;
  for (size_t ci = 0; ci < sizes.size(); done += sizes[ci], ++ci)
  {
    const uint32_t nq = sizes[ci];
    if (ci + 1 < sizes.size())
      epa_hint_next_chunk(s->ctx, query_rows + (done + nq) * s->sites, sizes[ci + 1]);
    const int rc = epa_place_chunk(s->ctx, query_rows + done * s->sites, nq, opts,
                                   out ? out + done * opts->filter_max : nullptr, counts ? counts + done : nullptr);
    if (rc)
    {
      std::string msg = epa_last_error(s->ctx);
      if (rc == EPA_ERR_QUERY) msg += " (chunk starting at query " + std::to_string(done) + ")";
      return host_fail(rc, msg);
    }
    if (s->tree.mapper.active && s->preserve_rooting && out && counts)
    {
      // rooted input: edge numbers and distal lengths of the rooted tree (rtree_mapper::in_rtree,
      // applied by the reference when it prints a placement, src/io/jplace_util.cpp:20-32)
      for (uint32_t q = 0; q < nq; ++q)
        for (uint32_t k = 0; k < counts[done + q]; ++k)
        {
          epa_placement & p = out[(done + q) * opts->filter_max + k];
          const auto tr = s->tree.mapper.in_rtree((uint32_t) p.branch_id, p.distal_length);
          p.branch_id = tr.first;
          p.distal_length = tr.second;
        }
    }
  }
  return EPA_OK;
}

extern "C" int epa_session_set_preserve_rooting(epa_session * s, int on)
{
  if (!s) return host_fail(EPA_ERR_ARG, "null argument");
  s->preserve_rooting = on != 0;
  s->newick_precision = -1;
  return EPA_OK;
}

extern "C" int epa_session_is_rooted(const epa_session * s) { return s && s->tree.mapper.active ? 1 : 0; }

// ----------------------------------------------------------------------------------------------
//  jplace
// ----------------------------------------------------------------------------------------------
static void write_pquery(FILE * fh, const char * name, const epa_placement * recs, uint32_t count, int precision, bool last)
{
  static_assert(sizeof(epa_host::PlacementFields) == sizeof(epa_placement), "record layout");
  std::string out;
  epa_host::append_pquery(out, name, std::strlen(name), reinterpret_cast<const epa_host::PlacementFields *>(recs), count, precision, last);
  std::fwrite(out.data(), 1, out.size(), fh);
}

static void jplace_begin(FILE * fh, const char * newick)
{
  std::string out = "{\n  \"tree\": \"";
  json_escape(out, newick, std::strlen(newick));
  out += "\",\n  \"placements\": \n  [\n";
  std::fwrite(out.data(), 1, out.size(), fh);
}

static void jplace_end(FILE * fh, const char * invocation)
{
  std::string out = "  ],\n  \"metadata\": {\"invocation\": \"";
  if (invocation) json_escape(out, invocation, std::strlen(invocation));
  out += "\"},\n  \"version\": 3,\n  \"fields\": [\"edge_num\", \"likelihood\", \"like_weight_ratio\", \"distal_length\", \"pendant_length\"]\n}\n";
  std::fwrite(out.data(), 1, out.size(), fh);
}

extern "C" int epa_write_jplace(const char * path, const char * numbered_newick, const char * invocation,
                                const char * const * query_names, uint64_t n_queries, const epa_placement * recs,
                                const uint32_t * counts, uint32_t stride, int precision)
{
  if (!path || !numbered_newick || !query_names || !recs || !counts) return host_fail(EPA_ERR_ARG, "null argument");
  FILE * fh = std::fopen(path, "w");
  if (!fh) return host_fail(EPA_ERR_ARG, std::string("cannot open ") + path);
  if (precision > 18) precision = 18;
  jplace_begin(fh, numbered_newick);
  for (uint64_t q = 0; q < n_queries; ++q)
    write_pquery(fh, query_names[q], recs + q * stride, counts[q], precision, q + 1 == n_queries);
  jplace_end(fh, invocation);
  std::fclose(fh);
  return EPA_OK;
}

// ----------------------------------------------------------------------------------------------
//  whole run (src/main.cpp:470-552 + src/core/place.cpp:173-251)
// ----------------------------------------------------------------------------------------------
extern "C" int epa_run_files(const char * tree_file, const char * ref_msa_file, const char * query_file,
                             const char * model, const char * outdir, const epa_options * opts, uint32_t chunk_size,
                             int precision, int device, const char * invocation)
{
  return epa_run_files_ex(tree_file, ref_msa_file, query_file, model, outdir, opts, chunk_size, precision, device,
                          invocation, 1);
}

extern "C" int epa_run_files_ex(const char * tree_file, const char * ref_msa_file, const char * query_file,
                                const char * model, const char * outdir, const epa_options * opts, uint32_t chunk_size,
                                int precision, int device, const char * invocation, int preserve_rooting)
{
  // one device: the same reader / device / writer pipeline as the multi-GPU run (pipeline.cpp)
  return epa_run_files_multi(tree_file, ref_msa_file, query_file, model, outdir, opts, chunk_size, precision, &device, 1,
                             invocation, preserve_rooting, 0, nullptr);
}

// ----------------------------------------------------------------------------------------------
//  device-free inspection (CPU tests of the host logic)
// ----------------------------------------------------------------------------------------------
extern "C" int epa_host_parse_tree(const char * newick, int precision, char * out_newick, size_t cap, uint32_t * n_tips,
                                   uint32_t * n_edges)
{
  if (!newick) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const Tree t = Tree::parse(newick);
    if (out_newick && cap)
    {
      const std::string s = t.numbered_newick(precision);
      std::snprintf(out_newick, cap, "%s", s.c_str());
    }
    if (n_tips) *n_tips = (uint32_t) t.num_tips();
    if (n_edges) *n_edges = (uint32_t) t.num_edges();
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_read_alignment(const char * path, uint32_t * n_sequences, uint32_t * sites, char * rows,
                                       size_t rows_cap, char * labels, size_t labels_cap)
{
  if (!path) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const Alignment a = read_alignment(path);
    if (n_sequences) *n_sequences = (uint32_t) a.size();
    if (sites) *sites = (uint32_t) a.sites;
    if (rows)
    {
      if (rows_cap < a.rows.size()) return host_fail(EPA_ERR_ARG, "rows buffer too small");
      std::memcpy(rows, a.rows.data(), a.rows.size());
    }
    if (labels && labels_cap)
    {
      std::string all;
      for (const auto & nm : a.names) { all += nm; all += '\n'; }
      if (all.size() + 1 > labels_cap) return host_fail(EPA_ERR_ARG, "labels buffer too small: " + std::to_string(all.size() + 1) + " bytes needed");
      std::memcpy(labels, all.c_str(), all.size() + 1);
    }
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_read_alignment_mt(const char * path, int threads, int want_mask, uint32_t * n_sequences, uint32_t * sites,
                                          char * rows, size_t rows_cap, char * labels, size_t labels_cap, uint8_t * gap_mask_out)
{
  if (!path) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const MappedFile file(path);
    const QueryIndex idx = index_queries(file, path, threads, want_mask != 0);
    if (n_sequences) *n_sequences = (uint32_t) idx.records.size();
    if (sites) *sites = (uint32_t) idx.sites;
    if (gap_mask_out) std::memcpy(gap_mask_out, idx.gap_mask.data(), idx.sites);
    if (rows)
    {
      if (rows_cap < idx.records.size() * idx.sites) return host_fail(EPA_ERR_ARG, "rows buffer too small");
      std::vector<uint32_t> keep(idx.sites);
      for (size_t s = 0; s < idx.sites; ++s) keep[s] = (uint32_t) s;
      decode_rows(idx, 0, idx.records.size(), keep, reinterpret_cast<uint8_t *>(rows), threads);
    }
    if (labels && labels_cap)
    {
      std::string all;
      for (const auto & r : idx.records) { all.append(r.name, r.name_len); all += '\n'; }
      if (all.size() + 1 > labels_cap) return host_fail(EPA_ERR_ARG, "labels buffer too small: " + std::to_string(all.size() + 1) + " bytes needed");
      std::memcpy(labels, all.c_str(), all.size() + 1);
    }
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_format_fixed(double value, int precision, char * out, size_t cap)
{
  if (!out || cap < 400) return host_fail(EPA_ERR_ARG, "buffer of at least 400 bytes expected");
  if (precision > 18) precision = 18;
  const size_t n = format_fixed(out, value, precision);
  out[n] = 0;
  return (int) n;
}

extern "C" int epa_host_fasta_to_bfast(const char * fasta_path, const char * out_dir, char * out_path, size_t cap)
{
  if (!fasta_path || !out_dir) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const std::string written = write_bfast(read_fasta(fasta_path), fasta_path, out_dir);
    if (out_path && cap) std::snprintf(out_path, cap, "%s", written.c_str());
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_empirical_frequencies(const char * model_desc, const uint32_t * tip_masks, uint32_t n_tips, uint32_t sites,
                                              double * freqs, double * eigenvals)
{
  if (!model_desc || !tip_masks || !freqs) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    Model m = Model::parse(model_desc);
    m.set_empirical_freqs(tip_masks, n_tips, sites);
    for (int k = 0; k < m.states; ++k) { freqs[k] = m.freqs[(size_t) k]; if (eigenvals) eigenvals[k] = m.eigenvals[(size_t) k]; }
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_model_from_file(const char * path, char * out, size_t cap)
{
  if (!path || !out || !cap) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const std::string desc = model_string_from_file(path);
    if (desc.size() + 1 > cap) return host_fail(EPA_ERR_ARG, "model string buffer too small");
    std::snprintf(out, cap, "%s", desc.c_str());
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_map_rooted(const char * newick, uint32_t * edges, double * distal, uint32_t count,
                                   char * out_unrooted_newick, size_t cap)
{
  if (!newick) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const Tree t = Tree::parse(newick);
    if (!t.mapper.active) return host_fail(EPA_ERR_ARG, "the tree is not rooted");
    for (uint32_t i = 0; i < count; ++i)
    {
      if (edges[i] >= t.num_edges()) return host_fail(EPA_ERR_ARG, "edge out of range");
      const auto tr = t.mapper.in_rtree(edges[i], distal[i]);
      edges[i] = tr.first;
      distal[i] = tr.second;
    }
    if (out_unrooted_newick && cap) std::snprintf(out_unrooted_newick, cap, "%s", t.numbered_newick(10, false).c_str());
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_tree_schedule(const char * newick, uint32_t * n_slots, epa_clv_op * ops, uint32_t ops_cap,
                                      uint32_t * n_ops, epa_edge_desc * edges, uint32_t edges_cap, uint32_t * n_edges,
                                      char * tip_labels, size_t labels_cap)
{
  if (!newick) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const Tree t = Tree::parse(newick);
    const Tree::Schedule s = t.schedule();
    if (s.ops.size() > ops_cap || s.edges.size() > edges_cap) return host_fail(EPA_ERR_ARG, "capacity too small");
    if (n_slots) *n_slots = s.n_slots;
    if (n_ops) *n_ops = (uint32_t) s.ops.size();
    if (n_edges) *n_edges = (uint32_t) s.edges.size();
    if (ops) std::copy(s.ops.begin(), s.ops.end(), ops);
    if (edges) std::copy(s.edges.begin(), s.edges.end(), edges);
    if (tip_labels && labels_cap)
    {
      std::string all;
      for (int v : t.tip_node) { all += t.nodes[v].label; all += '\n'; }
      std::snprintf(tip_labels, labels_cap, "%s", all.c_str());
    }
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}

extern "C" int epa_host_parse_model(const char * model, uint32_t * states, uint32_t * rate_cats, double * rates,
                                    double * weights, double * freqs, double * eigenvals, double * eigenvecs,
                                    double * inv_eigenvecs)
{
  if (!model) return host_fail(EPA_ERR_ARG, "null argument");
  try
  {
    const Model m = Model::parse(model);
    if (states) *states = (uint32_t) m.states;
    if (rate_cats) *rate_cats = (uint32_t) m.rate_cats;
    if (rates) std::copy(m.rates.begin(), m.rates.end(), rates);
    if (weights) std::copy(m.weights.begin(), m.weights.end(), weights);
    if (freqs) std::copy(m.freqs.begin(), m.freqs.end(), freqs);
    if (eigenvals) std::copy(m.eigenvals.begin(), m.eigenvals.end(), eigenvals);
    if (eigenvecs) std::copy(m.eigenvecs.begin(), m.eigenvecs.end(), eigenvecs);
    if (inv_eigenvecs) std::copy(m.inv_eigenvecs.begin(), m.inv_eigenvecs.end(), inv_eigenvecs);
    return EPA_OK;
  }
  catch (const std::exception & e) { return host_fail(EPA_ERR_ARG, e.what()); }
}
