// pipeline.cpp - files -> jplace as a three-stage pipeline over one or several GPUs of the box.
//
// What the reference overlaps with two helper tasks - the read of the next chunk
// (src/seq/MSA_Stream.cpp:79-85) and the write of the previous chunk's jplace text
// (src/io/jplace_writer.hpp:58-69) - and spreads over MPI ranks by query blocks
// (src/net/epa_mpi_util.cpp:10-30, shared output src/io/jplace_writer.hpp:92-132) runs here as
//
//   reader  : query file memory-mapped and indexed once by all host threads (record table, width
//             check, all-gap column mask = the reference's first pass, src/seq/MSA_Info.hpp:22-111);
//             then chunk k+1 is decoded (FASTA text or bfast nibbles -> upper-case rows, pre-mask
//             applied) straight into pinned staging memory
//   devices : one host thread + one epa_session per GPU; each takes the next decoded chunk and
//             runs the chunk loop on it (epa_session_place: H2D of the next device chunk and D2H of
//             the previous one overlap the kernels); records land in the chunk's pinned result
//             buffer - the D2H copy is the "gather", no collective is needed inside one process
//   writer  : formats chunk k-1 on several threads (printf-exact fixed-point digits without printf)
//             and appends it to the jplace in input order
//
// Results do not depend on the number of devices, threads or the chunk size.
#include "../../../include/epa_b200_host.h"

#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>

#include "fastio.hpp"
#include "model.hpp"
#include "seqio.hpp"
#include "session_internal.hpp"

using namespace epa_host;

namespace epa_host {
int host_fail_msg(int code, const std::string & msg);      // session.cpp
}

namespace {

// Process-wide pool of page-locked staging blocks. Page-locking runs at 2-3 GB/s and holds up other CUDA
// calls while it does, so the blocks of a finished run are kept for the next one (a service placing one
// query file after the other pays once); epa_host_release_pinned_pool() gives them back.
struct PinnedPool {
  std::mutex m;
  std::vector<std::pair<void *, size_t>> idle;
  void * get(size_t bytes, size_t * cap)
  {
    {
      std::lock_guard<std::mutex> lk(m);
      size_t best = idle.size();
      for (size_t i = 0; i < idle.size(); ++i)
        if (idle[i].second >= bytes && (best == idle.size() || idle[i].second < idle[best].second)) best = i;
      if (best != idle.size())
      {
        void * p = idle[best].first;
        *cap = idle[best].second;
        idle.erase(idle.begin() + (long) best);
        return p;
      }
    }
    void * p = nullptr;
    if (epa_pinned_alloc(&p, bytes) != EPA_OK) throw std::runtime_error("cannot allocate pinned host memory");
    *cap = bytes;
    return p;
  }
  void put(void * p, size_t cap) { std::lock_guard<std::mutex> lk(m); idle.emplace_back(p, cap); }
  void release()
  {
    std::lock_guard<std::mutex> lk(m);
    for (auto & b : idle) epa_pinned_free(b.first);
    idle.clear();
  }
  ~PinnedPool() { /* blocks die with the process: the CUDA runtime may already be gone here */ }
};
PinnedPool g_pinned;

struct Slot {
  uint8_t * rows = nullptr;            // pinned [cap][width]
  epa_placement * recs = nullptr;      // pinned [cap][filter_max]
  uint32_t * counts = nullptr;         // pinned [cap]
  size_t first = 0, count = 0;
  size_t index = 0;                    // chunk number
};

}  // namespace

extern "C" void epa_host_release_pinned_pool(void) { g_pinned.release(); }

extern "C" int epa_run_files_multi(const char * tree_file, const char * ref_msa_file, const char * query_file,
                                   const char * model, const char * outdir, const epa_options * opts, uint32_t chunk_size,
                                   int precision, const int * devices, uint32_t n_devices, const char * invocation,
                                   int preserve_rooting, int host_threads, epa_run_stats * stats)
{
  if (!tree_file || !ref_msa_file || !query_file || !model || !outdir || !opts || !devices || n_devices == 0)
    return host_fail_msg(EPA_ERR_ARG, "null argument");
  if (precision > 18) precision = 18;
  try
  {
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    std::string dir = outdir;
    if (dir.empty()) dir = ".";
    if (dir.back() != '/') dir += '/';
    ::mkdir(dir.c_str(), 0755);
    std::ofstream log(dir + "epa_info.log");
    std::mutex log_mutex;
    auto info = [&](const std::string & line)
    {
      std::lock_guard<std::mutex> lk(log_mutex);
      log << "INFO " << line << "\n";
      std::printf("INFO %s\n", line.c_str());
    };
    if (host_threads <= 0) host_threads = (int) std::max(1u, std::thread::hardware_concurrency());
    info("Selected: Output dir: " + dir);
    info(std::string("Selected: Query file: ") + query_file);
    info(std::string("Selected: Tree file: ") + tree_file);
    info(std::string("Selected: Reference MSA: ") + ref_msa_file);
    info(std::string("Selected: Specified model: ") + model);
    {
      std::string d;
      for (uint32_t i = 0; i < n_devices; ++i) d += (i ? "," : "") + std::to_string(devices[i]);
      info("Selected: device(s) cuda:" + d + " (libepa_b200, sm_100a), " + std::to_string(host_threads) + " host threads");
    }

    std::ifstream tf(tree_file);
    if (!tf) return host_fail_msg(EPA_ERR_ARG, std::string("Cannot open file: ") + tree_file);
    const std::string newick((std::istreambuf_iterator<char>(tf)), std::istreambuf_iterator<char>());
    Alignment ref = read_fasta(ref_msa_file);

    // ---- shared pipeline state ----
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Slot *> free_slots, ready;
    std::map<size_t, Slot *> placed;
    bool reader_done = false;
    int rc_all = EPA_OK;
    std::string err_all;
    auto failed = [&]() { return rc_all != EPA_OK; };
    auto set_error = [&](int rc, const std::string & msg)
    {
      std::lock_guard<std::mutex> lk(mu);
      if (rc_all == EPA_OK) { rc_all = rc; err_all = msg; }
      cv.notify_all();
    };


    // Page-locked staging memory. Its size only needs upper bounds (alignment width of the reference MSA, queries
    // estimated from the file size), so the allocation starts now and overlaps the indexing of the query file.
    MappedFile qfile(query_file);
    if (chunk_size == 0) chunk_size = 131072;
    const uint32_t fmax = opts->filter_max;
    const size_t ref_sites0 = ref.sites;               // (ref is column-masked below)
    const size_t q_upper = qfile.size() / std::max<size_t>(1, ref_sites0 / 2) + 1;
    const size_t slot_q = (size_t) std::min<uint64_t>(q_upper, chunk_size);
    const size_t slots_upper = std::min<size_t>((q_upper + slot_q - 1) / slot_q, 3 * (size_t) n_devices + 2);
    std::vector<Slot> slots(slots_upper);
    // until the query file is indexed only the slots every run needs are made; the final number follows the chunk count
    size_t slots_wanted = std::min<size_t>(slots_upper, (size_t) n_devices + 2);
    bool slots_final = false;
    struct PinnedGuard {
      std::mutex m;
      std::vector<std::pair<void *, size_t>> p;
      ~PinnedGuard() { for (auto & b : p) g_pinned.put(b.first, b.second); }
      void * get(size_t bytes)
      {
        size_t cap = 0;
        void * q = g_pinned.get(bytes, &cap);
        std::lock_guard<std::mutex> lk(m);
        p.emplace_back(q, cap);
        return q;
      }
    } pinned;
    std::thread allocator([&]()
    {
      try
      {
        for (size_t i = 0; i < slots.size(); ++i)
        {
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return i < slots_wanted || slots_final || rc_all != EPA_OK; });
            if (i >= slots_wanted || rc_all != EPA_OK) break;
          }
          Slot & sl = slots[i];
          sl.rows = static_cast<uint8_t *>(pinned.get(std::max<size_t>(1, slot_q * ref_sites0)));
          sl.recs = static_cast<epa_placement *>(pinned.get(std::max<size_t>(1, slot_q * fmax * sizeof(epa_placement))));
          sl.counts = static_cast<uint32_t *>(pinned.get(std::max<size_t>(1, slot_q * sizeof(uint32_t))));
          { std::lock_guard<std::mutex> lk(mu); free_slots.push_back(&sl); }
          cv.notify_all();
        }
      }
      catch (const std::exception & e) { set_error(EPA_ERR_NOMEM, e.what()); }
    });
    struct Joiner {
      std::thread & t; std::mutex & m; size_t & wanted;
      bool & final_; std::condition_variable & c;
      ~Joiner() { { std::lock_guard<std::mutex> lk(m); if (!final_) { wanted = 0; final_ = true; } } c.notify_all(); if (t.joinable()) t.join(); }
    } allocator_joiner{allocator, mu, slots_wanted, slots_final, cv};

    // first pass over the queries: records, width, all-gap columns
    const QueryIndex qidx = index_queries(qfile, query_file, host_threads, opts->premasking != 0);
    if (ref.sites != qidx.sites)
      return host_fail_msg(EPA_ERR_ARG, "reference and query MSA have different widths (" + std::to_string(ref.sites) + " vs " +
                                        std::to_string(qidx.sites) + ")");
    std::vector<uint32_t> keep;
    if (opts->premasking)
    {
      std::vector<uint8_t> mask = gap_mask(ref);                         // src/main.cpp:470-494
      for (size_t i = 0; i < mask.size(); ++i) mask[i] |= qidx.gap_mask[i];
      for (size_t s = 0; s < mask.size(); ++s) if (!mask[s]) keep.push_back((uint32_t) s);
      ref = apply_mask(ref, mask);
    }
    else
      for (size_t s = 0; s < qidx.sites; ++s) keep.push_back((uint32_t) s);
    const size_t width = keep.size();
    const double t_index = since(t0);

    // -m takes a model string or a RAxML 8 info / raxml-ng bestModel / IQ-TREE report file (src/main.cpp:433-436)
    std::string model_desc_str = model;
    {
      struct stat st;
      if (stat(model, &st) == 0 && S_ISREG(st.st_mode))
      {
        model_desc_str = model_string_from_file(model);
        info("Selected: Specified model file: " + std::string(model));
        info("  ==> model " + model_desc_str);
      }
    }
    const Model parsed = Model::parse(model_desc_str);
    info("Using model parameters:");
    info(parsed.describe());

    // reference state on every device (built concurrently)
    std::vector<const char *> names(ref.size());
    for (size_t i = 0; i < ref.size(); ++i) names[i] = ref.names[i].c_str();
    struct SessionDeleter { void operator()(epa_session * s) const { epa_session_close(s); } };
    std::vector<std::unique_ptr<epa_session, SessionDeleter>> sessions(n_devices);
    {
      std::vector<int> rcs(n_devices, EPA_OK);
      std::vector<std::string> msgs(n_devices);
      std::vector<std::thread> th;
      for (uint32_t d = 0; d < n_devices; ++d)
        th.emplace_back([&, d]()
        {
          epa_session * s = nullptr;
          rcs[d] = epa_session_open(&s, newick.c_str(), (uint32_t) ref.size(), names.data(),
                                    reinterpret_cast<const char *>(ref.rows.data()), (uint32_t) ref.sites, model_desc_str.c_str(), devices[d]);
          if (rcs[d]) msgs[d] = epa_host_last_error();
          else { sessions[d].reset(s); epa_session_set_preserve_rooting(s, preserve_rooting); }
        });
      for (auto & t : th) t.join();
      for (uint32_t d = 0; d < n_devices; ++d)
        if (rcs[d]) return host_fail_msg(rcs[d], msgs[d]);
    }
    epa_session * s0 = sessions[0].get();
    if (epa_session_is_rooted(s0))
      info(preserve_rooting ? "Selected: Preserving the root of the input tree" : "Selected: Unrooting the input tree");
    double tree_logl = 0.0;
    if (epa_session_tree_logl(s0, &tree_logl) == EPA_OK)
    {
      char buf[96];
      std::snprintf(buf, sizeof buf, "Reference tree log-likelihood: %.6f", tree_logl);
      info(buf);
    }
    const double t_setup = since(t0);

    const std::string jpath = dir + "epa_result.jplace";
    // the jplace goes out through positioned writes: the formatting threads of a chunk write their own parts side by
    // side (a single writer copying 156 MB per 10^6 queries into the page cache was the slowest host stage)
    struct Fd {
      int fd = -1;
      ~Fd() { if (fd >= 0) ::close(fd); }
    } fh;
    fh.fd = ::open(jpath.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fh.fd < 0) return host_fail_msg(EPA_ERR_ARG, "cannot open " + jpath);
    uint64_t file_off = 0;
    std::atomic<bool> write_failed{false};
    auto write_at = [&](const char * data, size_t n, uint64_t at)
    {
      while (n)
      {
        const ssize_t k = ::pwrite(fh.fd, data, n, (off_t) at);
        if (k <= 0) { if (k < 0 && errno == EINTR) continue; write_failed = true; return; }
        data += k; n -= (size_t) k; at += (uint64_t) k;
      }
    };
    info("Output file: " + jpath);
    {
      std::string head = "{\n  \"tree\": \"";
      const char * nw = epa_session_numbered_newick(s0, precision);
      json_escape(head, nw, std::strlen(nw));
      head += "\",\n  \"placements\": \n  [\n";
      write_at(head.data(), head.size(), file_off);
      file_off += head.size();
    }

    const uint64_t Q = qidx.records.size();
    if (!opts->prescoring)
    {
      // all-pairs mode: keep a chunk below 2^32 pairs and a few GB of results (as epa_session_place does)
      const uint64_t max_q = std::max<uint64_t>(1, (1ull << 28) / std::max<uint32_t>(1, epa_session_num_edges(s0)));
      chunk_size = (uint32_t) std::min<uint64_t>(chunk_size, max_q);
    }
    // one pipeline slot = one device chunk. Per device: one chunk in the kernels, one whose records are
    // still being copied out, one prefetched; plus one being decoded and one being formatted.
    const size_t pchunk = (size_t) std::min<uint64_t>(std::max<uint64_t>(Q, 1), (uint64_t) chunk_size);
    const size_t n_chunks = (size_t) ((Q + pchunk - 1) / pchunk);
    {
      std::lock_guard<std::mutex> lk(mu);
      slots_wanted = std::min<size_t>(slots_upper, std::max<size_t>(1, std::min<size_t>(n_chunks, 3 * (size_t) n_devices + 2)));
      slots_final = true;
    }
    cv.notify_all();

    const bool debug = std::getenv("EPA_B200_PIPE_DEBUG") != nullptr;
    const auto t1 = std::chrono::steady_clock::now();
    // the decode and format stages share the host threads (each is busy for a fraction of a chunk's device time)
    const int side_threads = std::max(1, std::min(host_threads, std::max(2, host_threads / 2)));
    double busy_read = 0.0, busy_write = 0.0;
    std::vector<double> busy_dev(n_devices, 0.0);

    std::thread reader([&]()
    {
      try
      {
        for (size_t c = 0; c < n_chunks; ++c)
        {
          Slot * sl = nullptr;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return !free_slots.empty() || failed(); });
            if (failed()) break;
            sl = free_slots.front(); free_slots.pop_front();
          }
          const auto ta = std::chrono::steady_clock::now();
          sl->index = c; sl->first = c * pchunk; sl->count = (size_t) std::min<uint64_t>(pchunk, Q - sl->first);
          decode_rows(qidx, sl->first, sl->count, keep, sl->rows, side_threads);
          busy_read += since(ta);
          if (debug) std::fprintf(stderr, "[pipe] decoded chunk %zu: %.1f ms (at %.3f s)\n", c, since(ta) * 1e3, since(t0));
          {
            std::lock_guard<std::mutex> lk(mu);
            ready.push_back(sl);
          }
          cv.notify_all();
        }
      }
      catch (const std::exception & e) { set_error(EPA_ERR_ARG, e.what()); }
      { std::lock_guard<std::mutex> lk(mu); reader_done = true; }
      cv.notify_all();
    });

    // rooted input: edge numbers and distal lengths of the rooted tree (rtree_mapper::in_rtree, applied by the
    // reference when it prints a placement, src/io/jplace_util.cpp:20-32)
    auto finish_slot = [&](epa_session * s, Slot * sl)
    {
      if (s->tree.mapper.active && s->preserve_rooting)
        for (size_t q = 0; q < sl->count; ++q)
          for (uint32_t k = 0; k < sl->counts[q]; ++k)
          {
            epa_placement & p = sl->recs[q * fmax + k];
            const auto tr = s->tree.mapper.in_rtree((uint32_t) p.branch_id, p.distal_length);
            p.branch_id = tr.first;
            p.distal_length = tr.second;
          }
      { std::lock_guard<std::mutex> lk(mu); placed[sl->index] = sl; }
      cv.notify_all();
    };

    std::vector<std::thread> workers;
    for (uint32_t d = 0; d < n_devices; ++d)
      workers.emplace_back([&, d]()
      {
        epa_session * s = sessions[d].get();
        epa_ctx * ctx = s->ctx;
        epa_set_deferred_results(ctx, 1);
        auto pop = [&](bool block) -> Slot *
        {
          std::unique_lock<std::mutex> lk(mu);
          if (block) cv.wait(lk, [&]() { return !ready.empty() || reader_done || failed(); });
          if (failed() || ready.empty()) return nullptr;
          Slot * sl = ready.front(); ready.pop_front();
          return sl;
        };
        Slot * cur = nullptr, * prev = nullptr;
        for (;;)
        {
          if (!cur) cur = pop(true);
          if (!cur) break;
          Slot * next = pop(false);               // already decoded: its upload overlaps this chunk's kernels
          const auto ta = std::chrono::steady_clock::now();
          if (next) epa_hint_next_chunk(ctx, reinterpret_cast<const char *>(next->rows), (uint32_t) next->count);
          int rc;
          if (debug && cur->index == 0)
          {
            // first chunk, stage by stage (developer timing: where do the first-use costs go?)
            auto lap = [&](const char * what) { epa_synchronize(ctx); std::fprintf(stderr, "[pipe]   %s done at %.3f s\n", what, since(t0)); };
            rc = epa_upload_queries(ctx, reinterpret_cast<const char *>(cur->rows), (uint32_t) cur->count, opts->premasking); lap("upload");
            if (!rc && opts->prescoring) { rc = epa_preplace(ctx); lap("preplace"); }
            if (!rc) { rc = epa_select(ctx, opts, nullptr); lap("select"); }
            if (!rc) { rc = epa_place_pairs(ctx, opts); lap("place_pairs"); }
            if (!rc) { rc = epa_collect(ctx, opts, cur->recs, cur->counts); lap("collect"); }
          }
          else
            rc = epa_place_chunk(ctx, reinterpret_cast<const char *>(cur->rows), (uint32_t) cur->count, opts, cur->recs, cur->counts);
          if (rc == EPA_OK && prev) rc = epa_wait_older_results(ctx);
          busy_dev[d] += since(ta);
          if (debug) std::fprintf(stderr, "[pipe] dev %u chunk %zu: %.1f ms (at %.3f s, next %s)\n", d, cur->index, since(ta) * 1e3, since(t0), next ? "prefetched" : "none");
          if (rc)
          {
            std::string msg = epa_last_error(ctx);
            if (rc == EPA_ERR_QUERY) msg += " (chunk starting at query " + std::to_string(cur->first) + ")";
            set_error(rc, msg);
            break;
          }
          if (prev) finish_slot(s, prev);          // its records were copied out under this chunk's kernels
          prev = cur; cur = next;
          if (!cur)
          {
            // nothing decoded yet (or the end of the file): do not sit on finished records
            if (epa_wait_results(ctx) == EPA_OK) finish_slot(s, prev);
            prev = nullptr;
          }
        }
        epa_wait_results(ctx);
        epa_set_deferred_results(ctx, 0);
      });

    // writer: this thread
    {
      std::vector<std::string> parts((size_t) side_threads);
      for (size_t c = 0; c < n_chunks && !failed(); ++c)
      {
        Slot * sl = nullptr;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&]() { return placed.count(c) || failed(); });
          if (failed()) break;
          sl = placed[c]; placed.erase(c);
        }
        const auto ta = std::chrono::steady_clock::now();
        const int nt = (int) std::max<size_t>(1, std::min<size_t>((size_t) side_threads, sl->count / 1024 + 1));
        std::vector<std::thread> th;
        std::vector<std::atomic<uint64_t>> sizes((size_t) nt);
        for (auto & z : sizes) z.store(UINT64_MAX, std::memory_order_relaxed);
        for (int t = 0; t < nt; ++t)
          th.emplace_back([&, t]()
          {
            std::string & out = parts[(size_t) t];
            out.clear();
            const size_t lo = sl->count * (size_t) t / (size_t) nt, hi = sl->count * (size_t) (t + 1) / (size_t) nt;
            try
            {
              out.reserve((hi - lo) * 160);
              for (size_t q = lo; q < hi; ++q)
                append_pquery(out, qidx.records[sl->first + q].name, qidx.records[sl->first + q].name_len,
                              reinterpret_cast<const PlacementFields *>(sl->recs + q * fmax), sl->counts[q], precision, sl->first + q + 1 == Q);
            }
            catch (...) { out.clear(); write_failed = true; }       // (the other threads wait for this part's size)
            // this part starts where the parts before it end: their sizes are published as they finish
            sizes[(size_t) t].store(out.size(), std::memory_order_release);
            uint64_t at = file_off;
            for (int u = 0; u < t; ++u)
            {
              uint64_t z;
              while ((z = sizes[(size_t) u].load(std::memory_order_acquire)) == UINT64_MAX) std::this_thread::yield();
              at += z;
            }
            write_at(out.data(), out.size(), at);
          });
        for (auto & t : th) t.join();
        for (int t = 0; t < nt; ++t) file_off += parts[(size_t) t].size();
        if (write_failed) set_error(EPA_ERR_ARG, "cannot write " + jpath);
        busy_write += since(ta);
        if (debug) std::fprintf(stderr, "[pipe] wrote chunk %zu: %.1f ms (at %.3f s)\n", c, since(ta) * 1e3, since(t0));
        info(std::to_string(sl->first + sl->count) + " Sequences done!");
        {
          std::lock_guard<std::mutex> lk(mu);
          free_slots.push_back(sl);
        }
        cv.notify_all();
      }
    }
    reader.join();
    for (auto & w : workers) w.join();
    if (allocator.joinable()) allocator.join();
    if (failed())
      return host_fail_msg(rc_all, err_all);
    {
      std::string tail = "  ],\n  \"metadata\": {\"invocation\": \"";
      if (invocation) json_escape(tail, invocation, std::strlen(invocation));
      tail += "\"},\n  \"version\": 3,\n  \"fields\": [\"edge_num\", \"likelihood\", \"like_weight_ratio\", \"distal_length\", \"pendant_length\"]\n}\n";
      write_at(tail.data(), tail.size(), file_off);
      file_off += tail.size();
    }
    if (write_failed || ::close(fh.fd) != 0) { fh.fd = -1; return host_fail_msg(EPA_ERR_ARG, "cannot write " + jpath); }
    fh.fd = -1;
    const double t_place = since(t1), t_total = since(t0);
    char buf[160];
    std::snprintf(buf, sizeof buf, "Time spent placing: %.3fs", t_place);
    info(buf);
    std::snprintf(buf, sizeof buf, "Elapsed Time: %.3fs", t_total);
    info(buf);
    if (stats)
    {
      stats->n_queries = Q;
      stats->seconds_total = t_total;
      stats->seconds_index = t_index;
      stats->seconds_setup = t_setup - t_index;
      stats->seconds_place = t_place;
      stats->busy_read = busy_read;
      stats->busy_write = busy_write;
      stats->busy_device_max = *std::max_element(busy_dev.begin(), busy_dev.end());
    }
    return EPA_OK;
  }
  catch (const std::exception & e)
  {
    return host_fail_msg(EPA_ERR_ARG, e.what());
  }
}
