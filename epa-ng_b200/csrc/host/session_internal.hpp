// session_internal.hpp - the host layer's session object, shared by session.cpp and pipeline.cpp.
#pragma once
#include <string>

#include "../../../include/epa_b200_host.h"
#include "model.hpp"
#include "tree.hpp"

struct epa_session {
  epa_host::Tree tree;
  epa_host::Model model;
  epa_ctx * ctx = nullptr;
  uint32_t sites = 0;
  std::string newick_cache;
  int newick_precision = -1;
  bool preserve_rooting = true;    // rooted input: report placements on the rooted tree
  ~epa_session() { if (ctx) epa_ctx_destroy(ctx); }
};
