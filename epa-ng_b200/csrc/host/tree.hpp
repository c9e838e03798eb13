// tree.hpp - reference tree of the host layer.
//
// The reference keeps libpll's ring-of-unodes structure; here the tree is a plain array of nodes
// hanging from the top-level trifurcation, because all the hot path needs from it are
//   (1) the edge numbering = jplace edge_num: post-order over the newick as written, the edge
//       above the i-th visited node is edge i (utree_query_branches, src/core/pll/pll_util.cpp:182-205,
//       golden strings test/src/pll_util.cpp:134-186),
//   (2) per edge the two directional CLVs looking away from it, i.e. the pruning schedule
//       (precompute_clvs, src/core/pll/epa_pll_util.cpp:62-107), and
//   (3) the numbered newick for the jplace header (src/core/pll/pll_util.cpp:207-352).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "epa_b200.h"

namespace epa_host {

constexpr double kDefaultBranchLength = 0.10536051565782630123;   // -ln(0.9), src/util/constants.hpp:12

struct TreeNode {
  int parent = -1;
  std::vector<int> children;      // file order
  std::string label;
  double length = 0.0;            // edge above the node
  int edge = -1;                  // post-order index of the edge above the node (= jplace edge_num)
  int tip = -1;                   // tip index (order of appearance in the post-order), -1 for inner nodes
};

// Translation of placements on the unrooted working tree back onto a ROOTED input tree
// (determine_edge_num_translation, src/io/file_io.cpp:60-118; rtree_mapper,
// src/core/pll/rtree_mapper.hpp:38-61; golden values test/src/rtree_mapper.cpp:58-102).
struct RootMapper {
  bool active = false;
  uint32_t utree_root_edge = 0;   // working-tree edge that holds the input tree's root
  uint32_t proximal_edge = 0;     // rooted edge numbers of the two root edges ...
  uint32_t distal_edge = 0;
  double proximal_length = -1.0;  // ... and their lengths
  double distal_length = -1.0;
  std::vector<uint32_t> map;      // working-tree edge -> rooted edge number

  // (edge_num, distal_length) of a placement in the rooted tree
  std::pair<uint32_t, double> in_rtree(uint32_t branch, double distal) const
  {
    if (branch != utree_root_edge) return {map[branch], distal};
    if (distal > distal_length) return {proximal_edge, proximal_length - (distal - distal_length)};
    return {distal_edge, distal};
  }
};

struct Tree {
  std::vector<TreeNode> nodes;
  int root = -1;                  // top-level trifurcation of the (unrooted) working tree
  // rooted input: the tree as written, with its own post-order edge numbers, and the mapper
  std::vector<TreeNode> rooted_nodes;
  int rooted_root = -1;
  RootMapper mapper;
  std::vector<int> edge_node;     // edge -> node below it
  std::vector<int> tip_node;      // tip index -> node

  size_t num_tips() const { return tip_node.size(); }
  size_t num_edges() const { return edge_node.size(); }

  // Parses a strictly binary newick tree: unrooted (top-level trifurcation) or rooted (top-level
  // bifurcation, unrooted the way pll_rtree_unroot + build_tree_from_file do, src/io/file_io.cpp:129-173).
  // Throws std::runtime_error.
  static Tree parse(const std::string & newick);
  // jplace tree string; for rooted input the ROOTED tree with rooted edge numbers unless
  // preserve_rooting is off (--preserve-rooting, src/main.cpp)
  std::string numbered_newick(int precision = 10, bool preserve_rooting = true) const;

  // Node ids handed to libepa_b200: tips 0..T-1; down-CLV of inner node v and up-CLV of node v get
  // consecutive slots. ops/edges are ready for epa_compute_clvs / epa_ctx_create.
  struct Schedule {
    uint32_t n_slots = 0;
    std::vector<epa_clv_op> ops;
    std::vector<epa_edge_desc> edges;
  };
  Schedule schedule() const;
};

}  // namespace epa_host
