#include "tree.hpp"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace epa_host {

namespace {

struct Parser {
  const std::string & s;
  size_t pos = 0;
  std::vector<TreeNode> & nodes;

  Parser(const std::string & text, std::vector<TreeNode> & n) : s(text), nodes(n) {}

  [[noreturn]] void error(const std::string & what) const
  {
    throw std::runtime_error("Treeparsing failed! " + what + " at position " + std::to_string(pos));
  }

  void skip()
  {
    for (;;)
    {
      while (pos < s.size() && std::isspace((unsigned char) s[pos])) ++pos;
      if (pos < s.size() && s[pos] == '[')
      {
        const size_t end = s.find(']', pos);
        if (end == std::string::npos) error("unterminated comment");
        pos = end + 1;
        continue;
      }
      break;
    }
  }

  std::string label()
  {
    skip();
    std::string out;
    if (pos < s.size() && (s[pos] == '\'' || s[pos] == '"'))
    {
      const char q = s[pos++];
      const size_t end = s.find(q, pos);
      if (end == std::string::npos) error("unterminated quoted label");
      out = s.substr(pos, end - pos);
      pos = end + 1;
      return out;
    }
    while (pos < s.size() && !std::isspace((unsigned char) s[pos]) && !std::strchr("()[],:;", s[pos])) out += s[pos++];
    return out;
  }

  // label and optional ":length" following a subtree
  void annotations(TreeNode & node)
  {
    node.label = label();
    skip();
    if (pos < s.size() && s[pos] == ':')
    {
      ++pos;
      skip();
      char * end = nullptr;
      node.length = std::strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) error("expected a branch length");
      pos = (size_t) (end - s.c_str());
    }
  }

  // iterative descent (caterpillar trees with 10^4+ levels must not overflow the stack)
  int parse_tree()
  {
    skip();
    if (pos >= s.size() || s[pos] != '(') error("expected '('");
    nodes.emplace_back();
    const int root = 0;
    int cur = root;
    ++pos;
    bool expect_subtree = true;
    while (true)
    {
      skip();
      if (pos >= s.size()) error("unexpected end of input");
      if (expect_subtree)
      {
        const int child = (int) nodes.size();
        nodes.emplace_back();
        nodes[child].parent = cur;
        nodes[cur].children.push_back(child);
        if (s[pos] == '(')
        {
          ++pos;
          cur = child;
          continue;           // descend, still expecting a subtree
        }
        annotations(nodes[child]);
        if (nodes[child].label.empty()) error("tip without a label");
        expect_subtree = false;
        continue;
      }
      const char c = s[pos];
      if (c == ',') { ++pos; expect_subtree = true; }
      else if (c == ')')
      {
        ++pos;
        annotations(nodes[cur]);
        if (cur == root) break;
        cur = nodes[cur].parent;
      }
      else error(std::string("unexpected character '") + c + "'");
    }
    skip();
    if (pos >= s.size() || s[pos] != ';') error("expected ';'");
    return root;
  }
};

}  // namespace

namespace {

// post-order over the file order: numbers the edge above every non-root node, collects the tips
void number_edges(std::vector<TreeNode> & nodes, int root, std::vector<int> * edge_node, std::vector<int> * tip_node)
{
  int next_edge = 0, next_tip = 0;
  std::vector<std::pair<int, size_t>> stack;
  stack.emplace_back(root, 0);
  while (!stack.empty())
  {
    auto & [v, k] = stack.back();
    TreeNode & node = nodes[v];
    if (k < node.children.size())
    {
      const int c = node.children[k++];
      stack.emplace_back(c, 0);
      continue;
    }
    if (v != root)
    {
      if (!node.children.empty() && node.children.size() != 2)
        throw std::invalid_argument("Input Tree contains multifurcations (polytomies)!");
      if (node.children.empty()) { node.tip = next_tip++; if (tip_node) tip_node->push_back(v); }
      node.edge = next_edge++;
      if (edge_node) edge_node->push_back(v);
      if (!node.length) node.length = kDefaultBranchLength;     // set_missing_branch_lengths
    }
    stack.pop_back();
  }
}

}  // namespace

Tree Tree::parse(const std::string & newick)
{
  Tree t;
  Parser p(newick, t.nodes);
  t.root = p.parse_tree();
  const size_t top = t.nodes[t.root].children.size();
  if (top == 2)
  {
    // Rooted input. Keep the tree as written (rooted edge numbers for the jplace), then merge the
    // two root edges: the working tree hangs from the left child of the root if that is an inner
    // node (children: its two children, then the right subtree across the merged edge), otherwise
    // from the right child (children: the left tip across the merged edge, then its two children).
    t.rooted_nodes = t.nodes;
    t.rooted_root = t.root;
    number_edges(t.rooted_nodes, t.rooted_root, nullptr, nullptr);
    const int L = t.nodes[t.root].children[0], R = t.nodes[t.root].children[1];
    const bool left = !t.nodes[L].children.empty();
    if (!left && t.nodes[R].children.empty()) throw std::runtime_error("Number of tip nodes too small");
    const double lenL = t.nodes[L].length, lenR = t.nodes[R].length;
    const int hub = left ? L : R, far = left ? R : L;
    TreeNode & h = t.nodes[hub];
    if (left) h.children.push_back(far);
    else h.children.insert(h.children.begin(), far);
    t.nodes[far].parent = hub;
    t.nodes[far].length = lenL + lenR;
    h.parent = -1;
    h.length = 0.0;
    t.nodes[t.root].children.clear();          // the old root node stays unused
    t.root = hub;
    t.mapper.active = true;
    // the lengths the reference hands to the mapper are the ones in the file (no default applied)
    t.mapper.distal_length = left ? lenR : lenL;
    t.mapper.proximal_length = left ? lenL : lenR;
  }
  else if (top != 3) throw std::invalid_argument("Input Tree contains multifurcations (polytomies)!");
  number_edges(t.nodes, t.root, &t.edge_node, &t.tip_node);
  if (t.tip_node.size() < 3) throw std::runtime_error("Number of tip nodes too small");
  if (t.mapper.active)
  {
    // working-tree edge -> rooted edge number: both numberings are post-orders of the same
    // subtrees, the rooted one has one extra edge (the second root edge)
    const int L = t.rooted_nodes[t.rooted_root].children[0], R = t.rooted_nodes[t.rooted_root].children[1];
    t.mapper.map.resize(t.edge_node.size());
    for (size_t e = 0; e < t.edge_node.size(); ++e) t.mapper.map[e] = (uint32_t) t.rooted_nodes[t.edge_node[e]].edge;
    const bool left = (t.root == L);
    const int far = left ? R : L;
    t.mapper.utree_root_edge = (uint32_t) t.nodes[far].edge;
    t.mapper.distal_edge = (uint32_t) t.rooted_nodes[far].edge;
    t.mapper.proximal_edge = (uint32_t) t.rooted_nodes[left ? L : R].edge;
  }
  return t;
}

std::string Tree::numbered_newick(int precision, bool preserve_rooting) const
{
  const bool as_rooted = mapper.active && preserve_rooting;
  const std::vector<TreeNode> & nodes = as_rooted ? this->rooted_nodes : this->nodes;
  const int root = as_rooted ? this->rooted_root : this->root;
  std::string out;
  char buf[64];
  // iterative: (node, state) with state = index of the next child to print
  std::vector<std::pair<int, size_t>> stack;
  stack.emplace_back(root, 0);
  while (!stack.empty())
  {
    auto & [v, k] = stack.back();
    const TreeNode & node = nodes[v];
    if (k == 0 && !node.children.empty()) out += '(';
    if (k < node.children.size())
    {
      if (k > 0) out += ',';
      const int c = node.children[k++];
      stack.emplace_back(c, 0);
      continue;
    }
    if (!node.children.empty()) out += ')';
    out += node.label;
    if (v != root)
    {
      std::snprintf(buf, sizeof buf, ":%.*f{%d}", precision, node.length, node.edge);
      out += buf;
    }
    stack.pop_back();
  }
  out += ';';
  return out;
}

Tree::Schedule Tree::schedule() const
{
  Schedule s;
  const uint32_t T = (uint32_t) num_tips();
  std::vector<uint32_t> down(nodes.size(), UINT32_MAX), up(nodes.size(), UINT32_MAX);
  uint32_t next = T;
  for (size_t v = 0; v < nodes.size(); ++v)
  {
    if ((int) v == root || nodes[v].edge < 0) continue;       // edge < 0: the detached root of a rooted input
    down[v] = nodes[v].tip >= 0 ? (uint32_t) nodes[v].tip : next++;
    up[v] = next++;
  }
  s.n_slots = next - T;
  for (size_t v = 0; v < nodes.size(); ++v)
  {
    if ((int) v == root || nodes[v].edge < 0) continue;
    const TreeNode & node = nodes[v];
    if (node.tip < 0)
    {
      const int a = node.children[0], b = node.children[1];
      s.ops.push_back(epa_clv_op{down[v], down[a], down[b], 0, nodes[a].length, nodes[b].length});
    }
    const int p = node.parent;
    if (p == root)
    {
      int o[2], k = 0;
      for (int c : nodes[p].children) if (c != (int) v) o[k++] = c;
      s.ops.push_back(epa_clv_op{up[v], down[o[0]], down[o[1]], 0, nodes[o[0]].length, nodes[o[1]].length});
    }
    else
    {
      const int sib = nodes[p].children[0] == (int) v ? nodes[p].children[1] : nodes[p].children[0];
      s.ops.push_back(epa_clv_op{up[v], down[sib], up[p], 0, nodes[sib].length, nodes[p].length});
    }
  }
  s.edges.resize(num_edges());
  for (size_t e = 0; e < num_edges(); ++e)
  {
    const int v = edge_node[e];
    s.edges[e] = epa_edge_desc{down[v], up[v], nodes[v].length};   // a tip, if any, is the distal side
  }
  return s;
}

}  // namespace epa_host
