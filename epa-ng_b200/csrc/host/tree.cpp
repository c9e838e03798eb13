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

Tree Tree::parse(const std::string & newick)
{
  Tree t;
  Parser p(newick, t.nodes);
  t.root = p.parse_tree();
  const size_t top = t.nodes[t.root].children.size();
  if (top == 2)
    throw std::runtime_error("Treeparsing failed! rooted reference trees (top-level bifurcation) are not supported yet");
  if (top != 3) throw std::invalid_argument("Input Tree contains multifurcations (polytomies)!");
  // post-order over the file order, iteratively: (node, next child index)
  std::vector<std::pair<int, size_t>> stack;
  stack.emplace_back(t.root, 0);
  while (!stack.empty())
  {
    auto & [v, k] = stack.back();
    TreeNode & node = t.nodes[v];
    if (k < node.children.size())
    {
      const int c = node.children[k++];
      stack.emplace_back(c, 0);
      continue;
    }
    if (v != t.root)
    {
      if (!node.children.empty() && node.children.size() != 2)
        throw std::invalid_argument("Input Tree contains multifurcations (polytomies)!");
      if (node.children.empty()) { node.tip = (int) t.tip_node.size(); t.tip_node.push_back(v); }
      node.edge = (int) t.edge_node.size();
      t.edge_node.push_back(v);
      if (!node.length) node.length = kDefaultBranchLength;     // set_missing_branch_lengths
    }
    stack.pop_back();
  }
  if (t.tip_node.size() < 3) throw std::runtime_error("Number of tip nodes too small");
  return t;
}

std::string Tree::numbered_newick(int precision) const
{
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
    if ((int) v == root) continue;
    down[v] = nodes[v].tip >= 0 ? (uint32_t) nodes[v].tip : next++;
    up[v] = next++;
  }
  s.n_slots = next - T;
  for (size_t v = 0; v < nodes.size(); ++v)
  {
    if ((int) v == root) continue;
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
