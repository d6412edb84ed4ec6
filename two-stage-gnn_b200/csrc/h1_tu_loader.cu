// H1 -- TU-format dataset loader straight into the packed corpus layout.  Host code only.
//
// Replaces `read_graphfile` (Code/sage+gat+diffpool/load_data.py:12-126 == Code/eigengcn/load_data.py): text ->
// python dicts -> networkx graphs -> (later) dense N x N matrices per graph, and PyG's TUDataset reader used by
// Code/sag (train*.py:161-163).  One pass over DS_graph_indicator.txt / DS_A.txt / DS_graph_labels.txt /
// DS_node_labels.txt / DS_node_attributes.txt, output = the ragged arrays the packer and K0/K1 consume
// (node_ptr, edge_ptr, local row/col, node labels, graph labels, optional attributes).
//
// mode TSG_TU_NETWORKX reproduces the reference's networkx semantics exactly:
//   * a graph's node set = the endpoints of ITS edges (graph of an edge = graph_indicator[first endpoint],
//     load_data.py:79-80); nodes without edges do not exist (nx.from_edgelist, :89);
//   * node order = first appearance in the graph's edge list, endpoint e0 before e1 (networkx insertion
//     order, relabelled 0.. in that order, :109-122);
//   * undirected simple graph: duplicates merged, a self loop (u,u) kept once;
//   * node label = value - 1, negative wraps like a python index (:31-33, :100-102); graph labels are
//     renumbered in order of first appearance (:54-66); graphs with more than max_nodes nodes are dropped (:90-91).
// mode TSG_TU_PYG reproduces torch_geometric.io.read_tu_data as used by TUDataset(root, name):
//   * every node of the indicator file exists, in file order; edges relabelled, self loops removed, coalesced;
//   * node label = value - min(value); graph label = rank among the sorted distinct values.
// Output edges are symmetric as stored in the files for PYG (TU files list both directions) and symmetrised
// for NETWORKX, always sorted lexicographically by (row, col) inside a graph.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

struct TuCorpus {
  std::vector<int64_t> node_ptr, edge_ptr, row, col, y;
  std::vector<int32_t> label;
  std::vector<float> attr;
  int64_t num_node_labels = 0, attr_dim = 0, num_classes = 0;
};

bool read_file(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize((size_t)n);
  size_t got = n > 0 ? fread(&(*out)[0], 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}

// every integer of the text, in order (separators: anything that is not part of a number)
void parse_ints(const std::string& s, std::vector<int64_t>* out) {
  const char* p = s.data();
  const char* e = p + s.size();
  while (p < e) {
    while (p < e && !((*p >= '0' && *p <= '9') || *p == '-' || *p == '+')) ++p;
    if (p >= e) break;
    char* q;
    long long v = strtoll(p, &q, 10);
    if (q == p) { ++p; continue; }
    out->push_back((int64_t)v);
    p = q;
  }
}

// rows of floats; returns the column count of the first non-empty line (0 if none)
int64_t parse_float_rows(const std::string& s, std::vector<float>* out) {
  int64_t dim = 0, cur = 0;
  const char* p = s.data();
  const char* e = p + s.size();
  while (p < e) {
    if (*p == '\n') { if (cur > 0 && dim == 0) dim = cur; cur = 0; ++p; continue; }
    if (*p == ',' || *p == ' ' || *p == '\t' || *p == '\r') { ++p; continue; }
    char* q;
    float v = strtof(p, &q);
    if (q == p) { ++p; continue; }
    out->push_back(v); ++cur;
    p = q;
  }
  if (cur > 0 && dim == 0) dim = cur;
  return dim;
}

}  // namespace

struct tsg_tu_handle { TuCorpus c; };

static int tu_load_impl(const char* prefix, int mode, int64_t max_nodes, tsg_tu_handle** out);

extern "C" int tsg_tu_load(const char* prefix, int mode, int64_t max_nodes, tsg_tu_handle** out) {
  try {
    return tu_load_impl(prefix, mode, max_nodes, out);
  } catch (...) {                                   // nothing may propagate across the C ABI
    tsg::set_error("tu_load: out of memory or internal error");
    return TSG_EINVAL;
  }
}

static int tu_load_impl(const char* prefix, int mode, int64_t max_nodes, tsg_tu_handle** out) {
  TSG_REQUIRE(prefix && out, "tu_load: null pointer");
  TSG_REQUIRE(mode == TSG_TU_NETWORKX || mode == TSG_TU_PYG, "tu_load: unknown mode %d", mode);
  const std::string pre(prefix);
  std::string txt;
  std::vector<int64_t> indic, glab, nlab, apairs;
  if (!read_file(pre + "_graph_indicator.txt", &txt)) { tsg::set_error("tu_load: cannot read %s_graph_indicator.txt", prefix); return TSG_EINVAL; }
  parse_ints(txt, &indic);
  if (!read_file(pre + "_graph_labels.txt", &txt)) { tsg::set_error("tu_load: cannot read %s_graph_labels.txt", prefix); return TSG_EINVAL; }
  parse_ints(txt, &glab);
  if (!read_file(pre + "_A.txt", &txt)) { tsg::set_error("tu_load: cannot read %s_A.txt", prefix); return TSG_EINVAL; }
  parse_ints(txt, &apairs);
  TSG_REQUIRE(apairs.size() % 2 == 0, "tu_load: odd number of integers in %s_A.txt", prefix);
  const bool has_nlab = read_file(pre + "_node_labels.txt", &txt);
  if (has_nlab) parse_ints(txt, &nlab);
  std::vector<float> nattr;
  int64_t attr_dim = 0;
  if (read_file(pre + "_node_attributes.txt", &txt)) attr_dim = parse_float_rows(txt, &nattr);
  const int64_t NN = (int64_t)indic.size(), G = (int64_t)glab.size(), M = (int64_t)apairs.size() / 2;
  TSG_REQUIRE(!has_nlab || (int64_t)nlab.size() == NN, "tu_load: %lld node labels for %lld nodes", (long long)nlab.size(), (long long)NN);
  TSG_REQUIRE(attr_dim == 0 || (int64_t)nattr.size() == NN * attr_dim, "tu_load: ragged node attribute file");
  for (int64_t i = 0; i < NN; ++i) TSG_REQUIRE(indic[i] >= 1 && indic[i] <= G, "tu_load: graph indicator %lld out of range", (long long)indic[i]);
  for (int64_t e = 0; e < 2 * M; ++e) TSG_REQUIRE(apairs[e] >= 1 && apairs[e] <= NN, "tu_load: node id %lld out of range", (long long)apairs[e]);

  tsg_tu_handle* h = new tsg_tu_handle();
  TuCorpus& c = h->c;
  c.attr_dim = attr_dim;
  // ---- labels
  std::vector<int64_t> ymap(G);
  if (mode == TSG_TU_NETWORKX) {
    std::unordered_map<int64_t, int64_t> first;
    for (int64_t g = 0; g < G; ++g) {
      auto it = first.find(glab[g]);
      if (it == first.end()) it = first.emplace(glab[g], (int64_t)first.size()).first;
      ymap[g] = it->second;
    }
    c.num_classes = (int64_t)first.size();
    if (has_nlab) {
      int64_t mx = -(1LL << 60);
      for (int64_t v : nlab) mx = std::max(mx, v - 1);
      c.num_node_labels = mx + 1;
      for (auto& v : nlab) { v -= 1; if (v < 0) v += c.num_node_labels; }       // python negative index
    }
  } else {
    std::vector<int64_t> uniq(glab);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    for (int64_t g = 0; g < G; ++g) ymap[g] = std::lower_bound(uniq.begin(), uniq.end(), glab[g]) - uniq.begin();
    c.num_classes = (int64_t)uniq.size();
    if (has_nlab && NN > 0) {
      int64_t mn = nlab[0], mx = nlab[0];
      for (int64_t v : nlab) { mn = std::min(mn, v); mx = std::max(mx, v); }
      for (auto& v : nlab) v -= mn;
      c.num_node_labels = mx - mn + 1;
    }
  }
  // ---- bucket the edges by graph (file order kept inside a graph)
  std::vector<int64_t> ecount(G + 1, 0);
  for (int64_t e = 0; e < M; ++e) ++ecount[indic[apairs[2 * e] - 1]];
  std::vector<int64_t> estart(G + 2, 0);
  for (int64_t g = 1; g <= G; ++g) estart[g + 1] = estart[g] + ecount[g];
  std::vector<int64_t> eorder(M), fill(estart.begin(), estart.end());
  for (int64_t e = 0; e < M; ++e) eorder[fill[indic[apairs[2 * e] - 1]]++] = e;
  // ---- PYG: node ranges per graph (file order; indicator must be grouped)
  std::vector<int64_t> nfirst(G + 2, -1), ncount(G + 2, 0);
  for (int64_t i = 0; i < NN; ++i) { const int64_t g = indic[i]; if (nfirst[g] < 0) nfirst[g] = i; ++ncount[g]; }
  if (mode == TSG_TU_PYG)
    for (int64_t i = 1; i < NN; ++i)
      if (indic[i] < indic[i - 1]) { delete h; tsg::set_error("tu_load: graph indicator is not sorted (needed for the PyG layout)"); return TSG_EINVAL; }

  c.node_ptr.push_back(0); c.edge_ptr.push_back(0);
  std::unordered_map<int64_t, int32_t> local;
  std::vector<int64_t> order;                       // global node id (0-based) of every local node
  std::vector<std::pair<int32_t, int32_t>> ed;
  for (int64_t g = 1; g <= G; ++g) {
    local.clear(); order.clear(); ed.clear();
    if (mode == TSG_TU_PYG) {
      for (int64_t k = 0; k < ncount[g]; ++k) order.push_back(nfirst[g] + k);
      for (int64_t k = estart[g]; k < estart[g + 1]; ++k) {
        const int64_t u = apairs[2 * eorder[k]] - 1, v = apairs[2 * eorder[k] + 1] - 1;
        if (u == v) continue;                                                    // remove_self_loops
        if (indic[v] != g) { delete h; tsg::set_error("tu_load: edge (%lld,%lld) crosses graphs", (long long)u + 1, (long long)v + 1); return TSG_EINVAL; }
        ed.emplace_back((int32_t)(u - nfirst[g]), (int32_t)(v - nfirst[g]));
      }
    } else {
      auto id_of = [&](int64_t u) {
        auto it = local.find(u);
        if (it == local.end()) { it = local.emplace(u, (int32_t)order.size()).first; order.push_back(u); }
        return it->second;
      };
      for (int64_t k = estart[g]; k < estart[g + 1]; ++k) {
        const int64_t u = apairs[2 * eorder[k]] - 1, v = apairs[2 * eorder[k] + 1] - 1;
        const int32_t a = id_of(u), b = id_of(v);
        ed.emplace_back(a, b);
        if (a != b) ed.emplace_back(b, a);                                       // undirected
      }
      if (max_nodes > 0 && (int64_t)order.size() > max_nodes) continue;          // load_data.py:90-91
    }
    std::sort(ed.begin(), ed.end());
    ed.erase(std::unique(ed.begin(), ed.end()), ed.end());
    for (auto& pr : ed) { c.row.push_back(pr.first); c.col.push_back(pr.second); }
    for (int64_t u : order) {
      c.label.push_back(has_nlab ? (int32_t)nlab[u] : 0);
      for (int64_t d = 0; d < attr_dim; ++d) c.attr.push_back(nattr[u * attr_dim + d]);
    }
    c.node_ptr.push_back(c.node_ptr.back() + (int64_t)order.size());
    c.edge_ptr.push_back(c.edge_ptr.back() + (int64_t)ed.size());
    c.y.push_back(ymap[g - 1]);
  }
  *out = h;
  return TSG_OK;
}

extern "C" int tsg_tu_sizes(const tsg_tu_handle* h, int64_t* sizes) {
  TSG_REQUIRE(h && sizes, "tu_sizes: null pointer");
  const TuCorpus& c = h->c;
  sizes[0] = (int64_t)c.y.size(); sizes[1] = c.node_ptr.back(); sizes[2] = c.edge_ptr.back();
  sizes[3] = c.num_node_labels; sizes[4] = c.attr_dim; sizes[5] = c.num_classes;
  return TSG_OK;
}

extern "C" int tsg_tu_fill(const tsg_tu_handle* h, int64_t* node_ptr, int64_t* edge_ptr, int64_t* row, int64_t* col,
                           int32_t* node_label, int64_t* y, float* attr) {
  TSG_REQUIRE(h && node_ptr && edge_ptr && y, "tu_fill: null pointer");
  const TuCorpus& c = h->c;
  memcpy(node_ptr, c.node_ptr.data(), c.node_ptr.size() * 8);
  memcpy(edge_ptr, c.edge_ptr.data(), c.edge_ptr.size() * 8);
  if (!c.row.empty()) { TSG_REQUIRE(row && col, "tu_fill: null edge buffers"); memcpy(row, c.row.data(), c.row.size() * 8); memcpy(col, c.col.data(), c.col.size() * 8); }
  if (!c.label.empty() && node_label) memcpy(node_label, c.label.data(), c.label.size() * 4);
  memcpy(y, c.y.data(), c.y.size() * 8);
  if (attr && !c.attr.empty()) memcpy(attr, c.attr.data(), c.attr.size() * 4);
  return TSG_OK;
}

extern "C" void tsg_tu_free(tsg_tu_handle* h) { delete h; }
