// Host-side model builder.  Everything the reference's Inference constructor derives from its inputs
// (reference source/functions.cpp:76-156) is rebuilt here on flat arrays, in the same floating-point order
// where integers depend on it (heights -> order -> nSub).  No recursion: a 10,000-leaf caterpillar tree is
// 10,000 levels deep.
#include "host_model.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <unordered_map>

namespace ppm {

static void finish_tree(Tree& t);

// RecTree(string) (objects.cpp:51-78) accepts: binary nodes only, ':length' after every node, an optional
// label after ')' which is skipped (firstPoint, objects.cpp:38-50), and ignores whatever follows the last ')'
// (the root's own ':0;').  Lengths go through stold then narrow to double (objects.cpp:71-72).
Tree parse_newick(const std::string& line) {
  Tree t;
  const char* s = line.c_str();
  size_t n = line.size(), pos = 0;
  std::vector<int> open;  // internal nodes whose children are still being read
  auto attach = [&](int v) {
    if (open.empty()) {
      if (v != 0) throw std::runtime_error("newick: more than one top-level node");
      return;
    }
    int p = open.back();
    t.parent[v] = p;
    if (t.left[p] < 0) t.left[p] = v;
    else if (t.right[p] < 0) t.right[p] = v;
    else throw std::runtime_error("newick: node with more than two children (the model needs a binary tree)");
  };
  auto add = [&]() {
    t.left.push_back(-1); t.right.push_back(-1); t.parent.push_back(-1);
    t.bl.push_back(0.); t.label.emplace_back();
    return (int)t.left.size() - 1;
  };
  auto read_len = [&](int v) {
    if (pos < n && s[pos] == ':') {
      char* end = nullptr;
      long double x = strtold(s + pos + 1, &end);
      if (end == s + pos + 1) throw std::runtime_error("newick: missing branch length");
      t.bl[v] = (double)x;
      pos = end - s;
    } else if (!open.empty()) {
      throw std::runtime_error("newick: every node needs ':length'");
    }
  };
  bool done = false;
  while (pos < n && !done) {
    char c = s[pos];
    if (c == '(') {
      int v = add(); attach(v); open.push_back(v); pos++;
    } else if (c == ',') {
      pos++;
    } else if (c == ')') {
      if (open.empty()) throw std::runtime_error("newick: unbalanced ')'");
      int v = open.back(); open.pop_back(); pos++;
      if (t.right[v] < 0) throw std::runtime_error("newick: node with fewer than two children");
      while (pos < n && s[pos] != ':' && s[pos] != ',' && s[pos] != ')' && s[pos] != ';') pos++;  // label
      read_len(v);
      if (open.empty()) done = true;
    } else if (c == ';' || c == ' ' || c == '\r' || c == '\t') {
      pos++;
    } else {
      size_t j = pos;
      while (j < n && s[j] != ':' && s[j] != ',' && s[j] != ')' && s[j] != '(' && s[j] != ';') j++;
      int v = add(); attach(v);
      t.label[v] = line.substr(pos, j - pos);
      pos = j;
      read_len(v);
      if (open.empty()) done = true;
    }
  }
  if (!open.empty()) throw std::runtime_error("newick: unbalanced '('");
  if (t.left.empty()) throw std::runtime_error("newick: empty tree");
  finish_tree(t);
  return t;
}

Tree tree_from_arrays(int n_nodes, const int32_t* left, const int32_t* right, const double* bl) {
  if (n_nodes < 1 || n_nodes % 2 == 0) throw std::runtime_error("tree: n_nodes must be odd and positive");
  Tree t;
  t.left.assign(left, left + n_nodes);
  t.right.assign(right, right + n_nodes);
  t.bl.assign(bl, bl + n_nodes);
  t.parent.assign(n_nodes, -1);
  t.label.assign(n_nodes, "");
  for (int v = 0; v < n_nodes; v++) {
    int a = t.left[v], b = t.right[v];
    if ((a < 0) != (b < 0)) throw std::runtime_error("tree: a node needs zero or two children");
    if (a >= 0) {
      if (a <= v || b <= a || b >= n_nodes) throw std::runtime_error("tree: ids are not a preorder numbering");
      t.parent[a] = v; t.parent[b] = v;
    }
  }
  for (int v = 1; v < n_nodes; v++)
    if (t.parent[v] < 0) throw std::runtime_error("tree: node without parent");
  finish_tree(t);
  // preorder check: left child is v+1, right child is last(left)+1
  for (int v = 0; v < n_nodes; v++)
    if (t.left[v] >= 0 && (t.left[v] != v + 1 || t.right[v] != t.last[t.left[v]] + 1))
      throw std::runtime_error("tree: ids are not a preorder numbering");
  return t;
}

Tree read_tree_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open tree file " + path);
  std::string line;
  std::getline(f, line);  // only the first line is read (objects.cpp:123-131)
  if (line.empty()) throw std::runtime_error("tree file is empty: " + path);
  return parse_newick(line);
}

double expected_gains(const Tree& t, double g2) {
  // sumGains(st) (functions.cpp:582-592) = sum over the internal nodes v of the subtree of exp(g2 * height[v]) minus its leaves;
  // children have larger preorder ids than their parent, so one reverse pass fills it
  std::vector<double> G(t.n_nodes, 0.);
  for (int v = t.n_nodes - 1; v >= 0; v--)
    G[v] = t.left[v] < 0 ? -1. : std::exp(g2 * t.height[v]) + G[t.left[v]] + G[t.right[v]];
  // a lineage st is in subtrees[rank] for every interval between its own height and its parent's (functions.cpp:130-141);
  // sum over those intervals of  w + G (e^{-g2 h_lo} - e^{-g2 h_hi}) / g2  telescopes to the two ends of its branch
  double sum = 0.;
  // (the reference starts at interval rank nLeaves-1, functions.cpp:598: anything below the last leaf is skipped)
  for (int v = 1; v < t.n_nodes; v++) {
    const int r0 = std::max(t.rank[v], t.n_leaves - 1), r1 = t.rank[t.parent[v]];
    if (r1 <= r0) continue;
    const double lo = t.height[t.order[r0]], hi = t.height[t.parent[v]];
    if (hi > lo) sum += (hi - lo) + G[v] * (std::exp(-g2 * lo) - std::exp(-g2 * hi)) / g2;
  }
  return sum / t.H;
}

static void finish_tree(Tree& t) {
  const int nn = (int)t.left.size();
  t.n_nodes = nn;
  t.n_leaves = (nn + 1) / 2;
  t.bl[0] = 0.;
  int n_leaf_found = 0;
  t.leaf_pos.assign(nn, -1);
  t.leaf_node.clear();
  for (int v = 0; v < nn; v++)
    if (t.left[v] < 0) { t.leaf_pos[v] = n_leaf_found++; t.leaf_node.push_back(v); }
  if (n_leaf_found != t.n_leaves) throw std::runtime_error("tree: not a full binary tree");

  // depths root -> leaf (functions.cpp:66-75), H = max depth, height = H - depth (:108-109)
  std::vector<double> depth(nn, 0.);
  for (int v = 0; v < nn; v++)
    if (t.left[v] >= 0) {
      depth[t.left[v]] = depth[v] + t.bl[t.left[v]];
      depth[t.right[v]] = depth[v] + t.bl[t.right[v]];
    }
  t.H = *std::max_element(depth.begin(), depth.end());
  t.height.resize(nn);
  for (int v = 0; v < nn; v++) t.height[v] = t.H - depth[v];

  // subtree extents and treeLength in the reference's summation order (objects.cpp:115-122)
  t.last.assign(nn, 0);
  std::vector<double> tl(nn, 0.);
  for (int v = nn - 1; v >= 0; v--) {
    if (t.left[v] < 0) { t.last[v] = v; continue; }
    t.last[v] = t.last[t.right[v]];
    tl[v] = ((t.bl[t.left[v]] + tl[t.left[v]]) + t.bl[t.right[v]]) + tl[t.right[v]];
  }
  t.L = tl[0];

  // order by increasing height (functions.cpp:112-114).  std::sort leaves ties unspecified; we break them
  // children-first (larger preorder id first), which every consistent sweep needs ("ties").
  t.order.resize(nn);
  for (int i = 0; i < nn; i++) t.order[i] = i;
  std::sort(t.order.begin(), t.order.end(), [&](int a, int b) {
    if (t.height[a] != t.height[b]) return t.height[a] < t.height[b];
    return a > b;
  });
  t.rank.resize(nn);
  for (int i = 0; i < nn; i++) t.rank[t.order[i]] = i;

  // level lists (counting sort by level)
  {
    std::vector<int> up(nn, 0), dn(nn, 0);
    for (int v = nn - 1; v >= 0; v--)
      if (t.left[v] >= 0) up[v] = 1 + std::max(up[t.left[v]], up[t.right[v]]);
    for (int v = 1; v < nn; v++) dn[v] = dn[t.parent[v]] + 1;
    auto bucket = [&](const std::vector<int>& lev, std::vector<int>& off, std::vector<int>& nodes) {
      int nl = *std::max_element(lev.begin(), lev.end()) + 1;
      off.assign(nl + 1, 0);
      for (int v = 0; v < nn; v++) off[lev[v] + 1]++;
      for (int l = 0; l < nl; l++) off[l + 1] += off[l];
      nodes.resize(nn);
      std::vector<int> pos(off.begin(), off.end() - 1);
      for (int v = 0; v < nn; v++) nodes[pos[lev[v]]++] = v;
    };
    bucket(up, t.lev_up_off, t.lev_up_nodes);
    bucket(dn, t.lev_dn_off, t.lev_dn_nodes);
    const int nld = (int)t.lev_dn_off.size() - 1;
    std::vector<int> n_internal(nld + 1, 0);
    for (int l = 0; l < nld; l++) {
      auto first = t.lev_dn_nodes.begin() + t.lev_dn_off[l], last = t.lev_dn_nodes.begin() + t.lev_dn_off[l + 1];
      n_internal[l] = (int)(std::stable_partition(first, last, [&](int v) { return t.left[v] >= 0; }) - first);
    }
    t.lev_dn_chain.assign(nld + 1, 0);
    for (int l = 0; l + 1 < nld; l++) t.lev_dn_chain[l] = (n_internal[l] == 1 && n_internal[l + 1] == 1) ? 1 : 0;
  }

  // breadth-first last node (what computeMRCA returns for an all-zero pattern, functions.cpp:52-64)
  {
    std::vector<int> q{0};
    for (size_t k = 0; k < q.size(); k++)
      if (t.left[q[k]] >= 0) { q.push_back(t.left[q[k]]); q.push_back(t.right[q[k]]); }
    t.bfs_last = q.back();
  }

  // Mobile integral grid (functions.cpp:250-260) and alive-lineage counts (functions.cpp:130-141)
  t.nsub.assign(std::max(nn - 1, 0), 0);
  t.alive.assign(std::max(nn - 1, 0), 0);
  t.slice_start.assign(nn, 0);
  t.slice_t.clear(); t.slice_wt.clear();
  {
    // alive count of interval r = lineages v (non-root) with rank[v] <= r < rank[parent[v]]
    std::vector<int> delta(nn + 1, 0);
    for (int v = 1; v < nn; v++) { delta[t.rank[v]]++; delta[t.rank[t.parent[v]]]--; }
    int run = 0;
    for (int r = 0; r < nn - 1; r++) { run += delta[r]; t.alive[r] = run; }
  }
  t.n_terms = 0;
  for (int r = 0; r < nn - 1; r++) {
    t.slice_start[r] = (int)t.slice_t.size();
    if (r < t.n_leaves - 1) continue;  // the integral starts at the last leaf (functions.cpp:250)
    double h0 = t.height[t.order[r]];
    double w = t.height[t.order[r + 1]] - h0;
    if (!(w > 0.)) continue;
    int ns = std::max(1, (int)std::ceil(w * 100 / t.H));
    t.nsub[r] = ns;
    for (int i = 0; i < ns; i++) {
      t.slice_t.push_back(h0 + (2 * i + 1) * w / (2 * ns));
      t.slice_wt.push_back(w / ns);
    }
    t.n_terms += (int64_t)ns * t.alive[r];
  }
  if (nn >= 1) t.slice_start[nn - 1] = (int)t.slice_t.size();
  t.n_slices = (int)t.slice_t.size();
}

// ---------------------------------------------------------------------------------------------------------
namespace {
struct RowHash {
  const uint32_t* base; int nw;
  size_t operator()(int64_t i) const {
    uint64_t h = 1469598103934665603ull;
    const uint32_t* p = base + (size_t)i * nw;
    for (int k = 0; k < nw; k++) { h ^= p[k]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
struct RowEq {
  const uint32_t* base; int nw;
  bool operator()(int64_t a, int64_t b) const {
    return std::memcmp(base + (size_t)a * nw, base + (size_t)b * nw, sizeof(uint32_t) * nw) == 0;
  }
};
}  // namespace

Patterns build_patterns(const Tree& t, const uint32_t* rows, const int32_t* counts, int64_t n_rows, bool dedupe) {
  Patterns p;
  const int nw = (t.n_leaves + 31) / 32;
  p.n_words = nw;
  p.n_genes = n_rows;
  if (n_rows <= 0) throw std::runtime_error("matrix: no gene columns");
  // bits beyond n_leaves must be clear
  if (t.n_leaves % 32) {
    uint32_t mask = ~0u << (t.n_leaves % 32);
    for (int64_t i = 0; i < n_rows; i++)
      if (rows[(size_t)i * nw + nw - 1] & mask) throw std::runtime_error("patterns: bits set beyond n_leaves");
  }
  p.gene_to_pattern.resize(n_rows);
  std::vector<int64_t> keep;
  if (dedupe) {  // PAmatrix::compress keeps first-occurrence order (objects.cpp:254-273)
    std::unordered_map<int64_t, int32_t, RowHash, RowEq> seen((size_t)n_rows * 2, RowHash{rows, nw}, RowEq{rows, nw});
    for (int64_t i = 0; i < n_rows; i++) {
      auto it = seen.find(i);
      if (it == seen.end()) {
        seen.emplace(i, (int32_t)keep.size());
        p.gene_to_pattern[i] = (int32_t)keep.size();
        keep.push_back(i);
        p.counts.push_back(counts ? counts[i] : 1);
      } else {
        p.gene_to_pattern[i] = it->second;
        p.counts[it->second] += counts ? counts[i] : 1;
      }
    }
  } else {
    keep.resize(n_rows);
    p.counts.resize(n_rows);
    for (int64_t i = 0; i < n_rows; i++) { keep[i] = i; p.gene_to_pattern[i] = (int32_t)i; p.counts[i] = counts ? counts[i] : 1; }
  }
  p.n_patterns = (int64_t)keep.size();
  p.words.resize((size_t)p.n_patterns * nw);
  for (int64_t j = 0; j < p.n_patterns; j++)
    std::memcpy(&p.words[(size_t)j * nw], rows + (size_t)keep[j] * nw, sizeof(uint32_t) * nw);

  // binary lifting over parents for the MRCA
  int LOG = 1;
  while ((1 << LOG) < t.n_nodes) LOG++;
  std::vector<std::vector<int>> up(LOG, std::vector<int>(t.n_nodes, 0));
  for (int v = 0; v < t.n_nodes; v++) up[0][v] = v ? t.parent[v] : 0;
  for (int k = 1; k < LOG; k++)
    for (int v = 0; v < t.n_nodes; v++) up[k][v] = up[k - 1][up[k - 1][v]];

  p.mrca.resize(p.n_patterns);
  p.freq.resize(p.n_patterns);
  p.n_obs = 0;
  long double gene_sum = 0;
  for (int64_t j = 0; j < p.n_patterns; j++) {
    const uint32_t* w = &p.words[(size_t)j * nw];
    int f = 0, lo = -1, hi = -1;
    for (int k = 0; k < nw; k++) {
      if (!w[k]) continue;
      f += __builtin_popcount(w[k]);
      if (lo < 0) lo = 32 * k + __builtin_ctz(w[k]);
      hi = 32 * k + 31 - __builtin_clz(w[k]);
    }
    p.freq[j] = f;  // functions.cpp:117-122
    // computeMRCA (functions.cpp:46-65): the last node in breadth-first order containing every present leaf,
    // i.e. their deepest common ancestor; leaves in preorder => it is the LCA of the first and last present leaf.
    int m;
    if (f == 0) m = t.bfs_last;
    else {
      int a = t.leaf_node[lo], b = t.leaf_node[hi];
      for (int k = LOG - 1; k >= 0; k--)
        if (t.last[up[k][a]] < b) a = up[k][a];
      m = (t.last[a] >= b) ? a : t.parent[a];
    }
    p.mrca[j] = m;
    p.n_obs += p.counts[j];                                  // functions.cpp:83
    gene_sum += (long double)f * p.counts[j];                // functions.cpp:123-126
  }
  p.mean_genes = (double)gene_sum / t.n_leaves;              // functions.cpp:127 (integers: exact in double)
  return p;
}

// PAmatrix(path) (objects.cpp:164-214): first line = header (ignored) or "ccc,count,count,..."; then one
// line per genome: name,v,v,...  Genome rows are matched to tree leaves by name (objects.cpp:240-249).
// 32 x 32 bit transpose: out[j] bit i = in[i] bit j
static void transpose32(const uint32_t* in, uint32_t* out) {
  uint32_t a[32];
  std::memcpy(a, in, sizeof(a));
  for (int j = 16, m = 0x0000ffff; j; j >>= 1, m ^= m << j)
    for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
      const uint32_t tt = ((a[k] >> j) ^ a[k + j]) & (uint32_t)m;
      a[k] ^= tt << j;
      a[k + j] ^= tt;
    }
  std::memcpy(out, a, sizeof(a));
}

// The matrix file is read one genome line at a time (2 MB per line at 1 M genes; the file is never held whole).  A genome's
// values are packed ALONG THE GENES (sequential writes: 8 text bytes "v,v,v,v," become 4 bits with one multiply); once every
// line is in, 32 x 32 bit tiles are transposed into the gene-major rows the patterns are built from.  The previous reader set one
// bit per present value at a 628-byte stride (1.3e9 read-modify-writes for 5,000 genomes x 1 M genes).
int64_t read_matrix_rows(const Tree& t, const std::string& path, std::vector<uint32_t>& rows, std::vector<int32_t>& counts) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open matrix file " + path);
  std::string line;
  if (!std::getline(f, line)) throw std::runtime_error("matrix file is empty: " + path);
  counts.clear();
  auto strip = [](std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); };
  strip(line);
  if (line.size() >= 3 && line[0] == 'c' && line[1] == 'c' && line[2] == 'c') {
    size_t pos = line.find(',');
    while (pos != std::string::npos) {
      counts.push_back((int32_t)strtol(line.c_str() + pos + 1, nullptr, 10));
      pos = line.find(',', pos + 1);
    }
  }
  std::unordered_map<std::string, int> leaf_of;
  for (int k = 0; k < t.n_leaves; k++) leaf_of[t.label[t.leaf_node[k]]] = k;
  const int nw = (t.n_leaves + 31) / 32;
  rows.clear();
  int64_t n_genes = -1, gw = 0;
  std::vector<uint32_t> gm;  // genome-major bits: [nw * 32][gw], bit g & 31 of word g >> 5 = the genome's value for gene g
  std::vector<char> seen(t.n_leaves, 0);
  while (std::getline(f, line)) {
    strip(line);
    if (line.empty()) continue;
    size_t c = line.find(',');
    if (c == std::string::npos) throw std::runtime_error("matrix: line without values");
    std::string name = line.substr(0, c);
    auto it = leaf_of.find(name);
    const char* p = line.c_str() + c + 1;
    const char* const end = line.c_str() + line.size();
    // count values on the first data line
    if (n_genes < 0) {
      n_genes = 1;
      for (const char* q = p; *q; q++) n_genes += (*q == ',');
      gw = (n_genes + 31) / 32;
      gm.assign((size_t)nw * 32 * gw, 0u);
    }
    // A genome that is no leaf of the tree: the reference keeps such a row in `matrix`, where it changes freq[], meanGenes (hence
    // the optimiser's bounds and start) and pattern identity, and leaves mrca uninitialised (functions.cpp:51,64,117-127).  Nothing
    // sensible can be reproduced from that, so the input is rejected.
    if (it == leaf_of.end()) throw std::runtime_error("matrix: genome row '" + name + "' is not a leaf of the tree");
    const int k = it->second;
    uint32_t* dst = gm.data() + (size_t)k * gw;
    if (seen[k]) std::fill(dst, dst + gw, 0u);  // a later row with the same name wins (std::map assignment, objects.cpp:195)
    seen[k] = 1;
    int64_t g = 0;
    // fast path: single-digit values separated by commas (what the reference's R scripts write), four values per 8 bytes;
    // anything else falls through to the general strtol loop below at the position reached
    while (g + 4 <= n_genes && end - p >= 8) {
      uint64_t x;
      std::memcpy(&x, p, 8);
      if ((x & 0xfffefffefffefffeull) != 0x2c302c302c302c30ull) break;  // "[01],[01],[01],[01],"
      const uint64_t nib = (((x & 0x0001000100010001ull) * 0x0001000200040008ull) >> 48) & 0xfull;
      dst[g >> 5] |= (uint32_t)nib << (g & 31);  // (g is a multiple of 4 here: the nibble never straddles a word)
      g += 4;
      p += 8;
    }
    while (g < n_genes && (p[0] == '0' || p[0] == '1') && (p[1] == ',' || p[1] == '\0')) {
      if (p[0] == '1') dst[g >> 5] |= 1u << (g & 31);
      g++;
      if (p[1] == '\0') { p += 1; break; }
      p += 2;
    }
    if (*p == '\0' && g > 0) {
      if (g != n_genes) throw std::runtime_error("matrix: row " + name + " is shorter than the first row");
      continue;
    }
    while (true) {
      char* e2;
      long v = strtol(p, &e2, 10);
      if (e2 == p) throw std::runtime_error("matrix: empty value in row " + name);
      if (v != 0 && v != 1) throw std::runtime_error("matrix: values must be 0 or 1 (row " + name + ")");
      if (g >= n_genes) throw std::runtime_error("matrix: row " + name + " is longer than the first row");
      if (v) dst[g >> 5] |= 1u << (g & 31);
      g++;
      if (*e2 == ',') p = e2 + 1; else break;
    }
    if (g != n_genes) throw std::runtime_error("matrix: row " + name + " is shorter than the first row");
  }
  if (n_genes <= 0) throw std::runtime_error("matrix: no genome rows");
  for (int k = 0; k < t.n_leaves; k++)
    if (!seen[k]) throw std::runtime_error("matrix: no row for tree leaf '" + t.label[t.leaf_node[k]] + "'");
  if (!counts.empty() && (int64_t)counts.size() != n_genes) throw std::runtime_error("matrix: ccc header length differs from the rows");
  // gene-major rows: rows[g][w] bit i = genome 32 w + i
  rows.assign((size_t)n_genes * nw, 0u);
  uint32_t tin[32], tout[32];
  for (int w = 0; w < nw; w++)
    for (int64_t gb = 0; gb < gw; gb++) {
      uint32_t any = 0;
      for (int i = 0; i < 32; i++) { tin[i] = gm[((size_t)w * 32 + i) * gw + gb]; any |= tin[i]; }
      if (!any) continue;
      transpose32(tin, tout);
      const int64_t g1 = std::min<int64_t>(32, n_genes - 32 * gb);
      for (int64_t j = 0; j < g1; j++) rows[(size_t)(32 * gb + j) * nw + w] = tout[j];
    }
  return n_genes;
}

Patterns read_matrix_file(const Tree& t, const std::string& path, bool dedupe) {
  std::vector<uint32_t> rows;
  std::vector<int32_t> counts;
  const int64_t n_genes = read_matrix_rows(t, path, rows, counts);
  return build_patterns(t, rows.data(), counts.empty() ? nullptr : counts.data(), n_genes, dedupe);
}

// The compressed matrix format of the reference (objects.cpp:175-187 reads it, simulation_aux.R:247-252 compress.rep makes
// it): first line "ccc,count,count,...", then one line per genome "name,v,v,..." with one column per unique pattern.
void write_ccc_file(const Tree& t, const Patterns& p, const std::string& path) {
  std::ofstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path + " for writing");
  std::string line = "ccc";
  for (int64_t j = 0; j < p.n_patterns; j++) { line += ','; line += std::to_string(p.counts[j]); }
  line += '\n';
  f << line;
  for (int k = 0; k < t.n_leaves; k++) {
    line = t.label[t.leaf_node[k]];
    line.reserve(line.size() + 2 * (size_t)p.n_patterns + 2);
    for (int64_t j = 0; j < p.n_patterns; j++) {
      line += ',';
      line += ((p.words[(size_t)j * p.n_words + (k >> 5)] >> (k & 31)) & 1u) ? '1' : '0';
    }
    line += '\n';
    f << line;
  }
  if (!f) throw std::runtime_error("write error on " + path);
}

// ---------------------------------------------------------------------------------------------------------
Program build_program(const Tree& t, int R, bool deep, bool leaf_pow) {
  Program pg;
  pg.R = R;
  pg.deep = deep;
  const int nn = t.n_nodes, S = t.n_slices, nl = t.n_leaves;
  const int nb = (S + R - 1) / R;
  pg.n_blocks = nb;
  pg.blk.assign(nb + 1, BlockDescHost{});
  for (int b = 0; b < nb; b++) { pg.blk[b].slice0 = b * R; pg.blk[b].nslices = std::min(R, S - b * R); }
  pg.blk[nb].slice0 = S; pg.blk[nb].nslices = 0;

  // alive slice range of every non-root lineage: intervals [max(rank, nl-1), rank(parent)-1]
  std::vector<int> js0(nn, S), js1(nn, S);
  for (int v = 1; v < nn; v++) {
    js0[v] = t.slice_start[std::max(t.rank[v], nl - 1)];
    js1[v] = t.slice_start[t.rank[t.parent[v]]];
  }
  // static leaf pairs for the u-form: cherries first (they merge together), then the remaining leaves by merge time.
  // The device patterns are stored in PROGRAM leaf order: pair p owns bits 2p and 2p+1 (so the two observations of a
  // pair are one 2-bit field), unpaired leaves follow.  leaf_perm: preorder leaf index -> device bit position.
  double max_leaf_h = 0;
  for (int k = 0; k < nl; k++) max_leaf_h = std::max(max_leaf_h, std::fabs(t.height[t.leaf_node[k]]));
  pg.leaf_u = max_leaf_h <= 1e-6 * t.H;
  pg.max_leaf_h = max_leaf_h;
  // leaf powers: leaf heights that are rounding noise of the tree file only (the height correction is first order), folded events
  pg.lp = leaf_pow && pg.leaf_u && max_leaf_h <= 1e-8 * t.H && std::getenv("PPM_NO_FOLD") == nullptr && nn > 1;
  for (int v = 1; v < nn && pg.lp; v++)
    if (t.left[v] < 0 && js0[v] != 0) pg.lp = false;  // (an internal node below a leaf's height: that leaf starts above the first slice)
  pg.has_s1 = pg.lp && max_leaf_h > 0.;
  pg.leaf_perm.assign(nl, -1);
  pg.n_pairs = 0;
  if (pg.leaf_u) {
    auto add_pair = [&](int a, int bq) { pg.leaf_perm[a] = 2 * pg.n_pairs; pg.leaf_perm[bq] = 2 * pg.n_pairs + 1; pg.n_pairs++; };
    for (int v = 0; v < nn; v++)
      if (t.left[v] >= 0 && t.left[t.left[v]] < 0 && t.left[t.right[v]] < 0) add_pair(t.leaf_pos[t.left[v]], t.leaf_pos[t.right[v]]);
    pg.n_cherries = pg.n_pairs;
    std::vector<int> rest;
    for (int k = 0; k < nl; k++) if (pg.leaf_perm[k] < 0 && nn > 1) rest.push_back(k);
    std::sort(rest.begin(), rest.end(), [&](int x, int y) {
      int dx = js1[t.leaf_node[x]], dy = js1[t.leaf_node[y]];
      return dx != dy ? dx < dy : x < y;
    });
    for (size_t i = 0; i + 1 < rest.size(); i += 2) add_pair(rest[i], rest[i + 1]);
    if (pg.n_pairs > 65535) throw std::runtime_error("program: too many leaf pairs for 16-bit references");
  }
  {
    int next = 2 * pg.n_pairs;
    for (int k = 0; k < nl; k++) if (pg.leaf_perm[k] < 0) pg.leaf_perm[k] = next++;
    pg.leaf_node_dev.assign(nl, 0);
    for (int k = 0; k < nl; k++) pg.leaf_node_dev[pg.leaf_perm[k]] = t.leaf_node[k];
  }
  auto dpos = [&](int v) { return pg.leaf_perm[t.leaf_pos[v]]; };  // device bit position of leaf node v
  auto pair_of = [&](int pos) { return pos < 2 * pg.n_pairs ? pos / 2 : -1; };
  // node ops: internal nodes in rank order; op block = block of the node's first alive slice
  std::vector<int> op_block(nn, nb);
  for (int v = 1; v < nn; v++)
    if (t.left[v] >= 0 && js0[v] < S) op_block[v] = js0[v] / R;  // (a node with an empty alive range still merges here)
  // Ops per block: internal nodes in rank order; cherries (both children leaves: they depend on nothing and the sweep serves
  // them from a per-parameter table indexed by the 2-bit observation pair) run first.
  // Slots: a node takes one at its op.  With folded events a child is last read by its parent's op (merge + its slices of
  // the block), so its slot is free for that very op's result and for the ops after it; without folding its partial entry
  // still needs it until the block ends.
  const bool fold = std::getenv("PPM_NO_FOLD") == nullptr;
  std::vector<std::vector<int>> block_nodes(nb + 1);
  {
    int cur = 0;
    for (int r = 0; r < nn; r++) {
      int v = t.order[r];
      if (t.left[v] < 0) continue;
      cur = std::max(cur, op_block[v]);  // children are formed in an earlier-or-equal block; keep the sequence monotone
      op_block[v] = cur;
      block_nodes[cur].push_back(v);
    }
  }
  auto is_cherry = [&](int v) { return pg.leaf_u && v != 0 && t.left[t.left[v]] < 0 && t.left[t.right[v]] < 0; };
  std::vector<int> slot(nn, -1);
  std::vector<int> free_slots;
  int n_slots = 0;
  for (int b = 0; b <= nb; b++) {
    std::stable_partition(block_nodes[b].begin(), block_nodes[b].end(), is_cherry);
    pg.blk[b].op0 = (int)pg.ops.size();
    pg.blk[b].opc = pg.blk[b].op0;
    std::vector<int> release;
    for (int v : block_nodes[b]) {
      NodeOp op{};
      op.node = v; op.lnode = t.left[v]; op.rnode = t.right[v];
      op.lref = t.left[op.lnode] < 0 ? -(dpos(op.lnode) + 1) : slot[op.lnode];
      op.rref = t.left[op.rnode] < 0 ? -(dpos(op.rnode) + 1) : slot[op.rnode];
      for (int c : {op.lnode, op.rnode})
        if (slot[c] >= 0) (fold ? free_slots : release).push_back(slot[c]);
      if (v == 0) { op.dst = -1; op.i0 = R; }
      else {
        if (free_slots.empty()) free_slots.push_back(n_slots++);
        slot[v] = free_slots.back(); free_slots.pop_back();
        op.dst = slot[v];
        op.i0 = (b < nb && js0[v] < std::min(S, (b + 1) * R)) ? std::max(0, js0[v] - b * R) : R;
      }
      op.pair = is_cherry(v) ? pair_of(-(op.lref + 1)) : -1;
      if (op.pair >= 0) pg.blk[b].opc = (int)pg.ops.size() + 1;
      pg.ops.push_back(op);
    }
    pg.blk[b].op1 = (int)pg.ops.size();
    for (int sfree : release) free_slots.push_back(sfree);
  }
  pg.cherry_op.assign(pg.n_cherries, 0);
  pg.sops.assign(pg.ops.size() + 1, SweepOp{});  // + one padding record: the sweep fetches one op ahead
  for (size_t o = 0; o < pg.ops.size(); o++) {
    const NodeOp& op = pg.ops[o];
    SweepOp& s = pg.sops[o];
    auto leaf_bit = [](int ref) { int pos = -(ref + 1); return ((pos >> 5) << 5) | (((pos & 31) - 4) & 31); };
    s.lslot = op.lref < 0 ? 0 : op.lref; s.rslot = op.rref < 0 ? 0 : op.rref;
    s.lbit = op.lref < 0 ? leaf_bit(op.lref) : 0; s.rbit = op.rref < 0 ? leaf_bit(op.rref) : 0;
    s.dst = op.dst; s.i0 = op.i0; s.flags = (op.lref < 0 ? 1 : 0) | (op.rref < 0 ? 2 : 0); s.pair = op.pair;
    if (op.pair >= 0) {
      pg.cherry_op[op.pair] = (int)o;
      s.lbit = (((2 * op.pair) >> 5) << 5) | ((((2 * op.pair) & 31) - 4) & 31);
    }
  }
  pg.n_slots = std::max(n_slots, 1);
  if (pg.lp) {  // leaf heights as integers: |h| / hunit <= 2^14, so the sum over 65,535 leaves stays inside 31 bits
    pg.hunit = pg.has_s1 ? max_leaf_h / 16384. : 0.;
    auto hq = [&](int v) { return pg.has_s1 ? (int32_t)std::lround(t.height[v] / pg.hunit) : 0; };
    pg.leaf_hq.assign(nl, 0);
    for (int v = 0; v < nn; v++) if (t.left[v] < 0 && nn > 1) pg.leaf_hq[dpos(v)] = hq(v);
    for (size_t o = 0; o < pg.ops.size(); o++) {
      const NodeOp& op = pg.ops[o];
      pg.sops[o].hlq = op.lref < 0 ? hq(op.lnode) : 0;
      pg.sops[o].hrq = op.rref < 0 ? hq(op.rnode) : 0;
    }
  }

  // entries: per block, four typed segments (see BlockDesc), each sorted by reference
  struct Ent { int ref, i0, i1, node; bool leaf; };
  const int kPadRef = 1 << 20;  // marks a padding record (sorts behind every real lineage)
  const int kPairRef = 1 << 21; // marks a leaf-pair entry (ref - kPairRef = pair id)
  if (deep && pg.n_slots + 1 > 65535) throw std::runtime_error("program: tree too large for 16-bit lineage references");
  std::vector<std::vector<Ent>> ent(nb);
  if (pg.lp) {
    pg.n_alive_leaves.assign(std::max(S, 1), 0);
    pg.htot.assign(std::max(S, 1), 0.);
    pg.leaf_h.assign(nl, 0.);
    for (int v = 1; v < nn; v++) {
      if (t.left[v] >= 0) continue;
      pg.leaf_h[dpos(v)] = t.height[v];
      for (int j = js0[v]; j < js1[v]; j++) { pg.n_alive_leaves[j]++; pg.htot[j] += t.height[v]; }
    }
  }
  for (int v = 1; v < nn; v++) {
    if (js0[v] >= js1[v]) continue;
    const bool leaf = t.left[v] < 0;
    if (leaf && pg.lp) continue;  // leaf powers: the leaves enter as one power per slice
    for (int b = js0[v] / R; b <= (js1[v] - 1) / R; b++) {
      int jb = b * R, je = std::min(S, jb + R);
      Ent e;
      e.i0 = std::max(js0[v], jb) - jb; e.i1 = std::min(js1[v], je) - jb;
      e.node = v; e.leaf = leaf; e.ref = leaf ? dpos(v) : slot[v];
      if (e.ref > 65535) throw std::runtime_error("program: tree too large for 16-bit lineage references");
      ent[b].push_back(e);
    }
  }
  double fp64 = 0;
  long long n_generic = 0;  // entries served by the generic loops (K/Y-form leaves, partial records)
  // Events folded into node ops: a lineage born inside block b multiplies its alive slices of b at its own op (its
  // partials are in registers there), a lineage that dies inside block b -- when its parent is formed -- multiplies its
  // prefix of b at the parent's op (the merge has just loaded its partials).  Those entries leave the block's lists; what
  // stays partial (leaves that start above the first slice: non-ultrametric trees only) keeps the generic partial loops.
  pg.op_tb.assign(pg.ops.size(), t.H);
  {
    std::vector<int> op_of(nn, -1);
    for (size_t o = 0; o < pg.ops.size(); o++) op_of[pg.ops[o].node] = (int)o;
    for (size_t o = 0; o < pg.ops.size(); o++) {
      const int b = op_block[pg.ops[o].node];
      if (b < nb) pg.op_tb[o] = t.slice_t[pg.blk[b].slice0];
    }
    for (int b = 0; b < nb && fold; b++) {
      const int nsl = pg.blk[b].nslices;
      std::vector<Ent> keep;
      for (const Ent& e : ent[b]) {
        const bool full = e.i0 == 0 && e.i1 == nsl;
        const uint32_t live = (((1u << e.i1) - 1u) & ~((1u << e.i0) - 1u));
        const int v = e.node, par = t.parent[v];
        bool folded = false;
        if (!full) {
          if (!e.leaf && op_of[v] >= 0 && op_block[v] == b) {           // born in this block: own op
            pg.sops[op_of[v]].mself = (int32_t)live; folded = true;
            fp64 += 1 + 2 * (e.i1 - e.i0);
          } else if (e.i0 == 0 && op_of[par] >= 0 && op_block[par] == b) {  // dies in this block: the parent's op
            SweepOp& sp = pg.sops[op_of[par]];
            if (sp.pair >= 0) {  // cherry: both leaves die together; one quadratic in u for the two (first member counts)
              if (sp.ml == 0) { sp.ml = (int32_t)live; fp64 += 3 * (e.i1 - e.i0); }
              folded = true;
            } else {
              (t.left[par] == v ? sp.ml : sp.mr) = (int32_t)live; folded = true;
              fp64 += 1 + 2 * (e.i1 - e.i0);
            }
          }
        }
        if (!folded) keep.push_back(e);
      }
      ent[b].swap(keep);
      // a child's prefix [0,i0) and the new node's suffix [i0,end) partition the block: one update, operands chosen per slice
      const uint32_t all = (1u << nsl) - 1u;
      for (int o = pg.blk[b].opc; o < pg.blk[b].op1 && deep; o++) {  // (used by the global-memory variant only)
        SweepOp& sp = pg.sops[o];
        if (!sp.mself || sp.dst < 0) continue;
        const uint32_t ms = (uint32_t)sp.mself, l = (uint32_t)sp.ml, r = (uint32_t)sp.mr;
        if (l && (l & ms) == 0 && ((l | ms) & all) == all) sp.comb = 1;
        else if (r && (r & ms) == 0 && ((r | ms) & all) == all) sp.comb = 2;
      }
    }
  }
  for (int b = 0; b < nb; b++) {
    const int nsl = pg.blk[b].nslices;
    auto is_full = [&](const Ent& e) { return e.i0 == 0 && e.i1 == nsl; };  // in a short last block the padding slices carry weight 0
    // pairs: both members alive in every slice of the block -> one entry for the two
    std::vector<Ent> list;
    if (pg.leaf_u) {
      std::vector<int> full_member(pg.n_pairs, 0);
      for (const Ent& e : ent[b]) if (e.leaf && is_full(e) && e.ref != kPadRef && pair_of(e.ref) >= 0) full_member[pair_of(e.ref)]++;
      std::vector<char> emitted(pg.n_pairs, 0);
      for (const Ent& e : ent[b]) {
        if (e.leaf && is_full(e) && pair_of(e.ref) >= 0 && full_member[pair_of(e.ref)] == 2) {
          if (emitted[pair_of(e.ref)]) continue;
          emitted[pair_of(e.ref)] = 1;
          Ent pe = e; pe.ref = kPairRef + pair_of(e.ref);
          list.push_back(pe);
        } else list.push_back(e);
      }
    } else list = ent[b];
    // 0 internal full, 1 leaf pair (u), 2 single leaf full (u), 3 single leaf full (K/Y), 4 leaf partial, 5 internal partial
    auto cls_of = [&](const Ent& e) {
      if (!e.leaf) return is_full(e) ? 0 : 5;
      if (e.ref >= kPairRef) return 1;
      if (!is_full(e)) return 4;
      return pg.leaf_u ? 2 : 3;
    };
    if (deep) {  // pad the lineage segment to a multiple of 4 with records on the constant row (slot n_slots, K irrelevant)
      int n0 = 0;
      for (const Ent& e : list) n0 += cls_of(e) == 0;
      for (; n0 % 4; n0++) { Ent e; e.i0 = 0; e.i1 = nsl; e.node = t.leaf_node[0]; e.leaf = false; e.ref = kPadRef; list.push_back(e); }
    }
    std::stable_sort(list.begin(), list.end(), [&](const Ent& x, const Ent& y) {
      int cx = cls_of(x), cy = cls_of(y);
      if (cx != cy) return cx < cy;
      return x.ref < y.ref;
    });
    BlockDescHost& bd = pg.blk[b];
    int seg0[7] = {-1, -1, -1, -1, -1, -1, -1};
    for (const Ent& e : list) {
      int c = cls_of(e), at = (int)pg.entry_ref.size();
      for (int k = 0; k <= c; k++) if (seg0[k] < 0) seg0[k] = at;
      if (c >= 4) pg.n_partial++; else pg.n_full++;
      if (c >= 3) n_generic++;
      const bool padrec = (c == 0 && e.ref == kPadRef);
      if (!padrec) fp64 += (c == 1) ? 3 * (e.i1 - e.i0) : (c == 2 ? 2 * (e.i1 - e.i0) : 1 + 2 * (e.i1 - e.i0));
      const uint32_t live = (((1u << e.i1) - 1u) & ~((1u << e.i0) - 1u));
      int ref = e.ref; uint32_t mk = live;
      if (padrec) ref = pg.n_slots;
      else if (c == 1) { ref = e.ref - kPairRef; mk = 0; }
      else if (c == 2) { mk = 0; }                        // ref = bit position
      else if (c == 3) { ref = e.ref >> 5; mk = 1u << (e.ref & 31); }
      else if (c >= 4) {
        // code in the low half: i0 for a suffix [i0,end), 16+i1 for a prefix [0,i1), 0 = use the mask in the high half
        const bool to_end = e.i1 == nsl;
        uint32_t code = 0;
        if (to_end && e.i0 > 0 && e.i0 < 16) code = (uint32_t)e.i0;
        else if (e.i0 == 0 && e.i1 > 0 && e.i1 < 16) code = 16u + (uint32_t)e.i1;
        mk = code | (live << 16);
      }
      pg.entry_ref.push_back(ref);
      pg.entry_mask.push_back(mk);
      pg.entry_node.push_back(e.node);
      pg.entry_block.push_back(b);
    }
    const int end = (int)pg.entry_ref.size();
    seg0[6] = end;
    for (int k = 5; k >= 0; k--) if (seg0[k] < 0) seg0[k] = seg0[k + 1];
    bd.e_slot0 = seg0[0]; bd.e_pair0 = seg0[1]; bd.e_leafu0 = seg0[2]; bd.e_leaf0 = seg0[3]; bd.e_pleaf0 = seg0[4]; bd.e_pslot0 = seg0[5]; bd.e_end = end;
    // prefetch hints: the upper half of ref names the next record's operand within the same segment
    // (0 = an always-valid dummy at the end of a segment); pair records keep the whole word for the pair id
    for (int k = 0; k < 6; k++) {
      if (k == 1) continue;
      for (int e = seg0[k]; e + 1 < seg0[k + 1]; e++) pg.entry_ref[e] |= (pg.entry_ref[e + 1] & 0xffff) << 16;
    }
  }
  {
    BlockDescHost& bd = pg.blk[nb];
    bd.e_slot0 = bd.e_pair0 = bd.e_leafu0 = bd.e_leaf0 = bd.e_pleaf0 = bd.e_pslot0 = bd.e_end = (int)pg.entry_ref.size();
  }
  // + general node ops (8 per merge; cherries come from a table), block ends (~3 per slice), epilogue (~60)
  int n_cherry_ops = 0;
  for (const NodeOp& op : pg.ops) n_cherry_ops += op.pair >= 0;
  pg.lean = n_generic == 0;
  pg.fp64_instr_per_eval = fp64 + 8.0 * ((double)pg.ops.size() - n_cherry_ops) + 3.0 * S + 60;
  if (pg.lp) pg.fp64_instr_per_eval += 19.0 * S + (pg.has_s1 ? 2.0 : 1.0) * R * nl * 0.67;  // one exp2 per slice; count updates at the leaves' deaths
  return pg;
}

std::vector<uint32_t> permute_patterns(const Program& pg, const uint32_t* rows, long long n, int n_words) {
  std::vector<uint32_t> out((size_t)n * n_words, 0u);
  const int nl = (int)pg.leaf_perm.size();
  for (long long j = 0; j < n; j++) {
    const uint32_t* src = rows + (size_t)j * n_words;
    uint32_t* dst = out.data() + (size_t)j * n_words;
    for (int w = 0; w < n_words; w++) {
      uint32_t bits = src[w];
      while (bits) {  // set bits only: pangenome rows are sparse
        const int k = w * 32 + __builtin_ctz(bits);
        bits &= bits - 1;
        if (k < nl) { const int d = pg.leaf_perm[k]; dst[d >> 5] |= 1u << (d & 31); }
      }
    }
  }
  return out;
}

}  // namespace ppm
