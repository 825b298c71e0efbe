// Builder of the band program (band_program.h): tiles of 32*R slices, typed record lists, one flat descriptor list per warp.
#include "band_program.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace ppm {

namespace {
struct MEnt { int ref, lo, hi; };
struct TileLists {
  std::vector<int> full_int, full_leaf;             // node-array index / leaf bit position
  std::vector<std::vector<int>> f_int, f_leaf;      // per sub-tile k
  std::vector<std::vector<MEnt>> m_int, m_leaf;     // per sub-tile k
};
inline uint32_t desc_code(int op, int variant, int k, int count) {
  return (uint32_t)op | ((uint32_t)variant << 4) | ((uint32_t)k << 6) | ((uint32_t)count << 9);
}
}  // namespace

BandProgram build_band_program(const Tree& t, const Program& pg, int R, int G) {
  BandProgram bp;
  bp.R = R; bp.G = G; bp.T = 32 * R;
  const int nn = t.n_nodes, nl = t.n_leaves, S = t.n_slices, T = bp.T;
  if (!pg.leaf_u) { bp.why = "leaves are not at height 0"; return bp; }
  if (nn > 65535) { bp.why = "more than 65,535 nodes"; return bp; }
  if (nl < 4 || S < 1) { bp.why = "tree too small"; return bp; }
  if (R < 1 || R > 8 || G < 1 || G > 32) { bp.why = "bad R/G"; return bp; }

  // alive slice range of every non-root lineage (functions.cpp:130-141, 250-262), as in build_program
  std::vector<int> js0(nn, S), js1(nn, S);
  for (int v = 1; v < nn; v++) {
    js0[v] = t.slice_start[std::max(t.rank[v], nl - 1)];
    js1[v] = t.slice_start[t.rank[t.parent[v]]];
  }
  auto dpos = [&](int v) { return pg.leaf_perm[t.leaf_pos[v]]; };
  auto pair_of = [&](int pos) { return pos < 2 * pg.n_pairs ? pos / 2 : -1; };

  // node records in birth (height) order; phase A ops grouped by level
  std::vector<int> idx(nn, -1), lev(nn, 0);
  for (int r = 0; r < nn; r++) {
    const int v = t.order[r];
    if (t.left[v] >= 0 && v != 0) idx[v] = bp.n_int++;
  }
  for (int v = nn - 1; v >= 0; v--)
    if (t.left[v] >= 0) lev[v] = 1 + std::max(lev[t.left[v]], lev[t.right[v]]);
  std::vector<int> by_level;
  for (int v = 1; v < nn; v++) if (t.left[v] >= 0) by_level.push_back(v);
  std::stable_sort(by_level.begin(), by_level.end(), [&](int a, int b) { return lev[a] != lev[b] ? lev[a] < lev[b] : idx[a] < idx[b]; });
  bp.lev_off.push_back(0);
  for (size_t i = 0; i < by_level.size(); i++) {
    const int v = by_level[i], l = t.left[v], r = t.right[v];
    if (i > 0 && lev[v] != lev[by_level[i - 1]]) bp.lev_off.push_back((int)i);
    BandOp op{};
    const bool ll = t.left[l] < 0, rl = t.left[r] < 0;
    op.lref = (uint16_t)(ll ? dpos(l) : idx[l]);
    op.rref = (uint16_t)(rl ? dpos(r) : idx[r]);
    op.dst = (uint16_t)idx[v];
    op.flags = (uint16_t)((ll ? 1 : 0) | (rl ? 2 : 0));
    bp.ops.push_back(op);
    bp.op_nodes.push_back(v); bp.op_nodes.push_back(l); bp.op_nodes.push_back(r);
  }
  bp.lev_off.push_back((int)by_level.size());
  bp.n_levels = (int)bp.lev_off.size() - 1;

  bp.n_tiles = (S + T - 1) / T;
  bp.S_pad = bp.n_tiles * T;
  bp.wt.assign(bp.S_pad, 0.);
  bp.slice_t.assign(bp.S_pad, t.H);
  std::copy(t.slice_wt.begin(), t.slice_wt.end(), bp.wt.begin());
  std::copy(t.slice_t.begin(), t.slice_t.end(), bp.slice_t.begin());
  {  // shifts that apply to slice j: nodes born at or before it (birth order = record index, monotone in js0)
    std::vector<int> born;
    for (int r = 0; r < nn; r++) { const int v = t.order[r]; if (idx[v] >= 0) born.push_back(js0[v]); }
    bp.nborn.assign(bp.S_pad, (uint16_t)bp.n_int);
    size_t k = 0;
    for (int j = 0; j < S; j++) {
      while (k < born.size() && born[k] <= j) k++;
      bp.nborn[j] = (uint16_t)k;
    }
  }

  // per tile: classify every lineage that is alive somewhere in it
  std::vector<TileLists> tiles(bp.n_tiles);
  for (auto& tl : tiles) { tl.f_int.resize(R); tl.f_leaf.resize(R); tl.m_int.resize(R); tl.m_leaf.resize(R); }
  for (int v = 1; v < nn; v++) {
    if (js0[v] >= js1[v]) continue;
    const bool leaf = t.left[v] < 0;
    const int ref = leaf ? dpos(v) : idx[v];
    for (int tau = js0[v] / T; tau <= (js1[v] - 1) / T; tau++) {
      const int s0 = std::max(js0[v], tau * T) - tau * T, s1 = std::min(js1[v], (tau + 1) * T) - tau * T;
      TileLists& tl = tiles[tau];
      if (s0 == 0 && s1 == T) { (leaf ? tl.full_leaf : tl.full_int).push_back(ref); continue; }
      for (int k = 0; k < R; k++) {
        const int lo = std::max(s0, 32 * k) - 32 * k, hi = std::min(s1, 32 * k + 32) - 32 * k;
        if (lo >= hi) continue;
        if (lo == 0 && hi == 32) (leaf ? tl.f_leaf : tl.f_int)[k].push_back(ref);
        else (leaf ? tl.m_leaf : tl.m_int)[k].push_back(MEnt{ref, lo, hi});
      }
    }
  }

  // descriptor chunks per tile
  struct Chunk { uint32_t code, off; double cost; };
  std::vector<std::vector<Chunk>> tchunks(bp.n_tiles);
  std::vector<double> tcost(bp.n_tiles, 0.);
  double fp64 = 10.0 * bp.n_int * 32.0 / 32.0;  // phase A: ~10 FP64 per merge, one lane each (counted per lane below)
  double fp64_warp = 0;                          // phase B: FP64 warp instructions
  std::vector<int> members(std::max(pg.n_pairs, 1));
  for (int tau = 0; tau < bp.n_tiles; tau++) {
    TileLists& tl = tiles[tau];
    auto emit = [&](int op, int variant, int k, const std::vector<uint32_t>& ents, double per_rec, double fp_per_rec) {
      for (size_t i = 0; i < ents.size(); i += 32) {
        const int cnt = (int)std::min<size_t>(32, ents.size() - i);
        Chunk c{desc_code(op, variant, k, cnt), (uint32_t)bp.entries.size(), cnt * per_rec + 14.};
        for (int j = 0; j < cnt; j++) bp.entries.push_back(ents[i + j]);
        while (bp.entries.size() % 4) bp.entries.push_back(0u);  // (chunks start 16-byte aligned)
        tchunks[tau].push_back(c);
        tcost[tau] += c.cost;
        fp64_warp += cnt * fp_per_rec;
      }
    };
    // leaves -> pairs (both members in the same list) and singles
    auto split_leaves = [&](std::vector<int>& leaves, std::vector<uint32_t>& pairs, std::vector<uint32_t>& singles) {
      std::sort(leaves.begin(), leaves.end());
      for (int pos : leaves) if (pair_of(pos) >= 0) members[pair_of(pos)] = 0;
      for (int pos : leaves) if (pair_of(pos) >= 0) members[pair_of(pos)]++;
      for (int pos : leaves) {
        const int pr = pair_of(pos);
        if (pr >= 0 && members[pr] == 2) { if ((pos & 1) == 0) pairs.push_back((uint32_t)pr); }
        else singles.push_back((uint32_t)pos);
      }
    };
    std::vector<uint32_t> ints, pairs, singles;
    std::sort(tl.full_int.begin(), tl.full_int.end());
    for (int r : tl.full_int) ints.push_back((uint32_t)r);
    split_leaves(tl.full_leaf, pairs, singles);
    bp.n_full16 += (long long)ints.size() + (long long)singles.size(); bp.n_full32 += (long long)pairs.size();
    emit(BD_INT, BV_FULL, 0, ints, 2.0 * R, 2.0 * R);
    emit(BD_PAIR, BV_FULL, 0, pairs, 3.0 * R, 3.0 * R);
    emit(BD_SGL, BV_FULL, 0, singles, 2.0 * R, 2.0 * R);
    for (int k = 0; k < R; k++) {
      ints.clear(); pairs.clear(); singles.clear();
      std::sort(tl.f_int[k].begin(), tl.f_int[k].end());
      for (int r : tl.f_int[k]) ints.push_back((uint32_t)r);
      split_leaves(tl.f_leaf[k], pairs, singles);
      bp.n_f16 += (long long)ints.size() + (long long)singles.size(); bp.n_f32 += (long long)pairs.size();
      emit(BD_INT, BV_F, k, ints, 2.2, 2.0);
      emit(BD_PAIR, BV_F, k, pairs, 3.2, 3.0);
      emit(BD_SGL, BV_F, k, singles, 2.2, 2.0);
      auto pack = [](const MEnt& e) { return (uint32_t)e.ref | ((uint32_t)e.lo << 16) | ((uint32_t)e.hi << 21); };
      ints.clear(); singles.clear();
      for (const MEnt& e : tl.m_int[k]) ints.push_back(pack(e));
      for (const MEnt& e : tl.m_leaf[k]) singles.push_back(pack(e));
      bp.n_m16 += (long long)ints.size() + (long long)singles.size();
      emit(BD_INT, BV_M, k, ints, 3.5, 2.0);
      emit(BD_SGL, BV_M, k, singles, 3.5, 2.0);
    }
    tcost[tau] += 12.0 * R + 40.;  // tile begin / end
    fp64_warp += 3.0 * R;
  }

  // tiles -> warps: longest-processing-time-first (the tiles near the leaves are the big ones), then ascending per warp
  std::vector<int> order(bp.n_tiles);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return tcost[a] > tcost[b]; });
  std::vector<double> load(G, 0.);
  std::vector<std::vector<int>> mine(G);
  for (int tau : order) {
    const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    load[w] += tcost[tau];
    mine[w].push_back(tau);
  }
  bp.cost_max = *std::max_element(load.begin(), load.end());
  bp.cost_sum = std::accumulate(load.begin(), load.end(), 0.);
  bp.warp_off.push_back(0);
  for (int w = 0; w < G; w++) {
    std::sort(mine[w].begin(), mine[w].end());
    for (int tau : mine[w]) {
      bp.desc.push_back(BandDesc{desc_code(BD_TILE_BEGIN, 0, 0, 0), (uint32_t)tau});
      for (const Chunk& c : tchunks[tau]) bp.desc.push_back(BandDesc{c.code, c.off});
      bp.desc.push_back(BandDesc{desc_code(BD_TILE_END, 0, 0, 0), (uint32_t)tau});
    }
    for (int k = 0; k < 3; k++) bp.desc.push_back(BandDesc{desc_code(BD_END, 0, 0, 0), 0u});  // (the kernel reads two descriptors ahead)
    bp.warp_off.push_back((int)bp.desc.size());
  }
  for (int k = 0; k < 96; k++) bp.desc.push_back(BandDesc{desc_code(BD_END, 0, 0, 0), 0u});  // the kernel's descriptor window reads ahead
  for (int k = 0; k < 64; k++) bp.entries.push_back(0u);  // the gather of the chunk after the last one reads (and ignores) these
  bp.fp64_instr_per_eval = fp64 + 32.0 * fp64_warp + 60;
  bp.ok = true;
  return bp;
}

}  // namespace ppm
