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
inline uint32_t desc_code(int op, int variant, int k, int count) {  // (count 128 -> field 0)
  return (uint32_t)op | ((uint32_t)variant << 4) | ((uint32_t)k << 6) | ((uint32_t)(count & 127) << 9);
}
}  // namespace

BandProgram build_band_program(const Tree& t, const Program& pg, int R, int G, bool stream, bool leaf_pow) {
  BandProgram bp;
  bp.R = R; bp.G = G; bp.T = 32 * R; bp.stream = stream;
  const int nn = t.n_nodes, nl = t.n_leaves, S = t.n_slices, T = bp.T;
  if (!pg.leaf_u) { bp.why = "leaves are not at height 0"; return bp; }
  if (nn > 65535) { bp.why = "more than 65,535 nodes"; return bp; }
  if (nl < 4 || S < 1) { bp.why = "tree too small"; return bp; }
  if (R < 1 || R > 8 || G < 1 || G > (stream ? 4096 : 32)) { bp.why = "bad R/G"; return bp; }
  for (int k = 0; k < nl; k++) bp.max_leaf_h = std::max(bp.max_leaf_h, std::fabs(t.height[t.leaf_node[k]]));
  // leaf powers: stream layout only, and only when the leaf heights are rounding noise of the tree file (the height correction is
  // first order; the prepass checks n (lam max_h)^2 per parameter vector)
  bp.leaf_pow = leaf_pow && stream && bp.max_leaf_h <= 1e-8 * t.H;
  bp.has_s1 = bp.leaf_pow && bp.max_leaf_h > 0.;
  const bool want_lp = bp.leaf_pow;

  // alive slice range of every non-root lineage (functions.cpp:130-141, 250-262), as in build_program
  std::vector<int> js0(nn, S), js1(nn, S);
  for (int v = 1; v < nn; v++) {
    js0[v] = t.slice_start[std::max(t.rank[v], nl - 1)];
    js1[v] = t.slice_start[t.rank[t.parent[v]]];
  }
  for (int v = 1; v < nn && bp.leaf_pow; v++)
    if (t.left[v] < 0 && js0[v] != 0) bp.leaf_pow = false;  // (an internal node below a leaf's height: that leaf starts above the first slice)
  if (want_lp && !bp.leaf_pow) bp.has_s1 = false;
  bp.chunk = bp.leaf_pow ? kBandStreamChunkLP : kBandStreamChunk;
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
  // stream layout: birth order (children are born before their parents), so that an op's record index is its own index
  if (stream) std::stable_sort(by_level.begin(), by_level.end(), [&](int a, int b) { return idx[a] < idx[b]; });
  else std::stable_sort(by_level.begin(), by_level.end(), [&](int a, int b) { return lev[a] != lev[b] ? lev[a] < lev[b] : idx[a] < idx[b]; });
  bp.lev_off.push_back(0);
  for (size_t i = 0; i < by_level.size(); i++) {
    const int v = by_level[i], l = t.left[v], r = t.right[v];
    if (!stream && i > 0 && lev[v] != lev[by_level[i - 1]]) bp.lev_off.push_back((int)i);
    BandOp op{};
    const bool ll = t.left[l] < 0, rl = t.left[r] < 0;
    op.lref = (uint16_t)(ll ? dpos(l) : idx[l]);
    op.rref = (uint16_t)(rl ? dpos(r) : idx[r]);
    op.dst = (uint16_t)idx[v];
    op.flags = (uint16_t)((ll ? 1 : 0) | (rl ? 2 : 0));
    bp.ops.push_back(op);
    bp.op_nodes.push_back(v); bp.op_nodes.push_back(l); bp.op_nodes.push_back(r);
    bp.op_leaf_h.push_back(ll ? t.height[l] : 0.); bp.op_leaf_h.push_back(rl ? t.height[r] : 0.);
  }
  bp.lev_off.push_back((int)by_level.size());
  bp.n_levels = (int)bp.lev_off.size() - 1;
  if (stream) { bp.n_levels = 0; for (int v = 0; v < nn; v++) bp.n_levels = std::max(bp.n_levels, lev[v]); }
  // phase A rounds: a level is served 32*G ops at a time; the CTA synchronises after the last round of a level only
  for (int l = 0; !stream && l < bp.n_levels; l++)
    for (int o = bp.lev_off[l]; o < bp.lev_off[l + 1]; o += 32 * G) {
      const int n = std::min(32 * G, bp.lev_off[l + 1] - o);
      const bool last = o + 32 * G >= bp.lev_off[l + 1];
      bp.rounds.push_back((uint32_t)o | ((uint32_t)n << 16) | (last ? 1u << 30 : 0u));
    }
  bp.n_rounds = (int)bp.rounds.size();
  bp.rounds.push_back(0u);  // (the kernel prefetches one round ahead)

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

  if (bp.leaf_pow) {
    bp.n_alive_leaves.assign(bp.S_pad, 0);
    bp.htot.assign(bp.S_pad, 0.);
    bp.leaf_h.assign(nl, 0.);
    for (int v = 1; v < nn; v++) {
      if (t.left[v] >= 0) continue;
      bp.leaf_h[dpos(v)] = t.height[v];
      for (int j = js0[v]; j < js1[v]; j++) { bp.n_alive_leaves[j]++; bp.htot[j] += t.height[v]; }
    }
  }

  // per tile: classify every lineage that is alive somewhere in it
  std::vector<TileLists> tiles(bp.n_tiles);
  for (auto& tl : tiles) { tl.f_int.resize(R); tl.f_leaf.resize(R); tl.m_int.resize(R); tl.m_leaf.resize(R); }
  for (int v = 1; v < nn; v++) {
    if (js0[v] >= js1[v]) continue;
    const bool leaf = t.left[v] < 0;
    if (leaf && bp.leaf_pow) continue;  // the leaves enter as one power per slice
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

  // descriptor chunks per tile.  Internal lineages and single leaves share the 16-byte record lists (entry bit 31 = leaf).
  struct Chunk { uint32_t code; std::vector<uint32_t> ents; double cost; };
  std::vector<std::vector<Chunk>> tchunks(bp.n_tiles);
  std::vector<double> tcost(bp.n_tiles, 0.);
  double fp64 = 10.0 * bp.n_int;  // phase A: ~10 FP64 per merge, one lane each
  double fp64_warp = 0;           // phase B: FP64 warp instructions
  std::vector<int> members(std::max(pg.n_pairs, 1));
  const uint32_t kLeaf = 1u << 31;
  for (int tau = 0; tau < bp.n_tiles; tau++) {
    TileLists& tl = tiles[tau];
    auto emit = [&](int op, int variant, int k, const std::vector<uint32_t>& ents, double per_rec, double fp_per_rec) {
      const size_t ch = stream ? (size_t)bp.chunk : 32;
      for (size_t i = 0; i < ents.size(); i += ch) {
        const int cnt = (int)std::min<size_t>(ch, ents.size() - i);
        Chunk c{desc_code(op, variant, k, cnt), std::vector<uint32_t>(ents.begin() + i, ents.begin() + i + cnt), cnt * per_rec + 60.};
        if (stream) c.ents.resize(ch, kBandPadEntry);  // every chunk owns a full set of entries: the gather needs no bounds
        tcost[tau] += c.cost;
        tchunks[tau].push_back(std::move(c));
        fp64_warp += cnt * fp_per_rec;
      }
    };
    // leaves -> pairs (both members in the same list) and singles
    auto split_leaves = [&](std::vector<int>& leaves, std::vector<uint32_t>& pairs, std::vector<uint32_t>& recs) {
      std::sort(leaves.begin(), leaves.end());
      for (int pos : leaves) if (pair_of(pos) >= 0) members[pair_of(pos)] = 0;
      for (int pos : leaves) if (pair_of(pos) >= 0) members[pair_of(pos)]++;
      for (int pos : leaves) {
        const int pr = pair_of(pos);
        if (pr >= 0 && members[pr] == 2) { if ((pos & 1) == 0) pairs.push_back((uint32_t)pr); }
        else recs.push_back((uint32_t)pos | kLeaf);
      }
    };
    std::vector<uint32_t> recs, pairs;
    std::sort(tl.full_int.begin(), tl.full_int.end());
    for (int r : tl.full_int) recs.push_back((uint32_t)r);
    split_leaves(tl.full_leaf, pairs, recs);
    bp.n_full16 += (long long)recs.size(); bp.n_full32 += (long long)pairs.size();
    emit(BD_PAIR, BV_FULL, 0, pairs, 3.0 * R + 2, 3.0 * R);
    emit(BD_REC, BV_FULL, 0, recs, 2.0 * R + 1, 2.0 * R);
    for (int k = 0; k < R; k++) {
      recs.clear(); pairs.clear();
      std::sort(tl.f_int[k].begin(), tl.f_int[k].end());
      for (int r : tl.f_int[k]) recs.push_back((uint32_t)r);
      split_leaves(tl.f_leaf[k], pairs, recs);
      bp.n_f16 += (long long)recs.size(); bp.n_f32 += (long long)pairs.size();
      emit(BD_PAIR, BV_F, k, pairs, 5.5, 3.0);
      emit(BD_REC, BV_F, k, recs, 3.5, 2.0);
      auto pack = [](const MEnt& e) { return (uint32_t)e.ref | ((uint32_t)e.lo << 16) | ((uint32_t)e.hi << 21); };
      recs.clear();
      for (const MEnt& e : tl.m_int[k]) recs.push_back(pack(e));
      for (const MEnt& e : tl.m_leaf[k]) recs.push_back(pack(e) | kLeaf);
      bp.n_m16 += (long long)recs.size();
      emit(BD_REC, BV_M, k, recs, 5.5, 2.0);
    }
    tcost[tau] += 12.0 * R + 60.;  // tile begin / end
    fp64_warp += 3.0 * R;
    if (bp.leaf_pow) { tcost[tau] += 30.0 * R + 60.; fp64_warp += 20.0 * R; }
  }

  // tiles -> warps: longest-processing-time-first (the tiles near the leaves are the big ones), then ascending per warp.
  // Per warp: one flat list of one-word descriptors and, in the same order, the entries of its chunks back to back.
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
  bp.warp_ent_off.push_back(0);
  for (int w = 0; w < G; w++) {
    std::sort(mine[w].begin(), mine[w].end());
    for (int tau : mine[w]) {
      bp.desc.push_back(desc_code(BD_TILE_BEGIN, 0, 0, 0) | ((uint32_t)tau << 16));
      if (bp.has_s1) bp.desc.push_back(desc_code(BD_TILE_LP, 1, 0, 0) | ((uint32_t)tau << 16));
      if (bp.leaf_pow) bp.desc.push_back(desc_code(BD_TILE_LP, 0, 0, 0) | ((uint32_t)tau << 16));
      for (const Chunk& c : tchunks[tau]) {
        bp.desc.push_back(c.code);
        bp.entries.insert(bp.entries.end(), c.ents.begin(), c.ents.end());
      }
      bp.desc.push_back(desc_code(BD_TILE_END, 0, 0, 0) | ((uint32_t)tau << 16));
    }
    for (int k = 0; k < 8; k++) bp.desc.push_back(desc_code(BD_END, 0, 0, 0));  // (the kernel looks five descriptors ahead)
    bp.warp_off.push_back((int)bp.desc.size());
    bp.warp_ent_off.push_back((int)bp.entries.size());
  }
  if (stream && S < 65536) {  // alive slices of every internal lineage (by record index) and leaf (by device bit position)
    bp.op_span.assign(bp.ops.size(), 0u);
    for (size_t i = 0; i < bp.ops.size(); i++) { const int v = bp.op_nodes[3 * i]; bp.op_span[i] = (uint32_t)js0[v] | ((uint32_t)js1[v] << 16); }
    bp.leaf_span.assign(nl, 0u);
    for (int v = 1; v < nn; v++) if (t.left[v] < 0) bp.leaf_span[dpos(v)] = (uint32_t)js0[v] | ((uint32_t)js1[v] << 16);
  }
  if (stream) {
    // device form of the lists: one block of chunk + 4 words per descriptor = its `chunk` entries + its code.
    // A tile-begin block carries, instead of entries, the shift-prefix index of each of its slices (nborn), packed so that
    // lane l finds slices l, 32 + l, ... in words 2l, 2l + 1.  Every list is followed by END blocks (the kernel reads ahead).
    if (R > 4) { bp.why = "stream layout: at most 4 slices per lane"; return bp; }
    for (int w = 0; w < G; w++) {
      bp.unit_blk_off.push_back((int32_t)(bp.blocks.size() / (bp.chunk + 4)));
      size_t eoff = bp.warp_ent_off[w];
      for (int di = bp.warp_off[w];; di++) {
        const uint32_t code = bp.desc[di];
        const int op = code & 15;
        std::vector<uint32_t> blk(bp.chunk + 4, kBandPadEntry);
        blk[bp.chunk] = code;
        if (op == BD_REC || op == BD_PAIR) {
          std::copy(bp.entries.begin() + eoff, bp.entries.begin() + eoff + bp.chunk, blk.begin());
          eoff += bp.chunk;
        } else if (op == BD_TILE_BEGIN || op == BD_TILE_LP) {
          const int tau = (int)(code >> 16);
          for (int l = 0; l < 32; l++) {
            uint32_t h[4] = {0, 0, 0, 0};
            for (int k = 0; k < R; k++) h[k] = bp.nborn[tau * T + 32 * k + l];
            blk[2 * l] = h[0] | (h[1] << 16);
            blk[2 * l + 1] = h[2] | (h[3] << 16);
          }
        }
        bp.blocks.insert(bp.blocks.end(), blk.begin(), blk.end());
        if (op == BD_END) break;
      }
      std::vector<uint32_t> endblk(bp.chunk + 4, kBandPadEntry);
      endblk[bp.chunk] = desc_code(BD_END, 0, 0, 0);
      for (int k = 0; k < 5; k++) bp.blocks.insert(bp.blocks.end(), endblk.begin(), endblk.end());
    }
  }
  for (int k = 0; k < 96; k++) bp.desc.push_back(desc_code(BD_END, 0, 0, 0));  // the kernel's descriptor window reads ahead
  for (int k = 0; k < 256; k++) bp.entries.push_back(stream ? kBandPadEntry : 0u);  // a chunk's entry load always reads a full set
  bp.fp64_instr_per_eval = fp64 + 32.0 * fp64_warp + 60;
  bp.ok = true;
  return bp;
}

}  // namespace ppm
