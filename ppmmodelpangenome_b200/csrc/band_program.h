// "Band" program of the slice-parallel sweep (big trees): topology only, built once per tree on the host.
//
// The Mobile integral of likRep2 (reference source/functions.cpp:245-279) is a band matrix: lineage s multiplies
// the slices [js0(s), js1(s)) of the time grid.  The band kernel gives one (pattern, parameter vector) to a CTA of G
// warps and maps LANES TO SLICES: a tile is 32*R consecutive slices, lane l / register k holds slice 32k + l of the
// tile.  Phase A prunes the tree level by level (objects.cpp:427-435 in the linear domain) and leaves one 16-byte record
// (a, -d e^{lam h}) per internal lineage in shared memory; phase B walks, per tile, typed lists of lineage records:
//   FULL   alive in every slice of the tile          -> all R registers, no masks
//   F_k    alive in every slice of sub-tile k only    -> register k, no mask
//   M_k    born and/or merged inside sub-tile k       -> register k, lanes [lo, hi)
// Leaves alive together as a static pair (host_model.h: Program::leaf_perm) enter as ONE quadratic in u = 1 - e^{-lam t}.
// Every warp's work is one flat descriptor list (tile begin / record chunks of <= 32 / tile end), so the kernel's
// control flow is a single loop with the next chunk's records gathered while the current one is multiplied in.
#pragma once
#include <cstdint>
#include <vector>

#include "band_codes.h"
#include "host_model.h"

namespace ppm {

struct BandOp {          // phase A: one internal, non-root node
  uint16_t lref, rref;   // child: index of its record in the node array, or (leaf) its pattern bit position
  uint16_t dst;          // index of this node's record = its birth order among the internal non-root nodes
  uint16_t flags;        // bit 0: left child is a leaf, bit 1: right child is a leaf
};

struct BandProgram {
  bool ok = false;         // false: the tree is not served by the band kernel (why: `why`)
  const char* why = "";
  int R = 4, G = 1, T = 128;
  bool stream = false;     // stream layout (band_stream.cu): ops in birth order (dst = own index), G = work units per item,
                           // every chunk owns exactly `chunk` entries (padded with kBandPadEntry)
  int n_tiles = 0, S_pad = 0;
  int n_int = 0;           // internal non-root nodes
  int n_levels = 0;
  std::vector<BandOp> ops;          // grouped by level (children before parents), [lev_off[l], lev_off[l+1])
  std::vector<int32_t> op_nodes;    // [n_ops][3] node, left child, right child (preorder ids; per-parameter constants)
  std::vector<int32_t> lev_off;
  std::vector<uint32_t> rounds;     // phase A: first op | ops << 16 | (last round of its level) << 30; one padding word
  int n_rounds = 0;
  std::vector<uint16_t> nborn;      // [S_pad] internal non-root nodes born at or before each slice (prefix index of the shifts)
  std::vector<double> wt, slice_t;  // [S_pad] slice weights (0 for padding slices) and midpoints
  std::vector<uint32_t> desc;       // the G flat descriptor lists (band_codes.h), each closed by BD_END
  std::vector<int32_t> warp_off;    // [G+1] first descriptor of each warp
  std::vector<int32_t> warp_ent_off;  // [G+1] first entry of each warp; a warp's chunks own consecutive entries, in list order
  std::vector<uint32_t> blocks;     // stream layout, device form: [block][chunk + 4] = a descriptor's entries + its code
  std::vector<int32_t> unit_blk_off;  // [G] first block of each unit's list
  std::vector<uint32_t> op_span, leaf_span;  // stream layout: alive slices js0 | js1 << 16 per internal lineage / per leaf bit (sparse lists)
  std::vector<uint32_t> entries;    // ref | lo << 16 | hi << 21 | leaf << 31 (M records: lanes [lo, hi), hi in 1..32)
  // Leaf-power program (stream layout only).  Every leaf of an ultrametric tree sits at height 0, so at slice j a leaf observed x
  // contributes f_x(u_j) whatever the leaf: the product over the alive leaves is f_0^n0 f_1^n1 with n1 = the number of PRESENT
  // alive leaves of the pattern -- one exp2 per slice instead of one record per leaf (pair) and tile.  The lists then hold internal
  // lineages only; the prune kernel counts the present leaves that have died before every op (and, when the leaf heights are only
  // numerically 0, the sum of their heights: first-order correction), the integral kernel turns the counts into the power.
  bool leaf_pow = false;
  int chunk = kBandStreamChunk;     // stream layout: entries per chunk (kBandStreamChunkLP for a leaf-power program); a block = chunk + 4 words
  bool has_s1 = false;              // some leaf height != 0: the first-order height correction is carried along
  std::vector<int32_t> n_alive_leaves;  // [S_pad] leaves alive at each slice
  std::vector<double> htot;             // [S_pad] sum of the heights of the leaves alive at each slice
  std::vector<double> op_leaf_h;        // [n_ops][2] heights of an op's leaf children (0 for an internal child)
  std::vector<double> leaf_h;           // [n_leaves] height by device bit position
  double max_leaf_h = 0;
  double fp64_instr_per_eval = 0;   // modelled FP64 instructions per pattern-eval (phase A + phase B)
  double cost_max = 0, cost_sum = 0;  // modelled phase B cost of the busiest warp / of all warps (balance = sum / (G max))
  long long n_full16 = 0, n_full32 = 0, n_f16 = 0, n_f32 = 0, n_m16 = 0;
};

// pg: the handle's sweep program (it fixes the device bit positions of the leaves and the static pairs).
BandProgram build_band_program(const Tree& t, const Program& pg, int R, int G, bool stream = false, bool leaf_pow = false);

}  // namespace ppm
