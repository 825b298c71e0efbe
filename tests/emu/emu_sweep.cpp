// TEST INFRASTRUCTURE ONLY.  Single-lane CPU emulation of the CUDA kernels: compiles the SAME device
// functions (ppmmodelpangenome_b200/csrc/ppm_device_fns.cuh) with g++, one "lane" per warp, so that the
// CPU-only test tier can check the kernel arithmetic and the sweep program against the oracle and the
// golden vectors.  Never part of libppm_b200.so; the product has no host compute path.
#include <cuda_runtime.h>  // vector types only (double2/double4)

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <vector>

#include "../../ppmmodelpangenome_b200/csrc/host_model.h"
#include "../../ppmmodelpangenome_b200/csrc/band_program.h"
#include "../../ppmmodelpangenome_b200/csrc/ppm_device_fns.cuh"
#include "../../ppmmodelpangenome_b200/csrc/band_fns.cuh"

using namespace ppm;

static std::string g_err;

extern "C" const char* emu_error() { return g_err.c_str(); }

// optional posterior outputs of the next emu_eval call (params[0] only): expectedTime1 per pattern [n_patterns],
// expectedTime2 rows [n_patterns][n_slices+1], expectedGains(g2)
static double* g_et1 = nullptr; static double* g_dist = nullptr; static double g_gains_g2 = 0.; static double* g_gains = nullptr;
extern "C" void emu_set_posterior_outputs(double* et1, double* dist, double gains_g2, double* gains) {
  g_et1 = et1; g_dist = dist; g_gains_g2 = gains_g2; g_gains = gains;
}

// Scalar interpreter of the band program (band_program.h) for one (pattern, parameter vector): the same records, masks,
// exponent bookkeeping and tables as the CUDA band kernel, lanes and registers as plain loops.
static double band_item_emu(const ModelDev& md, const TablesDev& t, const BandProgram& bp, const BandTables& bt, int p, long long pat,
                            double* comp3) {
  const double* sc = t.scal + (size_t)p * SC_COUNT;
  const double pi1 = sc[SC_PI1];
  auto bit = [&](int pos) { return (int)((md.words[(size_t)(pos >> 5) * md.n_pat_pad + pat] >> (pos & 31)) & 1u); };
  const double2* leaftab = reinterpret_cast<const double2*>(sc + SC_LA0);
  std::vector<double2> node(std::max(bp.n_int, 1));
  std::vector<int> xcum(bp.n_int + 1, 0), n1cum(bp.n_int + 1, 0);
  std::vector<double> s1cum(bp.n_int + 1, 0.);
  int n1_0 = 0; double s1_0 = 0.;
  if (bp.leaf_pow)
    for (int pos = 0; pos < md.n_leaves; pos++) if (bit(pos)) { n1_0++; s1_0 += bp.leaf_h[pos]; }
  const int n_ops = (int)bp.ops.size();
  for (int o = 0; o < n_ops; o++) {
    const BandOp& op = bp.ops[o];
    double xl, yl, xr, yr;
    if (op.flags & 1) { const double2 v = leaftab[bit(op.lref)]; xl = v.x; yl = v.y; } else { xl = node[op.lref].x; yl = node[op.lref].y; }
    if (op.flags & 2) { const double2 v = leaftab[bit(op.rref)]; xr = v.x; yr = v.y; } else { xr = node[op.rref].x; yr = node[op.rref].y; }
    int k;
    node[op.dst] = band_merge(xl, yl, xr, yr, bt.bop + ((size_t)p * n_ops + o) * BOP_DOUBLES, pi1, k);
    xcum[op.dst + 1] = k;
    if (bp.leaf_pow) {  // present leaves that die with this op (stream layout: dst = birth order = op index)
      const double* oc = bt.bop + ((size_t)p * n_ops + o) * BOP_DOUBLES;
      if ((op.flags & 1) && bit(op.lref)) { n1cum[op.dst + 1]++; s1cum[op.dst + 1] += oc[5]; }
      if ((op.flags & 2) && bit(op.rref)) { n1cum[op.dst + 1]++; s1cum[op.dst + 1] += oc[6]; }
    }
  }
  for (int i = 0; i < bp.n_int; i++) { xcum[i + 1] += xcum[i]; n1cum[i + 1] += n1cum[i]; s1cum[i + 1] += s1cum[i]; }
  const int R = bp.R, T = bp.T;
  double Im = 0.; int Ie = 0x3fffffff;
  std::vector<double> prod(32 * R), Y(32 * R), U(32 * R);
  std::vector<int> ex(32 * R);
  for (int w = 0; w < bp.G; w++) {
    int tau = -1;
    size_t eoff = bp.warp_ent_off[w];
    for (int di = bp.warp_off[w];; di++) {
      const uint32_t code = bp.desc[di];
      const int op = code & 15, var = (code >> 4) & 3, k = (code >> 6) & 7;
      int cnt = (code >> 9) & 127;
      if ((op == BD_REC || op == BD_PAIR) && cnt == 0) cnt = 128;  // (7-bit field)
      if (op == BD_END) break;
      if (op == BD_TILE_BEGIN) {
        tau = (int)(code >> 16);
        for (int kk = 0; kk < R; kk++)
          for (int l = 0; l < 32; l++) {
            const int j = tau * T + 32 * kk + l;
            prod[kk * 32 + l] = 1.; ex[kk * 32 + l] = xcum[bp.nborn[j]];
            Y[kk * 32 + l] = bt.bY[(size_t)p * bp.S_pad + j]; U[kk * 32 + l] = bt.bU[(size_t)p * bp.S_pad + j];
          }
        continue;
      }
      if (op == BD_TILE_LP) {
        if (var) continue;  // (the height tables are read together with the counts below)
        for (int kk = 0; kk < R; kk++)
          for (int l = 0; l < 32; l++) {
            const int j = tau * T + 32 * kk + l, nb = bp.nborn[j];
            const size_t tj = (size_t)p * bp.S_pad + j;
            const double s2 = fma((double)(n1_0 - n1cum[nb]), bt.bLc[tj], bt.bLb[tj]) + (bp.has_s1 ? bt.bLe[tj] * (s1_0 - s1cum[nb]) : 0.);
            int kx;
            prod[kk * 32 + l] *= lp_pow2(s2, kx);
            ex[kk * 32 + l] -= kx;
          }
        continue;
      }
      if (op == BD_TILE_END) {
        for (int kk = 0; kk < R; kk++)
          for (int l = 0; l < 32; l++) {
            const int j = tau * T + 32 * kk + l;
            scaled_add(Im, Ie, prod[kk * 32 + l] * bp.wt[j], ex[kk * 32 + l]);
          }
        continue;
      }
      for (int e = 0; e < cnt; e++) {
        const uint32_t ent = bp.entries[eoff + e];
        if (bp.stream && ent == kBandPadEntry) continue;
        const int ref = ent & 0xffff, lo = (ent >> 16) & 31, hi = (ent >> 21) & 63;
        double4 q = make_double4(1., 0., 0., 0.);
        double2 r = make_double2(1., 0.);
        if (op == BD_REC) r = (ent >> 31) ? bt.bleaf[((size_t)p * md.n_leaves + ref) * 2 + bit(ref)] : node[ref];
        else q = t.pairq[((size_t)p * md.n_pairs + ref) * 4 + (bit(2 * ref) | (bit(2 * ref + 1) << 1))];
        for (int kk = (var == BV_FULL ? 0 : k); kk < (var == BV_FULL ? R : k + 1); kk++)
          for (int l = (var == BV_M ? lo : 0); l < (var == BV_M ? hi : 32); l++) {
            const int x = kk * 32 + l;
            prod[x] *= (op == BD_PAIR) ? fma(fma(q.z, U[x], q.y), U[x], q.x) : fma(r.y, Y[x], r.x);
          }
      }
      eoff += bp.stream ? bp.chunk : cnt;
      for (int x = 0; x < 32 * R; x++) {
        const int be = biased_exp(prod[x]);
        if (be != 0 && be < 1023 - 400) { prod[x] *= pow2i(400); ex[x] += 400; }
      }
    }
  }
  return band_mixture(md, t, p, pat, Im, Ie, comp3);
}

// out_ll[P], out_i2[P], out_parts[P][4]; out_comp [n_patterns][3] / out_cat [n_patterns] for params[0] (may be NULL)
// i2_override: NULL, or the i2 to use for comp/cat (as assignCat receives MLEi2)
// R < 0 selects the deep-prefetch code path (global-memory variant) with |R| slices per block
extern "C" int emu_eval(const char* tree_path, const char* pa_path, int dedupe, int R, const double* params, int P,
                        double* out_ll, double* out_i2, double* out_parts, double* out_comp, signed char* out_cat,
                        long long* out_npat) {
  try {
    Tree tree = read_tree_file(tree_path);
    Patterns pat = read_matrix_file(tree, pa_path, dedupe != 0);
    // R >= 1000: the band program (slice-parallel sweep), R = 1000 + 100 * G + slices-per-lane
    // R >= 3000: the same with the stream layout (band_stream.cu), G = work units per item
    // R >= 5000: stream layout + leaf powers
    const bool band_lp = R >= 5000;
    const bool band_stream = R >= 3000;
    const int band_code = R >= 5000 ? R - 5000 : R >= 3000 ? R - 3000 : R >= 1000 ? R - 1000 : 0;
    if (band_code) R = 8;
    // |R| in [100, 1000): the sweep program with leaf powers (host_model.h: Program::lp), |R| - 100 slices per block
    bool sweep_lp = false;
    if (R >= 100) { sweep_lp = true; R -= 100; }
    if (R <= -100) { sweep_lp = true; R += 100; }
    const bool deep = R < 0;
    if (deep) R = -R;
    Program prog = build_program(tree, R, deep, sweep_lp);
    if (sweep_lp && !prog.lp) throw std::runtime_error("sweep program: leaf powers refused (leaf heights)");
    BandProgram bp;
    if (band_code) {
      bp = build_band_program(tree, prog, band_code % 100, band_code / 100, band_stream, band_lp);
      if (band_lp && !bp.leaf_pow) throw std::runtime_error("band program: leaf powers refused (leaf heights)");
      if (!bp.ok) throw std::runtime_error(std::string("band program: ") + bp.why);
    }
    const long long count = pat.n_patterns;
    if (out_npat) *out_npat = count;
    std::vector<uint32_t> wm((size_t)pat.n_words * count);
    {
      const std::vector<uint32_t> rows = permute_patterns(prog, pat.words.data(), count, pat.n_words);
      for (long long j = 0; j < count; j++)
        for (int w = 0; w < pat.n_words; w++) wm[(size_t)w * count + j] = rows[(size_t)j * pat.n_words + w];
    }
    std::vector<uint32_t> zeros(pat.n_words * 32, 0u);
    static_assert(sizeof(NodeOpDev) == sizeof(NodeOp), "layout");
    static_assert(sizeof(BlockDesc) == sizeof(BlockDescHost), "layout");
    std::vector<double> wt_pad((size_t)std::max(prog.n_blocks, 1) * prog.R, 0.);
    std::copy(tree.slice_wt.begin(), tree.slice_wt.end(), wt_pad.begin());
    ModelDev md{};
    md.lp_max_h = (band_code && bp.leaf_pow) ? bp.max_leaf_h : 0.;
    md.n_nodes = tree.n_nodes; md.n_leaves = tree.n_leaves; md.n_words = pat.n_words; md.n_slices = tree.n_slices;
    md.n_blocks = prog.n_blocks; md.n_slots = prog.n_slots; md.n_entries = (int)prog.entry_ref.size();
    md.n_ops = (int)prog.ops.size(); md.R = prog.R; md.H = tree.H; md.n_obs = pat.n_obs;
    md.left = tree.left.data(); md.right = tree.right.data(); md.parent = tree.parent.data();
    md.bl = tree.bl.data(); md.height = tree.height.data();
    md.n_lev_up = (int)tree.lev_up_off.size() - 1; md.n_lev_dn = (int)tree.lev_dn_off.size() - 1;
    md.lev_up_off = tree.lev_up_off.data(); md.lev_up_nodes = tree.lev_up_nodes.data();
    md.lev_dn_off = tree.lev_dn_off.data(); md.lev_dn_nodes = tree.lev_dn_nodes.data(); md.lev_dn_chain = tree.lev_dn_chain.data();
    md.slice_t = tree.slice_t.data(); md.slice_wt_pad = wt_pad.data();
    md.blk = reinterpret_cast<const BlockDesc*>(prog.blk.data());
    md.ops = reinterpret_cast<const NodeOpDev*>(prog.ops.data());
    static_assert(sizeof(SweepOpDev) == sizeof(SweepOp), "layout");
    md.sops = reinterpret_cast<const SweepOpDev*>(prog.sops.data());
    std::vector<double> op_tb(prog.op_tb); if (op_tb.empty()) op_tb.assign(1, 0.);
    md.op_tb = op_tb.data();
    std::vector<int> cherry_op(prog.cherry_op); if (cherry_op.empty()) cherry_op.assign(1, 0);
    md.cherry_op = cherry_op.data(); md.n_cherries = prog.n_cherries;
    md.entry_ref = prog.entry_ref.data(); md.entry_mask = prog.entry_mask.data();
    md.entry_node = prog.entry_node.data(); md.entry_block = prog.entry_block.data();
    md.n_pairs = prog.n_pairs; md.leaf_node = prog.leaf_node_dev.data();
    md.words = wm.data(); md.n_pat = count; md.n_pat_pad = count;
    md.counts = pat.counts.data(); md.mrca = pat.mrca.data(); md.freq = pat.freq.data(); md.zero_words = zeros.data();
    md.lp = prog.lp ? 1 : 0; md.has_s1 = prog.has_s1 ? 1 : 0;
    md.lp_alive = prog.n_alive_leaves.data(); md.lp_htot = prog.htot.data();
    std::vector<double> s1_0((size_t)std::max<long long>(count, 1), 0.);
    if (prog.has_s1)
      for (long long j = 0; j < count; j++)
        for (int pos = 0; pos < tree.n_leaves; pos++)
          if ((wm[(size_t)(pos >> 5) * count + j] >> (pos & 31)) & 1u) s1_0[j] += prog.leaf_h[pos];
    md.lp_s1_0 = s1_0.data();
    std::vector<int> s1q_0((size_t)std::max<long long>(count, 1), 0);
    if (prog.has_s1)
      for (long long j = 0; j < count; j++)
        for (int pos = 0; pos < tree.n_leaves; pos++)
          if ((wm[(size_t)(pos >> 5) * count + j] >> (pos & 31)) & 1u) s1q_0[j] += prog.leaf_hq[pos];
    md.lp_s1q_0 = s1q_0.data(); md.lp_hunit = prog.hunit;

    const size_t nn = tree.n_nodes;
    std::vector<double4> cls(P * nn);
    std::vector<double2> kap(P * nn), d1u(P * nn);
    std::vector<EntryRec> recs((size_t)P * (prog.entry_ref.size() + 1));
    std::vector<double> yy((size_t)P * std::max(prog.n_blocks, 1) * prog.R);
    std::vector<double> oprec((size_t)P * std::max<size_t>(prog.ops.size(), 1) * OPREC_DOUBLES);
    std::vector<double2> clsT((size_t)P * nn), a01((size_t)P * std::max<long long>(count, 1));
    std::vector<uint32_t> zbits((nn + 31) / 32);
    std::vector<uint16_t> frontier(nn);
    std::vector<double> uu((size_t)P * std::max(prog.n_blocks, 1) * prog.R);
    std::vector<double2> leafab((size_t)P * tree.n_leaves * 2);
    std::vector<double4> pairq((size_t)P * std::max<size_t>(prog.n_pairs, 1) * 4);
    std::vector<double2> cherry((size_t)P * std::max<size_t>(prog.n_cherries, 1) * 4);
    std::vector<double> scal((size_t)P * SC_COUNT), zeta((size_t)P * 6 * nn), i2ovr(P);
    TablesDev t{};
    t.P = P; t.params = params; t.cls = cls.data(); t.kap = kap.data(); t.d1u = d1u.data(); t.recs = recs.data(); t.oprec = oprec.data(); t.clsT = clsT.data(); t.a01 = a01.data(); t.uu = uu.data(); t.leafab = leafab.data(); t.pairq = pairq.data(); t.cherry = cherry.data();
    t.yy = yy.data(); t.scal = scal.data(); t.zeta = zeta.data();
    std::vector<double2> zrow((size_t)P * nn), zpart((size_t)P * std::max(tree.n_slices, 1));
    std::vector<int> zk((size_t)P * nn), zxb((size_t)P * (std::max(prog.n_blocks, 1) + 1));
    t.zrow = zrow.data(); t.zk = zk.data(); t.zxb = zxb.data(); t.zpart = zpart.data();
    std::vector<double> slb((size_t)P * std::max(prog.n_blocks, 1) * prog.R);
    std::vector<double2> slce(slb.size());
    t.slb = slb.data(); t.slce = slce.data();

    // two items per "warp" as on the device (NI = 2), one lane each
    const int NI = 2;
    std::vector<uint32_t> sw(NI * pat.n_words);
    std::vector<double2> slotmem(NI * (prog.n_slots + 1));
    ItemMem<SlotsMem> mem[2];
    for (int q = 0; q < NI; q++) { mem[q].sw_copy = sw.data() + q * pat.n_words; mem[q].sw = mem[q].sw_copy; mem[q].sws = 1; mem[q].slots.base = slotmem.data() + q * (prog.n_slots + 1); }
    std::vector<double> et1v(g_et1 ? (size_t)P * nn : 0);
    if (g_et1) t.et1 = et1v.data();
    if (g_gains) *g_gains = expected_gains(tree, g_gains_g2);
    auto run = [&](int mode, int nP, double* ll, double* comp, signed char* cat, bool force, double* dist = nullptr) {
      SweepArgs a{};
      a.dist = dist;
      a.m = md; a.t = t; a.mode = mode; a.first = 0; a.count = count;
      a.n_batches = a.count; a.n_items = a.n_batches * nP; a.comp = comp; a.cat = cat; a.force = force;
      for (int p = 0; p < nP; p++) {
        double s = 0;
        for (long long bt = 0; bt < a.n_batches; bt += NI) {
          ItemTables T;
          T.recs = recs.data() + (size_t)p * (md.n_entries + 1); T.blk = md.blk; T.ops = md.sops;
          T.oprec = oprec.data() + (size_t)p * md.n_ops * OPREC_DOUBLES;
          T.yy = yy.data() + (size_t)p * md.n_blocks * md.R; T.wt = md.slice_wt_pad;
          T.leaftab = reinterpret_cast<const double2*>(scal.data() + (size_t)p * SC_COUNT + SC_LA0);
          T.uu = uu.data() + (size_t)p * md.n_blocks * md.R;
          T.leafab = leafab.data() + (size_t)p * md.n_leaves * 2;
          T.pairq = pairq.data() + (size_t)p * md.n_pairs * 4;
          T.cherry = cherry.data() + (size_t)p * md.n_cherries * 4;
          T.lb = slb.data() + (size_t)p * md.n_blocks * md.R; T.lce = slce.data() + (size_t)p * md.n_blocks * md.R;
          double c[2] = {0., 0.};
          if (dist && prog.R == 8 && deep) sweep_items<8, 2, true, SlotsMem, true>(a, T, p, bt, 0, mem, c, nullptr);
          else if (dist && prog.R == 8) sweep_items<8, 2, false, SlotsMem, true>(a, T, p, bt, 0, mem, c, nullptr);
          else if (dist) throw std::runtime_error("per-slice output emulated for R = 8");
          else if (deep && prog.R == 8) sweep_items<8, 2, true, SlotsMem>(a, T, p, bt, 0, mem, c, nullptr);
          else if (deep && prog.R == 16) sweep_items<16, 2, true, SlotsMem>(a, T, p, bt, 0, mem, c, nullptr);
          else if (deep) throw std::runtime_error("deep emulation built for R in {8,16}");
          else if (prog.R == 8) sweep_items<8, 2, false, SlotsMem>(a, T, p, bt, 0, mem, c, nullptr);
          else if (prog.R == 4) sweep_items<4, 2, false, SlotsMem>(a, T, p, bt, 0, mem, c, nullptr);
          else if (prog.R == 12) sweep_items<12, 2, false, SlotsMem>(a, T, p, bt, 0, mem, c, nullptr);
          else if (prog.R == 16) sweep_items<16, 2, false, SlotsMem>(a, T, p, bt, 0, mem, c, nullptr);
          else throw std::runtime_error("emulator built for R in {4,8,12,16}");
          s += c[0]; s += c[1];
        }
        if (ll) ll[p] = s;
      }
    };
    for (int p = 0; p < P; p++) prepass_nodes_body(md, t, p, 0, 1);
    for (int p = 0; p < P; p++)
      for (int i = 0; i < md.n_entries + 1 + md.n_blocks * md.R + md.n_ops + md.n_leaves + md.n_pairs + md.n_cherries; i++) prepass_tables_body(md, t, p, i);
    for (long long j = 0; j < count; j++) {
      const int nf = class_frontier(md, md.words + j, (size_t)md.n_pat_pad, zbits.data(), frontier.data());
      for (int p = 0; p < P; p++) a01[(size_t)j * P + p] = class_sums(t, p, frontier.data(), nf);
    }
    for (int p = 0; p < P; p++) zero_row_body(md, t, p, 0, 1);
    for (int p = 0; p < P; p++) prepass_i2_body(md, t, nullptr, p);
    BandDev bd{}; BandTables bt{};
    std::vector<double> bop, bY, bU, bLb, bLc, bLe; std::vector<double2> bleaf;
    if (band_code) {
      static_assert(sizeof(BandOpDev) == sizeof(BandOp), "layout");
      bd.R = bp.R; bd.G = bp.G; bd.T = bp.T; bd.n_tiles = bp.n_tiles; bd.S_pad = bp.S_pad; bd.n_int = bp.n_int; bd.n_ops = (int)bp.ops.size();
      bd.n_levels = bp.n_levels; bd.op_nodes = bp.op_nodes.data(); bd.slice_t = bp.slice_t.data(); bd.wt = bp.wt.data();
      bop.resize((size_t)P * std::max(bd.n_ops, 1) * BOP_DOUBLES); bY.resize((size_t)P * bp.S_pad); bU.resize((size_t)P * bp.S_pad);
      bleaf.resize((size_t)P * tree.n_leaves * 2);
      bt.bop = bop.data(); bt.bY = bY.data(); bt.bU = bU.data(); bt.bleaf = bleaf.data();
      bd.lp = bp.leaf_pow ? 1 : 0; bd.has_s1 = bp.has_s1 ? 1 : 0;
      bd.n_alive_leaves = bp.n_alive_leaves.data(); bd.htot = bp.htot.data(); bd.op_leaf_h = bp.op_leaf_h.data();
      bLb.resize((size_t)P * bp.S_pad); bLc.resize((size_t)P * bp.S_pad); bLe.resize((size_t)P * bp.S_pad);
      bt.bLb = bLb.data(); bt.bLc = bLc.data(); bt.bLe = bLe.data();
      for (int p = 0; p < P; p++)
        for (int i = 0; i < bd.n_ops + bd.S_pad + md.n_leaves; i++) prepass_band_body(md, t, bd, bt, p, i);
      for (int p = 0; p < P; p++) {
        if (scal[(size_t)p * SC_COUNT + SC_BAND] == 0.) throw std::runtime_error("band emulation: parameter vector not eligible (pi1 > 0.95 or lam*H/2 >= 600)");
        double s = 0;
        if (!(scal[(size_t)p * SC_COUNT + SC_I2] < 0.))
          for (long long j = 0; j < count; j++) s += band_item_emu(md, t, bp, bt, p, j, nullptr);
        out_ll[p] = s;
      }
    } else
    run(0, P, out_ll, nullptr, nullptr, false);
    for (int p = 0; p < P; p++) {
      const double* s = &scal[(size_t)p * SC_COUNT];
      out_i2[p] = s[SC_I2];
      if (out_parts) { out_parts[4 * p] = s[SC_P0]; out_parts[4 * p + 1] = s[SC_P1]; out_parts[4 * p + 2] = s[SC_A]; out_parts[4 * p + 3] = s[SC_B]; }
      if (s[SC_I2] < 0.) out_ll[p] = -1e9 * 1. + s[SC_I2];
    }
    if (band_code && out_comp) {
      for (long long j = 0; j < count; j++) band_item_emu(md, t, bp, bt, 0, j, out_comp + 3 * j);
      out_comp = nullptr;
    }
    if (out_comp || out_cat || g_dist) {
      std::vector<double> dummy(1);
      run(2, 1, dummy.data(), out_comp, out_cat, true, g_dist);
    }
    if (g_et1) for (long long j = 0; j < count; j++) g_et1[j] = et1v[pat.mrca[j]];
    g_et1 = nullptr; g_dist = nullptr; g_gains = nullptr;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
