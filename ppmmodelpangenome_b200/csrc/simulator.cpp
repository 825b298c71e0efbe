// PPM gene simulator (input generation for the benchmarks; not part of the likelihood path).
//
// C++ restatement of the reference's R simulators (/root/reference/simulation_aux.R: sim.N :92-118, sim.I1 :121-170,
// sim.I2 :173-238, driven as in run_simulations.R:63-78) on a tree given as flat preorder arrays: a fraction of
// Persistent genes (present at the root, pure loss l0 along the tree), Private genes (one gain -- at the root with weight
// 1/l1 against the tree length L, else uniform on the branches -- then pure loss l1) and Mobile genes (arrival uniform
// on the tree height, then the two-state gain/loss chain g2, l2), observation errors s_10 (1 -> 0; all categories) and
// s_01 (0 -> 1; Mobile only), all-zero genes redrawn, genes kept in category order.  Output: bit-packed rows in the
// C-ABI format (bit k = k-th leaf in preorder).  Gene g is drawn from its own counter-based stream (seed, g), so any
// prefix or slice of the gene list can be regenerated alone, and the result does not depend on the thread count.
// SURVEY 8(d) asks for "run_simulations.R or an equivalent C++ simulator"; the numpy one in simulate.py (same model,
// different random streams) stays for the small seeded test inputs.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../include/ppm_b200.h"
#include "host_model.h"

namespace {

struct Rng {  // splitmix64 seeding + xoshiro256++
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t stream) {
    uint64_t x = seed * 0xd1342543de82ef95ull + stream * 0x2545f4914f6cdd1dull + 0x1234567ull;
    for (int i = 0; i < 4; i++) s[i] = splitmix(x);
  }
  static uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)
};

}  // namespace

extern "C" int ppm_simulate_genes(const ppm_tree* tree, int64_t n_genes, uint64_t seed, const double* frac3, const double* rates_L,
                                  double s_10, double s_01, int32_t n_threads, uint32_t* out_rows, double* out_truth10) {
  if (!tree || !frac3 || !rates_L || !out_rows || n_genes < 1) return PPM_EINVAL;
  try {
    const ppm::Tree t = ppm::tree_from_arrays(tree->n_nodes, tree->left, tree->right, tree->branch_len);
    const int nn = t.n_nodes, nl = t.n_leaves, nw = (nl + 31) / 32;
    const double H = t.H, L = t.L;
    const double l0 = rates_L[0] / L, l1 = rates_L[1] / L, g2 = rates_L[2] / L, l2 = rates_L[3] / L;
    const double lam = g2 + l2, pi1 = lam > 0 ? g2 / lam : 0.;
    const int64_t n0 = (int64_t)std::llround(n_genes * frac3[0]), n1 = (int64_t)std::llround(n_genes * frac3[1]);
    std::vector<double> depth(nn, 0.), keep0(nn), keep1(nn), stay1(nn), gain(nn), cum_bl(nn, 0.);
    for (int v = 1; v < nn; v++) depth[v] = depth[t.parent[v]] + t.bl[v];
    for (int v = 0; v < nn; v++) {
      keep0[v] = std::exp(-l0 * t.bl[v]);
      keep1[v] = std::exp(-l1 * t.bl[v]);
      const double E = std::exp(-lam * t.bl[v]);
      stay1[v] = pi1 + (1 - pi1) * E;  // P(1 -> 1) over the branch
      gain[v] = pi1 * (1 - E);         // P(0 -> 1)
      cum_bl[v] = (v ? cum_bl[v - 1] : 0.) + (v ? t.bl[v] : 0.);
    }
    const double p_root = (1 / l1) / (1 / l1 + L);
    const int T = std::max(1, n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency());
    std::memset(out_rows, 0, sizeof(uint32_t) * (size_t)n_genes * nw);
    auto work = [&](int tid) {
      std::vector<uint8_t> st(nn);
      for (int64_t g = tid; g < n_genes; g += T) {
        const int cat = g < n0 ? 0 : (g < n0 + n1 ? 1 : 2);
        uint32_t* row = out_rows + (size_t)g * nw;
        for (uint64_t attempt = 0;; attempt++) {
          Rng r(seed, (uint64_t)g * 64 + attempt);
          bool any = false;
          if (cat < 2) {
            // pure loss below the (single) gain point: a lost lineage loses its whole subtree -- skip it in preorder
            const double* keep = cat == 0 ? keep0.data() : keep1.data();
            int gain_node = 0;
            double frac_t = 0;
            if (cat == 1 && !(r.uni() < p_root)) {
              const double x = r.uni() * L;
              gain_node = (int)(std::upper_bound(cum_bl.begin() + 1, cum_bl.end(), x) - cum_bl.begin());
              gain_node = std::min(std::max(gain_node, 1), nn - 1);
              frac_t = r.uni();
            }
            // the gene exists in the subtree of gain_node only
            const int v0 = gain_node, v1 = t.last[gain_node];
            bool alive0 = true;
            if (gain_node != 0) alive0 = r.uni() < std::exp(-l1 * t.bl[gain_node] * (1 - frac_t));  // survives the rest of its branch
            if (alive0) {
              st[v0] = 1;
              for (int v = v0; v <= v1;) {
                if (v != v0) {
                  st[v] = st[t.parent[v]] && (r.uni() < keep[v]);
                  if (!st[v]) { v = t.last[v] + 1; continue; }
                }
                if (t.left[v] < 0 && !(r.uni() < s_10)) { const int k = t.leaf_pos[v]; row[k >> 5] |= 1u << (k & 31); any = true; }
                v++;
              }
            }
          } else {
            const double tarr = r.uni() * H;  // depth of arrival: every lineage alive then starts absent
            for (int v = 1; v < nn;) {
              const int pv = t.parent[v];
              uint8_t s;  // 0 / 1 state, 2 = above the arrival (not yet in play)
              if (depth[v] <= tarr) s = 2;
              else if (depth[pv] <= tarr) s = r.uni() < pi1 * (1 - std::exp(-lam * (depth[v] - tarr)));
              else s = r.uni() < (st[pv] == 1 ? stay1[v] : gain[v]);
              st[v] = s;
              if (t.left[v] < 0) {
                const bool obs = s == 1 ? !(r.uni() < s_10) : (r.uni() < s_01);
                if (obs) { const int k = t.leaf_pos[v]; row[k >> 5] |= 1u << (k & 31); any = true; }
              }
              v++;
            }
          }
          if (any) break;
          std::memset(row, 0, sizeof(uint32_t) * nw);  // all-zero gene: redrawn
        }
      }
    };
    std::vector<std::thread> th;
    for (int k = 1; k < T; k++) th.emplace_back(work, k);
    work(0);
    for (auto& x : th) x.join();
    if (out_truth10) {
      out_truth10[0] = (double)n0; out_truth10[1] = l0; out_truth10[2] = (double)n1 * l1 / (1 + l1 * L); out_truth10[3] = l1;
      out_truth10[4] = g2; out_truth10[5] = l2; out_truth10[6] = s_10; out_truth10[7] = s_01; out_truth10[8] = H; out_truth10[9] = L;
    }
    return PPM_OK;
  } catch (const std::bad_alloc&) { return PPM_ENOMEM; }
  catch (const std::exception&) { return PPM_EINVAL; }
}
