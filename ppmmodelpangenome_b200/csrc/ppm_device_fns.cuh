// Device functions of the PPM likelihood engine, written once as __host__ __device__ so that the very same
// source is (1) the body of the CUDA kernels in ppm_kernels.cu and (2) compiled by g++ into the single-lane
// emulator under tests/emu/, which lets the CPU-only test tier check the kernel arithmetic against the
// oracle.  The product library (libppm_b200.so) contains ONLY the CUDA instantiation: there is no host path.
#pragma once
#include "ppm_kernels.cuh"

#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define PPM_HD __host__ __device__ __forceinline__
#else
#define PPM_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PPM_LANES 32
#define ppm_syncwarp() __syncwarp()
#define ppm_ballot(pred) __ballot_sync(0xffffffffu, (pred))
#define ppm_any(pred) __any_sync(0xffffffffu, (pred))
#define ppm_all(pred) __all_sync(0xffffffffu, (pred))
#define ppm_hiloint2double(hi, lo) __hiloint2double((hi), (lo))
#define ppm_double2hiint(x) __double2hiint(x)
#else  // single-lane emulation
#define PPM_LANES 1
#define ppm_syncwarp() ((void)0)
#define ppm_ballot(pred) ((pred) ? 1u : 0u)
#define ppm_any(pred) (pred)
#define ppm_all(pred) (pred)
static inline double ppm_hiloint2double(int hi, int lo) {
  unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double d; std::memcpy(&d, &u, 8); return d;
}
static inline int ppm_double2hiint(double x) {
  unsigned long long u; std::memcpy(&u, &x, 8); return (int)(u >> 32);
}
#endif

#if defined(__CUDA_ARCH__)
#define ppm_cp_async16(dst, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define ppm_cp_commit() asm volatile("cp.async.commit_group;" ::: "memory")
#define ppm_cp_wait1() asm volatile("cp.async.wait_group 1;" ::: "memory")
#define ppm_cp_wait0() asm volatile("cp.async.wait_group 0;" ::: "memory")
#else
#define ppm_cp_async16(dst, src) std::memcpy((dst), (src), 16)
#define ppm_cp_commit() ((void)0)
#define ppm_cp_wait1() ((void)0)
#define ppm_cp_wait0() ((void)0)
#endif

namespace ppm {

// ------------------------------------------------------------------------------------------------ helpers
PPM_HD double lse2(double a, double b) {  // logSumExp (objects.cpp:21-33) without the +-100 cut
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double m = fmax(a, b), d = fmin(a, b) - m;
  return m + log1p(exp(d));
}
PPM_HD double pow2i(int k) {  // 2^k for k in [-1022, 1023]
  return ppm_hiloint2double((1023 + k) << 20, 0);
}
PPM_HD int biased_exp(double x) { return (ppm_double2hiint(x) >> 20) & 0x7ff; }
PPM_HD void scaled_add(double& Im, int& Ie, double v, int e) {  // (Im * 2^-Ie) += v * 2^-e ; Ie starts at INT_MAX/2
  if (v != 0.) {
    const int en = e < Ie ? e : Ie;
    int d1 = Ie - en, d2 = e - en;   // one of them is 0
    d1 = d1 > 1000 ? 1000 : d1; d2 = d2 > 1000 ? 1000 : d2;
    Im = fma(Im, pow2i(-d1), v * pow2i(-d2));
    Ie = en;
  }
}

// Leaf powers: 2^s = p * 2^k with p in [0.70, 1.42]; s is the log2 of a product of leaf terms (band_program.h: leaf_pow), |s| < 2^31
// (the prepass clamps the log2 of an impossible observation at -1e4 per leaf).  Round to the nearest integer with the 1.5 * 2^52
// trick, then 2^r on |r| <= 1/2 by the degree-10 Chebyshev interpolant (relative error 3.1e-16).
PPM_HD double lp_pow2(double s, int& k) {
  const double magic = 6755399441055744.0;
  const double t = s + magic;
#if defined(__CUDA_ARCH__)
  k = __double2loint(t);
#else
  { long long b; memcpy(&b, &t, 8); k = (int)(uint32_t)b; }
#endif
  const double r = s - (t - magic);
  double p = 7.072585949269223e-09;
  p = fma(p, r, 1.0208690299958306e-07);
  p = fma(p, r, 1.321544258792169e-06);
  p = fma(p, r, 1.5252657260200837e-05);
  p = fma(p, r, 0.0001540353044173605);
  p = fma(p, r, 0.0013333558230164974);
  p = fma(p, r, 0.009618129107606888);
  p = fma(p, r, 0.05550410866444772);
  p = fma(p, r, 0.24022650695910097);
  p = fma(p, r, 0.69314718055995);
  p = fma(p, r, 1.0);
  return p;
}
// int -> double on the FP64 pipe (one DADD): 2^52 + 2^31 + n assembled from its two words, minus the bias
PPM_HD double lp_i2d(int n) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(0x43300000, n ^ (int)0x80000000) - 4503601774854144.0;
#else
  return (double)n;
#endif
}
// acc += a * b on the slices at or after i0 (a predicated DFMA: no select instructions)
PPM_HD void lp_fma_ge(double& acc, double a, double b, int i, int i0) {
#if defined(__CUDA_ARCH__)
  asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %3, %4;\n\t@p fma.rn.f64 %0, %1, %2, %0;\n\t}" : "+d"(acc) : "d"(a), "d"(b), "r"(i), "r"(i0));
#else
  if (i >= i0) acc = fma(a, b, acc);
#endif
}

// Mobile merge of two child lineages (x, y) = (a, d): m0 = x - k0 y, m1 = x + k1 y along each child's branch, multiplied
// (objects.cpp:427-435 in the linear domain); orc = {k0_l, k1_l, k0_r, k1_r}.  Returns the parent's (a, d), not renormalised.
PPM_HD void merge_children(double2 L, double2 Rr, const double* orc, double pi1, double& av, double& d) {
  const double m0 = fma(-orc[0], L.y, L.x) * fma(-orc[2], Rr.y, Rr.x);
  const double m1 = fma(orc[1], L.y, L.x) * fma(orc[3], Rr.y, Rr.x);
  d = m1 - m0;
  av = fma(pi1, d, m0);
}

// ------------------------------------------------------------------------------------------------ prepass
#if defined(__CUDA_ARCH__)
#define PPM_CTA_BARRIER() __syncthreads()
#else
#define PPM_CTA_BARRIER() ((void)0)
#endif

// One CTA per parameter vector (tid/nt = thread and CTA size; the host emulation runs it with nt = 1): per-node
// tables that need a tree recursion -- zero-row pruning of the two pure-loss categories, getPobs1, likUpper1.
// Independent per-node transcendentals are spread over the threads; the recursions run level by level
// (m.lev_up: leaves first; m.lev_dn: root first), so the serial depth is the tree depth, not the node count.
PPM_HD void prepass_nodes_body(const ModelDev& m, const TablesDev& t, int p, int tid, int nt) {
  const double* par = t.params + 8 * (size_t)p;
  const double l0 = par[1], l1 = par[3], g2 = par[4], l2 = par[5], e2 = par[6], e1 = par[7];
  const int nn = m.n_nodes;
  double4* cls = t.cls + (size_t)p * nn;
  double2* kap = t.kap + (size_t)p * nn;
  double2* d1u = t.d1u + (size_t)p * nn;
  double* z0 = t.zeta + (size_t)p * 6 * nn;  // zero-row log p1 per node for l0; later log V
  double* z1 = z0 + nn;                      // ... for l1; later U
  double* ta = z1 + nn;                      // getPobs1 terms
  double* ss0 = ta + nn;                     // sum of the "active" factors S over the subtree of each node (l0)
  double* ss1 = ss0 + nn;                    // ... (l1)
  double* sc = t.scal + (size_t)p * SC_COUNT;

  const double lam = g2 + l2;
  const double pi1 = (g2 == 0.) ? 0. : g2 / lam;           // functions.cpp:227-231
  const double pi0 = (g2 == 0.) ? 1. : l2 / lam;
  const double log_e2 = log(e2), log_1me2 = log(1 - e2);

  // (A) per node, independent: branch survival logs and the Mobile branch factors.
  //     cls[v] = {S0, -, S1, -}; .y/.w become log T0 / log T1 after (B); d1u[v] = (sigma0, sigma1) until (D) ends
  for (int v = tid; v < nn; v += nt) {
    const bool leaf = m.left[v] < 0;
    const double bl = m.bl[v];
    double4 c;
    // active branch (a present leaf below): survive the branch; a leaf adds its own log(1-e2)   (objects.cpp:408,423)
    c.x = -l0 * bl + (leaf ? log_1me2 : 0.);
    c.z = -l1 * bl + (leaf ? log_1me2 : 0.);
    const double sg0 = exp(-l0 * bl), sg1 = exp(-l1 * bl);
    c.y = 0.; c.w = 0.;  // (T0 / T1: written after (B))
    cls[v] = c;
    d1u[v] = make_double2(sg0, sg1);  // branch survival, read by (B) (d1u is written for good at the end of (D))
    const double E = exp(-lam * bl);
    kap[v] = make_double2(pi1 * E, pi0 * E);
  }
  PPM_CTA_BARRIER();
  // (B) bottom-up by levels: the frontier factor T_v = (1 - sigma_v) + sigma_v Z_v, Z_v = T_a T_b (a leaf: Z = e2) -- the
  //     reference's LSE(lost on the branch, kept and unobserved below) of objects.cpp:404-426 on the all-zero row -- in the
  //     LINEAR domain with an integer exponent (value = tm 2^-te): a product over thousands of leaves leaves the double range,
  //     and a chain of 10,000 dependent log1p(exp()) pairs is what a caterpillar's node prepass used to wait for.  The logs
  //     (cls.y/.w, z0/z1) are taken afterwards, one node per thread.  During the walk z0/z1 hold tm, ta/sg hold te.
  // (levels of at most 32 nodes -- every level of a caterpillar -- are served by warp 0 alone; a run of them is kept in step by
  // __syncwarp instead of the CTA barrier)
  // A chain is walked by ONE thread: it keeps the node it has just finished in registers and forwards it to the parent, and
  // requests the next level's node and child ids one level ahead.
  double* sg = ss1 + nn;             // sixth scratch array
  int fv = -1;                       // the node this thread finished last ...
  double ftm0 = 0., ftm1 = 0.; int fte0 = 0, fte1 = 0;  // ... and what its parent reads
  double2 fzr = make_double2(0., 0.); double fs0 = 0., fs1 = 0.;
  int nk_ = -1, nv_ = 0, na_ = 0, nb_ = 0;  // ids requested ahead for list position nk_
  // T = (1 - sigma) + sigma Zm 2^-Ze, renormalised when it has become small
  auto frontier_lin = [&](double sig, double Zm, int Ze, double& tm, int& te) {
    const double A = 1 - sig;
    double Im = A; int Ie = (A != 0.) ? 0 : 0x3fffffff;
    scaled_add(Im, Ie, sig * Zm, Ze);
    if (Im == 0.) Ie = 0;
    const int be = biased_exp(Im);
    if (be != 0 && be < 1023 - 400) { Im *= pow2i(400); Ie += 400; }
    tm = Im; te = Ie;
  };
  // (the level offsets travel in registers, the next one requested a level ahead: a load of lev_up_off at the head of every
  //  level cost every warp an L2 round trip per level -- half of the 3 us a caterpillar's level took)
  int uo0 = m.lev_up_off[0], uo1 = m.lev_up_off[1], uo2 = m.n_lev_up >= 2 ? m.lev_up_off[2] : uo1;
  for (int lv = 0; lv < m.n_lev_up; lv++) {
    const int uo3 = lv + 3 <= m.n_lev_up ? m.lev_up_off[lv + 3] : uo2;  // (used from the next level on)
    const bool small_lv = PPM_LANES == 32 && uo1 - uo0 <= 32;
    const bool next_small = small_lv && lv + 1 < m.n_lev_up && uo2 - uo1 <= 32;
    if (!small_lv || tid < 32)
    for (int k = uo0 + tid; k < uo1; k += nt) {
      int v, a, b;
      if (k == nk_) { v = nv_; a = na_; b = nb_; }
      else { v = m.lev_up_nodes[k]; a = m.left[v]; b = m.right[v]; }
      if (lv + 1 < m.n_lev_up) {  // this thread's first node of the next level
        const int kn = uo1 + tid;
        if (kn < uo2) { nk_ = kn; nv_ = m.lev_up_nodes[kn]; na_ = m.left[nv_]; nb_ = m.right[nv_]; }
      }
      const bool leaf = a < 0;
      double Zm0 = e2, Zm1 = e2; int Ze0 = 0, Ze1 = 0;
      if (!leaf) {
        const bool fa = a == fv, fb = b == fv;
        Zm0 = (fa ? ftm0 : z0[a]) * (fb ? ftm0 : z0[b]); Ze0 = (fa ? fte0 : (int)ta[a]) + (fb ? fte0 : (int)ta[b]);
        Zm1 = (fa ? ftm1 : z1[a]) * (fb ? ftm1 : z1[b]); Ze1 = (fa ? fte1 : (int)sg[a]) + (fb ? fte1 : (int)sg[b]);
      }
      const double2 sgv = d1u[v];
      double tm0, tm1; int te0, te1;
      frontier_lin(sgv.x, Zm0, Ze0, tm0, te0);
      frontier_lin(sgv.y, Zm1, Ze1, tm1, te1);
      if (v == 0) { tm0 = 1.; tm1 = 1.; te0 = 0; te1 = 0; }  // (the root has no branch: T = 1)
      z0[v] = tm0; z1[v] = tm1; ta[v] = (double)te0; sg[v] = (double)te1;
      double2 zrv = make_double2(0., 0.);
      if (t.zrow) {
        // Mobile partials of the all-zero row (the zero-row integral no longer runs through the sweep): leaves observe 0
        // (objects.cpp:407-408), internal nodes merge their children as the sweep does, renormalised the same way
        double2* zr = t.zrow + (size_t)p * nn;
        int* zk = t.zk + (size_t)p * nn;
        if (leaf) {
          const double m0 = 1 - e1, m1 = e2;
          zrv = make_double2(m0 + pi1 * (m1 - m0), m1 - m0);
          zr[v] = zrv; zk[v] = 0;
        } else {
          const double orc[4] = {kap[a].x, kap[a].y, kap[b].x, kap[b].y};
          double av, d;
          merge_children(a == fv ? fzr : zr[a], b == fv ? fzr : zr[b], orc, pi1, av, d);
          const int ha = ppm_double2hiint(av) & 0x7fffffff, hd = ppm_double2hiint(d) & 0x7fffffff;
          const int h = ha > hd ? ha : hd;
          const bool small = (uint32_t)(h - 0x00100000) < (uint32_t)(((1023 - 40) << 20) - 0x00100000);
          const int kk = small ? 1022 - (h >> 20) : 0;
          const double sfac = pow2i(kk);
          zrv = make_double2(av * sfac, d * sfac);
          zr[v] = zrv; zk[v] = kk;
        }
      }
      // class-sum tables: with Stot = sum of S over ALL branches, log p1(root) = Stot + sum over the maximal all-zero
      // subtrees c of F_c, F_c = T_c - (sum of S over the subtree of c)
      const double4 c = cls[v];
      const double cx = v == 0 ? 0. : c.x, cz = v == 0 ? 0. : c.z;
      const double s0 = cx + (leaf ? 0. : (a == fv ? fs0 : ss0[a]) + (b == fv ? fs0 : ss0[b]));
      const double s1 = cz + (leaf ? 0. : (a == fv ? fs1 : ss1[a]) + (b == fv ? fs1 : ss1[b]));
      ss0[v] = s0; ss1[v] = s1;
      fv = v; ftm0 = tm0; ftm1 = tm1; fte0 = te0; fte1 = te1; fzr = zrv; fs0 = s0; fs1 = s1;
    }
    // (between two one-node levels nothing is exchanged between threads: thread 0 walks the chain in program order, and the
    //  fence a __syncwarp implies -- every store of the level performed before the next level's loads issue -- is what a
    //  10,000-level chain used to pay 3 us per level for)
    const bool chain = next_small && uo1 - uo0 == 1 && uo2 - uo1 == 1;
    if (chain) {} else if (next_small) { if (tid < 32) ppm_syncwarp(); } else PPM_CTA_BARRIER();
    uo0 = uo1; uo1 = uo2; uo2 = uo3;
  }
  // the logs, one node per thread: T0 / T1 into cls ...
  for (int v = tid; v < nn; v += nt) {
    double4 c = cls[v];
    c.y = log(z0[v]) - ta[v] * 0.693147180559945309417232;
    c.w = log(z1[v]) - sg[v] * 0.693147180559945309417232;
    if (v == 0) c = make_double4(0., 0., 0., 0.);
    cls[v] = c;
  }
  PPM_CTA_BARRIER();
  // ... then, per node, the zero-row log p1 (the children's T), the class-sum table entry and the getPobs1 term
  // (functions.cpp:284-299)
  for (int v = tid; v < nn; v += nt) {
    const int a = m.left[v], b = m.right[v];
    const bool leaf = a < 0;
    const double zl0 = leaf ? log_e2 : cls[a].y + cls[b].y;
    const double zl1 = leaf ? log_e2 : cls[a].w + cls[b].w;
    const double4 c = cls[v];
    t.clsT[(size_t)v * t.P + p] = make_double2(c.y - ss0[v], c.w - ss1[v]);
    const double lost = 1 - exp(-l1 * m.bl[v]);
    const double tav = leaf ? lost * (1 - e2) : lost * (1 - exp(zl1));
    z0[v] = zl0; z1[v] = zl1; ta[v] = tav;
  }
  PPM_CTA_BARRIER();
  // (C) computeI2 pieces that do not need the Mobile integral (functions.cpp:337-341), summed in the reference's order
  if (tid == 0) {
    double p0 = (l0 == 0.) ? 1 - pow(e2, (double)m.n_leaves) : 1 - exp(z0[0]);
    double p1 = 1 - exp(z1[0]);
    double A = 0;
    for (int v = 1; v < nn; v++) A += ta[v];
    A /= l1;
    sc[SC_PI1] = pi1;
    double m0 = 1 - e1, m1 = e2;           // leaf observed 0 (objects.cpp:407-408)
    sc[SC_LD0] = m1 - m0; sc[SC_LA0] = m0 + pi1 * (m1 - m0);
    m0 = e1; m1 = 1 - e2;                  // leaf observed 1
    sc[SC_LD1] = m1 - m0; sc[SC_LA1] = m0 + pi1 * (m1 - m0);
    sc[SC_P0] = p0; sc[SC_P1] = p1; sc[SC_A] = A;
    sc[SC_LOG1ME2] = log_1me2;
    sc[SC_STOT0] = ss0[0]; sc[SC_STOT1] = ss1[0];
    sc[SC_FACTORED] = (pi1 > 0.95) ? 1. : 0.;
    // band kernel (big trees): expanded leaf-pair quadratics need pi1 <= 0.95, and its exp tables are taken relative to
    // the mid-height of the tree, so lam * H / 2 must stay well inside the double range
    // (leaf-power program: the leaf heights enter to first order only)
    sc[SC_BAND] = (pi1 <= 0.95 && lam * m.H * 0.5 < 600. && (double)m.n_leaves * (lam * m.lp_max_h) * (lam * m.lp_max_h) <= 1e-12) ? 1. : 0.;
    sc[SC_I2] = 1.;  // placeholder until prepass_i2 (the zero-row sweep must run regardless)
    sc[SC_IZ_M] = 0.; sc[SC_IZ_E] = 0.;
    z0[0] = 0.; z1[0] = 0.;
    d1u[0] = make_double2(0., -INFINITY);
    // expectedTime1 (functions.cpp:637-690) as a top-down recursion: for a gain node g with parent p, branch b, sibling s,
    //   D(g) = (1 - e^{-l1 b}) + e^{-l1 b} T(s) D(p),   N(g) = (1/l1 - e^{-l1 b}(b + 1/l1)) + e^{-l1 b} T(s) (N(p) + b D(p)),
    // D(root) = 1, N(root) = 1/l1, expected time = N/D (the reference's sums over the root path, regrouped)
    if (t.et1) { t.et1[(size_t)p * nn] = 1. / l1; ss0[0] = 1.; }
  }
  PPM_CTA_BARRIER();
  // (D) top-down by levels: U (likUpper1, functions.cpp:176-209, as U(v) = lost_v + sigma_v*T(sib)*U(parent),
  //     U(child of root) = lost_v) and log V(v) = sum over the path of log sigma_c + T(sib c).
  //     z0 is reused as log V, z1 as U: a node's own zero-row values are dead once level (B) is complete.
  // The transcendentals of a node do not depend on its parent: sigma_v = exp(-l1 bl_v) and sigma_v T(sib v) are taken first, one
  // node per thread (sg: the sixth scratch array; ta is dead after (C)); the chain itself is two FMAs and two additions per node
  // (10,000 levels: a few ms instead of 14); log U and the table entry follow, again one node per thread.  Same operations on the
  // same operands as the one-pass form: bitwise the same tables.
  for (int v = 1 + tid; v < nn; v += nt) {
    const int par_ = m.parent[v];
    const int sib = (m.left[par_] == v) ? m.right[par_] : m.left[par_];
    const double sig = exp(-l1 * m.bl[v]);
    sg[v] = sig;
    ta[v] = sig * exp(cls[sib].w);
  }
  PPM_CTA_BARRIER();
  // (forwarding and requests one level ahead as in (B): here a node reads its PARENT, finished by the same thread on a chain)
  fv = -1; nk_ = -1;
  double fU = 0., fV = 0., fD = 0., fN = 0.;
  int np_ = 0, ns_ = 0;
  int do0 = m.n_lev_dn >= 1 ? m.lev_dn_off[1] : 0, do1 = m.n_lev_dn >= 2 ? m.lev_dn_off[2] : do0, do2 = m.n_lev_dn >= 3 ? m.lev_dn_off[3] : do1;
  int dc0 = m.n_lev_dn >= 2 ? m.lev_dn_chain[1] : 0;
  for (int lv = 1; lv < m.n_lev_dn; lv++) {
    const int do3 = lv + 3 <= m.n_lev_dn ? m.lev_dn_off[lv + 3] : do2;
    const int dc1 = lv + 1 < m.n_lev_dn ? m.lev_dn_chain[lv + 1] : 0;
    const bool small_lv = PPM_LANES == 32 && do1 - do0 <= 32;
    const bool next_small = small_lv && lv + 1 < m.n_lev_dn && do2 - do1 <= 32;
    if (!small_lv || tid < 32)
    for (int k = do0 + tid; k < do1; k += nt) {
      int v, par_, sib;
      if (k == nk_) { v = nv_; par_ = np_; sib = ns_; }
      else { v = m.lev_dn_nodes[k]; par_ = m.parent[v]; sib = (m.left[par_] == v) ? m.right[par_] : m.left[par_]; }
      if (lv + 1 < m.n_lev_dn) {
        const int kn = do1 + tid;
        if (kn < do2) {
          nk_ = kn; nv_ = m.lev_dn_nodes[kn]; np_ = m.parent[nv_];
          ns_ = (m.left[np_] == nv_) ? m.right[np_] : m.left[np_];
        }
      }
      if (m.left[v] < 0) continue;  // a leaf has no children to serve: it reads its parent after the walk
      const double blv = m.bl[v], Tsib = cls[sib].w;
      const double sig = sg[v], sT = ta[v];
      const double U = (1 - sig) + sT * (par_ == fv ? fU : z1[par_]);
      const double logV = (par_ == fv ? fV : z0[par_]) + (-l1 * blv) + Tsib;
      z1[v] = U; z0[v] = logV;
      double Dv = 0., Nv = 0.;
      if (t.et1) {  // ss0 is dead after level (B)/(C): reused as D
        const double Dp = par_ == fv ? fD : ss0[par_], Np = par_ == fv ? fN : t.et1[(size_t)p * nn + par_];
        Dv = (1 - sig) + sT * Dp;
        Nv = (1. / l1 - sig * (blv + 1. / l1)) + sT * (Np + blv * Dp);
        ss0[v] = Dv;
        t.et1[(size_t)p * nn + v] = Nv;
      }
      fv = v; fU = U; fV = logV; fD = Dv; fN = Nv;
    }
    // (the lists hold the internal nodes of a level first: one internal node per level = thread 0 walks a chain)
    const bool chain = next_small && dc0 != 0;
    if (chain) {} else if (next_small) { if (tid < 32) ppm_syncwarp(); } else PPM_CTA_BARRIER();
    do0 = do1; do1 = do2; do2 = do3; dc0 = dc1;
  }
  // the leaves, one per thread, from their parents' finished values (same formulas)
  for (int v = 1 + tid; v < nn; v += nt) {
    if (m.left[v] >= 0) continue;
    const int par_ = m.parent[v];
    const int sib = (m.left[par_] == v) ? m.right[par_] : m.left[par_];
    const double blv = m.bl[v], Tsib = cls[sib].w;
    const double sig = sg[v], sT = ta[v];
    z1[v] = (1 - sig) + sT * z1[par_];
    z0[v] = z0[par_] + (-l1 * blv) + Tsib;
    if (t.et1) {
      const double Dp = ss0[par_], Np = t.et1[(size_t)p * nn + par_];
      ss0[v] = (1 - sig) + sT * Dp;
      t.et1[(size_t)p * nn + v] = (1. / l1 - sig * (blv + 1. / l1)) + sT * (Np + blv * Dp);
    }
  }
  PPM_CTA_BARRIER();
  for (int v = 1 + tid; v < nn; v += nt) {
    const double logV = z0[v], logU = log(z1[v]);
    d1u[v] = make_double2(lse2(logV, logU) - logV, logU);
  }
  if (t.et1) {
    for (int v = tid; v < nn; v += nt) t.et1[(size_t)p * nn + v] /= ss0[v];
  }
}

// One thread per (parameter vector, entry or padded slice): the exp tables of the Mobile integral.
// recs[e] = { K = pi1*exp(-(g2+l2)*(t_block - h_lineage)), ref, slice mask }; yy[j] = exp(-(g2+l2)*(t_j - t_block)),
// zero for the padding slices of the last block.
// Mobile merge of two child lineages given as (a, d) = (m0 + pi1 (m1-m0), m1-m0) ... the children's stored form is
// (x, y) with m0 = x - k0 y, m1 = x + k1 y along their branches (objects.cpp:427-435 in the linear domain);
// orc = {k0_l, k1_l, k0_r, k1_r}.  Returns the parent's (a, d), not yet renormalised.

PPM_HD double pack2i(int lo, int hi) {
  const long long b = ((long long)hi << 32) | (long long)(uint32_t)lo;
  double d; memcpy(&d, &b, 8);
  return d;
}

PPM_HD void prepass_tables_body(const ModelDev& m, const TablesDev& t, int p, int i) {
  const double* par = t.params + 8 * (size_t)p;
  const double g2 = par[4], l2 = par[5];
  const double lam = g2 + l2;
  const double pi1 = (g2 == 0.) ? 0. : g2 / lam;
  const int n_pad = m.n_blocks * m.R;
  if (i < m.n_entries) {
    int b = m.entry_block[i];
    double tb = m.slice_t[m.blk[b].slice0];
    double hs = m.height[m.entry_node[i]];
    EntryRec r;
    r.K = pi1 * exp(-lam * (tb - hs));
    r.ref = m.entry_ref[i];
    r.mask = m.entry_mask[i];
    if (i >= m.blk[b].e_pair0 && i < m.blk[b].e_leafu0) {
      // leaf-pair record: K is not needed (u-form); its 8 bytes carry the byte offsets the sweep would otherwise
      // compute per entry -- low word: the pair's observation word row in the item's pattern block, high word: the pair's
      // row in pairq -- and mask the rotation that lands the 2-bit observation pair on bits 5..6 (x 32 bytes)
      const int pid = r.ref;
      r.K = pack2i((pid >> 4) * (m.word_stride ? m.word_stride : PPM_LANES) * 4, pid * 4 * (int)sizeof(double4));
      r.mask = (uint32_t)((2 * (pid & 15) - 5) & 31);
    }
    t.recs[(size_t)p * (m.n_entries + 1) + i] = r;
  } else if (i == m.n_entries) {
    EntryRec r; r.K = 0.; r.ref = 0; r.mask = 0u;  // padding record: the prefetch of the last entry reads it
    t.recs[(size_t)p * (m.n_entries + 1) + i] = r;
  } else if (i <= m.n_entries + n_pad) {
    int j = i - m.n_entries - 1;
    double v = 0.;
    if (j < m.n_slices) {
      double tb = m.slice_t[(j / m.R) * m.R];
      v = exp(-lam * (m.slice_t[j] - tb));
    }
    t.yy[(size_t)p * n_pad + j] = v;
    // u_j = 1 - exp(-(g2+l2) t_j) = p01(t_j)/pi1 (functions.cpp:227-231): the variable of the u-form leaf terms
    t.uu[(size_t)p * n_pad + j] = (j < m.n_slices) ? 1. - exp(-lam * m.slice_t[j]) : 0.;
    if (m.lp) {  // leaf powers: see prepass_band_body
      double lb = 0., lc = 0., le = 0.;
      if (m.n_slices > 0) {
        // (padding slices of the last block -- weight 0 -- copy the last real slice: the block's exponents stay close together)
        const int jj = j < m.n_slices ? j : m.n_slices - 1;
        const double e2 = par[6], e1 = par[7];
        const double E = exp(-lam * m.slice_t[jj]), u = 1. - E;
        double lg[2], g[2];
        for (int x = 0; x < 2; x++) {
          const double m0 = x ? e1 : 1 - e1, m1 = x ? 1 - e2 : e2, d = m1 - m0;  // objects.cpp:407-408
          const double f = fma(pi1 * d, u, m0);
          lg[x] = f > 0. ? fmax(log2(f), -1e4) : -1e4;
          g[x] = f > 0. ? -pi1 * d * lam * E / (f * 0.693147180559945309417232) : 0.;
        }
        lb = (double)m.lp_alive[jj] * lg[0] + g[0] * m.lp_htot[jj]; lc = lg[1] - lg[0]; le = g[1] - g[0];
      }
      t.slb[(size_t)p * n_pad + j] = lb; t.slce[(size_t)p * n_pad + j] = make_double2(lc, le * m.lp_hunit);  // (the sweep counts heights in units of lp_hunit)
    }
  } else if (i <= m.n_entries + n_pad + m.n_ops) {
    // per-op record: the branch tables of the two children, contiguous (no dependent index loads in the sweep)
    int o = i - m.n_entries - n_pad - 1;
    const NodeOpDev op = m.ops[o];
    const double2 kl = t.kap[(size_t)p * m.n_nodes + op.lnode], kr = t.kap[(size_t)p * m.n_nodes + op.rnode];
    double* r = t.oprec + ((size_t)p * m.n_ops + o) * OPREC_DOUBLES;
    r[0] = kl.x; r[1] = kl.y; r[2] = kr.x; r[3] = kr.y;
    // K of the events folded into the op (children dying, the node being born) for the op's block, as in the entry records
    const double tb = m.op_tb[o];
    r[4] = pi1 * exp(-lam * (tb - m.height[op.lnode]));
    r[5] = pi1 * exp(-lam * (tb - m.height[op.rnode]));
    r[6] = pi1 * exp(-lam * (tb - m.height[op.node]));
    r[7] = 0.;
    if (m.lp) {  // leaf powers: a leaf child has no folded event; its slot carries the leaf's height (count updates at its death)
      if (m.left[op.lnode] < 0) r[4] = m.height[op.lnode];
      if (m.left[op.rnode] < 0) r[5] = m.height[op.rnode];
    }
  } else if (i <= m.n_entries + n_pad + m.n_ops + m.n_leaves + m.n_pairs) {
    // u-form leaf terms.  A leaf s at height h_s (numerically 0) observed as x contributes, at time t,
    //   (1-q) m0_x + q m1_x,  q = pi1 (1 - exp(-lam (t - h_s)))                       (functions.cpp:232-240)
    // = alpha + beta * u(t),  alpha = m0_x - pi1 d_x expm1(lam h_s),  beta = pi1 d_x exp(lam h_s),  u = 1 - exp(-lam t).
    // A pair of leaves is the product of two such terms: a quadratic in u for each of the 4 observation pairs.
    const double e2 = par[6], e1 = par[7];
    int k = i - m.n_entries - n_pad - m.n_ops - 1;
    auto leaf_term = [&](int pos, int x, double& alpha, double& beta) {
      const double hs = m.height[m.leaf_node[pos]];
      const double m0 = x ? e1 : 1 - e1, m1 = x ? 1 - e2 : e2, d = m1 - m0;  // objects.cpp:407-408
      alpha = m0 - pi1 * d * expm1(lam * hs);
      beta = pi1 * d * exp(lam * hs);
    };
    if (k < m.n_leaves) {
      for (int x = 0; x < 2; x++) {
        double al, be;
        leaf_term(k, x, al, be);
        t.leafab[((size_t)p * m.n_leaves + k) * 2 + x] = make_double2(al, be);
      }
    } else {
      const int pr = k - m.n_leaves;
      for (int c = 0; c < 4; c++) {
        double a1, b1, a2, b2;
        leaf_term(2 * pr, c & 1, a1, b1);  // pair pr owns pattern bits 2pr, 2pr+1
        leaf_term(2 * pr + 1, c >> 1, a2, b2);
        t.pairq[((size_t)p * m.n_pairs + pr) * 4 + c] = make_double4(a1 * a2, a1 * b2 + a2 * b1, b1 * b2, 0.);
      }
    }
  } else if (i <= m.n_entries + n_pad + m.n_ops + m.n_leaves + m.n_pairs + m.n_cherries) {
    // cherry node = merge of two leaf tables (same arithmetic as merge_children in the sweep), for the 4 observation pairs
    const int c = i - m.n_entries - n_pad - m.n_ops - m.n_leaves - m.n_pairs - 1;
    const NodeOpDev op = m.ops[m.cherry_op[c]];
    const double2 kl = t.kap[(size_t)p * m.n_nodes + op.lnode], kr = t.kap[(size_t)p * m.n_nodes + op.rnode];
    const double orc[4] = {kl.x, kl.y, kr.x, kr.y};
    const double2* leaftab = reinterpret_cast<const double2*>(t.scal + (size_t)p * SC_COUNT + SC_LA0);
    for (int x = 0; x < 4; x++) {
      double av, d;
      merge_children(leaftab[x & 1], leaftab[x >> 1], orc, pi1, av, d);
      t.cherry[((size_t)p * m.n_cherries + c) * 4 + x] = make_double2(av, d);
    }
  }
}

// Band kernel tables (ppm_kernels.cuh: BandTables), one thread per (parameter vector, i).  Times are taken relative to
// T0 = H/2 so that exp(+-lam (t - T0)) stays in range: a lineage record holds nbn = -d exp(lam (h - T0)) and a slice
// Y' = pi1 exp(-lam (t - T0)), hence  a + nbn Y' = m0 + pi1 d (1 - exp(-lam (t - h)))  (functions.cpp:227-240, 269-270).
// A child's merge factors (objects.cpp:427-435: m0 = a - k0 d, m1 = a + k1 d along its branch, k0 = pi1 E, k1 = (1-pi1) E)
// are pre-divided by the child's exp(lam (h - T0)) when the child is internal (its record carries nbn, not d).
PPM_HD void prepass_band_body(const ModelDev& m, const TablesDev& t, const BandDev& b, const BandTables& bt, int p, int i) {
  const double* par = t.params + 8 * (size_t)p;
  const double g2 = par[4], l2 = par[5], e2 = par[6], e1 = par[7];
  const double lam = g2 + l2;
  const double pi1 = (g2 == 0.) ? 0. : g2 / lam;
  const double T0 = 0.5 * m.H;
  if (i < b.n_ops) {
    const int v = b.op_nodes[3 * i], l = b.op_nodes[3 * i + 1], r = b.op_nodes[3 * i + 2];
    double* o = bt.bop + ((size_t)p * b.n_ops + i) * BOP_DOUBLES;
    auto child = [&](int c, double& c0, double& c1) {
      if (m.left[c] < 0) {  // leaf: applied to d
        const double E = exp(-lam * m.bl[c]);
        c0 = -pi1 * E; c1 = (1 - pi1) * E;
      } else {              // internal: applied to nbn = -d exp(lam (h_c - T0))
        const double Es = exp(-lam * (m.bl[c] + (m.height[c] - T0)));
        c0 = pi1 * Es; c1 = -(1 - pi1) * Es;
      }
    };
    child(l, o[0], o[1]);
    child(r, o[2], o[3]);
    o[4] = -exp(lam * (m.height[v] - T0));
    o[5] = b.lp ? b.op_leaf_h[2 * i] : 0.;
    o[6] = b.lp ? b.op_leaf_h[2 * i + 1] : 0.;
    o[7] = 0.;
  } else if (i < b.n_ops + b.S_pad) {
    const int j = i - b.n_ops;
    const double tj = b.slice_t[j];
    bt.bY[(size_t)p * b.S_pad + j] = pi1 * exp(-lam * (tj - T0));
    bt.bU[(size_t)p * b.S_pad + j] = 1. - exp(-lam * tj);
    if (b.lp) {
      // Leaf powers.  A leaf at height h observed x contributes f_x = m0_x + pi1 d_x (1 - exp(-lam (t - h))) (functions.cpp:232-240);
      // at h = 0 that is the same for every leaf, and to first order in h: log2 f_x(t; h) = log2 f_x(t; 0) + g_x(t) h,
      // g_x = -pi1 d_x lam exp(-lam t) / (f_x ln 2).  Product over the alive leaves (n1 present, heights summing to S1 over the
      // present ones): 2^(bLb + n1 bLc + S1 bLe).
      const double E = exp(-lam * tj), u = 1. - E;
      double lg[2], g[2];
      for (int x = 0; x < 2; x++) {
        const double m0 = x ? e1 : 1 - e1, m1 = x ? 1 - e2 : e2, d = m1 - m0;  // objects.cpp:407-408
        const double f = fma(pi1 * d, u, m0);
        lg[x] = f > 0. ? fmax(log2(f), -1e4) : -1e4;  // (an impossible observation: any such leaf sends the product to ~0)
        g[x] = f > 0. ? -pi1 * d * lam * E / (f * 0.693147180559945309417232) : 0.;
      }
      const bool real = j < m.n_slices;
      bt.bLb[(size_t)p * b.S_pad + j] = real ? (double)b.n_alive_leaves[j] * lg[0] + g[0] * b.htot[j] : 0.;
      bt.bLc[(size_t)p * b.S_pad + j] = real ? lg[1] - lg[0] : 0.;
      bt.bLe[(size_t)p * b.S_pad + j] = real ? g[1] - g[0] : 0.;
    }
  } else if (i < b.n_ops + b.S_pad + m.n_leaves) {
    const int pos = i - b.n_ops - b.S_pad;
    const double hs = m.height[m.leaf_node[pos]];
    const double Eh = exp(lam * (hs - T0));
    for (int x = 0; x < 2; x++) {
      const double m0 = x ? e1 : 1 - e1, m1 = x ? 1 - e2 : e2, d = m1 - m0;  // objects.cpp:407-408
      bt.bleaf[((size_t)p * m.n_leaves + pos) * 2 + x] = make_double2(m0 + pi1 * d, -d * Eh);
    }
  }
}

// ------------------------------------------------------------------------------------------------ zero row
// Mobile integral of the all-zero row (getPobs2, functions.cpp:300-333 = likRep2 on the row add0() appends).  The row's
// partials are per-parameter constants per node (zrow/zk, node prepass), so its slices are independent: one thread per
// slice walks the entry records and the folded events of the slice's block and multiplies exactly the factors the sweep
// would (same tables, same K*Y form), with its own range checks.  One CTA per parameter vector; the slice values are
// reduced in a fixed order with exponent alignment.
PPM_HD void zero_row_slice(const ModelDev& m, const TablesDev& t, int p, int j) {
  const int R = m.R, b = j / R, i = j - b * R;
  const double* sc = t.scal + (size_t)p * SC_COUNT;
  const double2* zr = t.zrow + (size_t)p * m.n_nodes;
  const int* zk = t.zk + (size_t)p * m.n_nodes;
  const EntryRec* recs = t.recs + (size_t)p * (m.n_entries + 1);
  const int nbR = m.n_blocks * R;
  const double Y = t.yy[(size_t)p * nbR + j], U = t.uu[(size_t)p * nbR + j];
  const bool factored = sc[SC_FACTORED] != 0.;
  const double la0 = sc[SC_LA0], ld0 = sc[SC_LD0], pi1 = sc[SC_PI1];
  (void)pi1;
  const BlockDesc bd = m.blk[b];
  double prod = 1.;
  int ex = t.zxb[(size_t)p * (m.n_blocks + 1) + b];
  int nfac = 0;
  auto mul = [&](double f) {
    prod *= f;
    if ((++nfac & 15) == 0) {
      const int be = biased_exp(prod);
      if (be != 0 && be < 1023 - 480) { prod *= pow2i(480); ex += 480; }
    }
  };
  const uint32_t bit = 1u << i;
  // node ops of the block: shifts of the nodes alive from this slice on, and the folded events
  for (int o = bd.op0; o < bd.op1; o++) {
    const SweepOpDev op = m.sops[o];
    const NodeOpDev nd = m.ops[o];
    const double* orc = t.oprec + ((size_t)p * m.n_ops + o) * OPREC_DOUBLES;
    if (op.dst >= 0 && i >= op.i0) ex += zk[nd.node];
    if (op.pair >= 0) {  // cherry: the two leaves as one quadratic in u (observation pair 0)
      if ((uint32_t)op.ml & bit) {
        if (!factored) {
          const double4 q = t.pairq[((size_t)p * m.n_pairs + op.pair) * 4];
          mul(fma(fma(q.z, U, q.y), U, q.x));
        } else {
          const double2 fa = t.leafab[((size_t)p * m.n_leaves + 2 * op.pair) * 2], fb = t.leafab[((size_t)p * m.n_leaves + 2 * op.pair + 1) * 2];
          mul(fma(fa.y, U, fa.x) * fma(fb.y, U, fb.x));
        }
      }
    } else {
      if ((uint32_t)op.ml & bit) {
        const bool lf = op.flags & 1;
        const double a = lf ? la0 : zr[nd.lnode].x, d = lf ? ld0 : zr[nd.lnode].y;
        mul(fma(d * -orc[4], Y, a));
      }
      if ((uint32_t)op.mr & bit) {
        const bool lf = op.flags & 2;
        const double a = lf ? la0 : zr[nd.rnode].x, d = lf ? ld0 : zr[nd.rnode].y;
        mul(fma(d * -orc[5], Y, a));
      }
    }
    if ((uint32_t)op.mself & bit) mul(fma(zr[nd.node].y * -orc[6], Y, zr[nd.node].x));
  }
  // internal lineages alive in the whole block (padding records of the deep program refer to the constant row: skip)
  for (int e = bd.e_slot0; e < bd.e_pair0; e++) {
    if ((recs[e].ref & 0xffff) == m.n_slots) continue;
    const double2 ad = zr[m.entry_node[e]];
    mul(fma(ad.y * -recs[e].K, Y, ad.x));
  }
  // leaf pairs / single leaves in u-form, observation 0
  for (int e = bd.e_pair0; e < bd.e_leafu0; e++) {
    const int pid = m.entry_ref[e] & 0xffff;
    if (!factored) {
      const double4 q = t.pairq[((size_t)p * m.n_pairs + pid) * 4];
      mul(fma(fma(q.z, U, q.y), U, q.x));
    } else {
      const double2 fa = t.leafab[((size_t)p * m.n_leaves + 2 * pid) * 2], fb = t.leafab[((size_t)p * m.n_leaves + 2 * pid + 1) * 2];
      mul(fma(fa.y, U, fa.x) * fma(fb.y, U, fb.x));
    }
  }
  for (int e = bd.e_leafu0; e < bd.e_leaf0; e++) {
    const double2 f = t.leafab[((size_t)p * m.n_leaves + (m.entry_ref[e] & 0xffff)) * 2];
    mul(fma(f.y, U, f.x));
  }
  // leaves in K/Y form (trees with tall leaves), whole block or part of it; internal lineages still listed as partial
  for (int e = bd.e_leaf0; e < bd.e_pleaf0; e++) mul(fma(ld0 * -recs[e].K, Y, la0));
  for (int e = bd.e_pleaf0; e < bd.e_pslot0; e++)
    if ((recs[e].mask >> 16) & bit) mul(fma(ld0 * -recs[e].K, Y, la0));
  for (int e = bd.e_pslot0; e < bd.e_end; e++)
    if ((recs[e].mask >> 16) & bit) { const double2 ad = zr[m.entry_node[e]]; mul(fma(ad.y * -recs[e].K, Y, ad.x)); }
  if (m.lp) {  // leaf powers: every alive leaf observed 0
    int kx;
    prod *= lp_pow2(t.slb[(size_t)p * nbR + j], kx);
    ex -= kx;
  }
  t.zpart[(size_t)p * m.n_slices + j] = make_double2(prod * m.slice_wt_pad[j], (double)ex);
}

// (1) shifts accumulated before each block (one CTA per parameter vector), (2) the slices (any number of CTAs: slices are
// independent), (3) the fixed-order sum (one thread per parameter vector).  tid/nt as in the node prepass.
PPM_HD void zero_row_prefix_body(const ModelDev& m, const TablesDev& t, int p, int tid, int nt) {
  int* zxb = t.zxb + (size_t)p * (m.n_blocks + 1);
  const int* zk = t.zk + (size_t)p * m.n_nodes;
  for (int b = tid; b <= m.n_blocks; b += nt) {  // shifts of the nodes formed in each block ...
    int acc = 0;
    const BlockDesc bd = m.blk[b];
    for (int o = bd.op0; o < bd.op1; o++)
      if (m.sops[o].dst >= 0) acc += zk[m.ops[o].node];
    zxb[b] = acc;
  }
  PPM_CTA_BARRIER();
  if (tid == 0) {  // ... turned into the sum over all earlier blocks
    int run = 0;
    for (int b = 0; b <= m.n_blocks; b++) { const int v = zxb[b]; zxb[b] = run; run += v; }
  }
}
PPM_HD void zero_row_sum_body(const ModelDev& m, const TablesDev& t, int p) {  // sum of mantissa * 2^-exponent in slice order
  double Im = 0.; int Ie = 0x3fffffff;
  const double2* zp = t.zpart + (size_t)p * m.n_slices;
  for (int j = 0; j < m.n_slices; j++) scaled_add(Im, Ie, zp[j].x, (int)zp[j].y);
  double* scw = t.scal + (size_t)p * SC_COUNT;
  scw[SC_IZ_M] = Im; scw[SC_IZ_E] = (double)Ie;
}
PPM_HD void zero_row_body(const ModelDev& m, const TablesDev& t, int p, int tid, int nt) {  // all three on one CTA (host emulation)
  zero_row_prefix_body(m, t, p, tid, nt);
  PPM_CTA_BARRIER();
  for (int j = tid; j < m.n_slices; j += nt) zero_row_slice(m, t, p, j);
  PPM_CTA_BARRIER();
  if (tid == 0) zero_row_sum_body(m, t, p);
}

// ------------------------------------------------------------------------------------------------ class sums
// log p1(root) of the two pure-loss categories (Persistent l0, Private l1): objects.cpp:404-426 collapsed.
// Per branch the factor is S (a present leaf below), T (all-zero subtree under an active parent) or nothing; hence
//   log p1(root) = Stot + sum over the pattern's FRONTIER nodes c (roots of maximal all-zero subtrees) of F_c,
// with Stot and F_c = T_c - sum_{subtree c} S per parameter vector (prepass) and the frontier per pattern only.
// Step 1 (pattern only, warp-uniform): one pass over the node ops marks all-zero subtrees and lists the frontier.
// zbits: one bit per node; frontier: up to n_nodes node ids.  Returns the frontier length.
PPM_HD int class_frontier(const ModelDev& m, const uint32_t* pw, size_t pw_stride, uint32_t* zbits, uint16_t* frontier) {
  int nf = 0;
  for (int o = 0; o < m.n_ops; o++) {
    const NodeOpDev op = m.ops[o];
    bool zl, zr;
    if (op.lref < 0) { int pos = -(op.lref + 1); zl = !((pw[(size_t)(pos >> 5) * pw_stride] >> (pos & 31)) & 1u); }
    else zl = (zbits[op.lnode >> 5] >> (op.lnode & 31)) & 1u;
    if (op.rref < 0) { int pos = -(op.rref + 1); zr = !((pw[(size_t)(pos >> 5) * pw_stride] >> (pos & 31)) & 1u); }
    else zr = (zbits[op.rnode >> 5] >> (op.rnode & 31)) & 1u;
    const bool zv = zl && zr;
    const uint32_t bit = 1u << (op.node & 31);
    uint32_t w = zbits[op.node >> 5];
    w = zv ? (w | bit) : (w & ~bit);
    ppm_syncwarp();
    zbits[op.node >> 5] = w;   // (all lanes write the same word)
    const bool parent_active = !zv || op.node == 0;  // the root always counts as active
    if (parent_active && zl) frontier[nf++] = (uint16_t)op.lnode;
    if (parent_active && zr) frontier[nf++] = (uint16_t)op.rnode;
    ppm_syncwarp();
  }
  return nf;
}
// Step 2 (one parameter vector per lane): Stot + sum of F over the frontier.
PPM_HD double2 class_sums(const TablesDev& t, int p, const uint16_t* frontier, int nf) {
  const double* sc = t.scal + (size_t)p * SC_COUNT;
  double a0 = sc[SC_STOT0], a1 = sc[SC_STOT1];
  for (int k = 0; k < nf; k++) {
    const double2 f = t.clsT[(size_t)frontier[k] * t.P + p];
    a0 += f.x; a1 += f.y;
  }
  return make_double2(a0, a1);
}

// i2 and the three log prefactors (functions.cpp:342-343, 348, 354, 356)
PPM_HD void prepass_i2_body(const ModelDev& m, const TablesDev& t, const double* i2_override, int p) {
  const double* par = t.params + 8 * (size_t)p;
  double* sc = t.scal + (size_t)p * SC_COUNT;
  const double N0 = par[0], i1 = par[2], l1 = par[3];
  double IZ = sc[SC_IZ_M] == 0. ? 0. : ldexp(sc[SC_IZ_M], -(int)sc[SC_IZ_E]);
  double B = m.H - IZ;
  double i2 = ((double)m.n_obs - N0 * sc[SC_P0] - i1 / l1 * sc[SC_P1] - i1 * sc[SC_A]) / B;
  sc[SC_B] = B;
  if (i2_override) i2 = i2_override[p];
  sc[SC_I2] = i2;
  sc[SC_C0] = log(N0 / (double)m.n_obs);
  sc[SC_C1] = log(i1 / (l1 * (double)m.n_obs));
  sc[SC_C2] = log(i2 / (double)m.n_obs);
}

// ------------------------------------------------------------------------------------------------ sweep
// Per work item (32 patterns): its pattern words [n_words][lanes], the all-zero-subtree masks per slot, and the
// Mobile partials (a,d) of the alive internal lineages [slot][lanes] (shared memory in the staged variant).
// Lineage partials live either in ordinary memory (shared in the classic staged variant, global in the deep
// variant, host memory in the emulator) ...
struct SlotsMem {
  double2* base;  // [slot][lanes]
  PPM_HD double2 ld(int slot, int lane) const { return base[slot * PPM_LANES + lane]; }
  PPM_HD void st(int slot, int lane, double2 v) const { base[slot * PPM_LANES + lane] = v; }
  PPM_HD void ld_async(int slot, int lane, double2& dst) const { dst = base[slot * PPM_LANES + lane]; }
  PPM_HD void wait_async(double2&) const {}
};
#if defined(__CUDACC__)
// ... or in Blackwell tensor memory: TMEM is 128 lanes x 512 columns x 32 bit per SM and a warp reaches the 32 lanes of
// its quarter (warp % 4), i.e. it is per-lane private scratch -- exactly the shape of slots[slot][lane].  A slot is 4
// columns (16 bytes per lane), moved with tcgen05.st / tcgen05.ld .32x32b.x4 (SASS STTM / LDTM).  The first `nt` slots of
// a warp live in its column range, the rest (if any) in a shared-memory overflow.  This frees ~17 KB of shared memory
// per warp, which is what bounds the resident warps of the classic variant.
struct SlotsTmem {
  uint32_t taddr;  // lane base << 16 | first column of this warp
  int nt;          // slots held in TMEM
  double2* ovf;    // [slot - nt][32]
  __device__ __forceinline__ void ld_async(int slot, int lane, double2& dst) const {
    if (slot < nt) {
      uint32_t r0, r1, r2, r3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr + 4u * (uint32_t)slot));
      dst.x = __hiloint2double((int)r1, (int)r0);
      dst.y = __hiloint2double((int)r3, (int)r2);
    } else dst = ovf[(slot - nt) * 32 + lane];
  }
  // the loaded registers may only be consumed after the wait: tie them to it
  __device__ __forceinline__ void wait_async(double2& v) const {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+d"(v.x), "+d"(v.y));
  }
  __device__ __forceinline__ double2 ld(int slot, int lane) const {
    double2 v;
    ld_async(slot, lane, v);
    wait_async(v);
    return v;
  }
  __device__ __forceinline__ void st(int slot, int lane, double2 v) const {
    if (slot < nt) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                   :: "r"(taddr + 4u * (uint32_t)slot), "r"((uint32_t)__double2loint(v.x)), "r"((uint32_t)__double2hiint(v.x)),
                      "r"((uint32_t)__double2loint(v.y)), "r"((uint32_t)__double2hiint(v.y)));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else ovf[(slot - nt) * 32 + lane] = v;
  }
};
#endif

template <class Slots>
struct ItemMem {
  const uint32_t* sw;  // the item's patterns in program leaf order (pair p = bits 2p, 2p+1): word w of lane l at sw[w * sws + l];
  size_t sws;          // a shared-memory copy [n_words][lanes] (sws = lanes) or, for big trees, the device rows themselves
  uint32_t* sw_copy;   // the shared-memory copy to fill (null when the rows are read in place)
  Slots slots;
};

// Where one work item reads its topology program and per-parameter tables from (shared memory copies in the
// staged variant, global memory otherwise).
struct ItemTables {
  const EntryRec* recs;   // [n_entries + 1]
  const BlockDesc* blk;   // [n_blocks + 1]
  const SweepOpDev* ops;   // [n_ops]
  const double* oprec;    // [n_ops][OPREC_DOUBLES]
  const double* yy;       // [n_blocks * R]
  const double* wt;       // [n_blocks * R]
  const double2* leaftab; // [2] leaf partials (a,d) for observation 0 / 1
  const double* uu;       // [n_blocks * R]
  const double2* leafab;  // [n_leaves][2] u-form leaf terms {alpha, beta}
  const double4* pairq;   // [n_pairs][4] u-form pair quadratics
  const double2* cherry;  // [n_cherries][4] cherry node partials by observation pair
  const double* lb; const double2* lce;  // [n_blocks * R] leaf powers (ModelDev::lp): base | (count coefficient, height coefficient)
};

// the two 32-bit halves of a record's K field (leaf-pair records)
PPM_HD void unpack2i(double v, int& lo, int& hi) {
#if defined(__CUDA_ARCH__)
  lo = __double2loint(v); hi = __double2hiint(v);
#else
  long long b; memcpy(&b, &v, 8); lo = (int)(uint32_t)b; hi = (int)(b >> 32);
#endif
}
PPM_HD uint32_t rotr32(uint32_t v, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(v, v, s);
#else
  s &= 31u; return s ? (v >> s) | (v << (32u - s)) : v;
#endif
}

// One 16-byte load per record: K | ref | mask
PPM_HD void load_rec(const EntryRec* recs, int e, double& K, uint32_t& ref, uint32_t& mask) {
#if defined(__CUDA_ARCH__)
  const double2 v = reinterpret_cast<const double2*>(recs)[e];
  K = v.x; ref = (uint32_t)__double2loint(v.y); mask = (uint32_t)__double2hiint(v.y);
#else
  K = recs[e].K; ref = (uint32_t)recs[e].ref; mask = recs[e].mask;
#endif
}

template <int R>
PPM_HD void rescale_products(double (&prod)[R], int (&ex)[R]) {
  // keep the running products away from the denormal range (rare path)
#pragma unroll
  for (int i = 0; i < R; i++) {
    int be = biased_exp(prod[i]);
    if (be != 0 && be < 1023 - 480) { prod[i] *= pow2i(480); ex[i] += 480; }
  }
}
template <int R>
PPM_HD bool products_small(const double (&prod)[R]) {
  // all products are >= 0, so their high words order like the values; an exact zero also lands here (harmless)
  int mn = ppm_double2hiint(prod[0]);
#pragma unroll
  for (int i = 1; i < R; i++) { int h = ppm_double2hiint(prod[i]); mn = h < mn ? h : mn; }
  return mn < ((1023 - 480) << 20);
}

// prod[i] *= a + nbb*Y[i] for a warp-uniform set of slices.  Lineages born inside a block live on a suffix
// [i0,R) of its slices, lineages merged inside it on a prefix [0,i1): the record carries a small code and a
// jump table lands on straight-line code for exactly the live slices (no selects, no wasted FP64); anything
// else (born AND merged inside one block) falls back to a per-slice select on the mask.
template <int I0, int I1, int R>
PPM_HD void update_range(double (&prod)[R], const double (&Y)[R], double a, double nbb) {
#pragma unroll
  for (int i = 0; i < R; i++)
    if (i >= I0 && i < I1) prod[i] *= fma(nbb, Y[i], a);
}
template <int R, int NI>
PPM_HD void coded_update(double (&prod)[NI][R], const double (&Y)[R], const double (&a)[NI], const double (&nbb)[NI], uint32_t mk) {
#if defined(PPM_PARTIAL_JUMPTABLE)  // measured 5 % slower on B200 than the straight-line select below (branch latency)
  switch (mk & 0xffffu) {
#else
  switch (0u) {
#endif
#define PPM_SUF(k) case k: _Pragma("unroll") for (int q = 0; q < NI; q++) update_range<k, R, R>(prod[q], Y, a[q], nbb[q]); break;
#define PPM_PRE(k) case 16 + k: _Pragma("unroll") for (int q = 0; q < NI; q++) update_range<0, k, R>(prod[q], Y, a[q], nbb[q]); break;
    PPM_SUF(1) PPM_SUF(2) PPM_SUF(3) PPM_SUF(4) PPM_SUF(5) PPM_SUF(6) PPM_SUF(7) PPM_SUF(8)
    PPM_SUF(9) PPM_SUF(10) PPM_SUF(11) PPM_SUF(12) PPM_SUF(13) PPM_SUF(14) PPM_SUF(15)
    PPM_PRE(1) PPM_PRE(2) PPM_PRE(3) PPM_PRE(4) PPM_PRE(5) PPM_PRE(6) PPM_PRE(7) PPM_PRE(8)
    PPM_PRE(9) PPM_PRE(10) PPM_PRE(11) PPM_PRE(12) PPM_PRE(13) PPM_PRE(14) PPM_PRE(15)
#undef PPM_SUF
#undef PPM_PRE
    default: {
      const uint32_t live = mk >> 16;
#pragma unroll
      for (int q = 0; q < NI; q++)
#pragma unroll
        for (int i = 0; i < R; i++) {
          const double t = prod[q][i] * fma(nbb[q], Y[i], a[q]);
          prod[q][i] = (live & (1u << i)) ? t : prod[q][i];
        }
    }
  }
}

// prod[i] *= a + nbb*Y[i] on the slices of a warp-uniform mask (events folded into node ops)
template <int R, int NI>
PPM_HD void live_update(double (&prod)[NI][R], const double (&Y)[R], const double (&a)[NI], const double (&nbb)[NI], uint32_t live) {
#pragma unroll
  for (int q = 0; q < NI; q++)
#pragma unroll
    for (int i = 0; i < R; i++) {
      const double t = prod[q][i] * fma(nbb[q], Y[i], a[q]);
      prod[q][i] = (live & (1u << i)) ? t : prod[q][i];
    }
}


// NI work items at once = NI x PPM_LANES patterns (one per lane and item) x one parameter vector.
// Everything topology-driven (records, block descriptors, node-op descriptors, branches, address arithmetic) is
// warp-uniform and therefore shared by the NI items; only the per-lane loads and the FP64 work are per item.
// That halves (NI = 2) the non-FP64 instructions per pattern and gives every warp NI independent dependency
// chains through the latency-bound parts (node ops, born/merged-in-block lineages).
// Returns, per item, this lane's count*log-mixture in contrib[].
// DEEP (the variant whose tables and lineage slots live in global memory): the records of the fully-alive entries
// are streamed through a per-warp shared-memory ring (wrec, 3 chunks of 32 records, cp.async two chunks ahead) and
// lineage partials are fetched four entries ahead, so L2/HBM latency is covered by ~140 cycles of FP64 work each.
// HOT: the instantiation for the objective itself (mode 0) on programs without K/Y-form leaves and partial entry records
// (ultrametric trees with folded events) and parameter vectors that do not need the factored leaf pairs (pi1 <= 0.95): the
// generic loops, the factored paths and the per-pattern outputs are compiled out (12 KB less code, 14 fewer registers:
// +3 % on configs[1]).  Factored parameter vectors of the same launch go through the generic instantiation.
template <int R, int NI, bool DEEP, class Slots, bool DIST = false, bool WG = false, bool HOT = false>
PPM_HD void sweep_items(const SweepArgs& a, const ItemTables& T, int p, long long batch0, int lane,
                        const ItemMem<Slots> (&mem)[NI], double (&contrib)[NI], EntryRec* wrec) {
  const ModelDev& m = a.m;
  const double* sc = a.t.scal + (size_t)p * SC_COUNT;
  const bool skip = (a.mode == 0) && !a.force && (sc[SC_I2] < 0.);  // penalty branch: no pattern work (functions.cpp:362-364)
  long long pat[NI];
  bool active[NI];
#pragma unroll
  for (int q = 0; q < NI; q++) {
    pat[q] = a.first + (batch0 + q) * PPM_LANES + lane;
    active[q] = (batch0 + q) < a.n_batches && pat[q] < a.first + a.count;
    contrib[q] = 0.;
  }
  if (skip) return;

#pragma unroll
  for (int q = 0; q < NI; q++)
    if (mem[q].sw_copy)
      for (int w = 0; w < m.n_words; w++)
        mem[q].sw_copy[w * PPM_LANES + lane] = !active[q] ? 0u : m.words[(size_t)w * m.n_pat_pad + pat[q]];
  ppm_syncwarp();
  const double pi1 = sc[SC_PI1];
  const double la0 = sc[SC_LA0], ld0 = sc[SC_LD0], la1 = sc[SC_LA1], ld1 = sc[SC_LD1];
  if (DEEP) {
#pragma unroll
    for (int q = 0; q < NI; q++) mem[q].slots.st(m.n_slots, lane, make_double2(1., 0.));
  }

  int X[NI];                // sum of renormalisation shifts of all lineages formed so far
  double Im[NI]; int Ie[NI];  // Mobile integral = Im * 2^-Ie (exponent of an empty sum: +inf)
  double prod[NI][R], Y[R], U[R];
  int ex[NI][R];
  const bool factored = HOT ? false : (sc[SC_FACTORED] != 0.);
#pragma unroll
  for (int q = 0; q < NI; q++) { X[q] = 0; Im[q] = 0.; Ie[q] = 0x3fffffff; }
  // Leaf powers (Program::lp; HOT kernels serve such programs only): the product of the alive leaves of a slice is one exp2 of
  // count tables x (present alive leaves, sum of their heights).
  const bool lp = HOT ? true : (m.lp != 0);
  const bool hs1 = lp && m.has_s1 != 0;
  // n1 / s1q: the lane's present leaves still alive at the start of the block and the sum of their heights (integer units of
  // lp_hunit); nv / sv: the same per slice of the block, lowered as leaves die at the node ops -- integer work only, one exp2 per
  // slice at the block's end
  int n1[NI], s1q[NI], nv[NI][R], sv[NI][R];
#pragma unroll
  for (int q = 0; q < NI; q++) {
    n1[q] = (lp && active[q]) ? m.freq[pat[q]] : 0;
    s1q[q] = (hs1 && active[q]) ? m.lp_s1q_0[pat[q]] : 0;
#pragma unroll
    for (int i = 0; i < R; i++) { nv[q][i] = 0; sv[q][i] = 0; }
  }
  // a node op whose leaf children (present: bl / br) die at slice i0 of the block
  auto leaf_deaths = [&](int q, int i0, uint32_t bl, uint32_t br, int hlq, int hrq) {
    const int cnt = (int)(bl + br);
    const int dq = (bl ? hlq : 0) + (br ? hrq : 0);
    n1[q] -= cnt; s1q[q] -= dq;
#pragma unroll
    for (int i = 0; i < R; i++) {
      nv[q][i] -= i >= i0 ? cnt : 0;
      sv[q][i] -= i >= i0 ? dq : 0;
    }
  };

  for (int b = 0; b <= m.n_blocks; b++) {
    const BlockDesc bd = T.blk[b];
    const bool real = b < m.n_blocks;
    if (real) {
#pragma unroll
      for (int i = 0; i < R; i++) { Y[i] = T.yy[b * R + i]; U[i] = HOT ? 0. : T.uu[b * R + i]; }
#pragma unroll
      for (int q = 0; q < NI; q++)
#pragma unroll
        for (int i = 0; i < R; i++) { prod[q][i] = DIST ? 1. : T.wt[b * R + i]; ex[q][i] = X[q]; }  // the slice weight goes in first
        // (w/nSub of functions.cpp:273-275; padding slices carry weight 0); the DIST output wants the bare integrand
      if (lp) {
#pragma unroll
        for (int q = 0; q < NI; q++)
#pragma unroll
          for (int i = 0; i < R; i++) { nv[q][i] = n1[q]; sv[q][i] = s1q[q]; }
      }
    }
    // ---- node ops: form every node born before the end of this block.
    // store a new lineage (a, d); when it has become small its mantissa is brought back to [0.5,1) and the shift
    // goes to X / ex[] (one integer test per lane on the high words; the exact shift only on the rare path)
    auto store_node = [&](int q, int dst, int i0, double& av, double& d) {
      const int ha = ppm_double2hiint(av) & 0x7fffffff, hd = ppm_double2hiint(d) & 0x7fffffff;
      const int h = ha > hd ? ha : hd;
      const bool small = (uint32_t)(h - 0x00100000) < (uint32_t)(((1023 - 40) << 20) - 0x00100000);  // 0 < exponent < 983
      if (ppm_any(small)) {
        const int k = small ? 1022 - (h >> 20) : 0;
        const double sfac = pow2i(k);
        av *= sfac; d *= sfac;
        X[q] += k;
        if (real) {
#pragma unroll
          for (int i = 0; i < R; i++) if (i >= i0) ex[q][i] += k;
        }
      }
      mem[q].slots.st(dst, lane, make_double2(av, d));  // every lane reads back only its own slot column
    };
    // cherries: both children are leaves, so the node is one of four per-parameter values picked by the 2-bit
    // observation pair (bits 2p, 2p+1 of the pattern; the record carries word and rotation)
    SweepOpDev opn = T.ops[bd.op0];  // records are fetched one op ahead (the table ends with a padding record)
    for (int o = bd.op0; o < bd.opc; o++) {
      const SweepOpDev op = opn;
      opn = T.ops[o + 1];
      double av[NI], dd[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) {
        const uint32_t w = mem[q].sw[(WG ? (size_t)((op.lbit >> 5)) * mem[q].sws : (size_t)(((op.lbit >> 5)) * PPM_LANES)) + lane];
        const uint32_t sel = rotr32(w, (uint32_t)op.lbit) & 0x30u;  // observation pair x 16
        const double2 v = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(T.cherry + op.pair * 4) + sel);
        av[q] = v.x; dd[q] = v.y;
        if (lp) leaf_deaths(q, op.i0, (sel >> 4) & 1u, sel >> 5, op.hlq, op.hrq);
        if (!HOT && op.ml) {  // the two leaves die here: their slices of this block as one quadratic in u (as in the pair loop)
          const double4 qv = *reinterpret_cast<const double4*>(reinterpret_cast<const char*>(T.pairq + op.pair * 4) + 2 * sel);
          if (!factored) {
#pragma unroll
            for (int i = 0; i < R; i++) {
              const double t = prod[q][i] * fma(fma(qv.z, U[i], qv.y), U[i], qv.x);
              prod[q][i] = ((uint32_t)op.ml & (1u << i)) ? t : prod[q][i];
            }
          } else {
            const double2 fa = T.leafab[(2 * op.pair) * 2 + (int)((sel >> 4) & 1u)], fb = T.leafab[(2 * op.pair + 1) * 2 + (int)(sel >> 5)];
#pragma unroll
            for (int i = 0; i < R; i++) {
              const double t = prod[q][i] * (fma(fa.y, U[i], fa.x) * fma(fb.y, U[i], fb.x));
              prod[q][i] = ((uint32_t)op.ml & (1u << i)) ? t : prod[q][i];
            }
          }
        }
        store_node(q, op.dst, op.i0, av[q], dd[q]);
      }
      if (op.mself) {
        const double nK = -(T.oprec + (size_t)o * OPREC_DOUBLES)[6];
        double nbb[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) nbb[q] = dd[q] * nK;
        live_update<R, NI>(prod, Y, av, nbb, (uint32_t)op.mself);
      }
    }
    // general merges: a leaf child comes from the two-entry leaf table (picked by its pattern bit), an internal one
    // from its slot; both are fetched, the record's flags keep one (no branches in this latency-bound chain)
#pragma unroll 2
    for (int o = bd.opc; o < bd.op1; o++) {
      const SweepOpDev op = opn;
      opn = T.ops[o + 1];
      const double* orc = T.oprec + (size_t)o * OPREC_DOUBLES;
      const bool lleaf = op.flags & 1, rleaf = op.flags & 2;
      double2 L[NI], Rr[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) {
        double2 ls, rs;
        mem[q].slots.ld_async(op.lslot, lane, ls);
        mem[q].slots.ld_async(op.rslot, lane, rs);
        const uint32_t wl = mem[q].sw[(WG ? (size_t)((op.lbit >> 5)) * mem[q].sws : (size_t)(((op.lbit >> 5)) * PPM_LANES)) + lane], wr = mem[q].sw[(WG ? (size_t)((op.rbit >> 5)) * mem[q].sws : (size_t)(((op.rbit >> 5)) * PPM_LANES)) + lane];
        const double2 lt = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(T.leaftab) + (rotr32(wl, (uint32_t)op.lbit) & 0x10u));
        const double2 rt = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(T.leaftab) + (rotr32(wr, (uint32_t)op.rbit) & 0x10u));
        mem[q].slots.wait_async(ls);
        mem[q].slots.wait_async(rs);
        L[q].x = lleaf ? lt.x : ls.x;  L[q].y = lleaf ? lt.y : ls.y;
        Rr[q].x = rleaf ? rt.x : rs.x; Rr[q].y = rleaf ? rt.y : rs.y;
        if (lp && (op.flags & 3))
          leaf_deaths(q, op.i0, lleaf ? (rotr32(wl, (uint32_t)op.lbit) >> 4) & 1u : 0u, rleaf ? (rotr32(wr, (uint32_t)op.rbit) >> 4) & 1u : 0u, op.hlq, op.hrq);
      }
      // a child that dies with this op multiplies its slices of the block here (it has no entry record) -- unless its
      // prefix and the new node's suffix partition the block (op.comb): then one update serves both, below.  (Global-memory
      // variant only: +4 % at 1,000 genomes; in the staged kernel the extra code costs more than the 16 FP64 it saves.)
      if (op.ml && !(DEEP && op.comb == 1)) {
        double aa[NI], nbb[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) { aa[q] = L[q].x; nbb[q] = L[q].y * -orc[4]; }
        live_update<R, NI>(prod, Y, aa, nbb, (uint32_t)op.ml);
      }
      if (op.mr && !(DEEP && op.comb == 2)) {
        double aa[NI], nbb[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) { aa[q] = Rr[q].x; nbb[q] = Rr[q].y * -orc[5]; }
        live_update<R, NI>(prod, Y, aa, nbb, (uint32_t)op.mr);
      }
      if (op.dst >= 0) {
        double av[NI], dd[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) {
          merge_children(L[q], Rr[q], orc, pi1, av[q], dd[q]);
          store_node(q, op.dst, op.i0, av[q], dd[q]);
        }
        // per slice: the dying child's factor before the birth, the new node's factor from it on
        auto both = [&](const double2 (&C)[NI], double Kc, uint32_t cm) {
          const double nKc = -Kc, nKs = -orc[6];
#pragma unroll
          for (int q = 0; q < NI; q++) {
            const double cn = C[q].y * nKc, sn = dd[q] * nKs;
#pragma unroll
            for (int i = 0; i < R; i++) {
              const bool c = cm & (1u << i);
              prod[q][i] *= fma(c ? cn : sn, Y[i], c ? C[q].x : av[q]);
            }
          }
        };
        if (DEEP && op.comb == 1) both(L, orc[4], (uint32_t)op.ml);
        else if (DEEP && op.comb == 2) both(Rr, orc[5], (uint32_t)op.mr);
        else if (op.mself) {  // the new node alone, when it is born inside the block
          double nbb[NI];
#pragma unroll
          for (int q = 0; q < NI; q++) nbb[q] = dd[q] * -orc[6];
          live_update<R, NI>(prod, Y, av, nbb, (uint32_t)op.mself);
        }
      }
    }
    if (!real) break;
#pragma unroll
    for (int q = 0; q < NI; q++)
      if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);

    // ---- entries alive in every slice of the block: first the internal lineages (partials in slots), then the
    // leaves (partials selected by the pattern bit).
    if (DEEP) {
      // the constant row (1,0) behind the last slot makes padding records multiply by one
      const int E0 = bd.e_slot0, E1 = bd.e_pair0, ES = bd.e_pair0;  // (leaves: u-form loops below; K/Y leaves: generic loop)
      const int nch = (E1 - E0 + 31) >> 5;
      auto stage = [&](int c) {
        int idx = E0 + c * 32 + lane;
        idx = idx > m.n_entries ? m.n_entries : idx;  // the table ends with one padding record
        ppm_cp_async16(wrec + (c % 3) * 32 + lane, T.recs + idx);
        ppm_cp_commit();
      };
      if (PPM_LANES == 32) { stage(0); stage(1); }
      double2 ring[NI][4];
      if (nch > 0 && ES > E0) {  // preload the first four lineage partials (ES - E0 is a multiple of 4)
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int sl = T.recs[E0 + u].ref & 0xffff;
#pragma unroll
          for (int q = 0; q < NI; q++) ring[q][u] = mem[q].slots.ld(sl, lane);
        }
      }
      for (int c = 0; c < nch; c++) {
        const EntryRec* cur; const EntryRec* nxt;
        if (PPM_LANES == 32) {
          stage(c + 2);
          ppm_cp_wait1();
          ppm_syncwarp();
          cur = wrec + (c % 3) * 32; nxt = wrec + ((c + 1) % 3) * 32;
        } else {  // host emulation reads the table directly
          cur = T.recs + E0 + c * 32; nxt = cur + 32;
        }
        const int base = E0 + c * 32;
        const int n = (E1 - base < 32) ? E1 - base : 32;
        int ns = ES - base; ns = ns < 0 ? 0 : (ns > n ? n : ns);  // lineage entries of this chunk: [0, ns), a multiple of 4
        for (int j0 = 0; j0 < ns; j0 += 4) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int j = j0 + u;
            double K; uint32_t ref, mk;
            load_rec(cur, j, K, ref, mk);
            // slot of the entry four ahead (next chunk if needed); past the segment: the constant row
            const int ja = j + 4;
            const int gidx = base + ja;
            int sl = m.n_slots;
            if (gidx < ES) sl = (ja < 32) ? (cur[ja].ref & 0xffff) : (nxt[ja - 32].ref & 0xffff);
            const double nK = -K;
#pragma unroll
            for (int q = 0; q < NI; q++) {
              const double2 ad = ring[q][u];
              ring[q][u] = mem[q].slots.ld(sl, lane);
              const double nbb = ad.y * nK;
#pragma unroll
              for (int i = 0; i < R; i++) prod[q][i] *= fma(nbb, Y[i], ad.x);
            }
          }
          if ((j0 & 15) == 12) {
#pragma unroll
            for (int q = 0; q < NI; q++)
              if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
          }
        }
        if (ns < n) {  // leaf entries of this chunk
          double K; uint32_t ref, mk;
          load_rec(cur, ns, K, ref, mk);
          for (int j = ns; j < n; j++) {
            bool bit[NI];
#pragma unroll
            for (int q = 0; q < NI; q++) bit[q] = (mem[q].sw[(WG ? (size_t)((ref & 0xffffu)) * mem[q].sws : (size_t)(((ref & 0xffffu)) * PPM_LANES)) + lane] & mk) != 0u;
            const double nK = -K;
            if (j + 1 < 32) load_rec(cur, j + 1, K, ref, mk); else load_rec(nxt, 0, K, ref, mk);
#pragma unroll
            for (int q = 0; q < NI; q++) {
              const double aa = bit[q] ? la1 : la0;
              const double nbb = (bit[q] ? ld1 : ld0) * nK;
#pragma unroll
              for (int i = 0; i < R; i++) prod[q][i] *= fma(nbb, Y[i], aa);
            }
            if ((j & 15) == 15) {
#pragma unroll
              for (int q = 0; q < NI; q++)
                if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
            }
          }
        }
        ppm_syncwarp();  // the ring slot (c % 3) is refilled two iterations from now
      }
      if (PPM_LANES == 32) { ppm_cp_wait0(); ppm_syncwarp(); }
    } else {
    // Each record names the NEXT record's operand in its upper bits, so the next operand load is issued before the
    // current products (software prefetch); chunks of 16 entries end with a range check.
    // (1) internal lineages alive in every slice: partials in slots; ref = slot | next slot << 16
    for (int e = bd.e_slot0; e < bd.e_pair0;) {
      const int ce = (e + 16 < bd.e_pair0) ? e + 16 : bd.e_pair0;
      double K; uint32_t ref, mk;
      load_rec(T.recs, e, K, ref, mk);
      double2 ad[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) ad[q] = mem[q].slots.ld((int)(ref & 0xffffu), lane);
#pragma unroll 4
      for (; e < ce; e++) {
        double2 an[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) mem[q].slots.ld_async((int)(ref >> 16), lane, an[q]);
        const double nK = -K;
        load_rec(T.recs, e + 1, K, ref, mk);
#pragma unroll
        for (int q = 0; q < NI; q++) {
          const double nbb = ad[q].y * nK;
#pragma unroll
          for (int i = 0; i < R; i++) prod[q][i] *= fma(nbb, Y[i], ad[q].x);
        }
#pragma unroll
        for (int q = 0; q < NI; q++) { mem[q].slots.wait_async(an[q]); ad[q] = an[q]; }
      }
#pragma unroll
      for (int q = 0; q < NI; q++)
        if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
    }
    }
    // (2) single leaves alive in every slice in block-relative K/Y form (only trees whose leaves are not at height 0):
    // ref = word | next word << 16, mk = the leaf's bit in that word
    for (int e = bd.e_leaf0; e < bd.e_pleaf0 && !HOT;) {
      const int ce = (e + 16 < bd.e_pleaf0) ? e + 16 : bd.e_pleaf0;
      double K; uint32_t ref, mk;
      load_rec(T.recs, e, K, ref, mk);
      uint32_t wv[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) wv[q] = mem[q].sw[(WG ? (size_t)((ref & 0xffffu)) * mem[q].sws : (size_t)(((ref & 0xffffu)) * PPM_LANES)) + lane];
#pragma unroll 2
      for (; e < ce; e++) {
        uint32_t wn[NI];
        bool bit[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) { wn[q] = mem[q].sw[(WG ? (size_t)((ref >> 16)) * mem[q].sws : (size_t)(((ref >> 16)) * PPM_LANES)) + lane]; bit[q] = (wv[q] & mk) != 0u; }
        const double nK = -K;
        load_rec(T.recs, e + 1, K, ref, mk);  // the table carries one padding record
#pragma unroll
        for (int q = 0; q < NI; q++) {
          const double aa = bit[q] ? la1 : la0;
          const double nbb = (bit[q] ? ld1 : ld0) * nK;
#pragma unroll
          for (int i = 0; i < R; i++) prod[q][i] *= fma(nbb, Y[i], aa);
          wv[q] = wn[q];
        }
      }
#pragma unroll
      for (int q = 0; q < NI; q++)
        if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
    }
    // ---- u-form leaves (all variants).  Pairs: one record per two leaves, their two pattern bits pick one of four
    // quadratics q.x + q.y u + q.z u^2 (3 FP64 per slice for two leaf terms, no selects); the next record's bits and
    // table row are fetched while the current products run.
    if (!HOT && bd.e_pair0 < bd.e_leafu0) {
      const int pmax = m.n_pairs - 1;
      double K; uint32_t ref, mk;
      load_rec(T.recs, bd.e_pair0, K, ref, mk);
      uint32_t cb[NI];   // observation pair of the current record's pair, per item
      int pid = (int)(ref & 0xffffu);
#pragma unroll
      for (int q = 0; q < NI; q++) cb[q] = (mem[q].sw[(WG ? (size_t)((pid >> 4)) * mem[q].sws : (size_t)(((pid >> 4)) * PPM_LANES)) + lane] >> (2 * (pid & 15))) & 3u;
      if (!factored) {
        // record = {pattern word byte offset, pairq byte offset, pair id, rotation}: word, rotate, mask, row -- no index arithmetic
        const char* cwl[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) cwl[q] = reinterpret_cast<const char*>(mem[q].sw + lane);
        const char* pq = reinterpret_cast<const char*>(T.pairq);
        double4 qv[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) qv[q] = T.pairq[pid * 4 + (int)cb[q]];
        // three-stage software pipeline (record -> pattern word -> table row are dependent shared-memory loads):
        // iteration e fetches the record of e+3, the word of e+2 and the row of e+1 while pair e is multiplied in.
        // Global-memory variant: the pair records stream through the per-warp ring as well (chunks of 32, two chunks
        // ahead), so the loop never waits on L2 for a record.
        const int elast = bd.e_leafu0 - 1;
        const int B0 = bd.e_pair0;
        const bool ring = DEEP && PPM_LANES == 32;
        auto pstage = [&](int c) {
          int idx = B0 + c * 32 + lane;
          idx = idx > m.n_entries ? m.n_entries : idx;  // the table ends with one padding record
          ppm_cp_async16(wrec + (c % 3) * 32 + lane, T.recs + idx);
          ppm_cp_commit();
        };
        auto prec = [&](int idx, double& Kx, uint32_t& refx, uint32_t& mkx) {
          if (ring) { const int o = idx - B0; load_rec(wrec, ((o >> 5) % 3) * 32 + (o & 31), Kx, refx, mkx); }
          else load_rec(T.recs, idx, Kx, refx, mkx);
        };
        if (ring) { pstage(0); pstage(1); ppm_cp_wait1(); ppm_syncwarp(); }
        if constexpr (DEEP) {
          // Global-memory variant: the quadratics' rows come from L2 on big trees (the table exceeds L1), so the pipeline is
          // deeper: iteration e fetches the record of e+PD+2, the word of e+PD+1 and the row of e+PD (PD = 3 measured best:
          // 5,000 genomes +16 %, 10,000-leaf caterpillar +24 % over PD = 1; PD = 5 spills and loses at 1,000 genomes).
          constexpr int PD = 3;
          double Kr[PD + 2]; uint32_t rt[PD + 2];
          uint32_t W[NI][PD + 1];
          double4 Q[NI][PD + 1];
#pragma unroll
          for (int k = 0; k <= PD; k++) prec(B0 + 1 + k < elast ? B0 + 1 + k : elast, Kr[k], ref, rt[k]);
#pragma unroll
          for (int k = 0; k < PD; k++) {
            int wo, qo;
            unpack2i(Kr[k], wo, qo);
#pragma unroll
            for (int q = 0; q < NI; q++) W[q][k] = *reinterpret_cast<const uint32_t*>(cwl[q] + wo);
          }
#pragma unroll
          for (int q = 0; q < NI; q++) Q[q][0] = qv[q];
#pragma unroll
          for (int k = 1; k < PD; k++) {
            int wo, qo;
            unpack2i(Kr[k - 1], wo, qo);
#pragma unroll
            for (int q = 0; q < NI; q++) Q[q][k] = *reinterpret_cast<const double4*>(pq + qo + (rotr32(W[q][k - 1], rt[k - 1]) & 0x60u));
          }
          for (int e = B0; e <= elast;) {
            if (ring && ((e - B0) & 31) == 0) {
              // chunk boundary: the ring slot of the previous chunk is free again, the next chunk must have landed
              ppm_syncwarp();
              pstage(((e - B0) >> 5) + 2);
              ppm_cp_wait1();
              ppm_syncwarp();
            }
            const int ce = (e + 16 <= elast) ? e + 16 : elast + 1;
#pragma unroll 4
            for (; e < ce; e++) {
              prec(e + PD + 2 < elast ? e + PD + 2 : elast, Kr[PD + 1], ref, rt[PD + 1]);
              int woN, qoN, wo_, qo_;
              unpack2i(Kr[PD], woN, qo_);
              unpack2i(Kr[PD - 1], wo_, qoN);
#pragma unroll
              for (int q = 0; q < NI; q++) {
                Q[q][PD] = *reinterpret_cast<const double4*>(pq + qoN + (rotr32(W[q][PD - 1], rt[PD - 1]) & 0x60u));
                W[q][PD] = *reinterpret_cast<const uint32_t*>(cwl[q] + woN);
              }
#pragma unroll
              for (int q = 0; q < NI; q++) {
#pragma unroll
                for (int i = 0; i < R; i++) prod[q][i] *= fma(fma(Q[q][0].z, U[i], Q[q][0].y), U[i], Q[q][0].x);
#pragma unroll
                for (int k = 0; k < PD; k++) { Q[q][k] = Q[q][k + 1]; W[q][k] = W[q][k + 1]; }
              }
#pragma unroll
              for (int k = 0; k <= PD; k++) { Kr[k] = Kr[k + 1]; rt[k] = rt[k + 1]; }
            }
#pragma unroll
            for (int q = 0; q < NI; q++)
              if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
          }
        } else {
          double K1, K2; uint32_t rot1, rot2;
          prec(B0 + 1 < elast ? B0 + 1 : elast, K1, ref, rot1);  // (the record behind the segment is not a pair)
          prec(B0 + 2 < elast ? B0 + 2 : elast, K2, ref, rot2);
          uint32_t w1[NI];
          {
            int wo, qo;
            unpack2i(K1, wo, qo);
  #pragma unroll
            for (int q = 0; q < NI; q++) w1[q] = *reinterpret_cast<const uint32_t*>(cwl[q] + wo);
          }
          for (int e = B0; e <= elast;) {
            if (ring && ((e - B0) & 31) == 0) {
              // chunk boundary: the ring slot of the previous chunk is free again, the next chunk must have landed
              ppm_syncwarp();
              pstage(((e - B0) >> 5) + 2);
              ppm_cp_wait1();
              ppm_syncwarp();
            }
            const int ce = (e + 16 <= elast) ? e + 16 : elast + 1;
  #pragma unroll 4
            for (; e < ce; e++) {
              double K3; uint32_t rot3;
              prec(e + 3 < elast ? e + 3 : elast, K3, ref, rot3);
              int wo1, qo1, wo2, qo2;
              unpack2i(K1, wo1, qo1);
              unpack2i(K2, wo2, qo2);
              double4 qn[NI];
              uint32_t w2[NI];
  #pragma unroll
              for (int q = 0; q < NI; q++) {
                qn[q] = *reinterpret_cast<const double4*>(pq + qo1 + (rotr32(w1[q], rot1) & 0x60u));
                w2[q] = *reinterpret_cast<const uint32_t*>(cwl[q] + wo2);
              }
  #pragma unroll
              for (int q = 0; q < NI; q++) {
  #pragma unroll
                for (int i = 0; i < R; i++) prod[q][i] *= fma(fma(qv[q].z, U[i], qv[q].y), U[i], qv[q].x);
                qv[q] = qn[q]; w1[q] = w2[q];
              }
              K1 = K2; rot1 = rot2; K2 = K3; rot2 = rot3;
            }
  #pragma unroll
            for (int q = 0; q < NI; q++)
              if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
          }
        }
        if (ring) { ppm_cp_wait0(); ppm_syncwarp(); }
      } else {  // pi1 close to 1: two linear factors keep full accuracy
        for (int e = bd.e_pair0; e < bd.e_leafu0; e++) {
          const int pa = 2 * pid, pb = 2 * pid + 1;
#pragma unroll
          for (int q = 0; q < NI; q++) {
            const double2 fa = T.leafab[pa * 2 + (int)(cb[q] & 1u)], fb = T.leafab[pb * 2 + (int)(cb[q] >> 1)];
#pragma unroll
            for (int i = 0; i < R; i++) prod[q][i] *= fma(fa.y, U[i], fa.x) * fma(fb.y, U[i], fb.x);
          }
          load_rec(T.recs, e + 1, K, ref, mk);
          pid = (int)(ref & 0xffffu);
          pid = pid < pmax ? pid : pmax;
#pragma unroll
          for (int q = 0; q < NI; q++) cb[q] = (mem[q].sw[(WG ? (size_t)((pid >> 4)) * mem[q].sws : (size_t)(((pid >> 4)) * PPM_LANES)) + lane] >> (2 * (pid & 15))) & 3u;
          if (((e - bd.e_pair0) & 7) == 7) {
#pragma unroll
            for (int q = 0; q < NI; q++)
              if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
          }
        }
      }
      if (factored) {  // (the quadratic loop above ends every chunk with a range check; this one checks every 8 pairs)
#pragma unroll
        for (int q = 0; q < NI; q++)
          if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
      }
    }
    // single leaves in u-form: ref = bit position | next << 16; the term alpha + beta u comes from a per-lane table row
    if (!HOT && bd.e_leafu0 < bd.e_leaf0) {
      double K; uint32_t ref, mk;
      load_rec(T.recs, bd.e_leafu0, K, ref, mk);
      double2 fv[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) {
        const uint32_t ps = ref & 0xffffu;
        const uint32_t bt = (mem[q].sw[(WG ? (size_t)((ps >> 5)) * mem[q].sws : (size_t)(((ps >> 5)) * PPM_LANES)) + lane] >> (ps & 31u)) & 1u;
        fv[q] = T.leafab[ps * 2 + bt];
      }
      for (int e = bd.e_leafu0; e < bd.e_leaf0; e++) {
        double2 fn[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) {
          const uint32_t ps = ref >> 16;
          const uint32_t bt = (mem[q].sw[(WG ? (size_t)((ps >> 5)) * mem[q].sws : (size_t)(((ps >> 5)) * PPM_LANES)) + lane] >> (ps & 31u)) & 1u;
          fn[q] = T.leafab[ps * 2 + bt];
        }
        load_rec(T.recs, e + 1, K, ref, mk);
#pragma unroll
        for (int q = 0; q < NI; q++) {
#pragma unroll
          for (int i = 0; i < R; i++) prod[q][i] *= fma(fv[q].y, U[i], fv[q].x);
          fv[q] = fn[q];
        }
        if (((e - bd.e_leafu0) & 15) == 15) {
#pragma unroll
          for (int q = 0; q < NI; q++)
            if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < NI; q++)
        if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
    }
    // (3) leaves merged inside the block: ref = bit position | next << 16, mk = code | live slices << 16
    if (bd.e_pleaf0 < bd.e_pslot0 && !HOT) {
      double K; uint32_t ref, mk;
      load_rec(T.recs, bd.e_pleaf0, K, ref, mk);
      uint32_t wv[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) wv[q] = mem[q].sw[(WG ? (size_t)(((ref & 0xffffu) >> 5)) * mem[q].sws : (size_t)((((ref & 0xffffu) >> 5)) * PPM_LANES)) + lane];
      for (int e = bd.e_pleaf0; e < bd.e_pslot0; e++) {
        uint32_t wn[NI];
        double aa[NI], nbb[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) {
          wn[q] = mem[q].sw[(WG ? (size_t)(((ref >> 16) >> 5)) * mem[q].sws : (size_t)((((ref >> 16) >> 5)) * PPM_LANES)) + lane];
          const bool bit = (wv[q] >> (ref & 31u)) & 1u;
          aa[q] = bit ? la1 : la0;
          nbb[q] = (bit ? ld1 : ld0) * -K;
        }
        const uint32_t mcur = mk;
        load_rec(T.recs, e + 1, K, ref, mk);
        coded_update<R, NI>(prod, Y, aa, nbb, mcur);
#pragma unroll
        for (int q = 0; q < NI; q++) wv[q] = wn[q];
      }
    }
    // (4) internal lineages born and/or merged inside the block: ref = slot | next slot << 16
    if (bd.e_pslot0 < bd.e_end && !HOT) {
      double K; uint32_t ref, mk;
      load_rec(T.recs, bd.e_pslot0, K, ref, mk);
      double2 ad[NI];
#pragma unroll
      for (int q = 0; q < NI; q++) ad[q] = mem[q].slots.ld((int)(ref & 0xffffu), lane);
      for (int e = bd.e_pslot0; e < bd.e_end; e++) {
        double2 an[NI];
        double aa[NI], nbb[NI];
#pragma unroll
        for (int q = 0; q < NI; q++) {
          mem[q].slots.ld_async((int)(ref >> 16), lane, an[q]);
          aa[q] = ad[q].x;
          nbb[q] = ad[q].y * -K;
        }
        const uint32_t mcur = mk;
        load_rec(T.recs, e + 1, K, ref, mk);
        coded_update<R, NI>(prod, Y, aa, nbb, mcur);
#pragma unroll
        for (int q = 0; q < NI; q++) { mem[q].slots.wait_async(an[q]); ad[q] = an[q]; }
      }
    }
    // ---- block end: integral += w/nSub * product (functions.cpp:273-275), exponents aligned
#pragma unroll
    for (int q = 0; q < NI; q++) {
      // (every loop above ends with a range check of its own; the generic partial loops and the deep variant's tails do not)
      if (DEEP || (!HOT && bd.e_pleaf0 < bd.e_end))
        if (ppm_any(products_small<R>(prod[q]))) rescale_products<R>(prod[q], ex[q]);
      if (lp) {  // the alive leaves of every slice as one power: mantissa in [0.70, 1.42] x 2^k
#pragma unroll
        for (int i = 0; i < R; i++) {
          const double2 ce = T.lce[b * R + i];
          double s2 = fma(lp_i2d(nv[q][i]), ce.x, T.lb[b * R + i]);
          if (hs1) s2 = fma(lp_i2d(sv[q][i]), ce.y, s2);
          int kx;
          prod[q][i] *= lp_pow2(s2, kx);
          ex[q][i] -= kx;
        }
      }
      if (DIST) {  // expectedTime2_aux (functions.cpp:713-746): the integrand exp(f_aux) of every slice
        if (a.dist && active[q]) {
          double* row = a.dist + (size_t)(pat[q] - a.first) * (m.n_slices + 1);
#pragma unroll
          for (int i = 0; i < R; i++)
            if (b * R + i < m.n_slices) row[b * R + i] = ldexp(prod[q][i], -ex[q][i]);
        }
      }
      if (lp) {  // the slices' exponents differ as a rule: align them to the smallest (zero products -- padding slices -- aside)
        // (padding slices carry weight 0 and the tables of the last real slice, so their exponents do not stand out)
        int emin = ex[q][0];
#pragma unroll
        for (int i = 1; i < R; i++) emin = ex[q][i] < emin ? ex[q][i] : emin;
        double tsum = 0.;
#pragma unroll
        for (int i = 0; i < R; i++) {
          int d = ex[q][i] - emin;
          d = d > 1000 ? 1000 : d;
          tsum = fma(DIST ? prod[q][i] * T.wt[b * R + i] : prod[q][i], pow2i(-d), tsum);
        }
        scaled_add(Im[q], Ie[q], tsum, emin);
        continue;
      }
      bool same = true;
#pragma unroll
      for (int i = 1; i < R; i++) same = same && (ex[q][i] == ex[q][0]);
      if (ppm_all(same)) {  // usual case: one common exponent, one scaled add; pairwise sum keeps the chain short
        double t[R];
#pragma unroll
        for (int i = 0; i < R; i++) t[i] = DIST ? prod[q][i] * T.wt[b * R + i] : prod[q][i];
#pragma unroll
        for (int st = 1; st < R; st *= 2)
#pragma unroll
          for (int i = 0; i + st < R; i += 2 * st) t[i] += t[i + st];
        scaled_add(Im[q], Ie[q], t[0], ex[q][0]);
      } else {
#pragma unroll
        for (int i = 0; i < R; i++) scaled_add(Im[q], Ie[q], DIST ? prod[q][i] * T.wt[b * R + i] : prod[q][i], ex[q][i]);
      }
    }
  }

#pragma unroll
  for (int q = 0; q < NI; q++) {
    if (DIST) {  // last column: exp(likRep2) = the integral itself (functions.cpp:756)
      if (a.dist && active[q]) a.dist[(size_t)(pat[q] - a.first) * (m.n_slices + 1) + m.n_slices] = ldexp(Im[q], -Ie[q]);
    }
    // ---- mixture (functions.cpp:345-358) and category argmax (functions.cpp:497-509)
    if (active[q]) {
      // log p1(root) of the two pure-loss categories, from the class-sum kernel
      const double2 A01 = a.t.a01[(size_t)(pat[q] - a.first) * a.t.P + p];  // [pattern][P]: one strided read per item
      const double lik0 = sc[SC_C0] + A01.x;
      const int mr = m.mrca[pat[q]];
      const double2 du = a.t.d1u[(size_t)p * m.n_nodes + mr];
      double lik1;
      if (m.freq[pat[q]] == 0) lik1 = lse2(sc[SC_C1] + A01.y, sc[SC_C1] + (du.y + sc[SC_LOG1ME2]));  // all-zero gene row: mrca is a leaf (functions.cpp:216)
      else lik1 = sc[SC_C1] + (A01.y + du.x);
      const double lik2 = (Im[q] > 0.) ? sc[SC_C2] + (log(Im[q]) - (double)Ie[q] * 0.693147180559945309417232) : -INFINITY;
      double mx = fmax(lik0, fmax(lik1, lik2));
      double tot = (mx == -INFINITY) ? -INFINITY : mx + log(exp(lik0 - mx) + exp(lik1 - mx) + exp(lik2 - mx));
      contrib[q] = tot * (double)m.counts[pat[q]];
      if (!HOT && a.mode == 2) {
        int c = 0;
        if (lik1 > lik0) c = 1;
        if (lik2 > (c ? lik1 : lik0)) c = 2;
        long long r = pat[q] - a.first;
        if (a.cat) a.cat[r] = (signed char)c;
        if (a.comp) { a.comp[3 * r] = lik0; a.comp[3 * r + 1] = lik1; a.comp[3 * r + 2] = lik2; }
      }
    }
  }
}

}  // namespace ppm
