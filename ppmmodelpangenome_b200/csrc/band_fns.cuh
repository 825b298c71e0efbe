// Per-item arithmetic of the band kernel that is not lane-parallel: the phase A merge and the final mixture.  __host__
// __device__ like ppm_device_fns.cuh, so that the CPU test tier (tests/emu) runs the same arithmetic.
#pragma once
#include "ppm_device_fns.cuh"

namespace ppm {

// Phase A merge (objects.cpp:427-435, linear domain).  (xl, yl), (xr, yr): child operands -- an internal child's record
// (a, nbn) or a leaf's (a, d); o = the op's per-parameter constants (BOP_DOUBLES).  Returns the node's record
// (a, nbn = -d exp(lam (h - H/2))), mantissa-normalised: when max(|a|, |d|) < 2^-20 both are scaled by 2^k so that the
// larger lands in [0.5, 1); k is returned (the slice exponents add it for every slice at or after the node's birth).
PPM_HD double2 band_merge(double xl, double yl, double xr, double yr, const double* o, double pi1, int& k) {
  const double m0 = fma(o[0], yl, xl) * fma(o[2], yr, xr);
  const double m1 = fma(o[1], yl, xl) * fma(o[3], yr, xr);
  double d = m1 - m0;
  double a = fma(pi1, d, m0);
  const int ha = ppm_double2hiint(a) & 0x7fffffff, hd = ppm_double2hiint(d) & 0x7fffffff;
  const int h = ha > hd ? ha : hd;
  const bool small = (uint32_t)(h - 0x00100000) < (uint32_t)(((1023 - 20) << 20) - 0x00100000);  // 0 < exponent < 1003
  k = small ? 1022 - (h >> 20) : 0;
  if (small) { const double sf = pow2i(k); a *= sf; d *= sf; }
  return make_double2(a, d * o[4]);
}

// sparse evaluation: is this pattern served from its term list?  (0: dense; 1 / 2: against the all-zero / all-ones row)
PPM_HD int bs_sparse_use(const ModelDev& m, const SparseLists& l, int n_tiles, unsigned int max_len, long long pat) {
  const int base = bs_sparse_item(m, pat);
  if (base == 0) return 0;
  const unsigned int len = l.off[(size_t)(pat + 1) * n_tiles] - l.off[(size_t)pat * n_tiles];
  return len <= max_len ? base : 0;
}

// Mixture of one (pattern, parameter vector) given its Mobile integral Im * 2^-Ie: functions.cpp:345-358 with the
// pure-loss categories from the class-sum kernel (a01) and the Private below-root table (d1u).  Returns count * log-mixture.
PPM_HD double band_mixture(const ModelDev& m, const TablesDev& t, int p, long long pat, double Im, int Ie, double* comp3) {
  const double* sc = t.scal + (size_t)p * SC_COUNT;
  const double2 A01 = t.a01[(size_t)pat * t.P + p];
  const double lik0 = sc[SC_C0] + A01.x;
  const int mr = m.mrca[pat];
  const double2 du = t.d1u[(size_t)p * m.n_nodes + mr];
  double lik1;
  if (m.freq[pat] == 0) lik1 = lse2(sc[SC_C1] + A01.y, sc[SC_C1] + (du.y + sc[SC_LOG1ME2]));  // all-zero gene row (functions.cpp:216)
  else lik1 = sc[SC_C1] + (A01.y + du.x);
  const double lik2 = (Im > 0.) ? sc[SC_C2] + (log(Im) - (double)Ie * 0.693147180559945309417232) : -INFINITY;
  const double mx = fmax(lik0, fmax(lik1, lik2));
  const double tot = (mx == -INFINITY) ? -INFINITY : mx + log(exp(lik0 - mx) + exp(lik1 - mx) + exp(lik2 - mx));
  if (comp3) { comp3[0] = lik0; comp3[1] = lik1; comp3[2] = lik2; }
  return tot * (double)m.counts[pat];
}

}  // namespace ppm
