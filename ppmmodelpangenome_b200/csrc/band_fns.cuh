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

// Leaf powers: 2^s = p * 2^k with p in [0.70, 1.42]; s <= ~0 is the log2 of a product of leaf terms (band_program.h: leaf_pow).
// Round to nearest integer with the 1.5 * 2^52 trick, then exp(r ln 2) by its Taylor polynomial of degree 13 on |r| <= 1/2
// (remainder < 5e-18).  s below -1e9 (an impossible observation) is clamped: the slice then weighs 2^-1e9.
PPM_HD double lp_pow2(double s, int& k) {
  s = fmax(s, -1.0e9);
  const double magic = 6755399441055744.0;
  const double t = s + magic;
#if defined(__CUDA_ARCH__)
  k = __double2loint(t);
#else
  { long long b; memcpy(&b, &t, 8); k = (int)(uint32_t)b; }
#endif
  const double x = (s - (t - magic)) * 0.693147180559945309417232;
  double p = 1.6059043836821613e-10;           // 1/13!
  p = fma(p, x, 2.08767569878681e-09);         // 1/12!
  p = fma(p, x, 2.505210838544172e-08);        // 1/11!
  p = fma(p, x, 2.755731922398589e-07);        // 1/10!
  p = fma(p, x, 2.7557319223985893e-06);       // 1/9!
  p = fma(p, x, 2.48015873015873e-05);         // 1/8!
  p = fma(p, x, 0.0001984126984126984);        // 1/7!
  p = fma(p, x, 0.001388888888888889);         // 1/6!
  p = fma(p, x, 0.008333333333333333);         // 1/5!
  p = fma(p, x, 0.041666666666666664);         // 1/4!
  p = fma(p, x, 0.16666666666666666);          // 1/3!
  p = fma(p, x, 0.5);
  p = fma(p, x, 1.);
  p = fma(p, x, 1.);
  return p;
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
