// Analytic test objectives shared by the reference-optimiser driver (oracle/nm_ref_driver.cpp) and the driver
// of our batched nmkb (tests/emu/nm_driver.cpp).
#pragma once
#include <array>
#include <cmath>
inline void nm_test_setup(int which, std::array<double, 8>& start, std::array<double, 8>& lower, std::array<double, 8>& upper) {
  for (int i = 0; i < 8; i++) { lower[i] = 0.; upper[i] = 10. + i; start[i] = 1. + 0.7 * i; }
  if (which == 2) for (int i = 0; i < 8; i++) start[i] = 0.3 + 0.1 * i;
}
inline double nm_test_function(int which, const std::array<double, 8>& x) {
  double s = 0;
  if (which == 0) {  // smooth bowl with cross terms
    for (int i = 0; i < 8; i++) s += (1 + i) * (x[i] - 2. - 0.5 * i) * (x[i] - 2. - 0.5 * i);
    for (int i = 0; i + 1 < 8; i++) s += 0.3 * x[i] * x[i + 1];
  } else if (which == 1) {  // Rosenbrock chain
    for (int i = 0; i + 1 < 8; i++) s += 100 * (x[i + 1] - x[i] * x[i]) * (x[i + 1] - x[i] * x[i]) + (1 - x[i]) * (1 - x[i]);
  } else {  // kinked + plateaus: provokes failed contractions (shrink) and restarts
    for (int i = 0; i < 8; i++) s += std::fabs(x[i] - 3.) + 0.5 * std::floor(4 * std::fabs(x[i] - 5.)) + 0.01 * std::sin(37 * x[i]);
  }
  return s;
}
