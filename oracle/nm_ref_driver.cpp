// TEST INFRASTRUCTURE ONLY.  Runs the reference's own optimiser header (compiled from /root/reference/source,
// never copied) on analytic 8-D test functions and prints every objective evaluation as hex doubles, to pin
// the trajectory of ppmmodelpangenome_b200/csrc/nmkb.h (tests/golden/nm_trajectory.json).
#include <vector>
#include <string>
#include <cstdio>
#include <cmath>
#include <array>
#include <algorithm>
using namespace std;
#include "nelder_mead_dfoptim.h"
#include "../tests/emu/nm_functions.h"

int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  array<double, 8> start, lower, upper;
  nm_test_setup(which, start, lower, upper);
  auto fn = [&](const array<double, 8>& x, int k) {
    double y = nm_test_function(which, x);
    printf("eval %d", k);
    for (double v : x) printf(" %a", v);
    printf(" %a\n", y);
    return y;
  };
  nelder_mead_result<double, 8> r = nelder_mead<double, 8>(fn, start, lower, upper, 0.0001, 3000, 10);
  printf("result %d %a", r.icount, r.ymin);
  for (double v : r.xmin) printf(" %a", v);
  printf(" %s\n", r.message.c_str());
  return 0;
}
