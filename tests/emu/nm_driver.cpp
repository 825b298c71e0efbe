// TEST INFRASTRUCTURE.  Runs ppmmodelpangenome_b200/csrc/nmkb.h on the analytic test functions and prints the
// LOGICAL evaluation sequence in the same format as oracle/nm_ref_driver.cpp.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "../../ppmmodelpangenome_b200/csrc/nmkb.h"
#include "nm_functions.h"

int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  bool spec = argc > 2 ? atoi(argv[2]) != 0 : true;
  std::array<double, 8> start, lower, upper;
  nm_test_setup(which, start, lower, upper);
  using NM = ppm::Nmkb<8>;
  NM::BatchFn f = [&](const std::vector<NM::Point>& x, std::vector<double>& y) {
    for (size_t i = 0; i < x.size(); i++) y[i] = nm_test_function(which, x[i]);
  };
  NM::Observer obs = [&](const NM::Point& x, int k, double y) {
    printf("eval %d", k);
    for (double v : x) printf(" %a", v);
    printf(" %a\n", y);
  };
  const int level = argc > 3 ? atoi(argv[3]) : 2;
  if (argc > 4) {
    // multi-start mode: the three test functions as three trajectories whose requests are served round by round, interleaved
    // (what the CLI's PPM_STARTS driver does); prints trajectory `which` only -- it must equal the single run
    using ST = ppm::NmkbStepper<8>;
    std::vector<std::unique_ptr<ST>> st;
    for (int w = 0; w < 3; w++) {
      std::array<double, 8> s0, lo, up;
      nm_test_setup(w, s0, lo, up);
      st.emplace_back(new ST(w == which ? obs : NM::Observer(), lo, up, s0, 0.0001, 3000, 10, level));
    }
    int rounds = 0;
    while (true) {
      bool any = false;
      for (int w = 0; w < 3; w++)
        if (!st[w]->done()) {
          any = true;
          const std::vector<NM::Point>& x = st[w]->request();
          std::vector<double> y(x.size());
          for (size_t i = 0; i < x.size(); i++) y[i] = nm_test_function(w, x[i]);
          st[w]->supply(y.data());
        }
      if (!any) break;
      rounds++;
    }
    const ppm::NmkbResult<8>& r = st[which]->result();
    printf("result %d %a", r.icount, r.ymin);
    for (double v : r.xmin) printf(" %a", v);
    printf(" %s\n", r.message.c_str());
    fprintf(stderr, "evaluated %d launches %d\n", r.evaluated, r.launches);
    return 0;
  }
  NM nm(f, obs, lower, upper);
  nm.speculate = spec;
  nm.level = level;
  ppm::NmkbResult<8> r = nm.minimize(start, 0.0001, 3000, 10);
  printf("result %d %a", r.icount, r.ymin);
  for (double v : r.xmin) printf(" %a", v);
  printf(" %s\n", r.message.c_str());
  fprintf(stderr, "evaluated %d launches %d\n", r.evaluated, r.launches);
  return 0;
}
