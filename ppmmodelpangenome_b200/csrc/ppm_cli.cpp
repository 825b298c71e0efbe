// inference -- drop-in for the reference CLI (reference source/inference.cpp:17-58):
//   ./inference seed tree.nwk pa_matrix.txt mle_param.txt inf_cat.txt [N0 l0 i1 l1 g2 l2 s_10 s_01]
// Same positional arguments, same progress lines on stdout, same output files (mle_param.txt appended with
// 13 fields, inf_cat.txt one category per input gene followed by a blank).  The objective is evaluated on
// the GPU through the C ABI (include/ppm_b200.h); the optimiser is the batched nmkb of nmkb.h.
// Environment: PPM_DEVICE (CUDA ordinal, default 0), PPM_GPUS (default 1: number of GPUs, devices PPM_DEVICE ..
// PPM_DEVICE+n-1, patterns sharded over them behind one group handle; or PPM_GPU_IDS=0,1,5 for an explicit list), PPM_DEDUPE (default 1: unique patterns with
// multiplicities; 0 evaluates every gene column like the reference does), PPM_SPECULATE (default 1).
#include <array>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "../../include/ppm_b200.h"
#include "nmkb.h"

using std::cout;
using std::endl;
using Point = std::array<double, 8>;

static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

struct Fit {
  ppm_handle h = nullptr;
  ppm_info info{};
  Point lower{}, upper{};
  // MLE (functions.hpp:95-103)
  Point mle{};
  double mle_i2 = 0, mle_lik = 0;

  void check(int rc, const char* what) {
    if (rc) {
      std::cerr << "ppm_b200: " << what << " failed (" << rc << "): " << ppm_last_error(h) << std::endl;
      std::exit(2);
    }
  }
  double compute_i2(const Point& p, double* parts = nullptr) {
    double i2 = 0;
    check(ppm_compute_i2(h, p.data(), 1, &i2, parts), "ppm_compute_i2");
    return i2;
  }
  void bounds() {  // functions.cpp:395-396, 434-435
    const double L = info.L, mg = info.mean_genes, nObs = (double)info.n_obs;
    lower = {0., 0., 0., 0., 0., 0., 0., 0.};
    upper = {mg * 2., 10 / L, 3. * nObs / L, 500 / L, 500 / L, 25000 / L, 0.2, 0.2};
  }
  // functions.cpp:382-402: random start until computeI2 >= 0 and inside the box
  Point draw_start(int seed) {
    const double L = info.L, mg = info.mean_genes;
    const int nObs = (int)info.n_obs;
    std::mt19937 rng(seed);
    std::uniform_int_distribution<int> uniN0(mg / 3, mg);
    std::uniform_int_distribution<int> unii1(nObs / (L * 5), nObs / L);
    std::exponential_distribution<> expl0(L);
    std::exponential_distribution<> expl1(L / 50);
    std::exponential_distribution<> expl2(L / 2500);
    std::exponential_distribution<> expe(100.);
    auto draw = [&]() {
      Point s;
      s[0] = uniN0(rng) * 1.; s[1] = expl0(rng); s[2] = unii1(rng) * 1.; s[3] = expl1(rng);
      s[4] = expl1(rng); s[5] = expl2(rng); s[6] = expe(rng); s[7] = expe(rng);
      return s;
    };
    auto above = [&](const Point& s) {
      for (int i = 0; i < 8; i++) if (s[i] > upper[i]) return true;
      return false;
    };
    Point s = draw();
    while (compute_i2(s) < 0. || above(s)) s = draw();
    return s;
  }
  void optim(const Point& start) {  // functions.cpp:404-428 / 437-461
    bool first_line = true;
    ppm::Nmkb<8>::BatchFn f = [&](const std::vector<Point>& x, std::vector<double>& y) {
      std::vector<double> flat(x.size() * 8), ll(x.size());
      for (size_t i = 0; i < x.size(); i++) std::copy(x[i].begin(), x[i].end(), flat.begin() + 8 * i);
      check(ppm_loglik(h, flat.data(), (int32_t)x.size(), ll.data(), nullptr), "ppm_loglik");
      for (size_t i = 0; i < x.size(); i++) y[i] = -ll[i];
    };
    ppm::Nmkb<8>::Observer obs = [&](const Point& par, int count, double y) {
      cout << count << "th iteration: ";
      cout << par[0] << ' ' << par[1] << ' ' << par[2] << ' ' << par[3] << ' ' << par[4] << ' ' << par[5] << ' ' << par[6]
           << ' ' << par[7] << ' ';
      cout << std::setprecision(9) << -y << endl;  // the manipulator persists, as in the reference
      (void)first_line;
    };
    ppm::Nmkb<8> nm(f, obs, lower, upper);
    nm.speculate = env_int("PPM_SPECULATE", 1) != 0;
    ppm::NmkbResult<8> res = nm.minimize(start, 0.0001, 3000, 10);
    mle = res.xmin;
    mle_i2 = compute_i2(mle);
    mle_lik = -res.ymin;
    cout << "Found minimum: " << mle[0] << ' ' << mle[1] << ' ' << mle[2] << ' ' << mle[3] << ' ' << mle_i2 << ' ' << mle[4] << ' '
         << mle[5] << ' ' << mle[6] << ' ' << mle[7] << endl;
    cout << "upper limit:   " << upper[0] << ' ' << upper[1] << ' ' << upper[2] << ' ' << upper[3] << '\t' << upper[4] << ' '
         << upper[5] << ' ' << upper[6] << ' ' << upper[7] << endl;
    cout << "logLik value: " << mle_lik << endl;
    cout << "Number of function evaluation: " << res.icount << endl;
    cout << "Message: " << res.message << endl;
    if (env_int("PPM_VERBOSE", 0))
      std::cerr << "[ppm_b200] " << res.evaluated << " points evaluated in " << res.launches << " launches" << std::endl;
  }
  void write_param(const std::string& seed, const std::string& file) {  // functions.cpp:469-487
    std::fstream o;
    o.open(file, std::ios::app);
    o << seed << " " << mle[0] << " " << mle[1] << " " << mle[2] << " " << mle[3] << " " << mle_i2 << " " << mle[4] << " "
      << mle[5] << " " << mle[6] << " " << mle[6] << " " << mle[6] << " " << mle[7] << " " << mle_lik << endl;
    o.close();
  }
  void assign_cat(const std::string& file) {  // functions.cpp:489-520
    std::vector<int8_t> cat((size_t)info.n_genes);
    check(ppm_assign_cat(h, mle.data(), mle_i2, cat.data()), "ppm_assign_cat");
    std::fstream o;
    o.open(file, std::ios::out);
    for (int8_t c : cat) o << (int)c << " ";
    o.close();
  }
  void print_compo() {  // functions.cpp:522-549
    double parts[4];
    compute_i2(mle, parts);
    cout << "Expected pangenome composition:" << endl;
    cout << mle[0] * parts[0] << " Persistent genes" << endl;
    cout << mle[2] * parts[1] / mle[3] + mle[2] * parts[2] << " Private genes" << endl;
    cout << mle_i2 * parts[3] << " Mobile genes" << endl;
  }
};

int main(int argc, char* argv[]) {
  if (argc != 6 && argc != 14) {
    std::cerr << "usage: " << argv[0] << " seed tree.nwk pa_matrix.txt mle_param.txt inf_cat.txt [N0 l0 i1 l1 g2 l2 s_10 s_01]\n";
    return 1;
  }
  time_t debut = time(NULL);
  Fit fit;
  const int dedupe = env_int("PPM_DEDUPE", 1);
  const int n_gpus = env_int("PPM_GPUS", 1), dev0 = env_int("PPM_DEVICE", 0);
  int rc;
  std::vector<int32_t> ids;
  if (const char* list = std::getenv("PPM_GPU_IDS")) {
    for (const char* p = list; *p;) {
      char* end;
      long v = std::strtol(p, &end, 10);
      if (end == p) break;
      ids.push_back((int32_t)v);
      p = (*end == ',') ? end + 1 : end;
    }
  } else if (n_gpus > 1) {
    for (int k = 0; k < n_gpus; k++) ids.push_back(dev0 + k);
  }
  if (!ids.empty()) {
    rc = ppm_create_group_from_files(&fit.h, argv[2], argv[3], dedupe, ids.data(), (int32_t)ids.size());
  } else {
    rc = ppm_create_from_files(&fit.h, argv[2], argv[3], dedupe, dev0, 0, 1);
  }
  if (rc) {
    std::cerr << "ppm_b200: cannot build the model (" << rc << "): " << ppm_last_error(nullptr) << std::endl;
    return 2;
  }
  ppm_get_info(fit.h, &fit.info);
  cout << "tree OK" << endl;
  cout << "matrix OK, size " << fit.info.n_genes << "x" << fit.info.n_leaves << endl;
  cout << "inference initialization OK" << endl;
  fit.bounds();
  Point start;
  if (std::stoi(argv[1]) != 0) start = fit.draw_start(std::stoi(argv[1]));
  else for (int i = 0; i < 8; i++) start[i] = std::stod(argv[6 + i]);
  fit.optim(start);
  fit.write_param(argv[1], argv[4]);
  fit.assign_cat(argv[5]);
  fit.print_compo();
  ppm_destroy(fit.h);
  time_t fin = time(NULL);
  cout << "DONE :)    Time duration: " << fin - debut << " seconds" << endl;
  return 0;
}
