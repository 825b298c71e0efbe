// inference -- drop-in for the reference CLI (reference source/inference.cpp:17-58):
//   ./inference seed tree.nwk pa_matrix.txt mle_param.txt inf_cat.txt [N0 l0 i1 l1 g2 l2 s_10 s_01]
// Same positional arguments, same progress lines on stdout, same output files (mle_param.txt appended with
// 13 fields, inf_cat.txt one category per input gene followed by a blank).  The objective is evaluated on
// the GPU through the C ABI (include/ppm_b200.h); the optimiser is the batched nmkb of nmkb.h.
// Environment: PPM_DEVICE (CUDA ordinal, default 0), PPM_GPUS (default 1: number of GPUs, devices PPM_DEVICE ..
// PPM_DEVICE+n-1, patterns sharded over them behind one group handle; or PPM_GPU_IDS=0,1,5 for an explicit list), PPM_DEDUPE (default 1: unique patterns with
// multiplicities; 0 evaluates every gene column like the reference does), PPM_SPECULATE (0 serial, 1 the 4 trial points of an
// iteration per launch, 2 also its 8 shrink points; default by the number of starts), PPM_STARTS (multi-start: S seeds together).
#include <array>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
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
  // One fitted trajectory: what the reference prints and stores for one seed (functions.cpp:404-428 / 437-461)
  struct Run {
    std::string seed;
    Point start{};
    std::ostringstream out;   // its stdout block, in a stream of its own (the reference's setprecision(9) persists per process)
    Point mle{};
    double mle_i2 = 0, mle_lik = 0;
    int icount = 0, evaluated = 0, launches = 0;
    std::string message;
  };
  // Fits every run; the requests of all live trajectories share one ppm_loglik launch per round (multi-start).  Each
  // trajectory is the serial algorithm's, bit for bit: a point's objective does not depend on what else is in the launch.
  void optim_all(std::vector<Run>& runs, int level) {
    using Stepper = ppm::NmkbStepper<8>;
    std::vector<std::unique_ptr<Stepper>> st;
    for (Run& r : runs) {
      Run* rp = &r;
      Stepper::Observer obs = [rp](const Point& par, int count, double y) {
        std::ostringstream& o = rp->out;
        o << count << "th iteration: ";
        o << par[0] << ' ' << par[1] << ' ' << par[2] << ' ' << par[3] << ' ' << par[4] << ' ' << par[5] << ' ' << par[6] << ' '
          << par[7] << ' ';
        o << std::setprecision(9) << -y << endl;  // the manipulator persists, as in the reference
      };
      st.emplace_back(new Stepper(obs, lower, upper, r.start, 0.0001, 3000, 10, level));
    }
    std::vector<double> flat, ll;
    long long rounds = 0, points = 0;
    while (true) {
      flat.clear();
      for (auto& t : st)
        if (!t->done())
          for (const Point& x : t->request()) flat.insert(flat.end(), x.begin(), x.end());
      if (flat.empty()) break;
      const int32_t P = (int32_t)(flat.size() / 8);
      ll.assign(P, 0.);
      check(ppm_loglik(h, flat.data(), P, ll.data(), nullptr), "ppm_loglik");
      for (double& v : ll) v = -v;
      size_t off = 0;
      for (auto& t : st)
        if (!t->done()) {
          const size_t n = t->request().size();
          t->supply(ll.data() + off);
          off += n;
        }
      rounds++; points += P;
    }
    // the post-fit i2 of every run in one call (functions.cpp:420, 453)
    std::vector<double> pm(runs.size() * 8), i2(runs.size());
    for (size_t k = 0; k < runs.size(); k++) {
      const ppm::NmkbResult<8>& res = st[k]->result();
      runs[k].mle = res.xmin; runs[k].mle_lik = -res.ymin; runs[k].icount = res.icount; runs[k].message = res.message;
      runs[k].evaluated = res.evaluated; runs[k].launches = res.launches;
      std::copy(res.xmin.begin(), res.xmin.end(), pm.begin() + 8 * k);
    }
    check(ppm_compute_i2(h, pm.data(), (int32_t)runs.size(), i2.data(), nullptr), "ppm_compute_i2");
    for (size_t k = 0; k < runs.size(); k++) {
      Run& r = runs[k];
      r.mle_i2 = i2[k];
      std::ostringstream& o = r.out;
      o << "Found minimum: " << r.mle[0] << ' ' << r.mle[1] << ' ' << r.mle[2] << ' ' << r.mle[3] << ' ' << r.mle_i2 << ' ' << r.mle[4]
        << ' ' << r.mle[5] << ' ' << r.mle[6] << ' ' << r.mle[7] << endl;
      o << "upper limit:   " << upper[0] << ' ' << upper[1] << ' ' << upper[2] << ' ' << upper[3] << '\t' << upper[4] << ' '
        << upper[5] << ' ' << upper[6] << ' ' << upper[7] << endl;
      o << "logLik value: " << r.mle_lik << endl;
      o << "Number of function evaluation: " << r.icount << endl;
      o << "Message: " << r.message << endl;
    }
    if (env_int("PPM_VERBOSE", 0))
      std::cerr << "[ppm_b200] " << runs.size() << " trajectories, speculation level " << level << ": " << points << " points evaluated in "
                << rounds << " launches" << std::endl;
  }
  void adopt(const Run& r) { mle = r.mle; mle_i2 = r.mle_i2; mle_lik = r.mle_lik; }
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
  void print_compo(std::ostream& o) {  // functions.cpp:522-549
    double parts[4];
    compute_i2(mle, parts);
    o << "Expected pangenome composition:" << endl;
    o << mle[0] * parts[0] << " Persistent genes" << endl;
    o << mle[2] * parts[1] / mle[3] + mle[2] * parts[2] << " Private genes" << endl;
    o << mle_i2 * parts[3] << " Mobile genes" << endl;
  }
};

int main(int argc, char* argv[]) {
  // seed 0 = start from the 8 values on the command line (inference.cpp:39-46): they must be there
  const bool args_ok = (argc == 6 || argc == 14) && (std::atoi(argv[1]) != 0 || argc == 14);
  if (!args_ok) {
    std::cerr << "usage: " << argv[0] << " seed tree.nwk pa_matrix.txt mle_param.txt inf_cat.txt [N0 l0 i1 l1 g2 l2 s_10 s_01]\n"
              << "  seed 0 needs the 8 start values; PPM_STARTS=S fits the seeds seed .. seed+S-1 together (multi-start)\n";
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
  // PPM_STARTS=S (seed != 0): the seeds seed .. seed+S-1, fitted together; the output is that of S consecutive runs of the
  // reference CLI (S blocks on stdout, S lines appended to mle_param.txt, inf_cat.txt of the last seed)
  const int seed0 = std::stoi(argv[1]);
  const int n_starts = seed0 != 0 ? std::max(1, env_int("PPM_STARTS", 1)) : 1;
  std::vector<Fit::Run> runs(n_starts);
  for (int k = 0; k < n_starts; k++) {
    if (seed0 != 0) { runs[k].seed = std::to_string(seed0 + k); runs[k].start = fit.draw_start(seed0 + k); }
    else { runs[k].seed = argv[1]; for (int i = 0; i < 8; i++) runs[k].start[i] = std::stod(argv[6 + i]); }
  }
  // speculation: 2 = the 4 trial points and the 8 shrink points of an iteration in one launch, 1 = the 4 trial points,
  // 0 = serial.  Default: whatever keeps a launch near 256 parameter vectors (what fills one B200 on a 100-genome tree).
  int level = n_starts <= 21 ? 2 : (n_starts <= 64 ? 1 : 0);
  if (std::getenv("PPM_SPECULATE")) level = std::max(0, std::min(2, env_int("PPM_SPECULATE", 2)));
  fit.optim_all(runs, level);
  for (int k = 0; k < n_starts; k++) {
    fit.adopt(runs[k]);
    fit.write_param(runs[k].seed, argv[4]);
    if (k == n_starts - 1) fit.assign_cat(argv[5]);
    fit.print_compo(runs[k].out);
    cout << runs[k].out.str();
  }
  if (n_starts > 1) {
    int best = 0;
    for (int k = 1; k < n_starts; k++) if (runs[k].mle_lik > runs[best].mle_lik) best = k;
    std::cerr << "[ppm_b200] best of " << n_starts << " starts: seed " << runs[best].seed << ", logLik " << std::setprecision(12)
              << runs[best].mle_lik << std::endl;
  }
  ppm_destroy(fit.h);
  time_t fin = time(NULL);
  cout << "DONE :)    Time duration: " << fin - debut << " seconds" << endl;
  return 0;
}
