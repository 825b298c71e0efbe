// Bound-constrained Nelder-Mead ("nmkb" of the R package dfoptim) with a BATCHED objective, as a resumable stepper.
//
// The reference drives its fit with source/nelder_mead_dfoptim.h, a dfoptim-nmkb clone: tanh box transform
// (:33-52), regular initial simplex (:86-105), reflect / expand / contract (:162-222), Armijo-type "orth"
// restart (:224-253), shrink (:254-271), stop on tol / maxfeval / restarts / simplex size (:148-149).
// This is an independent restatement with the same arithmetic (same formulas, same operation order, so the
// same trajectory in IEEE double -- which is why the variable names follow the reference), built around one
// difference: the objective is asked for SEVERAL points per call.  Every candidate of an iteration (reflection,
// expansion, outside and inside contraction: :162-222) and the N points of a shrink step (:254-271) are affine in
// the current simplex, i.e. known before the reflection value is -- so up to 4 + N points are evaluated
// speculatively in ONE GPU launch and consumed in the reference's order.  The logical evaluation count (`icount`)
// and the evaluation callback order are those of the serial algorithm, including its quirk of not counting the
// evaluations of a shrink step (:270).
//
// NmkbStepper exposes one trajectory as request() / supply(): a multi-start driver collects the requests of many
// trajectories (reference start sampler: functions.cpp:382-402) into one launch; every trajectory stays bit-identical
// to its serial run because the objective of a point does not depend on what else shares the launch.
// Nmkb is the single-trajectory convenience wrapper used by the CLI and the tests.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

namespace ppm {

template <int N>
struct NmkbResult {
  std::array<double, N> xmin{};
  double ymin = 0;
  int icount = 0;        // logical evaluations, counted as the serial algorithm counts them
  int evaluated = 0;     // points actually evaluated (speculative ones included)
  int launches = 0;      // batched objective calls
  std::string message;
};

// speculation level: 0 = the reflection only (serial order), 1 = + expansion and both contractions (4 points per
// iteration), 2 = + the N shrink points (4 + N points per iteration)
template <int N>
class NmkbStepper {
 public:
  using Point = std::array<double, N>;
  // called once per LOGICAL evaluation, in the serial algorithm's order: (bounded point, count, value)
  using Observer = std::function<void(const Point& x, int count, double y)>;

  NmkbStepper(Observer obs, const Point& lower, const Point& upper, const Point& start, double tol, int maxfeval,
              int restarts_max, int level)
      : obs_(obs), lo_(lower), up_(upper), tol_(tol), maxfeval_(maxfeval), restarts_max_(restarts_max), level_(level) {
    Point x0 = to_free(start);
    double aux = 0.;
    for (int i = 0; i < N; i++) aux += std::pow(x0[i], 2);
    const double scale = std::max(1.0, std::sqrt(aux));
    const double alpha1 = scale / (N * std::sqrt(2.)) * (std::sqrt(N + 1.) + N - 1.);
    const double alpha2 = scale / (N * std::sqrt(2.)) * (std::sqrt(N + 1.) - 1.);
    p[0] = x0;
    for (int i = 1; i < N + 1; i++) {
      for (int j = 0; j < N; j++) p[i][j] = x0[j] + alpha2;
      p[i][i - 1] = x0[i - 1] + alpha1;
    }
    seq.resize(N + 1);
    std::iota(seq.begin(), seq.end(), 0);
    order = seq;
    ask(std::vector<Point>(p, p + N + 1));  // the N+1 vertices of the initial simplex: one launch
    phase_ = INIT;
  }

  bool done() const { return phase_ == DONE; }
  // points to evaluate next, in the ORIGINAL (bounded) coordinates; empty once done
  const std::vector<Point>& request() const { return req_bounded_; }
  // values of request(), in order
  void supply(const double* vals) {
    res_.evaluated += (int)req_free_.size();
    res_.launches += 1;
    if (phase_ == INIT) {
      for (int i = 0; i < N + 1; i++) { y[i] = vals[i]; consume(req_free_[i], vals[i], true); }
      std::sort(order.begin(), order.end(), [&](int a, int b) { return y[a] < y[b]; });
      refresh();
      double aux = 0.;
      for (int i = 0; i < N; i++) aux += std::pow(sgrad[i], 2);
      alpha3 = 1e-04 * (*std::max_element(diam, diam + N)) / std::sqrt(aux);
      begin_iteration();
      return;
    }
    for (size_t i = 0; i < req_slot_.size(); i++) { val_[req_slot_[i]] = vals[i]; have_[req_slot_[i]] = true; }
    decide();
  }
  const NmkbResult<N>& result() const { return res_; }

 private:
  enum Phase { INIT, ITER, DONE };
  enum { C_STAR = 0, C_EXP = 1, C_OUT = 2, C_IN = 3, C_SHRINK = 4, NC = 4 + N };

  void ask(const std::vector<Point>& free_pts) {
    req_free_ = free_pts;
    req_bounded_.resize(free_pts.size());
    for (size_t i = 0; i < free_pts.size(); i++) req_bounded_[i] = to_bounded(free_pts[i]);
  }
  void ask_slots(const std::vector<int>& slots) {
    req_slot_ = slots;
    std::vector<Point> pts;
    for (int s : slots) pts.push_back(cand_[s]);
    ask(pts);
  }
  void finish() {
    res_.xmin = to_bounded(p[order[0]]);
    res_.ymin = y[order[0]];
    if (y[order[N]] - y[order[0]] <= tol_ || simplex_size <= 1e-06) res_.message = "Successful convergence";
    else if (res_.icount >= maxfeval_) res_.message = "Maximum number of fevals exceeded";
    else if (restarts >= restarts_max_) res_.message = "Stagnation in Nelder-Mead";
    req_free_.clear(); req_bounded_.clear(); req_slot_.clear();
    phase_ = DONE;
  }
  void refresh() {
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < N; j++) v[i][j] = p[order[i + 1]][j] - p[order[0]][j];
      delf[i] = y[order[i + 1]] - y[order[0]];
    }
    double a2 = 0., a3 = 0.;
    for (int i = 0; i < N; i++) {
      double a1 = 0., a4 = 0.;
      for (int j = 0; j < N; j++) {
        a1 += std::pow(v[i][j], 2);
        a2 += std::abs(v[i][j]);
        a4 += v[i][j] * delf[j];
      }
      diam[i] = std::sqrt(a1);
      a3 += std::abs(p[order[0]][i]);
      sgrad[i] = a4;
    }
    simplex_size = a2 / std::max(1.0, a3);
  }
  void begin_iteration() {
    if (!(res_.icount < maxfeval_ && y[order[N]] - y[order[0]] > tol_ && restarts < restarts_max_ && simplex_size > 1e-6)) {
      finish();
      return;
    }
    phase_ = ITER;
    const double rho = 1., gamma = 0.5, chi = 2., sigma = 0.5;
    Point pbar;
    for (int i = 0; i < N; i++) {
      double s = 0.;
      for (int j = 0; j < N + 1; j++) s += p[j][i];
      s -= p[order[N]][i];
      pbar[i] = s / N;
    }
    const Point& worst = p[order[N]];
    for (int i = 0; i < N; i++) {
      cand_[C_STAR][i] = (1. + rho) * pbar[i] - rho * worst[i];
      cand_[C_EXP][i] = (1. + rho * chi) * pbar[i] - rho * chi * worst[i];
      cand_[C_OUT][i] = (1. + rho * gamma) * pbar[i] - rho * gamma * worst[i];
      cand_[C_IN][i] = (1. - gamma) * pbar[i] + gamma * worst[i];
    }
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) cand_[C_SHRINK + i][j] = p[order[0]][j] - sigma * (p[order[i + 1]][j] - p[order[0]][j]);
    for (int k = 0; k < NC; k++) have_[k] = false;
    star_consumed_ = false; extra_consumed_ = false;
    std::vector<int> slots{C_STAR};
    if (level_ >= 1) { slots.push_back(C_EXP); slots.push_back(C_OUT); slots.push_back(C_IN); }
    if (level_ >= 2) for (int i = 0; i < N; i++) slots.push_back(C_SHRINK + i);
    ask_slots(slots);
  }
  // consumes what the serial algorithm would evaluate next; asks for whatever of it has not been evaluated yet
  void decide() {
    const double ystar = val_[C_STAR];
    if (!star_consumed_) { consume(cand_[C_STAR], ystar, true); star_consumed_ = true; }
    int extra = -1;  // the second evaluation of the iteration, if any
    bool reflect_only = false;
    if (y[order[0]] <= ystar && ystar < y[order[N - 1]]) reflect_only = true;
    else if (ystar < y[order[0]]) extra = C_EXP;
    else if (y[order[N - 1]] < ystar && ystar < y[order[N]]) extra = C_OUT;
    else extra = C_IN;
    if (extra >= 0 && !have_[extra]) { ask_slots({extra}); return; }
    bool happy = false;
    Point pnew{};
    double ynew = 0.;
    if (reflect_only) { happy = true; pnew = cand_[C_STAR]; ynew = ystar; }
    else {
      const double y2 = val_[extra];
      if (!extra_consumed_) { consume(cand_[extra], y2, true); extra_consumed_ = true; }
      if (extra == C_EXP) {
        if (y2 < ystar) { pnew = cand_[C_EXP]; ynew = y2; } else { pnew = cand_[C_STAR]; ynew = ystar; }
        happy = true;
      } else if (extra == C_OUT) {
        if (y2 < ystar) { pnew = cand_[C_OUT]; ynew = y2; happy = true; }
      } else {
        if (y2 < y[order[N]]) { pnew = cand_[C_IN]; ynew = y2; happy = true; }
      }
    }
    if (happy) {
      double armtst = 0;
      for (int i = 0; i < N; i++) armtst += std::pow(sgrad[i], 2);
      armtst *= alpha3;
      if ((y[order[N]] - ynew) / (N + 1.) < armtst / N) {
        // stagnation: restart from an axis-aligned simplex around the best vertex ("orth" step)
        restarts++;
        const double diams = *std::min_element(diam, diam + N);
        Point p2[N + 1];
        double y2v[N + 1];
        p2[0] = p[order[0]];
        y2v[0] = y[order[0]];
        for (int i = 0; i < N; i++) {
          p2[i + 1] = p[order[0]];
          p2[i + 1][i] -= diams * sign(sgrad[i]);
          y2v[i + 1] = y[order[i + 1]];
        }
        for (int i = 0; i < N + 1; i++) { p[i] = p2[i]; y[i] = y2v[i]; }
        order = seq;
      }
      p[order[N]] = pnew;
      y[order[N]] = ynew;
    } else if (restarts < restarts_max_) {
      // shrink towards the best vertex; these evaluations are not counted (reference quirk)
      if (!have_[C_SHRINK]) {
        std::vector<int> slots;
        for (int i = 0; i < N; i++) slots.push_back(C_SHRINK + i);
        ask_slots(slots);
        return;
      }
      Point p2[N + 1];
      p2[0] = p[order[0]];
      for (int i = 0; i < N; i++) p2[i + 1] = cand_[C_SHRINK + i];
      const double ybest = y[order[0]];
      for (int i = 0; i < N + 1; i++) p[i] = p2[i];
      y[0] = ybest;
      for (int i = 1; i < N + 1; i++) { y[i] = val_[C_SHRINK + i - 1]; consume(p[i], y[i], false); }
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return y[a] < y[b]; });
    refresh();
    begin_iteration();
  }

  static int sign(double v) { return (0. < v) - (v < 0.); }
  Point to_free(const Point& x) const {
    Point r;
    for (int j = 0; j < N; j++) r[j] = std::atanh(2. * (x[j] - lo_[j]) / (up_[j] - lo_[j]) - 1.);
    return r;
  }
  Point to_bounded(const Point& x) const {
    Point r;
    for (int j = 0; j < N; j++) r[j] = lo_[j] + (up_[j] - lo_[j]) / 2. * (1. + std::tanh(x[j]));
    return r;
  }
  // `counted`: the serial algorithm passes icount+1 to the objective and then increments it; in a shrink step
  // it passes icount+1 without incrementing
  void consume(const Point& free_pt, double yv, bool counted) {
    if (obs_) obs_(to_bounded(free_pt), res_.icount + 1, yv);
    if (counted) res_.icount++;
  }

  Observer obs_;
  Point lo_, up_;
  double tol_;
  int maxfeval_, restarts_max_, level_;
  Phase phase_ = INIT;
  NmkbResult<N> res_;
  // simplex state (names as in the reference header)
  Point p[N + 1];
  double y[N + 1];
  std::vector<int> seq, order;
  Point v[N], sgrad;
  double delf[N], diam[N], simplex_size = 0, alpha3 = 0;
  int restarts = 0;
  // candidates of the current iteration and their values
  Point cand_[NC];
  double val_[NC];
  bool have_[NC];
  bool star_consumed_ = false, extra_consumed_ = false;
  std::vector<Point> req_free_, req_bounded_;
  std::vector<int> req_slot_;
};

template <int N>
class Nmkb {
 public:
  using Point = std::array<double, N>;
  // evaluates f at every point (in the ORIGINAL, bounded coordinates)
  using BatchFn = std::function<void(const std::vector<Point>& x, std::vector<double>& y)>;
  using Observer = typename NmkbStepper<N>::Observer;

  Nmkb(BatchFn f, Observer obs, const Point& lower, const Point& upper) : f_(f), obs_(obs), lo_(lower), up_(upper) {}

  bool speculate = true;   // false: one point per call, the serial order
  int level = 2;           // speculation level when `speculate` (see NmkbStepper)

  NmkbResult<N> minimize(const Point& start, double tol, int maxfeval, int restarts_max) {
    NmkbStepper<N> st(obs_, lo_, up_, start, tol, maxfeval, restarts_max, speculate ? level : 0);
    std::vector<double> y;
    while (!st.done()) {
      const std::vector<Point>& x = st.request();
      y.assign(x.size(), 0.);
      f_(x, y);
      st.supply(y.data());
    }
    return st.result();
  }

 private:
  BatchFn f_;
  Observer obs_;
  Point lo_, up_;
};

}  // namespace ppm
