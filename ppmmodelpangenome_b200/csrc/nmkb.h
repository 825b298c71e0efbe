// Bound-constrained Nelder-Mead ("nmkb" of the R package dfoptim) with a BATCHED objective.
//
// The reference drives its fit with source/nelder_mead_dfoptim.h, a dfoptim-nmkb clone: tanh box transform
// (:33-52), regular initial simplex (:86-105), reflect / expand / contract (:162-222), Armijo-type "orth"
// restart (:224-253), shrink (:254-271), stop on tol / maxfeval / restarts / simplex size (:148-149).
// This is an independent restatement with the same arithmetic (same formulas, same operation order, so the
// same trajectory in IEEE double), built around one difference: the objective is asked for SEVERAL points per
// call.  Every candidate of an iteration (reflection, expansion, outside and inside contraction) is an affine
// function of the centroid and the worst vertex, i.e. known before the reflection value is -- so they are
// evaluated speculatively in ONE GPU launch and consumed in the reference's order.  The logical evaluation
// count (`icount`) and the evaluation callback order are those of the serial algorithm, including its quirk of
// not counting the evaluations of a shrink step (nelder_mead_dfoptim.h:270).
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

template <int N>
class Nmkb {
 public:
  using Point = std::array<double, N>;
  // evaluates f at every point (in the ORIGINAL, bounded coordinates)
  using BatchFn = std::function<void(const std::vector<Point>& x, std::vector<double>& y)>;
  // called once per LOGICAL evaluation, in the serial algorithm's order: (bounded point, count, value)
  using Observer = std::function<void(const Point& x, int count, double y)>;

  Nmkb(BatchFn f, Observer obs, const Point& lower, const Point& upper) : f_(f), obs_(obs), lo_(lower), up_(upper) {}

  bool speculate = true;

  NmkbResult<N> minimize(const Point& start, double tol, int maxfeval, int restarts_max) {
    NmkbResult<N> res;
    res_ = &res;
    Point p[N + 1];
    double y[N + 1];
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
    {  // the N+1 vertices of the initial simplex: one launch
      std::vector<Point> pts(p, p + N + 1);
      std::vector<double> vals = evaluate(pts);
      for (int i = 0; i < N + 1; i++) { y[i] = vals[i]; consume(pts[i], vals[i], true); }
    }
    std::vector<int> seq(N + 1);
    std::iota(seq.begin(), seq.end(), 0);
    std::vector<int> order = seq;
    auto by_value = [&](int a, int b) { return y[a] < y[b]; };
    std::sort(order.begin(), order.end(), by_value);

    const double rho = 1., gamma = 0.5, chi = 2., sigma = 0.5;
    int restarts = 0;
    Point v[N], sgrad;
    double delf[N], diam[N], simplex_size = 0;
    auto refresh = [&]() {
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
    };
    refresh();
    aux = 0.;
    for (int i = 0; i < N; i++) aux += std::pow(sgrad[i], 2);
    const double alpha3 = 1e-04 * (*std::max_element(diam, diam + N)) / std::sqrt(aux);

    while (res.icount < maxfeval && y[order[N]] - y[order[0]] > tol && restarts < restarts_max && simplex_size > 1e-6) {
      bool happy = false;
      Point pbar, pstar, pexp, pout, pin, pnew{};
      double ynew = 0.;
      for (int i = 0; i < N; i++) {
        double s = 0.;
        for (int j = 0; j < N + 1; j++) s += p[j][i];
        s -= p[order[N]][i];
        pbar[i] = s / N;
      }
      const Point& worst = p[order[N]];
      for (int i = 0; i < N; i++) {
        pstar[i] = (1. + rho) * pbar[i] - rho * worst[i];
        pexp[i] = (1. + rho * chi) * pbar[i] - rho * chi * worst[i];
        pout[i] = (1. + rho * gamma) * pbar[i] - rho * gamma * worst[i];
        pin[i] = (1. - gamma) * pbar[i] + gamma * worst[i];
      }
      // one launch for the reflection and the three points that may follow it
      std::vector<Point> cand = speculate ? std::vector<Point>{pstar, pexp, pout, pin} : std::vector<Point>{pstar};
      std::vector<double> cy = evaluate(cand);
      auto value_of = [&](int k, const Point& pt) {
        if (k < (int)cy.size()) return cy[k];
        return evaluate(std::vector<Point>{pt})[0];
      };
      const double ystar = cy[0];
      consume(pstar, ystar, true);
      if (y[order[0]] <= ystar && ystar < y[order[N - 1]]) {
        happy = true; pnew = pstar; ynew = ystar;
      } else if (ystar < y[order[0]]) {
        const double y2 = value_of(1, pexp);
        consume(pexp, y2, true);
        if (y2 < ystar) { pnew = pexp; ynew = y2; } else { pnew = pstar; ynew = ystar; }
        happy = true;
      } else if (y[order[N - 1]] < ystar && ystar < y[order[N]]) {
        const double y2 = value_of(2, pout);
        consume(pout, y2, true);
        if (y2 < ystar) { pnew = pout; ynew = y2; happy = true; }
      } else {
        const double y2 = value_of(3, pin);
        consume(pin, y2, true);
        if (y2 < y[order[N]]) { pnew = pin; ynew = y2; happy = true; }
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
      } else if (restarts < restarts_max) {
        // shrink towards the best vertex; these evaluations are not counted (reference quirk)
        Point p2[N + 1];
        p2[0] = p[order[0]];
        for (int i = 0; i < N; i++)
          for (int j = 0; j < N; j++) p2[i + 1][j] = p[order[0]][j] - sigma * (p[order[i + 1]][j] - p[order[0]][j]);
        const double ybest = y[order[0]];
        for (int i = 0; i < N + 1; i++) p[i] = p2[i];
        y[0] = ybest;
        std::vector<Point> pts(p + 1, p + N + 1);
        std::vector<double> vals = evaluate(pts);
        for (int i = 1; i < N + 1; i++) { y[i] = vals[i - 1]; consume(pts[i - 1], vals[i - 1], false); }
      }
      std::sort(order.begin(), order.end(), by_value);
      refresh();
    }
    res.xmin = to_bounded(p[order[0]]);
    res.ymin = y[order[0]];
    if (y[order[N]] - y[order[0]] <= tol || simplex_size <= 1e-06) res.message = "Successful convergence";
    else if (res.icount >= maxfeval) res.message = "Maximum number of fevals exceeded";
    else if (restarts >= restarts_max) res.message = "Stagnation in Nelder-Mead";
    return res;
  }

 private:
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
  std::vector<double> evaluate(const std::vector<Point>& free_pts) {
    std::vector<Point> b(free_pts.size());
    for (size_t i = 0; i < b.size(); i++) b[i] = to_bounded(free_pts[i]);
    std::vector<double> y(b.size());
    f_(b, y);
    res_->evaluated += (int)b.size();
    res_->launches += 1;
    return y;
  }
  // `counted`: the serial algorithm passes icount+1 to the objective and then increments it; in a shrink step
  // it passes icount+1 without incrementing
  void consume(const Point& free_pt, double y, bool counted) {
    if (obs_) obs_(to_bounded(free_pt), res_->icount + 1, y);
    if (counted) res_->icount++;
  }

  BatchFn f_;
  Observer obs_;
  Point lo_, up_;
  NmkbResult<N>* res_ = nullptr;
};

}  // namespace ppm
