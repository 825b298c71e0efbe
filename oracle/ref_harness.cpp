// TEST INFRASTRUCTURE ONLY (never linked into the product).
// White-box C-ABI harness around the UNMODIFIED reference sources, which are compiled from
// /root/reference/source by oracle/Makefile into oracle/_ref/libppm_ref.so. Used (1) to pin the
// C restatement in oracle/ppm_oracle.c, (2) to generate tests/golden/*.json (oracle/gen_golden.py),
// (3) as the CPU baseline of bench.py (cpu_baseline.kind == "reference").
#include <vector>
#include <map>
#include <string>
#include <array>
#include <functional>
#include <cmath>
#include <cstring>
#include <climits>
#include <algorithm>
#include <numeric>
#include <sstream>
#include <fstream>
#include <iostream>
#include <cstdio>
#include <omp.h>
#define private public
#include "functions.hpp"
#undef private

struct RefHandle {
  RecTree* tree;
  Inference* inf;
};

static std::vector<double> svec(const double* p) { return {p[6], p[6], p[6], p[7]}; }

extern "C" {

void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }

void* ref_create(const char* tree_path, const char* pa_path, int compress) {
  std::ifstream f(tree_path);
  std::string line;
  std::getline(f, line);
  if (line.empty()) return nullptr;
  RefHandle* h = new RefHandle;
  h->tree = new RecTree(line);
  PAmatrix mat{std::string(pa_path)};
  if (compress) mat.compress();
  h->inf = new Inference(*h->tree, mat);
  return h;
}
void ref_destroy(void* hv) {
  RefHandle* h = (RefHandle*)hv;
  delete h->inf;
  delete h->tree;
  delete h;
}
// out_i: nRep nObs nNodes nLeaves ; out_d: H L meanGenes
void ref_dims(void* hv, int* out_i, double* out_d) {
  Inference* inf = ((RefHandle*)hv)->inf;
  out_i[0] = inf->nRep; out_i[1] = inf->nObs; out_i[2] = inf->nNodes; out_i[3] = inf->nLeaves;
  out_d[0] = inf->H; out_d[1] = inf->L; out_d[2] = inf->meanGenes;
}
// which: 0 mrca[nRep] 1 freq[nRep] 2 order[nNodes] 3 counts[nRep] 4 left child[nNodes](-1 leaf) 5 right child
void ref_get_int(void* hv, int which, int* out) {
  Inference* inf = ((RefHandle*)hv)->inf;
  switch (which) {
    case 0: std::copy(inf->mrca.begin(), inf->mrca.end(), out); break;
    case 1: std::copy(inf->freq.begin(), inf->freq.end(), out); break;
    case 2: std::copy(inf->order.begin(), inf->order.end(), out); break;
    case 3: std::copy(inf->mat.getCounts().begin(), inf->mat.getCounts().begin() + inf->nRep, out); break;
    case 4: case 5:
      for (int i = 0; i < inf->nNodes; i++) {
        LikFunction& lf = inf->tLik.getLikFunctions()[i];
        out[i] = lf.isLeaf() ? -1 : (which == 4 ? lf.getChildren().first : lf.getChildren().second);
      }
      break;
  }
}
// which: 0 height[nNodes] 1 branchLength[nNodes]
void ref_get_double(void* hv, int which, double* out) {
  Inference* inf = ((RefHandle*)hv)->inf;
  if (which == 0) std::copy(inf->height.begin(), inf->height.end(), out);
  else std::copy(inf->branchLength.begin(), inf->branchLength.end(), out);
}
// leaf observation of pattern rep at preorder node id (must be a leaf)
void ref_get_pattern(void* hv, int rep, int* out /*[nNodes], -1 for internal*/) {
  Inference* inf = ((RefHandle*)hv)->inf;
  for (int i = 0; i < inf->nNodes; i++) {
    LikFunction& lf = inf->tLik.getLikFunctions()[i];
    out[i] = lf.isLeaf() ? lf.getLeafValue(rep) : -1;
  }
}
int ref_subtrees(void* hv, int rank, int* out) {
  Inference* inf = ((RefHandle*)hv)->inf;
  const std::vector<int>& v = inf->subtrees[rank];
  std::copy(v.begin(), v.end(), out);
  return (int)v.size();
}
// exactly what the objective lambda does (functions.cpp:406-409)
double ref_loglik(void* hv, const double* p) {
  Inference* inf = ((RefHandle*)hv)->inf;
  inf->tLik.clearLikValues();
  return inf->logLik(p[0], p[1], p[2], p[3], p[4], p[5], svec(p));
}
double ref_compute_i2(void* hv, const double* p, double* p0p1AB) {
  Inference* inf = ((RefHandle*)hv)->inf;
  inf->tLik.clearLikValues();
  double i2 = inf->computeI2(p[0], p[1], p[2], p[3], p[4], p[5], svec(p));
  if (p0p1AB) {
    std::vector<double> s = svec(p);
    p0p1AB[0] = (p[1] == 0.) ? 1 - pow(s[0], inf->nLeaves)
                             : 1 - exp(inf->tLik.evaluateRootRep(0, inf->nRep, {0, 1}, 0., p[1], 0., s[0], 0));
    p0p1AB[1] = 1 - exp(inf->tLik.evaluateRootRep(0, inf->nRep, {0, 1}, 0., p[3], 0., s[1], 1));
    p0p1AB[2] = inf->getPobs1(p[3], s[1]);
    p0p1AB[3] = inf->getPobs2(p[4], p[5], s[3], s[2]);
  }
  return i2;
}
// per pattern: lik0r lik1r lik1 lik2 (the four log terms of likRepTot, functions.cpp:345-358), fresh memo
void ref_components(void* hv, const double* p, double* out /*[nRep*4]*/) {
  Inference* inf = ((RefHandle*)hv)->inf;
  inf->tLik.clearLikValues();
  std::vector<double> s = svec(p);
  double N0 = p[0], l0 = p[1], i1 = p[2], l1 = p[3], g2 = p[4], l2 = p[5];
  double i2 = inf->computeI2(N0, l0, i1, l1, g2, l2, s);
  int nRep = inf->nRep, nObs = inf->nObs, nLeaves = inf->nLeaves;
#pragma omp parallel for schedule(dynamic, 8)
  for (int rep = 0; rep < nRep; rep++) {
    double lik0r = log(N0 / nObs);
    if (l0 == 0. && (s[0] > 0. || inf->freq[rep] < nLeaves))
      lik0r += inf->freq[rep] * log(1 - s[0]) + (nLeaves - inf->freq[rep]) * log(s[0]);
    else if (l0 > 0.)
      lik0r += inf->tLik.evaluateRootRep(0, rep, {0, 1}, 0., l0, 0., s[0], 0);
    out[4 * rep + 0] = lik0r;
    out[4 * rep + 1] = log(i1 / (l1 * nObs)) + inf->tLik.evaluateRootRep(0, rep, {0, 1}, 0., l1, 0., s[1], 1);
    out[4 * rep + 2] = log(i1 / (l1 * nObs)) + inf->likRep1(rep, l1, s[1]);
    out[4 * rep + 3] = log(i2 / nObs) + inf->likRep2(rep, g2, l2, s[3], s[2]);
  }
}
// categories with a FRESH memo at params p and the given i2 (i.e. without the stale-memo quirk):
// runs the reference's own assignCat (functions.cpp:489-520) into a temp file and parses it back.
int ref_assign_cat(void* hv, const double* p, double i2, int* out, const char* tmpfile) {
  Inference* inf = ((RefHandle*)hv)->inf;
  inf->tLik.clearLikValues();
  inf->MLEN0 = p[0]; inf->MLEl0 = p[1]; inf->MLEi1 = p[2]; inf->MLEl1 = p[3];
  inf->MLEg2 = p[4]; inf->MLEl2 = p[5]; inf->MLEs = svec(p); inf->MLEi2 = i2;
  // zero row must be evaluated first, as logLik does through computeI2
  inf->computeI2(p[0], p[1], p[2], p[3], p[4], p[5], svec(p));
  inf->assignCat(std::string(tmpfile));
  std::ifstream f(tmpfile);
  int n = 0; double v;
  while (f >> v) out[n++] = (int)v;
  std::remove(tmpfile);
  return n;
}

// computeProbaCat (functions.cpp:556-581) with a fresh memo, parsed back from the file the reference writes
int ref_proba_cat(void* hv, const double* p, double i2, double* out /*[nRep*3]*/, const char* tmpfile) {
  Inference* inf = ((RefHandle*)hv)->inf;
  inf->tLik.clearLikValues();
  inf->computeI2(p[0], p[1], p[2], p[3], p[4], p[5], svec(p));
  inf->computeProbaCat(std::string(tmpfile), p[0], p[1], p[2], p[3], i2, p[4], p[5], p[7], p[6]);
  std::ifstream f(tmpfile);
  int n = 0; std::string tok;
  while (f >> tok) out[n++] = std::stod(tok) ;
  std::remove(tmpfile);
  return n / 3;
}

// expectedTime1 (functions.cpp:637-690) per pattern, fresh memo (the zero row is evaluated first, serially)
void ref_expected_time1(void* hv, const double* p, double* out /*[nRep]*/) {
  Inference* inf = ((RefHandle*)hv)->inf;
  inf->tLik.clearLikValues();
  inf->computeI2(p[0], p[1], p[2], p[3], p[4], p[5], svec(p));
  for (int rep = 0; rep < inf->nRep; rep++) out[rep] = inf->expectedTime1(rep, p[3], p[6]);
}
// expectedTime2_times (:709-725) and, per pattern, expectedTime2_aux (:727-757) followed by exp(likRep2) as
// expectedTime2_store appends it (:756).  Returns the number of slices S; dist is [nRep][S+1] (may be NULL: count only)
int ref_expected_time2(void* hv, const double* p, double* times, double* dist) {
  Inference* inf = ((RefHandle*)hv)->inf;
  std::vector<double> tt = inf->expectedTime2_times();
  const int S = (int)tt.size();
  if (times) std::copy(tt.begin(), tt.end(), times);
  if (!dist) return S;
  inf->tLik.clearLikValues();
  inf->computeI2(p[0], p[1], p[2], p[3], p[4], p[5], svec(p));
#pragma omp parallel for schedule(dynamic, 8)
  for (int rep = 0; rep < inf->nRep; rep++) {
    std::vector<double> f = inf->expectedTime2_aux(rep, p[4], p[5], p[7], p[6]);
    std::copy(f.begin(), f.end(), dist + (size_t)rep * (S + 1));
    dist[(size_t)rep * (S + 1) + S] = exp(inf->likRep2(rep, p[4], p[5], p[7], p[6]));
  }
  return S;
}
double ref_expected_gains(void* hv, double g2) { return ((RefHandle*)hv)->inf->expectedGains(g2); }

}  // extern "C"
