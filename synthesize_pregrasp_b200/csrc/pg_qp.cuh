// Wrench-closure QP of utils/dyn_feasibility_check.py:28-75 for one contact triple:
//     minimise 1/2 z^T Q z   s.t.  G z <= h,      z = 3 contact forces in their frames [n, t0, t1]           (9 variables)
//     Q = W W^T (net force and net torque about the origin), friction pyramid |t0|, |t1| <= -mu f_n, f_n <= -f_min.
// The reference hands it to cvxpy (OSQP).  Here it is solved exactly by an active-set method in GENERATOR space: every
// feasible finger force is a non-negative combination of the four extreme rays r_k = (-1, +-mu, +-mu) of its pyramid whose
// coefficients sum to at least f_min, so with g_ik = W_i^T r_k (12 wrenches in R^6)
//     objective = 1/2 min || sum b_ik g_ik ||^2,   b >= 0,  sum_k b_ik - s_i = f_min,  s >= 0,
// a non-negative least-squares problem with three equality rows, handled by row weighting (Lawson & Hanson, "Solving Least
// Squares Problems", ch. 22-23) and solved by their NNLS algorithm with Householder QR on the passive set.
// Parity note: nothing in the reference pins the QP optimum (no recorded objectives, solver not installable); the tests check
// this solver against an independent SLSQP solve of the reference's own (Q, G, h) formulation.
#pragma once
#include "pg_math.cuh"

namespace pg {
namespace qp {

enum { ROWS = 9, COLS = 15 };

// least squares on the passive columns: min || A[:, idx] s - y || by Householder QR (ROWS x np, np <= ROWS)
PG_HDN void ls_passive(const double (*A)[COLS], const double* y, const int* idx, int np, double* s) {
  double R[ROWS][ROWS], q[ROWS];
  for (int r = 0; r < ROWS; r++) { q[r] = y[r]; for (int c = 0; c < np; c++) R[r][c] = A[r][idx[c]]; }
  for (int c = 0; c < np; c++) {
    double nrm = 0;
    for (int r = c; r < ROWS; r++) nrm += R[r][c] * R[r][c];
    nrm = sqrt(nrm);
    if (nrm == 0) continue;
    double alpha = R[c][c] > 0 ? -nrm : nrm;
    double v[ROWS], vn = 0;
    for (int r = c; r < ROWS; r++) { v[r] = R[r][c]; }
    v[c] -= alpha;
    for (int r = c; r < ROWS; r++) vn += v[r] * v[r];
    if (vn == 0) continue;
    for (int cc = c; cc < np; cc++) {
      double d = 0;
      for (int r = c; r < ROWS; r++) d += v[r] * R[r][cc];
      d = 2 * d / vn;
      for (int r = c; r < ROWS; r++) R[r][cc] -= d * v[r];
    }
    double d = 0;
    for (int r = c; r < ROWS; r++) d += v[r] * q[r];
    d = 2 * d / vn;
    for (int r = c; r < ROWS; r++) q[r] -= d * v[r];
  }
  // the four rays of one pyramid are linearly dependent (r1 + r4 = r2 + r3): a passive set holding all of them is rank
  // deficient; a negligible pivot means "this column adds nothing": basic solution with that coefficient at zero
  double dmax = 0;
  for (int c = 0; c < np; c++) if (fabs(R[c][c]) > dmax) dmax = fabs(R[c][c]);
  for (int c = np - 1; c >= 0; c--) {
    double t = q[c];
    for (int cc = c + 1; cc < np; cc++) t -= R[c][cc] * s[cc];
    s[c] = (fabs(R[c][c]) > 1e-11 * dmax) ? t / R[c][c] : 0.0;
  }
}

// Lawson-Hanson NNLS: min || A x - y ||, x >= 0.  Returns the number of outer iterations (negative: iteration cap hit).
PG_HDN int nnls(const double (*A)[COLS], const double* y, double* x) {
  bool passive[COLS];
  for (int j = 0; j < COLS; j++) { passive[j] = false; x[j] = 0; }
  double norm1 = 0;                                            // ||A||_1: the scale of the tolerances
  for (int j = 0; j < COLS; j++) { double c = 0; for (int r = 0; r < ROWS; r++) c += fabs(A[r][j]); if (c > norm1) norm1 = c; }
  // dual-feasibility tolerance: the gradient components a_j . r carry rounding noise of order eps * |a_j| * |A| |x| ~ 3e-8 with
  // the 1e4 row weights, and a column that is only noise-violated is (nearly) dependent on the passive ones
  const double tol = 2.5e-11 * norm1;
  int outer = 0, inner_total = 0;
  for (; outer < 3 * COLS; outer++) {
    double resid[ROWS];
    for (int r = 0; r < ROWS; r++) { double t = y[r]; for (int j = 0; j < COLS; j++) t -= A[r][j] * x[j]; resid[r] = t; }
    int best = -1, np = 0, idx[COLS]; double wbest = tol;
    for (int j = 0; j < COLS; j++) {
      if (passive[j]) { idx[np++] = j; continue; }
      double w = 0;
      for (int r = 0; r < ROWS; r++) w += A[r][j] * resid[r];
      if (w > wbest) { wbest = w; best = j; }
    }
    if (best < 0 || np >= ROWS) return outer;                   // dual feasible (or the passive columns already span the rows)
    passive[best] = true;
    double s[COLS];
    for (;;) {
      np = 0;
      for (int j = 0; j < COLS; j++) if (passive[j]) idx[np++] = j;
      double sp[ROWS];
      ls_passive(A, y, idx, np, sp);
      for (int j = 0; j < COLS; j++) s[j] = 0;
      double smin = 0;
      for (int c = 0; c < np; c++) { s[idx[c]] = sp[c]; if (sp[c] < smin) smin = sp[c]; }
      if (!(smin < 0) || inner_total >= 3 * COLS) break;
      inner_total++;
      double alpha = 1.0;
      for (int c = 0; c < np; c++) { int j = idx[c]; if (s[j] < 0) { double a = x[j] / (x[j] - s[j]); if (a < alpha) alpha = a; } }
      for (int j = 0; j < COLS; j++) x[j] = x[j] * (1.0 - alpha) + alpha * s[j];
      for (int c = 0; c < np; c++) if (x[idx[c]] <= tol) passive[idx[c]] = false;
    }
    for (int j = 0; j < COLS; j++) x[j] = s[j] > 0 ? s[j] : 0.0;
  }
  return -outer;
}

// objective of the wrench-closure QP for contact points p[3], raw normals n[3] (used as given for the frames, as
// find_dynamically_feasible_contacts :264-272 does; callers normalise first where the reference does, :94)
PG_HDN double wrench_closure_objective(const V3* p, const V3* n, double mu, double f_min, int* iters_out) {
  const double WEIGHT = 1e4;                 // sqrt of the penalty on the three equality rows
  double A[ROWS][COLS], y[ROWS], x[COLS];
  for (int r = 0; r < ROWS; r++) { y[r] = 0; for (int c = 0; c < COLS; c++) A[r][c] = 0; }
  for (int i = 0; i < 3; i++) {
    V3 nn = n[i], t0(nn.y, -nn.x, 0.0), t1 = cross(nn, t0);
    int k = 0;
    for (int s1 = 1; s1 >= -1; s1 -= 2) for (int s2 = 1; s2 >= -1; s2 -= 2, k++) {
      V3 f = nn * (-1.0) + t0 * (s1 * mu) + t1 * (s2 * mu);      // world force of the extreme ray
      V3 tq = cross(p[i], f);
      const int c = 4 * i + k;
      A[0][c] = f.x; A[1][c] = f.y; A[2][c] = f.z; A[3][c] = tq.x; A[4][c] = tq.y; A[5][c] = tq.z;
      A[6 + i][c] = WEIGHT;
    }
    A[6 + i][12 + i] = -WEIGHT;
    y[6 + i] = WEIGHT * f_min;
  }
  int it = nnls(A, y, x);
  if (iters_out) *iters_out = it;
  double w[6] = {0, 0, 0, 0, 0, 0};
  for (int c = 0; c < 12; c++) for (int r = 0; r < 6; r++) w[r] += A[r][c] * x[c];
  double o = 0;
  for (int r = 0; r < 6; r++) o += w[r] * w[r];
  return 0.5 * o;
}

}  // namespace qp
}  // namespace pg
