// Batched rigid-body stepping: one rollout per thread, all state in registers / thread-local memory.
//
// Replaces, for K rollouts at once, what the reference obtains from one PyBullet world per CPU process:
//   envs/plate_contact_graph_env.py:176-250 (reset), :454-482 (pre_simulate), :504-533 (servo loop),
//   :556-603 (settle) -> pybullet.stepSimulation / createConstraint / changeConstraint /
//   resetBasePositionAndOrientation / setCollisionFilterPair.
// The stepping scheme is the one the recorded reference trajectories pin (DESIGN.md "Physics"):
// gravity + proportional damping + gyroscopic term, box-box SAT narrowphase with persistent 4-point
// manifolds, fixed-iteration projected Gauss-Seidel (fingertip servo rows in alternating order, normal
// rows, paired cone-friction rows, residual early-out), semi-implicit Euler with exponential-map
// orientation update, dt = 1/250 s.  fp64 throughout: contact-rich rollouts amplify rounding.
// The box-box detector (box_box, rect_quad, cull_points, line_closest), the persistent-manifold rules (cache_entry, sort_cached,
// refresh) and the solver row order are written after bullet3's published btBoxBoxDetector / btPersistentManifold /
// btSequentialImpulseConstraintSolver (zlib licence, (c) Erwin Coumans et al.; btBoxBoxDetector after ODE's dBoxBox2, (c) Russell
// L. Smith): bit-level parity with that engine requires its operation order.
#pragma once
#include "pg_env.cuh"
#include "pg_convex.cuh"
#ifndef PG_BLOCK_THREADS
#define PG_BLOCK_THREADS 64
#endif

namespace pg {

#if defined(__CUDACC__)
__shared__ double pg_hull_sm[MAX_HULL * 3];       // hull vertices of the env, staged once per CTA (hull kernel only)
#endif
enum { B_OBJ = MAX_STATICS, B_TIP0 = MAX_STATICS + 1, B_TIP1 = MAX_STATICS + 2, NBODY = MAX_STATICS + 3 };
enum { MAXM = 12, MAXPTS = 24 };

#define PG_EPS 2.220446049250313e-16
#define PG_LARGE 1e30

struct DynBody {
  V3 pos, v, w;
  Q4 q;
  M3 R;
  V3 inertia, half;
  double mass, friction;
};

struct MPoint {
  V3 localA, localB, normalB;
  double dist;
};

struct Manifold {
  MPoint p[4];
  double breaking;
  short a, b, n, pad;
};

struct FixedCons {
  V3 pivotB;
  M3 frameB;
  double max_impulse;
  int active;
};

struct Sim {
  DynBody dyn[3];     // obj, tip0, tip1
  Manifold man[MAXM];
  int nman;
  // The dispatcher's manifold array as Bullet holds it (append on creation, swap-with-last on release): entries < MAXM index
  // `man`, entries >= DISP_IDLE are the three manifolds between the env's IDLE fingertips (see idle_pair_alive).
  unsigned char disp[MAXM + 3];
  int ndisp;
  int step;              // world steps since reset() (the stepSimulation of reset is step 1)
  FixedCons cons[2];
  unsigned filter_off;   // bit i: tip i <-> object collision disabled
  unsigned early;        // bit i: the re-created object proxy found tip i at creation (tight AABBs overlapping)
  V3 plo[2], phi[2];     // tight AABB each tip's broadphase proxy was last (re)created with
  int obj_first;         // reset substep: object proxy is older than the walls'
  int last_iterations;
  int sync_cta;          // hull kernel: all threads of the CTA run rollouts, so substeps may start on a CTA barrier
#if defined(PG_PHASE_CLOCKS)
  long long clk[6];      // profiling builds: cycles in collide / external / solver set-up / sweeps / integrate / rest
#endif
};
// profiling builds (-DPG_PHASE_CLOCKS, csrc/build_variant.sh + tools/gpu_phase_clocks.py): where a rollout's cycles go
#if defined(PG_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
#define PG_CLK(S, slot, t0) { long long _t = clock64(); (S).clk[slot] += _t - (t0); (t0) = _t; }
#define PG_CLK0(t0) long long t0 = clock64()
#define PG_CLK_RESET(t0) t0 = clock64()
#else
#define PG_CLK(S, slot, t0)
#define PG_CLK0(t0)
#define PG_CLK_RESET(t0)
#endif

// -------------------------------------------------------------------------------------------------
PG_HD bool is_static(int b) { return b < MAX_STATICS; }

// The env creates five fingertip bodies (plate_env:215-228, `all_fingers` = 0..4) and only ever moves the two active ones: fingers
// 2-4 are all created at (100, 100, 100), collide with each other, are pushed apart (0.4 m/s) and fall freely.  They touch nothing
// else, but their three manifolds are the OLDEST entries of the dispatcher's array and take part in the unstable quickSort by
// island id that fixes the solver's manifold order, until their fat AABBs separate: pair (2,4) during world step 20, pairs (2,3)
// and (3,4) during world step 39 -- the same for every env and every rollout (fingertip x half-extent 0.01, mass 0.01, dt 1/250,
// contact erp 0.08; oracle/envsim.py simulates the three bodies for real, tests/test_oracle_cpu.py pins these two numbers).
// When they go, the last manifold of the array moves into the freed slot and the sorted order of everybody else's manifolds flips:
// found from keyboard_20.0 (row 37 = world step 39), confirmed by cardboard_20.0 (all 100 rows at 3e-12 with it).
enum { DISP_IDLE = 100, IDLE_DROP_STEP_OUTER = 20, IDLE_DROP_STEP_INNER = 39, IDLE_ISLAND_KEY = 5 };
PG_HD bool idle_pair_alive(int entry, int step) {    // entry DISP_IDLE + 0: fingers (2,3), + 1: (2,4), + 2: (3,4)
  return step < (entry == DISP_IDLE + 1 ? IDLE_DROP_STEP_OUTER : IDLE_DROP_STEP_INNER);
}
// release of the dispatcher entry at position p (btCollisionDispatcher::releaseManifold: swap with the last entry, pop); a real
// manifold also frees its storage slot (the last storage slot moves into it)
PG_HD void disp_release_at(Sim& S, int p) {
  const int e = S.disp[p];
  S.disp[p] = S.disp[S.ndisp - 1];
  S.ndisp--;
  if (e < DISP_IDLE) {
    const int last = S.nman - 1;
    if (e != last) {
      S.man[e] = S.man[last];
      for (int q = 0; q < S.ndisp; q++) if (S.disp[q] == last) { S.disp[q] = (unsigned char)e; break; }
    }
    S.nman--;
  }
}

struct BodyView { V3 pos; M3 R; V3 half; double friction; };

PG_HD BodyView view(const EnvDev& E, const Sim& S, const M3* statR, int b) {
  BodyView v;
  if (is_static(b)) {
    const BoxBody& s = E.statics[b];
    v.pos = V3(s.pos[0], s.pos[1], s.pos[2]);
    v.R = statR[b];
    v.half = V3(s.half[0], s.half[1], s.half[2]);
    v.friction = s.friction;
  } else {
    const DynBody& d = S.dyn[b - MAX_STATICS];
    v.pos = d.pos; v.R = d.R; v.half = d.half; v.friction = d.friction;
  }
  return v;
}

PG_HD void aabb_of(const BodyView& b, double extra, V3& lo, V3& hi) {
  V3 we;
  for (int i = 0; i < 3; i++)
    we[i] = fabs(b.R(i, 0)) * b.half.x + fabs(b.R(i, 1)) * b.half.y + fabs(b.R(i, 2)) * b.half.z;
  lo = b.pos - we - V3(extra, extra, extra);
  hi = b.pos + we + V3(extra, extra, extra);
}

PG_HD bool overlap(V3 alo, V3 ahi, V3 blo, V3 bhi) {
  for (int k = 0; k < 3; k++) if (alo[k] > bhi[k] || ahi[k] < blo[k]) return false;
  return true;
}

// ---------------------------------------------------------------- persistent manifold ---------------
PG_HD int cache_entry(const Manifold& m, V3 localA) {
  double shortest = m.breaking * m.breaking;
  int nearest = -1;
  for (int i = 0; i < m.n; i++) {
    V3 d = m.p[i].localA - localA;
    double dd = dot(d, d);
    if (dd < shortest) { shortest = dd; nearest = i; }
  }
  return nearest;
}

PG_HD int sort_cached(const Manifold& m, V3 localA, double dist) {
  int maxPenIdx = -1;
  double maxPen = dist;
  for (int i = 0; i < 4; i++) if (m.p[i].dist < maxPen) { maxPenIdx = i; maxPen = m.p[i].dist; }
  double res[4] = {0, 0, 0, 0};
  if (maxPenIdx != 0) { V3 c = cross(localA - m.p[1].localA, m.p[3].localA - m.p[2].localA); res[0] = dot(c, c); }
  if (maxPenIdx != 1) { V3 c = cross(localA - m.p[0].localA, m.p[3].localA - m.p[2].localA); res[1] = dot(c, c); }
  if (maxPenIdx != 2) { V3 c = cross(localA - m.p[0].localA, m.p[3].localA - m.p[1].localA); res[2] = dot(c, c); }
  if (maxPenIdx != 3) { V3 c = cross(localA - m.p[0].localA, m.p[2].localA - m.p[1].localA); res[3] = dot(c, c); }
  int best = -1; double bv = -PG_LARGE;
  for (int i = 0; i < 4; i++) if (fabs(res[i]) > bv) { best = i; bv = fabs(res[i]); }
  return best;
}

struct Sink { Manifold* m; V3 posA, posB; M3 RA, RB; };

PG_HD void add_contact(Sink& s, V3 normalOnB, V3 pointOnB, double depth) {
  Manifold& m = *s.m;
  if (depth > m.breaking) return;
  V3 pointA = pointOnB + normalOnB * depth;
  MPoint np;
  np.localA = mulT(s.RA, pointA - s.posA);
  np.localB = mulT(s.RB, pointOnB - s.posB);
  np.normalB = normalOnB; np.dist = depth;
  int idx = cache_entry(m, np.localA);
  if (idx < 0) {
    idx = m.n;
    if (idx == 4) idx = sort_cached(m, np.localA, depth); else m.n++;
    if (idx < 0) idx = 0;
  }
  m.p[idx] = np;
}

PG_HD void refresh(Manifold& m, const BodyView& A, const BodyView& B) {
  V3 pa[4], pb[4];
  for (int i = m.n - 1; i >= 0; i--) {
    pa[i] = mul(A.R, m.p[i].localA) + A.pos;
    pb[i] = mul(B.R, m.p[i].localB) + B.pos;
    m.p[i].dist = dot(pa[i] - pb[i], m.p[i].normalB);
  }
  for (int i = m.n - 1; i >= 0; i--) {
    bool rm = false;
    if (!(m.p[i].dist <= m.breaking)) rm = true;
    else {
      V3 proj = pa[i] - m.p[i].normalB * m.p[i].dist;
      V3 d = pb[i] - proj;
      if (dot(d, d) > m.breaking * m.breaking) rm = true;
    }
    if (rm) {
      int last = m.n - 1;
      if (i != last) { m.p[i] = m.p[last]; pa[i] = pa[last]; pb[i] = pb[last]; }
      m.n--;
    }
  }
}

// ---------------------------------------------------------------- box-box narrowphase ----------------
PG_HD void line_closest(V3 pa, V3 ua, V3 pb, V3 ub, double* alpha, double* beta) {
  V3 p = pb - pa;
  double uaub = dot(ua, ub), q1 = dot(ua, p), q2 = -dot(ub, p);
  double d = 1 - uaub * uaub;
  if (d <= (double)0.0001f) { *alpha = 0; *beta = 0; }
  else { d = 1.0 / d; *alpha = (q1 + uaub * q2) * d; *beta = (uaub * q1 + q2) * d; }
}

PG_HDN int rect_quad(const double h[2], const double p[8], double ret[16]) {
  int nq = 4, nr = 0;
  double bufA[16], bufB[16];
  for (int i = 0; i < 8; i++) bufA[i] = p[i];
  double* q = bufA;
  double* r = bufB;
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      const double* pq = q;
      double* pr = r;
      nr = 0;
      bool full = false;
      for (int i = nq; i > 0 && !full; i--) {
        if (sign * pq[dir] < h[dir]) {
          pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++;
          if (nr & 8) { full = true; break; }
        }
        const double* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          pr[dir] = sign * h[dir];
          pr += 2; nr++;
          if (nr & 8) { full = true; break; }
        }
        pq += 2;
      }
      double* t = q; q = r; r = t;
      if (full) { dir = 2; break; }
      nq = nr;
    }
  }
  for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
  return nr;
}

PG_HDN void cull_points(int n, const double p[], int m, int i0, int iret[]) {
  const double MPI = (double)3.14159265f;
  double a, cx, cy, q;
  if (n == 1) { cx = p[0]; cy = p[1]; }
  else if (n == 2) { cx = 0.5 * (p[0] + p[2]); cy = 0.5 * (p[1] + p[3]); }
  else {
    a = 0; cx = 0; cy = 0;
    for (int i = 0; i < (n - 1); i++) {
      q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
      a += q;
      cx += q * (p[i * 2] + p[i * 2 + 2]);
      cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
    }
    q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
    if (fabs(a + q) > PG_EPS) a = 1.0 / (3.0 * (a + q)); else a = PG_LARGE;
    cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
    cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
  }
  double A[8];
  int avail[8];
  for (int i = 0; i < n; i++) { A[i] = atan2(p[i * 2 + 1] - cy, p[i * 2] - cx); avail[i] = 1; }
  avail[i0] = 0;
  iret[0] = i0;
  int k = 1;
  for (int j = 1; j < m; j++) {
    a = double(j) * (2 * MPI / m) + A[i0];
    if (a > MPI) a -= 2 * MPI;
    double maxdiff = 1e9, diff;
    iret[k] = i0;
    for (int i = 0; i < n; i++) {
      if (avail[i]) {
        diff = fabs(A[i] - a);
        if (diff > MPI) diff = 2 * MPI - diff;
        if (diff < maxdiff) { maxdiff = diff; iret[k] = i; }
      }
    }
    avail[iret[k]] = 0;
    k++;
  }
}

// btBoxBoxDetector (ODE dBoxBox2).  Box 1 = manifold body A, box 2 = body B.
PG_HDN void box_box(const BodyView& b1, const BodyView& b2, Sink& out) {
  const double fudge_factor = 1.05;
  const M3& R1 = b1.R;
  const M3& R2 = b2.R;
  V3 p = b2.pos - b1.pos, pp = mulT(R1, p), normalC;
  double A[3], B[3], R[3][3], Q[3][3], s, s2, l;
  int nbox = 0, ncol = -1, invert_normal = 0, code = 0;
  for (int i = 0; i < 3; i++) { A[i] = (2.0 * b1.half[i]) * 0.5; B[i] = (2.0 * b2.half[i]) * 0.5; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R[i][j] = dot(R1.col(i), R2.col(j)); Q[i][j] = fabs(R[i][j]); }
  s = -INFINITY;
#define PG_TST1(expr1, expr2, box, col, cc) \
  { double e1 = (expr1); s2 = fabs(e1) - (expr2); if (s2 > 0) return; \
    if (s2 > s) { s = s2; nbox = box; ncol = col; invert_normal = (e1 < 0); code = (cc); } }
  PG_TST1(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), 1, 0, 1);
  PG_TST1(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), 1, 1, 2);
  PG_TST1(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), 1, 2, 3);
  PG_TST1(dot(R2.col(0), p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), 2, 0, 4);
  PG_TST1(dot(R2.col(1), p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), 2, 1, 5);
  PG_TST1(dot(R2.col(2), p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), 2, 2, 6);
#undef PG_TST1
#define PG_TST2(expr1, expr2, n1, n2, n3, cc) \
  { double e1 = (expr1); s2 = fabs(e1) - (expr2); if (s2 > PG_EPS) return; \
    l = sqrt((n1) * (n1) + (n2) * (n2) + (n3) * (n3)); \
    if (l > PG_EPS) { s2 /= l; if (s2 * fudge_factor > s) { s = s2; ncol = -1; \
      normalC = V3((n1) / l, (n2) / l, (n3) / l); invert_normal = (e1 < 0); code = (cc); } } }
  const double fudge2 = (double)1.0e-5f;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] += fudge2;
  PG_TST2(pp[2] * R[1][0] - pp[1] * R[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -R[2][0], R[1][0], 7);
  PG_TST2(pp[2] * R[1][1] - pp[1] * R[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -R[2][1], R[1][1], 8);
  PG_TST2(pp[2] * R[1][2] - pp[1] * R[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -R[2][2], R[1][2], 9);
  PG_TST2(pp[0] * R[2][0] - pp[2] * R[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), R[2][0], 0, -R[0][0], 10);
  PG_TST2(pp[0] * R[2][1] - pp[2] * R[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), R[2][1], 0, -R[0][1], 11);
  PG_TST2(pp[0] * R[2][2] - pp[2] * R[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), R[2][2], 0, -R[0][2], 12);
  PG_TST2(pp[1] * R[0][0] - pp[0] * R[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -R[1][0], R[0][0], 0, 13);
  PG_TST2(pp[1] * R[0][1] - pp[0] * R[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -R[1][1], R[0][1], 0, 14);
  PG_TST2(pp[1] * R[0][2] - pp[0] * R[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -R[1][2], R[0][2], 0, 15);
#undef PG_TST2
  if (!code) return;
  V3 normal;
  if (ncol >= 0) normal = (nbox == 1 ? R1 : R2).col(ncol);
  else normal = mul(R1, normalC);
  if (invert_normal) normal = -normal;
  double depth = -s;

  if (code > 6) {
    V3 pa = b1.pos;
    for (int j = 0; j < 3; j++) {
      double sign = (dot(normal, R1.col(j)) > 0) ? 1.0 : -1.0;
      pa += R1.col(j) * (sign * A[j]);
    }
    V3 pb = b2.pos;
    for (int j = 0; j < 3; j++) {
      double sign = (dot(normal, R2.col(j)) > 0) ? -1.0 : 1.0;
      pb += R2.col(j) * (sign * B[j]);
    }
    double alpha, beta;
    V3 ua = R1.col((code - 7) / 3), ub = R2.col((code - 7) % 3);
    line_closest(pa, ua, pb, ub, &alpha, &beta);
    pb += ub * beta;
    add_contact(out, -normal, pb, -depth);
    return;
  }

  const M3& Ra = (code <= 3) ? R1 : R2;
  const M3& Rb = (code <= 3) ? R2 : R1;
  V3 pa = (code <= 3) ? b1.pos : b2.pos, pb = (code <= 3) ? b2.pos : b1.pos;
  const double* Sa = (code <= 3) ? A : B;
  const double* Sb = (code <= 3) ? B : A;
  V3 normal2 = (code <= 3) ? normal : -normal;
  V3 nr = mulT(Rb, normal2), anr(fabs(nr.x), fabs(nr.y), fabs(nr.z));
  int lanr, a1, a2;
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  V3 center;
  if (nr[lanr] < 0) center = pb - pa + Rb.col(lanr) * Sb[lanr];
  else center = pb - pa - Rb.col(lanr) * Sb[lanr];
  int codeN = (code <= 3) ? code - 1 : code - 4, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double quad[8], c1, c2, m11, m12, m21, m22;
  c1 = dot(center, Ra.col(code1));
  c2 = dot(center, Ra.col(code2));
  m11 = dot(Ra.col(code1), Rb.col(a1));
  m12 = dot(Ra.col(code1), Rb.col(a2));
  m21 = dot(Ra.col(code2), Rb.col(a1));
  m22 = dot(Ra.col(code2), Rb.col(a2));
  {
    double k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = {Sa[code1], Sa[code2]};
  double ret[16];
  int n = rect_quad(rect, quad, ret);
  if (n < 1) return;
  V3 point[8];
  double dep[8];
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    double k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    double k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    point[cnum] = center + Rb.col(a1) * k1 + Rb.col(a2) * k2;
    dep[cnum] = Sa[codeN] - dot(normal2, point[cnum]);
    if (dep[cnum] >= 0) { ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1]; cnum++; }
  }
  if (cnum < 1) return;
  int maxc = 4;
  if (maxc > cnum) maxc = cnum;
  if (cnum <= maxc) {
    for (int j = 0; j < cnum; j++) {
      V3 pw = point[j] + pa;
      if (code >= 4) pw = pw - normal * dep[j];
      add_contact(out, -normal, pw, -dep[j]);
    }
  } else {
    int i1 = 0;
    double maxdepth = dep[0];
    for (int i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
    int iret[8];
    cull_points(cnum, ret, maxc, i1, iret);
    for (int j = 0; j < maxc; j++) {
      V3 pw = point[iret[j]] + pa;
      if (code >= 4) pw = pw - normal * dep[iret[j]];
      add_contact(out, -normal, pw, -dep[iret[j]]);
    }
  }
}

// ---------------------------------------------------------------- collision pipeline -----------------
PG_HD double bs_radius(const EnvDev& E, int b) {
  if (is_static(b)) return E.static_radius_bs[b];
  if (b == B_OBJ) return E.obj_radius_bs;
  return E.tip_radius_bs[b - B_TIP0];
}

PG_HD bool pair_allowed(const EnvDev& E, const Sim& S, int i, int j) {   // i < j in body-id order
  if (is_static(i) && is_static(j)) return false;
  if (is_static(i) && i >= E.n_static) return false;
  // tip <-> floor collisions are disabled at reset (plate_env:228)
  if (i == E.floor_index && (j == B_TIP0 || j == B_TIP1)) return false;
  if (i == B_OBJ && j >= B_TIP0 && ((S.filter_off >> (j - B_TIP0)) & 1u)) return false;
  return true;
}

// tight world AABB of a body's collision shape (hull objects: local AABB of the vertex cloud + margin)
template <int SHAPE>
PG_HD void tight_aabb(const EnvDev& E, const BodyView& b, int body, V3& lo, V3& hi) {
  if (SHAPE == SHAPE_HULL && body == B_OBJ) {
    V3 c(E.obj_aabb_c[0], E.obj_aabb_c[1], E.obj_aabb_c[2]), e(E.obj_aabb_e[0], E.obj_aabb_e[1], E.obj_aabb_e[2]);
    V3 wc = mul(b.R, c) + b.pos, we;
    for (int i = 0; i < 3; i++) we[i] = fabs(b.R(i, 0)) * e.x + fabs(b.R(i, 1)) * e.y + fabs(b.R(i, 2)) * e.z;
    lo = wc - we; hi = wc + we;
  } else aabb_of(b, 0.0, lo, hi);
}

template <int SHAPE>
PG_HDN void collide(const EnvDev& E, Sim& S, const M3* statR) {
  BodyView bv[NBODY];
  V3 lo[NBODY], hi[NBODY];
  const double am = E.aabb_margin;
  for (int d = 0; d < 3; d++) S.dyn[d].R = qmat(S.dyn[d].q);
  for (int b = 0; b < NBODY; b++) {
    if (is_static(b) && b >= E.n_static) continue;
    bv[b] = view(E, S, statR, b);
    tight_aabb<SHAPE>(E, bv[b], b, lo[b], hi[b]);
    lo[b] = lo[b] - V3(am, am, am); hi[b] = hi[b] + V3(am, am, am);
  }
  auto is_compound = [&](int b) -> bool { return is_static(b) ? (E.statics[b].compound != 0) : (SHAPE == SHAPE_HULL && b == B_OBJ && E.obj_compound); };
  auto pair_overlap = [&](int a, int b) -> bool {
    if (!overlap(lo[a], hi[a], lo[b], hi[b])) return false;
    if (is_compound(a) || is_compound(b)) {   // btCompound(Compound)CollisionAlgorithm: child manifold lives only while the TIGHT aabbs overlap
      double ea = am + ((SHAPE == SHAPE_HULL && a == B_OBJ) ? E.obj_compound_margin : 0.0);
      double eb = am + ((SHAPE == SHAPE_HULL && b == B_OBJ) ? E.obj_compound_margin : 0.0);
      for (int k = 0; k < 3; k++)
        if (lo[a][k] + ea > hi[b][k] - eb || hi[a][k] - ea < lo[b][k] + eb) return false;
    }
    return true;
  };
  // drop manifolds whose pair separated, in the order of the dispatcher's array (swap-remove: the entry that moves in is looked
  // at next)
  for (int p = 0; p < S.ndisp;) {
    const int e = S.disp[p];
    const bool keep = e >= DISP_IDLE ? idle_pair_alive(e, S.step) : pair_overlap(S.man[e].a, S.man[e].b);
    if (!keep) disp_release_at(S, p);
    else p++;
  }
  // New manifolds are created in broadphase pair-cache order.  After PyBullet's setCollisionFilterPair sequence
  // (every call re-creates both proxies) the cache holds: the surviving tip-tip and wall-tip pairs, then what the
  // re-created object proxy found at creation, then the tip pairs found with fat AABBs (DESIGN.md "Row order").
  int ca[MAXM + 6], cb[MAXM + 6], nc = 0;
  // (the fingertip-fingertip pair comes after the object's pairs even when the two tight proxies overlapped -- two parked
  // fingertips, or both on the same rim: plate_20.0, start of env step 2)
  for (int s = 0; s < E.n_static; s++) if (!E.statics[s].compound) { ca[nc] = s; cb[nc++] = B_TIP1; ca[nc] = s; cb[nc++] = B_TIP0; }
  // what the re-created object proxy finds at creation: the floor, the fingertips whose tight AABB overlaps (index, thumb), the
  // walls youngest first (fitted to the seven recordings with the idle fingertips in the dispatcher's array; round 1, without
  // them, needed a first-wall exception and a special rule for a merely touching fingertip)
  for (int s = 0; s < E.n_static; s++) if (E.statics[s].compound) { ca[nc] = s; cb[nc++] = B_OBJ; }
  for (int t = 1; t >= 0; t--) if ((S.early >> t) & 1u) { ca[nc] = B_OBJ; cb[nc++] = B_TIP0 + t; }
  for (int s = E.n_static - 1; s >= 0; s--) if (!E.statics[s].compound) { ca[nc] = s; cb[nc++] = B_OBJ; }
  for (int t = 1; t >= 0; t--) if (!((S.early >> t) & 1u)) { ca[nc] = B_OBJ; cb[nc++] = B_TIP0 + t; }
  ca[nc] = B_TIP0; cb[nc++] = B_TIP1;
  for (int c = 0; c < nc; c++) {
    int i = ca[c] < cb[c] ? ca[c] : cb[c], j = ca[c] < cb[c] ? cb[c] : ca[c];
    if (!pair_allowed(E, S, i, j)) continue;
    if (!pair_overlap(i, j)) continue;
    bool have = false;
    for (int k = 0; k < S.nman; k++) if ((S.man[k].a == i && S.man[k].b == j) || (S.man[k].a == j && S.man[k].b == i)) have = true;
    if (have || S.nman >= MAXM) continue;
#if defined(PG_HOST_DEBUG) && !defined(__CUDA_ARCH__)
    fprintf(stderr, "host new manifold %d-%d  lo_i %.5f %.5f %.5f hi_i %.5f %.5f %.5f | lo_j %.5f %.5f %.5f hi_j %.5f %.5f %.5f\n", i, j, lo[i].x, lo[i].y, lo[i].z, hi[i].x, hi[i].y, hi[i].z, lo[j].x, lo[j].y, lo[j].z, hi[j].x, hi[j].y, hi[j].z);
#endif
    S.disp[S.ndisp++] = (unsigned char)S.nman;
    Manifold& m = S.man[S.nman++];
    m.n = 0;
    // body A: a compound shape always (floor; hull object), otherwise the older broadphase proxy.  After the first
    // pre_simulate the object's proxy is the youngest, so a box object is body B of every pair.
    if (is_static(i) && !E.statics[i].compound && j == B_OBJ && (S.obj_first || is_compound(B_OBJ))) { m.a = (short)j; m.b = (short)i; }
    else { m.a = (short)i; m.b = (short)j; }
    if (i == B_OBJ && !is_compound(B_OBJ)) { m.a = (short)j; m.b = (short)i; }      // (tip, obj): tip proxy is older
    double ta = bs_radius(E, m.a) * E.breaking_factor, tb = bs_radius(E, m.b) * E.breaking_factor;
    m.breaking = ta < tb ? ta : tb;
  }
  // btCompoundCollisionAlgorithm refreshes the contact points of its child manifolds BEFORE it runs the child algorithm (which
  // refreshes them again afterwards): drifted points are gone before the new ones arrive, and a new point that has to displace
  // one of four cached points is compared with their CURRENT depths (keyboard_20.0 row 43: the first time the floor's box-box
  // result has five points; with this the whole recording is reproduced).  A loop of its own: the same call placed inside the
  // barrier-aligned loop below produced a wrong device build (host build fine) -- profiles/README.md, round 2.
  for (int k = 0; k < S.nman; k++) {
    Manifold& m = S.man[k];
    if (m.n > 0 && (is_compound(m.a) || is_compound(m.b))) refresh(m, bv[m.a], bv[m.b]);
  }
  for (int k = 0;; k++) {
    const bool live = k < S.nman;
#if defined(__CUDA_ARCH__) && defined(PG_SUBSTEP_BARRIER)
    if (S.sync_cta) { if (!__syncthreads_or(live)) break; }     // narrowphase queries of the CTA start together (as in substep)
    else if (!live) break;
    if (!live) continue;
#else
    if (!live) break;
#endif
    Manifold& m = S.man[k];
    Sink sink;
    sink.m = &m; sink.posA = bv[m.a].pos; sink.posB = bv[m.b].pos; sink.RA = bv[m.a].R; sink.RB = bv[m.b].R;
    if (SHAPE == SHAPE_HULL && (m.a == B_OBJ || m.b == B_OBJ)) {
      cvx::Scratch scratch;
      cvx::Shape sh[2];
      for (int side = 0; side < 2; side++) {
        int b = side ? m.b : m.a;
        sh[side].hull = (b == B_OBJ); sh[side].half = bv[b].half; sh[side].nverts = E.n_hull; sh[side].margin = E.obj_margin;
#if defined(__CUDA_ARCH__)
        sh[side].verts = pg_hull_sm;       // every support query walks all vertices: 30 % of the hull kernel's stalls were these loads from L2
#else
        sh[side].verts = E.hull;
#endif
      }
      V3 n, pb; double dist;
      if (cvx::closest_points(sh[0], bv[m.a].pos, bv[m.a].R, sh[1], bv[m.b].pos, bv[m.b].R, m.breaking, scratch, n, pb, dist))
        add_contact(sink, n, pb, dist);
    } else {
      box_box(bv[m.a], bv[m.b], sink);
    }
    refresh(m, bv[m.a], bv[m.b]);
  }
}

// ---------------------------------------------------------------- dynamics ---------------------------
PG_HD void apply_external(const EnvDev& E, Sim& S) {
  for (int d = 0; d < 3; d++) {
    DynBody& b = S.dyn[d];
    V3 wl = mulT(b.R, b.w), vl = mulT(b.R, b.v);
    V3 force_l = mulT(b.R, V3(0, 0, E.gravity_z * b.mass));
    V3 zero_ang, zero_lin = -force_l;
    double wn = len(wl), vn = len(vl);
    zero_ang += mulc(b.inertia, wl) * (E.ang_damping + E.ang_damping * wn);
    zero_lin += vl * (b.mass * (E.lin_damping + E.lin_damping * vn));
    zero_ang += cross(wl, mulc(b.inertia, wl));
    zero_lin += cross(wl, vl) * b.mass;
    V3 acc_ang(-zero_ang.x / b.inertia.x, -zero_ang.y / b.inertia.y, -zero_ang.z / b.inertia.z);
    V3 acc_lin(-zero_lin.x / b.mass, -zero_lin.y / b.mass, -zero_lin.z / b.mass);
    V3 wdot = mul(b.R, acc_ang);
    V3 vdot = mul(b.R, acc_lin + cross(wl, vl));
    for (int k = 0; k < 3; k++) {
      b.w[k] += wdot[k] * E.dt; b.w[k] = fmax(-E.max_coord_vel, fmin(E.max_coord_vel, b.w[k]));
    }
    for (int k = 0; k < 3; k++) {
      b.v[k] += vdot[k] * E.dt; b.v[k] = fmax(-E.max_coord_vel, fmin(E.max_coord_vel, b.v[k]));
    }
  }
}

PG_HD V3 ang_response(const DynBody& b, V3 t) {
  V3 tl = mulT(b.R, t);
  V3 al(tl.x / b.inertia.x, tl.y / b.inertia.y, tl.z / b.inertia.z);
  return mul(b.R, al);
}

// ---------------------------------------------------------------- solver ------------------------------
// Memory plan of the projected Gauss-Seidel sweep (the dominant cost of a rollout):
//   * delta velocities of the 3 dynamic bodies (18 doubles) and their world inverse inertias (18 doubles) live
//     ON CHIP: a per-thread slice of shared memory on the GPU (element i of thread t at sm[i*stride + t],
//     conflict-free), plain arrays in the host build;
//   * one record per contact POINT (normal row + both friction rows share the lever arms), 208 B, in
//     thread-local memory; jacobian cross products and responses are recomputed instead of stored.
enum { SM_DV = 0, SM_IW = 18, SM_DOUBLES = 36 };

struct PRow {
  V3 relA, relB, n, l1, l2;
  double invN, inv1, inv2, denN, den1, den2, rhsN, rhs1, rhs2, impN, imp1, imp2, mu;
  short a, b, pad0, pad1;
};

struct NcSet {                 // fingertip servo rows: 6 per active constraint, in solver order
  double rhs[12], inv[12], den[12], imp[12];
  V3 col[2][3];                // tip frame axes (angular jacobians) per constraint in solver order
  short body[2];               // dynamic index (1 or 2) of the constrained tip
  int n;                       // rows
};

struct RowSet {
  PRow pt[MAXPTS];
  NcSet nc;
  int n_pts;
};

// device: the slice lives in dynamic shared memory, addressed directly (LDS/STS, no generic pointer, no aliasing
// with the thread-local row records); host test build: a plain array
#if defined(__CUDA_ARCH__)
extern __shared__ double pg_solver_sm[];
#if defined(PG_RUNTIME_STRIDE)
#define PG_SM(i) pg_solver_sm[(i) * blockDim.x + threadIdx.x]      // slices strided by the launched CTA size: small launches keep the L1
#else
#define PG_SM(i) pg_solver_sm[(i) * PG_BLOCK_THREADS + threadIdx.x]
#endif
#else
#define PG_SM(i) sm[(i) * ss]
#endif

PG_HD V3 sm_w(const double* sm, int ss, int d) { return V3(PG_SM(SM_DV + 6 * d), PG_SM(SM_DV + 6 * d + 1), PG_SM(SM_DV + 6 * d + 2)); }
PG_HD V3 sm_v(const double* sm, int ss, int d) { return V3(PG_SM(SM_DV + 6 * d + 3), PG_SM(SM_DV + 6 * d + 4), PG_SM(SM_DV + 6 * d + 5)); }

// world inverse inertia (symmetric, 6 doubles: xx xy xz yy yz zz) times t
PG_HD V3 iw_mul(const double* sm, int ss, int d, V3 t) {
  const int o = SM_IW + 6 * d;
  double xx = PG_SM(o), xy = PG_SM(o + 1), xz = PG_SM(o + 2), yy = PG_SM(o + 3), yz = PG_SM(o + 4), zz = PG_SM(o + 5);
  return V3(xx * t.x + xy * t.y + xz * t.z, xy * t.x + yy * t.y + yz * t.z, xz * t.x + yz * t.y + zz * t.z);
}

// J . dv for one side of a row:  t . dw + ax . dv   (summed in jacobian order, as the reference engine does)
PG_HD double side_dot(const double* sm, int ss, int d, V3 t, V3 ax) {
  V3 w = sm_w(sm, ss, d), v = sm_v(sm, ss, d);
  double s = 0;
  s += t.x * w.x; s += t.y * w.y; s += t.z * w.z;
  s += ax.x * v.x; s += ax.y * v.y; s += ax.z * v.z;
  return s;
}

// `imass` = 1/mass: the inner loop multiplies by reciprocals (the reference divides; <= 1 ulp apart per term)
PG_HD void side_apply(double* sm, int ss, int d, V3 t, V3 ax, double imass, double dimp) {
  V3 r = iw_mul(sm, ss, d, t);
  double dl = imass * dimp;
  PG_SM(SM_DV + 6 * d) += r.x * dimp; PG_SM(SM_DV + 6 * d + 1) += r.y * dimp; PG_SM(SM_DV + 6 * d + 2) += r.z * dimp;
  PG_SM(SM_DV + 6 * d + 3) += ax.x * dl; PG_SM(SM_DV + 6 * d + 4) += ax.y * dl; PG_SM(SM_DV + 6 * d + 5) += ax.z * dl;
}

// apply an already-scaled angular impulse `tq` and linear impulse `li`
PG_HD void side_apply2(double* sm, int ss, int d, V3 tq, V3 li, double imass) {
  V3 r = iw_mul(sm, ss, d, tq);
  PG_SM(SM_DV + 6 * d) += r.x; PG_SM(SM_DV + 6 * d + 1) += r.y; PG_SM(SM_DV + 6 * d + 2) += r.z;
  PG_SM(SM_DV + 6 * d + 3) += li.x * imass; PG_SM(SM_DV + 6 * d + 4) += li.y * imass; PG_SM(SM_DV + 6 * d + 5) += li.z * imass;
}

// effective-mass denominator and relative velocity of one contact row (setup)
PG_HD void contact_row_setup(const EnvDev& E, const Sim& S, const double* sm, int ss, int bA, int bB, V3 relA, V3 relB, V3 axis,
                             double dist, bool is_friction, double& inv, double& den, double& rhs) {
  double denom0 = 0, denom1 = 0, rel_vel = 0;
  if (!is_static(bA)) {
    const DynBody& A = S.dyn[bA - MAX_STATICS];
    V3 tA = cross(relA, axis), wA = iw_mul(sm, ss, bA - MAX_STATICS, tA), lin = axis / A.mass;
    denom0 += tA.x * wA.x; denom0 += tA.y * wA.y; denom0 += tA.z * wA.z;
    denom0 += axis.x * lin.x; denom0 += axis.y * lin.y; denom0 += axis.z * lin.z;
    double s = 0;
    s += A.w.x * tA.x; s += A.w.y * tA.y; s += A.w.z * tA.z;
    s += A.v.x * axis.x; s += A.v.y * axis.y; s += A.v.z * axis.z;
    rel_vel += s;
  }
  if (!is_static(bB)) {
    const DynBody& B = S.dyn[bB - MAX_STATICS];
    V3 nax = -axis;
    V3 tB = cross(relB, nax), wB = iw_mul(sm, ss, bB - MAX_STATICS, tB), lin = nax / B.mass;
    denom1 += tB.x * wB.x; denom1 += tB.y * wB.y; denom1 += tB.z * wB.z;
    denom1 += nax.x * lin.x; denom1 += nax.y * lin.y; denom1 += nax.z * lin.z;
    double s = 0;
    s += B.w.x * tB.x; s += B.w.y * tB.y; s += B.w.z * tB.z;
    s += B.v.x * nax.x; s += B.v.y * nax.y; s += B.v.z * nax.z;
    rel_vel += s;
  }
  double d = denom0 + denom1;
  inv = (d > PG_EPS) ? 1.0 / d : 0.0;
  den = d;
  double distance = is_friction ? 0.0 : dist + E.slop;
  double positionalError = 0, velocityError = 0 - rel_vel;
  if (is_friction) positionalError = -distance * E.erp2 / E.dt;
  else {
    if (distance > 0) { positionalError = 0; velocityError -= distance / E.dt; }
    else positionalError = -distance * E.erp2 / E.dt;
  }
  rhs = positionalError * inv + velocityError * inv;
}

PG_HD void matrix_to_euler_xyz(const M3& mat, V3& xyz) {
  // element(idx) = mat[idx % 3][idx / 3]
#define PG_EL(idx) mat((idx) % 3, (idx) / 3)
  double fi = PG_EL(2);
  if (fi < 1.0) {
    if (fi > -1.0) { xyz.x = atan2(-PG_EL(5), PG_EL(8)); xyz.y = asin(PG_EL(2)); xyz.z = atan2(-PG_EL(1), PG_EL(0)); }
    else { xyz.x = -atan2(PG_EL(3), PG_EL(4)); xyz.y = -1.57079632679489661923; xyz.z = 0; }
  } else { xyz.x = atan2(PG_EL(3), PG_EL(4)); xyz.y = 1.57079632679489661923; xyz.z = 0; }
#undef PG_EL
}

PG_HD void setup_fixed_rows(const EnvDev& E, const Sim& S, const double* sm, int ss, int slot, NcSet& nc) {
  const FixedCons& c = S.cons[slot];
  const int d = 1 + slot;
  const DynBody& A = S.dyn[d];
  M3 relRot;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) relRot(i, j) = dot(A.R.col(i), c.frameB.col(j));
  V3 angleDiff;
  matrix_to_euler_xyz(relRot, angleDiff);
  int k = nc.n / 6;
  nc.body[k] = (short)d;
  for (int i = 0; i < 6; i++) {
    int r = nc.n++;
    V3 nl, na;
    double posError;
    if (i < 3) { nl[i] = 1; posError = dot(A.pos - c.pivotB, nl); }
    else { na = A.R.col(i % 3); nc.col[k][i % 3] = na; posError = angleDiff[i % 3]; }
    V3 wA = iw_mul(sm, ss, d, na), lin = nl / A.mass;
    double dd = 0;
    dd += na.x * wA.x; dd += na.y * wA.y; dd += na.z * wA.z;
    dd += nl.x * lin.x; dd += nl.y * lin.y; dd += nl.z * lin.z;
    double inv = (dd > PG_EPS) ? 1.0 / dd : 0.0;
    double rel_vel = 0;
    rel_vel += A.w.x * na.x; rel_vel += A.w.y * na.y; rel_vel += A.w.z * na.z;
    rel_vel += A.v.x * nl.x; rel_vel += A.v.y * nl.y; rel_vel += A.v.z * nl.z;
    double velocityError = (0 - rel_vel) * 1.0;
    double positionalError = -posError * E.erp / E.dt;
    nc.rhs[r] = positionalError * inv + velocityError * inv;
    nc.inv[r] = inv;
    nc.den[r] = dd;
    nc.imp[r] = 0;
  }
}

// btAlignedObjectArray::quickSortInternal on (key, payload) pairs, explicit stack
PG_HD void bt_quicksort(int* key, int* val, int n) {
  if (n < 2) return;
  int stack_lo[16], stack_hi[16], sp = 0;
  stack_lo[0] = 0; stack_hi[0] = n - 1; sp = 1;
  while (sp > 0) {
    sp--;
    int lo = stack_lo[sp], hi = stack_hi[sp];
    int i = lo, j = hi, x = key[(lo + hi) / 2];
    do {
      while (key[i] < x) i++;
      while (x < key[j]) j--;
      if (i <= j) { int t = key[i]; key[i] = key[j]; key[j] = t; t = val[i]; val[i] = val[j]; val[j] = t; i++; j--; }
    } while (i <= j);
    // the recursion order of the original (left first) does not change the result: the halves are disjoint
    if (i < hi) { stack_lo[sp] = i; stack_hi[sp] = hi; sp++; }
    if (lo < j) { stack_lo[sp] = lo; stack_hi[sp] = j; sp++; }
  }
}

PG_HDN void solve(const EnvDev& __restrict__ E, Sim& __restrict__ S, const M3* __restrict__ statR, RowSet& __restrict__ rs, double* sm, int ss) {
  PG_CLK0(S_t0);
  // islands over the (overlapping-pair) manifolds; tags: obj 0, tip0 1, tip1 2
  int uf[3] = {0, 1, 2};
  for (int k = 0; k < S.nman; k++) {
    int a = S.man[k].a, b = S.man[k].b;
    if (a >= MAX_STATICS && b >= MAX_STATICS) {
      int i = a - MAX_STATICS, j = b - MAX_STATICS;
      while (uf[i] != i) i = uf[i];
      while (uf[j] != j) j = uf[j];
      if (i != j) uf[i] = j;
    }
  }
  int tag[3];
  for (int d = 0; d < 3; d++) { int i = d; while (uf[i] != i) i = uf[i]; tag[d] = i; }
  // the whole dispatcher array, idle fingertips' manifolds included (their island id is larger than any of ours), goes through the
  // island manager's quickSort; the order of OUR manifolds in the result depends on how many idle entries are still there
  int mkey[MAXM + 3], mval[MAXM + 3];
  for (int p = 0; p < S.ndisp; p++) {
    const int e = S.disp[p];
    if (e >= DISP_IDLE) { mkey[p] = IDLE_ISLAND_KEY; mval[p] = -1; continue; }
    int a = S.man[e].a, b = S.man[e].b;
    mkey[p] = (a >= MAX_STATICS) ? tag[a - MAX_STATICS] : tag[b - MAX_STATICS];
    mval[p] = e;
  }
  bt_quicksort(mkey, mval, S.ndisp);
#if defined(PG_HOST_DEBUG) && !defined(__CUDA_ARCH__)
  { fprintf(stderr, "host step %d order:", S.step); for (int k = 0; k < S.ndisp; k++) if (mval[k] >= 0) fprintf(stderr, " [%d](%d-%d n%d)", mval[k], S.man[mval[k]].a, S.man[mval[k]].b, S.man[mval[k]].n); fprintf(stderr, "\n"); }
#endif
  int ckey[2], cval[2], ncons = 0;
  for (int c = 0; c < 2; c++) if (S.cons[c].active) { ckey[ncons] = tag[1 + c]; cval[ncons] = c; ncons++; }
  bt_quicksort(ckey, cval, ncons);

  // on-chip state: zero delta velocities, world inverse inertias  R diag(1/I) R^T
  for (int i = 0; i < 18; i++) PG_SM(SM_DV + i) = 0.0;
  for (int d = 0; d < 3; d++) {
    const DynBody& b = S.dyn[d];
    double ix = 1.0 / b.inertia.x, iy = 1.0 / b.inertia.y, iz = 1.0 / b.inertia.z;
    const M3& R = b.R;
    const int o = SM_IW + 6 * d;
    PG_SM(o) = R(0, 0) * R(0, 0) * ix + R(0, 1) * R(0, 1) * iy + R(0, 2) * R(0, 2) * iz;
    PG_SM(o + 1) = R(0, 0) * R(1, 0) * ix + R(0, 1) * R(1, 1) * iy + R(0, 2) * R(1, 2) * iz;
    PG_SM(o + 2) = R(0, 0) * R(2, 0) * ix + R(0, 1) * R(2, 1) * iy + R(0, 2) * R(2, 2) * iz;
    PG_SM(o + 3) = R(1, 0) * R(1, 0) * ix + R(1, 1) * R(1, 1) * iy + R(1, 2) * R(1, 2) * iz;
    PG_SM(o + 4) = R(1, 0) * R(2, 0) * ix + R(1, 1) * R(2, 1) * iy + R(1, 2) * R(2, 2) * iz;
    PG_SM(o + 5) = R(2, 0) * R(2, 0) * ix + R(2, 1) * R(2, 1) * iy + R(2, 2) * R(2, 2) * iz;
  }

  rs.nc.n = 0; rs.n_pts = 0;
  // Islands are solved one after the other, each with its own early exit (PyBullet's server sets minimumSolverBatchSize = 0,
  // so btMultiBodyDynamicsWorld hands every island to the solver separately; the cardboard recording, whose two parked
  // fingertips collide with each other far away, is reproduced to 1e-16 only this way).  Manifolds and constraints are sorted
  // by island id, so the rows of island t are the contiguous ranges [p_lo[t], p_hi[t]) and [c_lo[t], c_hi[t]).
  int p_lo[3] = {0, 0, 0}, p_hi[3] = {0, 0, 0}, c_lo[3] = {0, 0, 0}, c_hi[3] = {0, 0, 0};
  for (int o = 0; o < S.ndisp; o++) {
    if (mval[o] < 0) continue;                     // an idle fingertips' manifold: not ours to solve
    const Manifold& m = S.man[mval[o]];
    const int isl = mkey[o];
    if (p_hi[isl] == p_lo[isl]) p_lo[isl] = rs.n_pts;
    BodyView A = view(E, S, statR, m.a), B = view(E, S, statR, m.b);
    double f = A.friction * B.friction;
    if (f < -10.0) f = -10.0;
    if (f > 10.0) f = 10.0;
    for (int k = 0; k < m.n && rs.n_pts < MAXPTS; k++) {
      const MPoint& cp = m.p[k];
      V3 posA = mul(A.R, cp.localA) + A.pos, posB = mul(B.R, cp.localB) + B.pos;
      PRow& r = rs.pt[rs.n_pts++];
      r.relA = posA - A.pos; r.relB = posB - B.pos; r.n = cp.normalB; r.a = m.a; r.b = m.b; r.mu = f;
      V3 l1, l2;
      plane_space1(cp.normalB, l1, l2);
      r.l1 = l1 / len(l1);
      r.l2 = l2 / len(l2);
      contact_row_setup(E, S, sm, ss, m.a, m.b, r.relA, r.relB, r.n, cp.dist, false, r.invN, r.denN, r.rhsN);
      contact_row_setup(E, S, sm, ss, m.a, m.b, r.relA, r.relB, r.l1, 0.0, true, r.inv1, r.den1, r.rhs1);
      contact_row_setup(E, S, sm, ss, m.a, m.b, r.relA, r.relB, r.l2, 0.0, true, r.inv2, r.den2, r.rhs2);
      r.impN = 0; r.imp1 = 0; r.imp2 = 0;
    }
    p_hi[isl] = rs.n_pts;
  }
  for (int c = 0; c < ncons; c++) {
    const int isl = ckey[c];
    if (c_hi[isl] == c_lo[isl]) c_lo[isl] = rs.nc.n;
    setup_fixed_rows(E, S, sm, ss, cval[c], rs.nc);
    c_hi[isl] = rs.nc.n;
  }
  S.last_iterations = 0;
  PG_CLK(S, 2, S_t0);
#if defined(PG_HOST_STATS) && !defined(__CUDA_ARCH__)
  { extern long g_stat_pts[32], g_stat_nc[16]; g_stat_pts[rs.n_pts < 31 ? rs.n_pts : 31]++; g_stat_nc[rs.nc.n < 15 ? rs.nc.n : 15]++; }
#endif
  if (rs.nc.n == 0 && rs.n_pts == 0) return;
  double max_imp[2] = {0, 0};
  for (int c = 0; c < ncons; c++) max_imp[c] = S.cons[cval[c]].max_impulse;
  const double m_obj = 1.0 / S.dyn[0].mass, m_tip = 1.0 / S.dyn[1].mass;   // inverse masses

  // Register-resident copies of one row, loaded one row AHEAD of their use (software pipelining): the row records
  // live in thread-local memory (L1/L2), and an exposed L2 round trip per row was the top stall in the ncu source view.
  struct NcRegs { double rhs, inv, den, imp, mi; V3 col; int d, i; };
  struct NRegs { V3 relA, relB, n; double rhs, inv, den, imp; int da, db; };
#if defined(__CUDA_ARCH__)
#define PG_PIN(x) asm volatile("" : "+d"(x))
#define PG_PIN3(v) { PG_PIN((v).x); PG_PIN((v).y); PG_PIN((v).z); }
#else
#define PG_PIN(x)
#define PG_PIN3(v)
#endif
  // host-only FLOP audit of the Gauss-Seidel sweep (tools/audit_substep_flops.py); compiles to nothing in the CUDA build
#if defined(PG_HOST_STATS) && !defined(__CUDA_ARCH__)
  extern double g_stat_flops[4];
#define PG_AUDIT(slot, n) g_stat_flops[slot] += (n)
#else
#define PG_AUDIT(slot, n)
#endif
  const double mi0 = max_imp[0], mi1 = max_imp[1];
  const int ncb0 = rs.nc.body[0], ncb1 = rs.nc.body[1];
  auto load_nc = [&](int index) -> NcRegs {
    NcRegs q;
    int c = index >= 6 ? 1 : 0;
    q.i = index - 6 * c; q.d = c ? ncb1 : ncb0; q.mi = c ? mi1 : mi0;
    q.rhs = rs.nc.rhs[index]; q.inv = rs.nc.inv[index]; q.den = rs.nc.den[index]; q.imp = rs.nc.imp[index];
    q.col = (q.i >= 3) ? rs.nc.col[c][q.i - 3] : V3();
    PG_PIN(q.rhs); PG_PIN(q.inv); PG_PIN(q.den); PG_PIN(q.imp); PG_PIN3(q.col);
    return q;
  };
  auto load_n = [&](int k) -> NRegs {
    const PRow& r = rs.pt[k];
    NRegs q;
    q.relA = r.relA; q.relB = r.relB; q.n = r.n; q.rhs = r.rhsN; q.inv = r.invN; q.den = r.denN; q.imp = r.impN;
    q.da = r.a - MAX_STATICS; q.db = r.b - MAX_STATICS;
    PG_PIN3(q.relA); PG_PIN3(q.relB); PG_PIN3(q.n); PG_PIN(q.rhs); PG_PIN(q.inv); PG_PIN(q.den); PG_PIN(q.imp);
    return q;
  };

  int it_max = 0;
  for (int isl = 0; isl < 3; isl++) {
  const int n0 = c_lo[isl], nn = c_hi[isl] - c_lo[isl], p0 = p_lo[isl], p1 = p_hi[isl];
  if (nn == 0 && p1 == p0) continue;
  int it;
  for (it = 0; it < E.iterations; it++) {
    double lsr = 0;
    // ---- fingertip servo rows, alternating sweep direction
    if (nn > 0) {
      NcRegs cur = load_nc(n0 + ((it & 1) ? 0 : nn - 1));
      for (int j = 0; j < nn; j++) {
        const int index = n0 + ((it & 1) ? j : nn - 1 - j);
        NcRegs nxt = cur;
        if (j + 1 < nn) nxt = load_nc(n0 + ((it & 1) ? j + 1 : nn - 2 - j));
        const int d = cur.d, i = cur.i;
        double delta = cur.rhs, imp = cur.imp;
        if (i < 3) {
          double a = 0;
          a += PG_SM(SM_DV + 6 * d + 3 + i);            // jacobian = unit vector e_i on the linear part
          delta -= a * cur.inv;
        } else {
          delta -= side_dot(sm, ss, d, cur.col, V3()) * cur.inv;
        }
        double sum = imp + delta;
        if (sum < -cur.mi) { delta = -cur.mi - imp; imp = -cur.mi; }
        else if (sum > cur.mi) { delta = cur.mi - imp; imp = cur.mi; }
        else imp = sum;
        rs.nc.imp[index] = imp;
        if (i < 3) PG_SM(SM_DV + 6 * d + 3 + i) += m_tip * delta;
        else side_apply(sm, ss, d, cur.col, V3(), m_tip, delta);
        double res = delta * cur.den;            // = delta / inv, the velocity change of this row
        lsr = fmax(lsr, res * res);
        PG_AUDIT(0, i < 3 ? 7 : 45);             // linear row: 7 flops; angular row: 12 (dot) + 28 (inertia product, axpy) + 5
        cur = nxt;
      }
    }
    // ---- normal rows
    if (p1 > p0) {
      NRegs cur = load_n(p0);
      for (int k = p0; k < p1; k++) {
        NRegs nxt = cur;
        if (k + 1 < p1) nxt = load_n(k + 1);
        const int da = cur.da, db = cur.db;
        V3 n = cur.n, tA = cross(cur.relA, n), tB = cross(cur.relB, -n);
        double delta = cur.rhs, a = 0, b = 0;
        if (da >= 0) a = side_dot(sm, ss, da, tA, n);
        if (db >= 0) b = side_dot(sm, ss, db, tB, -n);
        delta -= a * cur.inv;
        delta -= b * cur.inv;
        double sum = cur.imp + delta, impn;
        if (sum < 0.0) { delta = 0.0 - cur.imp; impn = 0.0; }
        else if (sum > 1e10) { delta = 1e10 - cur.imp; impn = 1e10; }
        else impn = sum;
        rs.pt[k].impN = impn;
        if (da >= 0) side_apply(sm, ss, da, tA, n, da == 0 ? m_obj : m_tip, delta);
        if (db >= 0) side_apply(sm, ss, db, tB, -n, db == 0 ? m_obj : m_tip, delta);
        double res = delta * cur.den;
        lsr = fmax(lsr, res * res);
        PG_AUDIT(1, 25 + 40 * ((da >= 0) + (db >= 0)));   // 2 cross products + clamp, per dynamic side a 6-dot and an impulse application
        cur = nxt;
      }
    }
    // ---- friction pairs with the implicit cone
    for (int k = p0; k < p1; k++) {
      PRow& r = rs.pt[k];
      // no normal impulse and no stored friction impulse: the cone clips both rows to zero, nothing changes
      // (identical result, zero residual) -- skip the row pair
      const double impN = r.impN, imp1 = r.imp1, imp2 = r.imp2;
      if (impN == 0.0 && imp1 == 0.0 && imp2 == 0.0) continue;
      const int da = r.a - MAX_STATICS, db = r.b - MAX_STATICS;
      V3 l1 = r.l1, l2 = r.l2, relA = r.relA, relB = r.relB;
      double inv1 = r.inv1, inv2 = r.inv2, den1 = r.den1, den2 = r.den2, rhs1 = r.rhs1, rhs2 = r.rhs2, mu = r.mu;
      PG_PIN3(l1); PG_PIN3(l2); PG_PIN3(relA); PG_PIN3(relB);
      PG_PIN(inv1); PG_PIN(inv2); PG_PIN(den1); PG_PIN(den2); PG_PIN(rhs1); PG_PIN(rhs2); PG_PIN(mu);
      V3 tA1 = cross(relA, l1), tB1 = cross(relB, -l1), tA2 = cross(relA, l2), tB2 = cross(relB, -l2);
      double lim = mu * impN, lo = -lim;
      double dB = rhs2, dA = rhs1, a1 = 0, b1 = 0, a2 = 0, b2 = 0;
      if (da >= 0) { a2 = side_dot(sm, ss, da, tA2, l2); a1 = side_dot(sm, ss, da, tA1, l1); }
      if (db >= 0) { b2 = side_dot(sm, ss, db, tB2, -l2); b1 = side_dot(sm, ss, db, tB1, -l1); }
      dB -= a2 * inv2; dB -= b2 * inv2;
      dA -= a1 * inv1; dA -= b1 * inv1;
      double sumB = imp2 + dB, sumA = imp1 + dA, o1, o2;
      if (sumA * sumA + sumB * sumB >= lo * lo) {
        // |lo*sin(atan2(a,b))| = |lo*a|/hypot, |lo*cos(atan2(a,b))| = |lo*b|/hypot: same clip without transcendentals
        double rinv = 1.0 / sqrt(sumA * sumA + sumB * sumB);
        double ca = fabs(lo * sumA * rinv), cb = fabs(lo * sumB * rinv);
        if (sumA < -ca) { dA = -ca - imp1; o1 = -ca; }
        else if (sumA > ca) { dA = ca - imp1; o1 = ca; }
        else o1 = sumA;
        if (sumB < -cb) { dB = -cb - imp2; o2 = -cb; }
        else if (sumB > cb) { dB = cb - imp2; o2 = cb; }
        else o2 = sumB;
      } else { o1 = sumA; o2 = sumB; }
      r.imp1 = o1; r.imp2 = o2;
      // both deltas were computed against the same velocities (Jacobi inside the pair), so the two impulses of a
      // body are applied as ONE combined linear + angular impulse (one inverse-inertia product per body)
      if (da >= 0) side_apply2(sm, ss, da, tA1 * dA + tA2 * dB, l1 * dA + l2 * dB, da == 0 ? m_obj : m_tip);
      if (db >= 0) side_apply2(sm, ss, db, tB1 * dA + tB2 * dB, -(l1 * dA + l2 * dB), db == 0 ? m_obj : m_tip);
      double res = dA * den1 + dB * den2;
      lsr = fmax(lsr, res * res);
      PG_AUDIT(2, 67 + 66 * ((da >= 0) + (db >= 0)));     // 4 cross products, cone clip, per dynamic side two 6-dots and one combined impulse
    }
    PG_AUDIT(3, 1);                                       // iterations
    if (lsr <= E.residual_threshold || it >= E.iterations - 1) { it++; break; }
  }
  if (it > it_max) it_max = it;
  }
  S.last_iterations = it_max;
  PG_CLK(S, 3, S_t0);
  for (int d = 0; d < 3; d++) {
    DynBody& b = S.dyn[d];
    V3 w = sm_w(sm, ss, d), v = sm_v(sm, ss, d);
    b.w.x += w.x; b.w.y += w.y; b.w.z += w.z;
    b.v.x += v.x; b.v.y += v.y; b.v.z += v.z;
  }
}

PG_HD void integrate(const EnvDev& E, Sim& S) {
  for (int d = 0; d < 3; d++) {
    DynBody& b = S.dyn[d];
    b.pos += b.v * E.dt;
    V3 angvel = b.w;
    double fAngle = len(angvel);
    const double THRESH = 0.5 * 1.57079632679489661923;
    if (fAngle * E.dt > THRESH) fAngle = 0.5 * 1.57079632679489661923 / E.dt;
    V3 axis;
    if (fAngle < 0.001) axis = angvel * (0.5 * E.dt - (E.dt * E.dt * E.dt) * 0.020833333333 * fAngle * fAngle);
    else axis = angvel * (sin(0.5 * fAngle * E.dt) / fAngle);
    Q4 dqi; dqi.x = -axis.x; dqi.y = -axis.y; dqi.z = -axis.z; dqi.w = cos(fAngle * E.dt * 0.5);
    Q4 inv; inv.x = -b.q.x; inv.y = -b.q.y; inv.z = -b.q.z; inv.w = b.q.w;
    Q4 ninv = qnormalize(qmul(inv, dqi));
    b.q.x = -ninv.x; b.q.y = -ninv.y; b.q.z = -ninv.z; b.q.w = ninv.w;
    b.R = qmat(b.q);
  }
}

template <int SHAPE>
PG_HDN void substep(const EnvDev& E, Sim& S, const M3* statR, RowSet& rs, double* sm, int ss) {
#if defined(__CUDA_ARCH__)
  // The hull kernel stalls on instruction fetch: ~15 k rarely-executed instructions of collision code per substep, missed
  // separately by every warp that reaches them at a different time.  All rollouts run the same number of substeps, so the
  // warps of a CTA are lined up here, before every narrowphase query (collide) and before the solver set-up, and take
  // those misses together.
#if defined(PG_SUBSTEP_BARRIER)
  if (S.sync_cta) __syncthreads();
#endif
#endif
  PG_CLK0(t0);
  S.step++;
  collide<SHAPE>(E, S, statR);
#if defined(__CUDA_ARCH__) && defined(PG_SUBSTEP_BARRIER)
  if (S.sync_cta) __syncthreads();
#endif
  PG_CLK(S, 0, t0);
  apply_external(E, S);
  PG_CLK(S, 1, t0);
  solve(E, S, statR, rs, sm, ss);
  PG_CLK_RESET(t0);
  integrate(E, S);
  PG_CLK(S, 4, t0);
}

PG_HD void drop_manifolds_of(Sim& S, int body) {
  for (int p = 0; p < S.ndisp;) {
    const int e = S.disp[p];
    if (e < DISP_IDLE && (S.man[e].a == body || S.man[e].b == body)) disp_release_at(S, p);
    else p++;
  }
}

// PyBullet's setCollisionFilterPair re-creates the broadphase proxies of both bodies (plate_env:466, :482): every
// manifold of the body is dropped, and the new proxy is collided against the trees with its TIGHT AABB at once --
// which decides the order of the pair cache, hence of the contact rows (see collide)
template <int SHAPE>
PG_HD void refresh_proxy(const EnvDev& E, Sim& S, int body) {
  drop_manifolds_of(S, body);
  DynBody& d = S.dyn[body - MAX_STATICS];
  d.R = qmat(d.q);
  BodyView v; v.pos = d.pos; v.R = d.R; v.half = d.half; v.friction = d.friction;
  V3 lo, hi;
  tight_aabb<SHAPE>(E, v, body, lo, hi);
  if (body != B_OBJ) { S.plo[body - B_TIP0] = lo; S.phi[body - B_TIP0] = hi; return; }
  double depth[2];
  for (int t = 0; t < 2; t++) {
    double dd = 1e300;
    for (int k = 0; k < 3; k++) { double a = S.phi[t][k] - lo[k], b = hi[k] - S.plo[t][k]; double m = a < b ? a : b; if (m < dd) dd = m; }
    depth[t] = ((S.filter_off >> t) & 1u) ? -1.0 : dd;
  }
  S.early = 0;
  for (int t = 0; t < 2; t++) {
    if (depth[t] >= 0) S.early |= (1u << t);
  }
}

}  // namespace pg
