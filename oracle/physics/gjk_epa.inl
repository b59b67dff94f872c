// Convex-convex narrowphase for hull / box pairs (the plate env).  Restates, from the published algorithm,
//   btConvexConvexAlgorithm::processCollision (non-SAT path: ONE btGjkPairDetector query per substep, no perturbation
//   iterations), btGjkPairDetector::getClosestPointsNonVirtual (margins removed, origin re-centred between the two
//   bodies, REL_ERROR2 = 1e-12 progress tests (double-precision build), first axis (0,1,0), the final normal checks d0 / d1 / d2
//   with the GJK normal kept as orgNormalInB), and btVoronoiSimplexSolver (Ericson closest-point sub-simplex tests,
//   reduceVertices order, double-precision degenerate-tetrahedron and equal-vertex thresholds).
// Overlapping cores (or a degenerate simplex closer than gGjkEpaPenetrationTolerance = 1e-12) go to the EPA restatement in
// epa.inl, as btGjkPairDetector does.
namespace pgo {

static long g_epa_events = 0, g_gjk_calls = 0;
static double g_pen_tie = 0.0;   // debugging: bias of the EPA-vs-GJK depth comparison (Bullet: none)
// btGjkPairDetector.cpp in a BT_USE_DOUBLE_PRECISION build (PyBullet's): REL_ERROR2 = 1e-12 and gGjkEpaPenetrationTolerance = 1e-12
// (1e-6 / 0.001 are the single-precision values: with them a touching pair whose cores are less than 1 mm apart went to EPA and
// plate_20.0 left the recording at row 87; with the double-precision values it stays within 1e-5 over all 100 rows).  The knobs
// remain for the what-if runs of tools/oracle_vs_recordings.py.
static double g_gjk_rel_error2 = 1.0e-12, g_gjk_pen_tol = 1.0e-12; static int g_gjk_org_rule = 1;
static double g_cc_nudge[7] = {0, 0, 0, 0, 0, 0, 0};   // debugging: at step dbg_step, fingertip pairs: point += [0..2], distance += [3], normal += [4..6] (renormalised)

static V3 support_core_local(const Shape& sh, V3 d) {
  if (sh.type == SH_BOX) {
    V3 h = sh.half - V3(sh.margin, sh.margin, sh.margin);          // getImplicitShapeDimensions
    return V3(d.x >= 0 ? h.x : -h.x, d.y >= 0 ? h.y : -h.y, d.z >= 0 ? h.z : -h.z);   // btFsels
  }
  if (sh.type == SH_CYL) {
    // btConvexShape::localGetSupportVertexWithoutMarginNonVirtual, cylinder with up axis Z (URDF <cylinder>)
    double radius = sh.radius - sh.margin, hh = sh.halfh - sh.margin;
    double s = std::sqrt(d.x * d.x + d.y * d.y);
    if (s != 0.0) { double k = radius / s; return V3(d.x * k, d.y * k, d.z < 0.0 ? -hh : hh); }
    return V3(radius, 0.0, d.z < 0.0 ? -hh : hh);
  }
  double best = -1e300; int bi = 0;                                  // btVector3::maxDot: first strict maximum
  for (size_t i = 0; i < sh.verts.size(); i++) { double t = dot(sh.verts[i], d); if (t > best) { best = t; bi = (int)i; } }
  return sh.verts[bi];
}

struct VSimplex {
  int n = 0;
  V3 W[5], P[5], Q[5];
  V3 lastW = V3(1e30, 1e30, 1e30);
  V3 cP1, cP2, cV;
  double bc[4] = {0, 0, 0, 0};
  bool used[4] = {false, false, false, false};
  bool degenerate = false, valid = false, needs_update = true;

  void add(V3 w, V3 p, V3 q) { lastW = w; needs_update = true; W[n] = w; P[n] = p; Q[n] = q; n++; }
  void remove(int i) { n--; W[i] = W[n]; P[i] = P[n]; Q[i] = Q[n]; }
  void reduce() {
    if (n >= 4 && !used[3]) remove(3);
    if (n >= 3 && !used[2]) remove(2);
    if (n >= 2 && !used[1]) remove(1);
    if (n >= 1 && !used[0]) remove(0);
  }
  bool in_simplex(V3 w) const {
    bool found = false;
    for (int i = 0; i < n; i++) if (len2(W[i] - w) <= 1e-12) found = true;   // VORONOI_DEFAULT_EQUAL_VERTEX_THRESHOLD (double)
    if (w.x == lastW.x && w.y == lastW.y && w.z == lastW.z) return true;
    return found;
  }
  bool bc_valid() const { return bc[0] >= 0 && bc[1] >= 0 && bc[2] >= 0 && bc[3] >= 0; }
};

struct SubRes { V3 closest; bool u[4]; double bc[4]; bool degenerate; };

static void closest_pt_triangle(V3 p, V3 a, V3 b, V3 c, SubRes& r) {
  for (int i = 0; i < 4; i++) { r.u[i] = false; r.bc[i] = 0; }
  V3 ab = b - a, ac = c - a, ap = p - a;
  double d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { r.closest = a; r.u[0] = true; r.bc[0] = 1; return; }
  V3 bp = p - b; double d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { r.closest = b; r.u[1] = true; r.bc[1] = 1; return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); r.closest = a + ab * v; r.u[0] = r.u[1] = true; r.bc[0] = 1 - v; r.bc[1] = v; return; }
  V3 cp = p - c; double d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { r.closest = c; r.u[2] = true; r.bc[2] = 1; return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); r.closest = a + ac * w; r.u[0] = r.u[2] = true; r.bc[0] = 1 - w; r.bc[2] = w; return; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    r.closest = b + (c - b) * w; r.u[1] = r.u[2] = true; r.bc[1] = 1 - w; r.bc[2] = w; return;
  }
  double denom = 1.0 / (va + vb + vc), v = vb * denom, w = vc * denom;
  r.closest = a + ab * v + ac * w; r.u[0] = r.u[1] = r.u[2] = true; r.bc[0] = 1 - v - w; r.bc[1] = v; r.bc[2] = w;
}

static int outside_of_plane(V3 p, V3 a, V3 b, V3 c, V3 d) {
  V3 nrm = cross(b - a, c - a);
  double signp = dot(p - a, nrm), signd = dot(d - a, nrm);
  if (signd * signd < 1e-8 * 1e-8) return -1;          // CATCH_DEGENERATE_TETRAHEDRON, double precision
  return signp * signd < 0 ? 1 : 0;
}

static bool closest_pt_tetrahedron(V3 p, V3 a, V3 b, V3 c, V3 d, SubRes& fr) {
  fr.closest = p; fr.degenerate = false;
  for (int i = 0; i < 4; i++) { fr.u[i] = true; fr.bc[i] = 0; }
  int oABC = outside_of_plane(p, a, b, c, d), oACD = outside_of_plane(p, a, c, d, b);
  int oADB = outside_of_plane(p, a, d, b, c), oBDC = outside_of_plane(p, b, d, c, a);
  if (oABC < 0 || oACD < 0 || oADB < 0 || oBDC < 0) { fr.degenerate = true; return false; }
  if (!oABC && !oACD && !oADB && !oBDC) return false;
  double best = 3.402823466e+38;
  SubRes t;
  if (oABC) { closest_pt_triangle(p, a, b, c, t); V3 q = t.closest; double sd = dot(q - p, q - p);
    if (sd < best) { best = sd; fr.closest = q; fr.u[0] = t.u[0]; fr.u[1] = t.u[1]; fr.u[2] = t.u[2]; fr.u[3] = false;
      fr.bc[0] = t.bc[0]; fr.bc[1] = t.bc[1]; fr.bc[2] = t.bc[2]; fr.bc[3] = 0; } }
  if (oACD) { closest_pt_triangle(p, a, c, d, t); V3 q = t.closest; double sd = dot(q - p, q - p);
    if (sd < best) { best = sd; fr.closest = q; fr.u[0] = t.u[0]; fr.u[1] = false; fr.u[2] = t.u[1]; fr.u[3] = t.u[2];
      fr.bc[0] = t.bc[0]; fr.bc[1] = 0; fr.bc[2] = t.bc[1]; fr.bc[3] = t.bc[2]; } }
  if (oADB) { closest_pt_triangle(p, a, d, b, t); V3 q = t.closest; double sd = dot(q - p, q - p);
    if (sd < best) { best = sd; fr.closest = q; fr.u[0] = t.u[0]; fr.u[1] = t.u[2]; fr.u[2] = false; fr.u[3] = t.u[1];
      fr.bc[0] = t.bc[0]; fr.bc[1] = t.bc[2]; fr.bc[2] = 0; fr.bc[3] = t.bc[1]; } }
  if (oBDC) { closest_pt_triangle(p, b, d, c, t); V3 q = t.closest; double sd = dot(q - p, q - p);
    if (sd < best) { best = sd; fr.closest = q; fr.u[0] = false; fr.u[1] = t.u[0]; fr.u[2] = t.u[2]; fr.u[3] = t.u[1];
      fr.bc[0] = 0; fr.bc[1] = t.bc[0]; fr.bc[2] = t.bc[2]; fr.bc[3] = t.bc[1]; } }
  return true;
}

// btVoronoiSimplexSolver::updateClosestVectorAndPoints
static bool simplex_update(VSimplex& s) {
  if (!s.needs_update) return s.valid;
  s.needs_update = false;
  for (int i = 0; i < 4; i++) { s.bc[i] = 0; s.used[i] = false; }
  s.degenerate = false;
  switch (s.n) {
    case 0: s.valid = false; break;
    case 1:
      s.cP1 = s.P[0]; s.cP2 = s.Q[0]; s.cV = s.cP1 - s.cP2; s.bc[0] = 1; s.valid = s.bc_valid(); break;
    case 2: {
      V3 from = s.W[0], to = s.W[1];
      V3 diff = V3() - from, v = to - from;
      double t = dot(v, diff);
      if (t > 0) {
        double dvv = dot(v, v);
        if (t < dvv) { t /= dvv; diff = diff - v * t; s.used[0] = s.used[1] = true; }
        else { t = 1; diff = diff - v; s.used[1] = true; }
      } else { t = 0; s.used[0] = true; }
      s.bc[0] = 1 - t; s.bc[1] = t;
      s.cP1 = s.P[0] + (s.P[1] - s.P[0]) * t; s.cP2 = s.Q[0] + (s.Q[1] - s.Q[0]) * t; s.cV = s.cP1 - s.cP2;
      s.reduce(); s.valid = s.bc_valid(); break;
    }
    case 3: {
      SubRes r; closest_pt_triangle(V3(), s.W[0], s.W[1], s.W[2], r);
      for (int i = 0; i < 4; i++) { s.bc[i] = r.bc[i]; s.used[i] = r.u[i]; }
      s.cP1 = s.P[0] * s.bc[0] + s.P[1] * s.bc[1] + s.P[2] * s.bc[2];
      s.cP2 = s.Q[0] * s.bc[0] + s.Q[1] * s.bc[1] + s.Q[2] * s.bc[2];
      s.cV = s.cP1 - s.cP2; s.reduce(); s.valid = s.bc_valid(); break;
    }
    case 4: {
      SubRes r; bool sep = closest_pt_tetrahedron(V3(), s.W[0], s.W[1], s.W[2], s.W[3], r);
      for (int i = 0; i < 4; i++) { s.bc[i] = r.bc[i]; s.used[i] = r.u[i]; }
      s.degenerate = r.degenerate;
      if (sep) {
        s.cP1 = s.P[0] * s.bc[0] + s.P[1] * s.bc[1] + s.P[2] * s.bc[2] + s.P[3] * s.bc[3];
        s.cP2 = s.Q[0] * s.bc[0] + s.Q[1] * s.bc[1] + s.Q[2] * s.bc[2] + s.Q[3] * s.bc[3];
        s.cV = s.cP1 - s.cP2; s.reduce();
      } else {
        if (s.degenerate) s.valid = false; else { s.valid = true; s.cV = V3(); }
        break;
      }
      s.valid = s.bc_valid(); break;
    }
  }
  return s.valid;
}

}  // namespace pgo
#include "epa.inl"
namespace pgo {

static void convex_convex(World& w, Body& A, Body& B, ContactSink& out) {
  Manifold& m = *out.m;
  g_gjk_calls++;
  const double REL_ERROR2 = g_gjk_rel_error2;
  double marginA = A.shape.margin, marginB = B.shape.margin, margin = marginA + marginB;
  double maxd = margin + m.breaking, max_dist2 = maxd * maxd;
  V3 offset = (A.pos + B.pos) * 0.5;
  V3 oA = A.pos - offset, oB = B.pos - offset;
  V3 axis(0, 1, 0), pointOnA, pointOnB, normalInB, orgNormalInB;
  double sqd = 1e30, distance = 0;
  bool is_valid = false, check_simplex = false;
  int degenerate = 0, iter = 0;
  VSimplex sx;
  for (;;) {
    V3 dA = mulT(A.R, -axis), dB = mulT(B.R, axis);
    V3 pW = mul(A.R, support_core_local(A.shape, dA)) + oA;
    V3 qW = mul(B.R, support_core_local(B.shape, dB)) + oB;
    V3 wv = pW - qW;
    double delta = dot(axis, wv);
    if (delta > 0 && delta * delta > sqd * max_dist2) { degenerate = 10; check_simplex = true; break; }
    if (sx.in_simplex(wv)) { degenerate = 1; check_simplex = true; break; }
    double f0 = sqd - delta, f1 = sqd * REL_ERROR2;
    if (f0 <= f1) { degenerate = (f0 <= 0) ? 2 : 11; check_simplex = true; break; }
    sx.add(wv, pW, qW);
    bool ok = simplex_update(sx);
    V3 nv = sx.cV;
    if (!ok) { degenerate = 3; check_simplex = true; break; }
    if (len2(nv) < REL_ERROR2) { axis = nv; degenerate = 6; check_simplex = true; break; }
    double prev = sqd;
    sqd = len2(nv);
    if (prev - sqd <= kEps * prev) { check_simplex = true; degenerate = 12; break; }
    axis = nv;
    if (iter++ > 1000) break;
    if (sx.n == 4) { degenerate = 13; break; }
  }
  if (check_simplex) {
    simplex_update(sx);
    pointOnA = sx.cP1; pointOnB = sx.cP2;
    normalInB = axis;
    double lsq = len2(axis);
    if (lsq < 1e-4) degenerate = 5;
    if (lsq > kEps * kEps) {
      double rlen = 1.0 / std::sqrt(lsq);
      normalInB = normalInB * rlen;
      double s = std::sqrt(lsq);
      pointOnA = pointOnA - axis * (marginA / s);
      pointOnB = pointOnB + axis * (marginB / s);
      distance = (1.0 / rlen) - margin;
      is_valid = true;
      orgNormalInB = normalInB;
    }
  }
  bool catch_degenerate = degenerate && ((distance + margin) < g_gjk_pen_tol);      // gGjkEpaPenetrationTolerance
  int pen_info = 0;
  if (!is_valid || catch_degenerate) {
    // btGjkPairDetector penetration branch: EPA on the margin-inclusive shapes (epa.inl)
    g_epa_events++;
    epa2::g_epa_debug = (w.P.debug > 1); epa2::g_epa_trace = (w.P.debug == 8 && w.step_count == w.P.dbg_step);
    V3 tA, tB;
    axis = V3();
    bool valid2 = epa2::calc_pen_depth(A.shape, A.R, oA, B.shape, B.R, oB, axis, tA, tB, pen_info);
    if (len2(axis) != 0) {
      if (valid2) {
        V3 tn = tB - tA; double lsq = len2(tn);
        if (lsq <= kEps * kEps) { tn = axis; lsq = len2(axis); }
        if (lsq > kEps * kEps) {
          tn = epa2::vdiv(tn, std::sqrt(lsq));
          double d2 = -len(tA - tB);
          if (w.P.debug > 1) fprintf(stderr, "  pen compare: epa %.17g gjk %.17g valid %d diff %.3g\n", d2, distance, (int)is_valid, d2 - distance);
          if (!is_valid || d2 < distance + g_pen_tie) { distance = d2; pointOnA = tA; pointOnB = tB; normalInB = tn; is_valid = true; }
        }
      } else {
        double d2 = len(tA - tB) - margin;
        if (!is_valid || d2 < distance) {
          distance = d2; pointOnA = tA - axis * marginA; pointOnB = tB + axis * marginB;
          normalInB = epa2::vdiv(axis, len(axis)); is_valid = true;
        }
      }
    }
  }
  if (is_valid && (distance < 0 || distance * distance < max_dist2)) {
    // "EPA degeneracy" guard of the detector: flip the normal when the opposite direction separates more
    auto sep = [&](V3 n) {
      V3 p = mul(A.R, support_core_local(A.shape, mulT(A.R, -n))) + oA;
      V3 q = mul(B.R, support_core_local(B.shape, mulT(B.R, n))) + oB;
      return dot(n, p - q) - margin;
    };
    double d0 = sep(normalInB), d1 = sep(-normalInB);
    if (d1 > d0) normalInB = -normalInB;
    if (g_gjk_org_rule && len2(orgNormalInB) != 0) {      // the GJK normal wins when it separates more than the EPA result does
      double d2 = sep(orgNormalInB);
      if (d2 > d0 && d2 > d1 && d2 > distance) { if (w.P.debug > 1) fprintf(stderr, "  org rule: d0 %.9g d1 %.9g d2 %.9g distance %.9g\n", d0, d1, d2, distance); normalInB = orgNormalInB; distance = d2; }
    }
  }
  if (w.P.debug > 1)
    fprintf(stderr, "gjk step %d pair %d-%d iters %d degen %d n %d valid %d pen %d dist %.9g nrm %.6f %.6f %.6f pB %.6f %.6f %.6f\n", w.step_count,
            (int)(&A - &w.bodies[0]), (int)(&B - &w.bodies[0]), iter, degenerate, sx.n, (int)is_valid, pen_info, distance, normalInB.x, normalInB.y, normalInB.z,
            pointOnB.x + offset.x, pointOnB.y + offset.y, pointOnB.z + offset.z);
  if (w.P.dbg_step > 0 && w.step_count == w.P.dbg_step && (&B - &w.bodies[0] == w.tip[0] || &A - &w.bodies[0] == w.tip[0] || &B - &w.bodies[0] == w.tip[1] || &A - &w.bodies[0] == w.tip[1])) {
    pointOnB = pointOnB + V3(g_cc_nudge[0], g_cc_nudge[1], g_cc_nudge[2]); distance += g_cc_nudge[3];
    normalInB = normalInB + V3(g_cc_nudge[4], g_cc_nudge[5], g_cc_nudge[6]); normalInB = normalInB / len(normalInB);
  }
  if (is_valid && (distance < 0 || distance * distance < max_dist2))
    add_contact(out, normalInB, pointOnB + offset, distance);
  (void)w;
}

}  // namespace pgo
