// Convex-convex narrowphase for hull objects (the plate env): one closest-point / penetration query per pair and
// substep, with the semantics of Bullet's btConvexConvexAlgorithm (reference: envs/plate_contact_graph_env.py:198
// loads plate_cvx_simple.urdf, whose mesh becomes a btConvexHullShape inside a btCompoundShape).
//   * separated cores: GJK on the margin-less shapes with a Voronoi-region simplex solver (double-precision tolerances), margins added back
//   * overlapping cores (degenerate simplex within 1e-12): expanding-polytope penetration query on the margin-inclusive
//     shapes, nine guess directions; the detector's final normal checks (d0 / d1, and the GJK normal when it separates more)
// Index-based pools and explicit stacks (no recursion, no pointers) so one thread can run it out of local memory.
// Written after bullet3's published btGjkPairDetector / btVoronoiSimplexSolver / btGjkEpa2 (zlib licence, (c) Erwin Coumans,
// Nathanael Presson et al.): bit-level parity with that engine requires its thresholds and operation order.
#pragma once
#include "pg_math.cuh"

namespace pg {
namespace cvx {

struct Shape {
  int hull;              // 0: box (half = full half extents, margin inside), 1: vertex cloud (margin outside)
  V3 half;
  const double* verts;
  int nverts;
  double margin;
};
struct Xf { V3 o; M3 R; };

PG_HD double len2(V3 a) { return dot(a, a); }
PG_HD V3 vdiv(V3 a, double s) { return a * (1.0 / s); }

PG_HDN V3 support_core(const Shape& s, V3 d) {
  if (!s.hull) {
    V3 h(s.half.x - s.margin, s.half.y - s.margin, s.half.z - s.margin);
    return V3(d.x >= 0 ? h.x : -h.x, d.y >= 0 ? h.y : -h.y, d.z >= 0 ? h.z : -h.z);
  }
  double best = -1e300; int bi = 0;
  for (int i = 0; i < s.nverts; i++) {
    double t = s.verts[3 * i] * d.x + s.verts[3 * i + 1] * d.y + s.verts[3 * i + 2] * d.z;
    if (t > best) { best = t; bi = i; }
  }
  return V3(s.verts[3 * bi], s.verts[3 * bi + 1], s.verts[3 * bi + 2]);
}

// ------------------------------------------------------------------ Voronoi simplex (separated cores)
struct VSimplex {
  int n;
  V3 W[4], P[4], Q[4], lastW, cP1, cP2, cV;
  double bc[4];
  bool used[4], degenerate, valid, dirty;
};

PG_HD void vs_remove(VSimplex& s, int i) { s.n--; s.W[i] = s.W[s.n]; s.P[i] = s.P[s.n]; s.Q[i] = s.Q[s.n]; }
PG_HD void vs_reduce(VSimplex& s) {
  if (s.n >= 4 && !s.used[3]) vs_remove(s, 3);
  if (s.n >= 3 && !s.used[2]) vs_remove(s, 2);
  if (s.n >= 2 && !s.used[1]) vs_remove(s, 1);
  if (s.n >= 1 && !s.used[0]) vs_remove(s, 0);
}
PG_HD bool vs_bc_valid(const VSimplex& s) { return s.bc[0] >= 0 && s.bc[1] >= 0 && s.bc[2] >= 0 && s.bc[3] >= 0; }

struct SubRes { V3 closest; bool u[4]; double bc[4]; bool degenerate; };

PG_HDN void closest_on_triangle(V3 p, V3 a, V3 b, V3 c, SubRes& r) {
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

PG_HD int outside_plane(V3 p, V3 a, V3 b, V3 c, V3 d) {
  V3 nrm = cross(b - a, c - a);
  double signp = dot(p - a, nrm), signd = dot(d - a, nrm);
  if (signd * signd < 1e-8 * 1e-8) return -1;
  return signp * signd < 0 ? 1 : 0;
}

PG_HDN bool closest_on_tetrahedron(V3 p, V3 a, V3 b, V3 c, V3 d, SubRes& fr) {
  fr.closest = p; fr.degenerate = false;
  for (int i = 0; i < 4; i++) { fr.u[i] = true; fr.bc[i] = 0; }
  int oABC = outside_plane(p, a, b, c, d), oACD = outside_plane(p, a, c, d, b);
  int oADB = outside_plane(p, a, d, b, c), oBDC = outside_plane(p, b, d, c, a);
  if (oABC < 0 || oACD < 0 || oADB < 0 || oBDC < 0) { fr.degenerate = true; return false; }
  if (!oABC && !oACD && !oADB && !oBDC) return false;
  double best = 3.402823466e+38;
  SubRes t;
  // face -> which tetrahedron vertex each triangle slot maps to
  const int map[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
  const int on[4] = {oABC, oACD, oADB, oBDC};
  V3 vv[4] = {a, b, c, d};
  for (int f = 0; f < 4; f++) {
    if (!on[f]) continue;
    closest_on_triangle(p, vv[map[f][0]], vv[map[f][1]], vv[map[f][2]], t);
    V3 q = t.closest; double sd = dot(q - p, q - p);
    if (sd < best) {
      best = sd; fr.closest = q;
      for (int i = 0; i < 4; i++) { fr.u[i] = false; fr.bc[i] = 0; }
      for (int k = 0; k < 3; k++) { fr.u[map[f][k]] = t.u[k]; fr.bc[map[f][k]] = t.bc[k]; }
    }
  }
  return true;
}

PG_HDN bool vs_update(VSimplex& s) {
  if (!s.dirty) return s.valid;
  s.dirty = false;
  for (int i = 0; i < 4; i++) { s.bc[i] = 0; s.used[i] = false; }
  s.degenerate = false;
  if (s.n == 0) s.valid = false;
  else if (s.n == 1) { s.cP1 = s.P[0]; s.cP2 = s.Q[0]; s.cV = s.cP1 - s.cP2; s.bc[0] = 1; s.valid = vs_bc_valid(s); }
  else if (s.n == 2) {
    V3 from = s.W[0], to = s.W[1];
    V3 diff = V3() - from, v = to - from;
    double t = dot(v, diff);
    if (t > 0) {
      double dvv = dot(v, v);
      if (t < dvv) { t /= dvv; s.used[0] = s.used[1] = true; }
      else { t = 1; s.used[1] = true; }
    } else { t = 0; s.used[0] = true; }
    s.bc[0] = 1 - t; s.bc[1] = t;
    s.cP1 = s.P[0] + (s.P[1] - s.P[0]) * t; s.cP2 = s.Q[0] + (s.Q[1] - s.Q[0]) * t; s.cV = s.cP1 - s.cP2;
    vs_reduce(s); s.valid = vs_bc_valid(s);
  } else if (s.n == 3) {
    SubRes r; closest_on_triangle(V3(), s.W[0], s.W[1], s.W[2], r);
    for (int i = 0; i < 4; i++) { s.bc[i] = r.bc[i]; s.used[i] = r.u[i]; }
    s.cP1 = s.P[0] * s.bc[0] + s.P[1] * s.bc[1] + s.P[2] * s.bc[2];
    s.cP2 = s.Q[0] * s.bc[0] + s.Q[1] * s.bc[1] + s.Q[2] * s.bc[2];
    s.cV = s.cP1 - s.cP2; vs_reduce(s); s.valid = vs_bc_valid(s);
  } else {
    SubRes r; bool sep = closest_on_tetrahedron(V3(), s.W[0], s.W[1], s.W[2], s.W[3], r);
    for (int i = 0; i < 4; i++) { s.bc[i] = r.bc[i]; s.used[i] = r.u[i]; }
    s.degenerate = r.degenerate;
    if (sep) {
      s.cP1 = s.P[0] * s.bc[0] + s.P[1] * s.bc[1] + s.P[2] * s.bc[2] + s.P[3] * s.bc[3];
      s.cP2 = s.Q[0] * s.bc[0] + s.Q[1] * s.bc[1] + s.Q[2] * s.bc[2] + s.Q[3] * s.bc[3];
      s.cV = s.cP1 - s.cP2; vs_reduce(s);
      s.valid = vs_bc_valid(s);
    } else if (s.degenerate) s.valid = false;
    else { s.valid = true; s.cV = V3(); }
  }
  return s.valid;
}

// ------------------------------------------------------------------ penetration query (margin-inclusive shapes)
enum { EPA_MAXV = 128, EPA_MAXF = 256, SV_GJK = 4, SV_TOTAL = SV_GJK + EPA_MAXV, EXPAND_STACK = 96 };
#define PG_EPA_ACC 1e-12
#define PG_EPA_PLANE_EPS 1e-14

struct Mink {
  const Shape* s0; const Shape* s1;
  M3 to1, to0; V3 to0_o;
  bool margins;
};
PG_HDN V3 mink_ls(const Mink& md, const Shape& s, V3 d) {
  if (!md.margins) return support_core(s, d);
  V3 n = d;
  if (len2(n) < 2.220446049250313e-16 * 2.220446049250313e-16) n = V3(-1, -1, -1);
  n = vdiv(n, len(n));
  return support_core(s, n) + n * s.margin;
}
PG_HD V3 mink_support0(const Mink& md, V3 d) { return mink_ls(md, *md.s0, d); }
PG_HDN V3 mink_support1(const Mink& md, V3 d) { return mul(md.to0, mink_ls(md, *md.s1, mul(md.to1, d))) + md.to0_o; }
PG_HD V3 mink_support(const Mink& md, V3 d) { return mink_support0(md, d) - mink_support1(md, -d); }
PG_HD M3 mulTM(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = dot(a.col(i), b.col(j));
  return r;
}
PG_HDN void mink_init(Mink& md, const Shape& s0, const Xf& x0, const Shape& s1, const Xf& x1, bool margins) {
  md.s0 = &s0; md.s1 = &s1;
  md.to1 = mulTM(x1.R, x0.R); md.to0 = mulTM(x0.R, x1.R); md.to0_o = mulT(x0.R, x1.o - x0.o);
  md.margins = margins;
}

struct SV { V3 d, w; };
struct Face { V3 n; double d; short f[3], prev, next; unsigned char c[3], e[3], pass; };

// all per-query scratch of one thread
struct Scratch {
  SV sv[SV_TOTAL];
  Face fc[EPA_MAXF];
  short stk_f[EXPAND_STACK]; unsigned char stk_e[EXPAND_STACK], stk_s[EXPAND_STACK];
};

struct GSimplex { unsigned char c[4]; double p[4]; int rank; };
struct Gjk2 {
  const Mink* md; SV* sv;
  V3 ray;
  GSimplex sx[2];
  unsigned char free_[4]; int nfree, current;
  int status;   // 0 valid, 1 inside, 2 failed
};

PG_HD double det3(V3 a, V3 b, V3 c) {
  return a.y * b.z * c.x + a.z * b.x * c.y - a.x * b.z * c.y - a.y * b.x * c.z + a.x * b.y * c.z - a.z * b.y * c.x;
}
PG_HDN void g_support(const Gjk2& g, V3 d, SV& sv) { sv.d = vdiv(d, len(d)); sv.w = mink_support(*g.md, sv.d); }
PG_HD void g_remove(Gjk2& g, GSimplex& s) { g.free_[g.nfree++] = s.c[--s.rank]; }
PG_HD void g_append(Gjk2& g, GSimplex& s, V3 v) { s.p[s.rank] = 0; s.c[s.rank] = g.free_[--g.nfree]; g_support(g, v, g.sv[s.c[s.rank]]); s.rank++; }

PG_HD double project2(V3 a, V3 b, double* w, unsigned& m) {
  V3 d = b - a; double l = len2(d);
  if (l > 0) {
    double t = -dot(a, d) / l;
    if (t >= 1) { w[0] = 0; w[1] = 1; m = 2; return len2(b); }
    else if (t <= 0) { w[0] = 1; w[1] = 0; m = 1; return len2(a); }
    else { w[1] = t; w[0] = 1 - t; m = 3; return len2(a + d * t); }
  }
  return -1;
}
PG_HDN double project3(V3 a, V3 b, V3 c, double* w, unsigned& m) {
  const int imd3[3] = {1, 2, 0};
  V3 vt[3] = {a, b, c};
  V3 dl[3] = {a - b, b - c, c - a};
  V3 n = cross(dl[0], dl[1]);
  double l = len2(n);
  if (l > 0) {
    double mindist = -1, subw[2] = {0, 0}; unsigned subm = 0;
    for (int i = 0; i < 3; i++) {
      if (dot(vt[i], cross(dl[i], n)) > 0) {
        int j = imd3[i];
        double subd = project2(vt[i], vt[j], subw, subm);
        if (mindist < 0 || subd < mindist) {
          mindist = subd;
          m = ((subm & 1) ? 1u << i : 0) + ((subm & 2) ? 1u << j : 0);
          w[i] = subw[0]; w[j] = subw[1]; w[imd3[j]] = 0;
        }
      }
    }
    if (mindist < 0) {
      double d = dot(a, n), s = sqrt(l);
      V3 p = n * (d / l);
      mindist = len2(p); m = 7;
      w[0] = len(cross(dl[1], b - p)) / s;
      w[1] = len(cross(dl[2], c - p)) / s;
      w[2] = 1 - (w[0] + w[1]);
    }
    return mindist;
  }
  return -1;
}
PG_HDN double project4(V3 a, V3 b, V3 c, V3 d, double* w, unsigned& m) {
  const int imd3[3] = {1, 2, 0};
  V3 vt[4] = {a, b, c, d};
  V3 dl[3] = {a - d, b - d, c - d};
  double vl = det3(dl[0], dl[1], dl[2]);
  bool ng = (vl * dot(a, cross(b - c, a - b))) <= 0;
  if (ng && fabs(vl) > 0) {
    double mindist = -1, subw[3] = {0, 0, 0}; unsigned subm = 0;
    for (int i = 0; i < 3; i++) {
      int j = imd3[i];
      double s = vl * dot(d, cross(dl[i], dl[j]));
      if (s > 0) {
        double subd = project3(vt[i], vt[j], d, subw, subm);
        if (mindist < 0 || subd < mindist) {
          mindist = subd;
          m = ((subm & 1) ? 1u << i : 0) + ((subm & 2) ? 1u << j : 0) + ((subm & 4) ? 8 : 0);
          w[i] = subw[0]; w[j] = subw[1]; w[imd3[j]] = 0; w[3] = subw[2];
        }
      }
    }
    if (mindist < 0) {
      mindist = 0; m = 15;
      w[0] = det3(c, b, d) / vl; w[1] = det3(a, c, d) / vl; w[2] = det3(b, a, d) / vl;
      w[3] = 1 - (w[0] + w[1] + w[2]);
    }
    return mindist;
  }
  return -1;
}

PG_HDN int gjk2_evaluate(Gjk2& g, const Mink& md, SV* sv, V3 guess) {
  int iterations = 0; double alpha = 0; V3 lastw[4]; int clastw = 0;
  g.md = &md; g.sv = sv;
  for (int i = 0; i < 4; i++) g.free_[i] = (unsigned char)i;
  g.nfree = 4; g.current = 0; g.status = 0;
  g.sx[0].rank = 0; g.ray = guess;
  double sqrl = len2(g.ray);
  g_append(g, g.sx[0], sqrl > 0 ? -g.ray : V3(1, 0, 0));
  g.sx[0].p[0] = 1; g.ray = sv[g.sx[0].c[0]].w;
  lastw[0] = lastw[1] = lastw[2] = lastw[3] = g.ray;
  do {
    int next = 1 - g.current;
    GSimplex& cs = g.sx[g.current]; GSimplex& ns = g.sx[next];
    double rl = len(g.ray);
    if (rl < 1e-12) { g.status = 1; break; }
    g_append(g, cs, -g.ray);
    V3 w = sv[cs.c[cs.rank - 1]].w;
    bool found = false;
    for (int i = 0; i < 4; i++) if (len2(w - lastw[i]) < 1e-12) { found = true; break; }
    if (found) { g_remove(g, cs); break; }
    clastw = (clastw + 1) & 3; lastw[clastw] = w;
    double omega = dot(g.ray, w) / rl;
    alpha = omega > alpha ? omega : alpha;
    if (((rl - alpha) - (1e-12 * rl)) <= 0) { g_remove(g, cs); break; }
    double weights[4], sqdist = -1; unsigned mask = 0;
    if (cs.rank == 2) sqdist = project2(sv[cs.c[0]].w, sv[cs.c[1]].w, weights, mask);
    else if (cs.rank == 3) sqdist = project3(sv[cs.c[0]].w, sv[cs.c[1]].w, sv[cs.c[2]].w, weights, mask);
    else if (cs.rank == 4) sqdist = project4(sv[cs.c[0]].w, sv[cs.c[1]].w, sv[cs.c[2]].w, sv[cs.c[3]].w, weights, mask);
    if (sqdist >= 0) {
      ns.rank = 0; g.ray = V3(); g.current = next;
      for (int i = 0, ni = cs.rank; i < ni; i++) {
        if (mask & (1u << i)) { ns.c[ns.rank] = cs.c[i]; ns.p[ns.rank++] = weights[i]; g.ray += sv[cs.c[i]].w * weights[i]; }
        else g.free_[g.nfree++] = cs.c[i];
      }
      if (mask == 15) g.status = 1;
    } else { g_remove(g, cs); break; }
    g.status = ((++iterations) < 128) ? g.status : 2;
  } while (g.status == 0);
  return g.status;
}

// does the simplex (grown with axis-aligned / normal probes) enclose the origin?
PG_HD bool enclose4(Gjk2& g, GSimplex& s) {
  const SV* sv = g.sv;
  return fabs(det3(sv[s.c[0]].w - sv[s.c[3]].w, sv[s.c[1]].w - sv[s.c[3]].w, sv[s.c[2]].w - sv[s.c[3]].w)) > 0;
}
PG_HDN bool enclose3(Gjk2& g, GSimplex& s) {
  const SV* sv = g.sv;
  V3 n = cross(sv[s.c[1]].w - sv[s.c[0]].w, sv[s.c[2]].w - sv[s.c[0]].w);
  if (len2(n) > 0) {
    g_append(g, s, n); if (enclose4(g, s)) return true; g_remove(g, s);
    g_append(g, s, -n); if (enclose4(g, s)) return true; g_remove(g, s);
  }
  return false;
}
PG_HDN bool enclose2(Gjk2& g, GSimplex& s) {
  const SV* sv = g.sv;
  V3 d = sv[s.c[1]].w - sv[s.c[0]].w;
  for (int i = 0; i < 3; i++) {
    V3 axis; axis[i] = 1;
    V3 p = cross(d, axis);
    if (len2(p) > 0) {
      g_append(g, s, p); if (enclose3(g, s)) return true; g_remove(g, s);
      g_append(g, s, -p); if (enclose3(g, s)) return true; g_remove(g, s);
    }
  }
  return false;
}
PG_HD bool enclose_origin(Gjk2& g, GSimplex& s) {       // callers only reach this with rank >= 2
  if (s.rank == 2) return enclose2(g, s);
  if (s.rank == 3) return enclose3(g, s);
  if (s.rank == 4) return enclose4(g, s);
  return false;
}

struct Epa {
  SV* sv; Face* fc; Scratch* sc;
  short hull_root, stock_root; int hull_count, nextsv, fresh;
  int status;            // 0 valid .. 9 failed (7 accuracy reached, 6 out of vertices, 8 fallback)
  V3 normal; double depth;
  int rrank; unsigned char rc[3]; double rp[3];
};
enum { EPA_VALID = 0, EPA_DEGENERATED = 2, EPA_NONCONVEX = 3, EPA_INVALIDHULL = 4, EPA_OUTOFFACES = 5, EPA_OUTOFVERTICES = 6,
       EPA_ACCURACY = 7, EPA_FALLBACK = 8, EPA_FAILED = 9 };

PG_HD void fl_append(Epa& e, short& root, short f) {
  e.fc[f].prev = -1; e.fc[f].next = root;
  if (root >= 0) e.fc[root].prev = f;
  root = f;
}
PG_HD void fl_remove(Epa& e, short& root, short f) {
  Face& F = e.fc[f];
  if (F.next >= 0) e.fc[F.next].prev = F.prev;
  if (F.prev >= 0) e.fc[F.prev].next = F.next;
  if (f == root) root = F.next;
}
PG_HD void epa_bind(Epa& e, short fa, int ea, short fb, int eb) {
  e.fc[fa].e[ea] = (unsigned char)eb; e.fc[fa].f[ea] = fb; e.fc[fb].e[eb] = (unsigned char)ea; e.fc[fb].f[eb] = fa;
}
PG_HDN bool edge_dist(const Epa& e, const Face& face, int ia, int ib, double& dist) {
  V3 a = e.sv[ia].w, b = e.sv[ib].w;
  V3 ba = b - a, n_ab = cross(ba, face.n);
  if (dot(a, n_ab) < 0) {
    double ba_l2 = len2(ba), a_dot_ba = dot(a, ba), b_dot_ba = dot(b, ba);
    if (a_dot_ba > 0) dist = len(a);
    else if (b_dot_ba < 0) dist = len(b);
    else { double ab = dot(a, b), t = (len2(a) * len2(b) - ab * ab) / ba_l2; dist = sqrt(t > 0 ? t : 0); }
    return true;
  }
  return false;
}
PG_HDN short epa_newface(Epa& e, int a, int b, int c, bool forced) {
  // free faces: the list of returned ones first (most recently returned on top), then the never-used slots in index
  // order -- the same pop order as a stock list pre-filled with all 256 slots, without writing 256 links per query
  if (e.stock_root >= 0 || e.fresh < EPA_MAXF) {
    short f;
    if (e.stock_root >= 0) { f = e.stock_root; fl_remove(e, e.stock_root, f); }
    else f = (short)e.fresh++;
    fl_append(e, e.hull_root, f); e.hull_count++;
    Face& F = e.fc[f];
    F.pass = 0; F.c[0] = (unsigned char)a; F.c[1] = (unsigned char)b; F.c[2] = (unsigned char)c;
    F.n = cross(e.sv[b].w - e.sv[a].w, e.sv[c].w - e.sv[a].w);
    double l = len(F.n);
    if (l > PG_EPA_ACC) {
      if (!(edge_dist(e, F, a, b, F.d) || edge_dist(e, F, b, c, F.d) || edge_dist(e, F, c, a, F.d)))
        F.d = dot(e.sv[a].w, F.n) / l;
      F.n = vdiv(F.n, l);
      if (forced || F.d >= -PG_EPA_PLANE_EPS) return f;
      e.status = EPA_NONCONVEX;
    } else e.status = EPA_DEGENERATED;
    fl_remove(e, e.hull_root, f); e.hull_count--; fl_append(e, e.stock_root, f);
    return -1;
  }
  e.status = EPA_OUTOFFACES;
  return -1;
}
PG_HDN short epa_findbest(const Epa& e) {
  short minf = e.hull_root; double mind = e.fc[minf].d * e.fc[minf].d;
  for (short f = e.fc[minf].next; f >= 0; f = e.fc[f].next) { double sq = e.fc[f].d * e.fc[f].d; if (sq < mind) { minf = f; mind = sq; } }
  return minf;
}
struct Horizon { short cf, ff; int nf; };

// silhouette walk from face f0 across edge e0 (explicit stack instead of recursion)
PG_HDN bool epa_expand(Epa& e, unsigned pass, int w, short f0, int e0, Horizon& h) {
  const int i1m3[3] = {1, 2, 0}, i2m3[3] = {2, 0, 1};
  Scratch& sc = *e.sc;
  int sp = 0; bool ret = false;
  sc.stk_f[0] = f0; sc.stk_e[0] = (unsigned char)e0; sc.stk_s[0] = 0; sp = 1;
  while (sp > 0) {
    short f = sc.stk_f[sp - 1]; int ed = sc.stk_e[sp - 1]; int stage = sc.stk_s[sp - 1];
    Face& F = e.fc[f];
    if (stage == 0) {
      if (F.pass == (unsigned char)pass) { ret = false; sp--; continue; }
      int e1 = i1m3[ed];
      if ((dot(F.n, e.sv[w].w) - F.d) < -PG_EPA_PLANE_EPS) {
        short nf = epa_newface(e, F.c[e1], F.c[ed], w, false);
        if (nf >= 0) {
          epa_bind(e, nf, 0, f, ed);
          if (h.cf >= 0) epa_bind(e, h.cf, 1, nf, 2); else h.ff = nf;
          h.cf = nf; ++h.nf;
          ret = true;
        } else ret = false;
        sp--; continue;
      }
      F.pass = (unsigned char)pass;
      sc.stk_s[sp - 1] = 1;
      if (sp >= EXPAND_STACK) { ret = false; sp = 0; break; }
      sc.stk_f[sp] = F.f[e1]; sc.stk_e[sp] = F.e[e1]; sc.stk_s[sp] = 0; sp++;
    } else if (stage == 1) {
      if (!ret) { sp--; continue; }
      int e2 = i2m3[ed];
      sc.stk_s[sp - 1] = 2;
      if (sp >= EXPAND_STACK) { ret = false; sp = 0; break; }
      sc.stk_f[sp] = F.f[e2]; sc.stk_e[sp] = F.e[e2]; sc.stk_s[sp] = 0; sp++;
    } else {
      if (ret) { fl_remove(e, e.hull_root, f); e.hull_count--; fl_append(e, e.stock_root, f); }
      sp--;
    }
  }
  return ret;
}

PG_HDN int epa_evaluate(Epa& e, Gjk2& g, Scratch& sc, V3 guess) {
  e.sc = &sc; e.sv = sc.sv; e.fc = sc.fc;
  GSimplex& s = g.sx[g.current];
  e.status = EPA_FAILED;
  if (s.rank > 1 && enclose_origin(g, s)) {
    e.hull_root = -1; e.stock_root = -1; e.hull_count = 0; e.fresh = 0;
    e.status = EPA_VALID; e.nextsv = 0;
    const SV* sv = e.sv;
    if (det3(sv[s.c[0]].w - sv[s.c[3]].w, sv[s.c[1]].w - sv[s.c[3]].w, sv[s.c[2]].w - sv[s.c[3]].w) < 0) {
      unsigned char tc = s.c[0]; s.c[0] = s.c[1]; s.c[1] = tc;
      double tp = s.p[0]; s.p[0] = s.p[1]; s.p[1] = tp;
    }
    short t0 = epa_newface(e, s.c[0], s.c[1], s.c[2], true), t1 = epa_newface(e, s.c[1], s.c[0], s.c[3], true);
    short t2 = epa_newface(e, s.c[2], s.c[1], s.c[3], true), t3 = epa_newface(e, s.c[0], s.c[2], s.c[3], true);
    if (e.hull_count == 4) {
      short best = epa_findbest(e);
      Face outer = e.fc[best];
      unsigned pass = 0; int iterations = 0;
      epa_bind(e, t0, 0, t1, 0); epa_bind(e, t0, 1, t2, 0); epa_bind(e, t0, 2, t3, 0);
      epa_bind(e, t1, 1, t3, 2); epa_bind(e, t1, 2, t2, 1); epa_bind(e, t2, 2, t3, 1);
      e.status = EPA_VALID;
      for (; iterations < 255; ++iterations) {
        if (e.nextsv < EPA_MAXV) {
          Horizon h; h.cf = -1; h.ff = -1; h.nf = 0;
          int w = SV_GJK + e.nextsv++;
          bool valid = true;
          e.fc[best].pass = (unsigned char)(++pass);
          g_support(g, e.fc[best].n, e.sv[w]);
          double wdist = dot(e.fc[best].n, e.sv[w].w) - e.fc[best].d;
          if (wdist > PG_EPA_ACC) {
            for (int j = 0; j < 3 && valid; ++j) valid = valid && epa_expand(e, pass, w, e.fc[best].f[j], e.fc[best].e[j], h);
            if (valid && h.nf >= 3) {
              epa_bind(e, h.cf, 1, h.ff, 2);
              fl_remove(e, e.hull_root, best); e.hull_count--; fl_append(e, e.stock_root, best);
              best = epa_findbest(e); outer = e.fc[best];
            } else { e.status = EPA_INVALIDHULL; break; }
          } else { e.status = EPA_ACCURACY; break; }
        } else { e.status = EPA_OUTOFVERTICES; break; }
      }
      V3 projection = outer.n * outer.d;
      e.normal = outer.n; e.depth = outer.d;
      e.rrank = 3; e.rc[0] = outer.c[0]; e.rc[1] = outer.c[1]; e.rc[2] = outer.c[2];
      e.rp[0] = len(cross(e.sv[outer.c[1]].w - projection, e.sv[outer.c[2]].w - projection));
      e.rp[1] = len(cross(e.sv[outer.c[2]].w - projection, e.sv[outer.c[0]].w - projection));
      e.rp[2] = len(cross(e.sv[outer.c[0]].w - projection, e.sv[outer.c[1]].w - projection));
      double sum = e.rp[0] + e.rp[1] + e.rp[2];
      e.rp[0] /= sum; e.rp[1] /= sum; e.rp[2] /= sum;
      return e.status;
    }
  }
  e.status = EPA_FALLBACK;
  e.normal = -guess;
  double nl = len(e.normal);
  if (nl > 0) e.normal = vdiv(e.normal, nl); else e.normal = V3(1, 0, 0);
  e.depth = 0; e.rrank = 1; e.rc[0] = s.c[0]; e.rp[0] = 1;
  return e.status;
}

PG_HD V3 xf_apply(const Xf& x, V3 p) { return mul(x.R, p) + x.o; }

PG_HDN bool penetration(const Shape& s0, const Xf& x0, const Shape& s1, const Xf& x1, V3 guess, Scratch& sc, V3& w0out, V3& w1out, V3& nrm) {
  Mink md; mink_init(md, s0, x0, s1, x1, true);
  Gjk2 g;
  int st = gjk2_evaluate(g, md, sc.sv, -guess);
  if (st == 1) {
    Epa e;
    int es = epa_evaluate(e, g, sc, -guess);
    if (es != EPA_FAILED) {
      V3 w0;
      for (int i = 0; i < e.rrank; i++) w0 += mink_support0(md, sc.sv[e.rc[i]].d) * e.rp[i];
      w0out = xf_apply(x0, w0);
      w1out = xf_apply(x0, w0 - e.normal * e.depth);
      nrm = -e.normal;
      return true;
    }
  }
  return false;
}

PG_HDN bool core_distance(const Shape& s0, const Xf& x0, const Shape& s1, const Xf& x1, V3 guess, Scratch& sc, V3& w0out, V3& w1out, V3& nrm) {
  Mink md; mink_init(md, s0, x0, s1, x1, false);
  Gjk2 g;
  int st = gjk2_evaluate(g, md, sc.sv, guess);
  if (st == 0) {
    V3 w0, w1;
    const GSimplex& s = g.sx[g.current];
    for (int i = 0; i < s.rank; i++) {
      w0 += mink_support0(md, sc.sv[s.c[i]].d) * s.p[i];
      w1 += mink_support1(md, -sc.sv[s.c[i]].d) * s.p[i];
    }
    w0out = xf_apply(x0, w0); w1out = xf_apply(x0, w1);
    nrm = w0 - w1;
    double dist = len(nrm);
    nrm = vdiv(nrm, dist > 1e-12 ? dist : 1);
    return true;
  }
  return false;
}

PG_HD V3 safe_normalize(V3 v) {
  double l2 = len2(v);
  if (l2 >= 2.220446049250313e-16 * 2.220446049250313e-16) return vdiv(v, sqrt(l2));
  return V3(1, 0, 0);
}

PG_HDN bool pen_depth(const Shape& sA, const Xf& xA, const Shape& sB, const Xf& xB, Scratch& sc, V3& v, V3& wA, V3& wB) {
  for (int i = 0; i < 9; i++) {
    V3 gdir;
    switch (i) {
      case 0: gdir = safe_normalize(xB.o - xA.o); break;
      case 1: gdir = safe_normalize(xA.o - xB.o); break;
      case 2: gdir = V3(0, 0, 1); break;
      case 3: gdir = V3(0, 1, 0); break;
      case 4: gdir = V3(1, 0, 0); break;
      case 5: gdir = vdiv(V3(1, 1, 0), sqrt(2.0)); break;
      case 6: gdir = vdiv(V3(1, 1, 1), sqrt(3.0)); break;
      case 7: gdir = vdiv(V3(0, 1, 1), sqrt(2.0)); break;
      default: gdir = vdiv(V3(1, 0, 1), sqrt(2.0)); break;
    }
    if (penetration(sA, xA, sB, xB, gdir, sc, wA, wB, v)) return true;
    if (core_distance(sA, xA, sB, xB, gdir, sc, wA, wB, v)) return false;
  }
  wA = V3(); wB = V3(); v = V3();
  return false;
}

// ------------------------------------------------------------------ one contact point per pair and substep
// returns true and (normal on B in world, point on B in world, signed distance) when a point should be added
PG_HDN bool closest_points(const Shape& sA, V3 posA, const M3& RA, const Shape& sB, V3 posB, const M3& RB, double breaking,
                           Scratch& sc, V3& normal_out, V3& point_b_out, double& dist_out) {
  // REL_ERROR2 and gGjkEpaPenetrationTolerance of a BT_USE_DOUBLE_PRECISION build (btGjkPairDetector.cpp): 1e-12 both (the
  // single-precision 1e-6 / 0.001 sent touching pairs with cores less than 1 mm apart to EPA; plate_20.0 says no)
  const double REL2 = 1.0e-12, PEN_TOL = 1.0e-12, EPS = 2.220446049250313e-16;
  double mA = sA.margin, mB = sB.margin, margin = mA + mB;
  double maxd = margin + breaking, max_dist2 = maxd * maxd;
  V3 offset = (posA + posB) * 0.5;
  Xf xA, xB; xA.o = posA - offset; xA.R = RA; xB.o = posB - offset; xB.R = RB;
  V3 axis(0, 1, 0), pointOnA, pointOnB, normalInB, orgNormalInB;
  double sqd = 1e30, distance = 0;
  bool is_valid = false, check_simplex = false;
  int degenerate = 0, iter = 0;
  VSimplex sx; sx.n = 0; sx.lastW = V3(1e30, 1e30, 1e30); sx.dirty = true; sx.valid = false; sx.degenerate = false;
  for (int i = 0; i < 4; i++) { sx.bc[i] = 0; sx.used[i] = false; }
  for (;;) {
    V3 pW = xf_apply(xA, support_core(sA, mulT(RA, -axis)));
    V3 qW = xf_apply(xB, support_core(sB, mulT(RB, axis)));
    V3 wv = pW - qW;
    double delta = dot(axis, wv);
    if (delta > 0 && delta * delta > sqd * max_dist2) { degenerate = 10; check_simplex = true; break; }
    bool in_simplex = false;
    for (int i = 0; i < sx.n; i++) if (len2(sx.W[i] - wv) <= 1e-12) in_simplex = true;
    if (wv.x == sx.lastW.x && wv.y == sx.lastW.y && wv.z == sx.lastW.z) in_simplex = true;
    if (in_simplex) { degenerate = 1; check_simplex = true; break; }
    double f0 = sqd - delta, f1 = sqd * REL2;
    if (f0 <= f1) { degenerate = (f0 <= 0) ? 2 : 11; check_simplex = true; break; }
    sx.lastW = wv; sx.dirty = true; sx.W[sx.n] = wv; sx.P[sx.n] = pW; sx.Q[sx.n] = qW; sx.n++;
    bool ok = vs_update(sx);
    V3 nv = sx.cV;
    if (!ok) { degenerate = 3; check_simplex = true; break; }
    if (len2(nv) < REL2) { axis = nv; degenerate = 6; check_simplex = true; break; }
    double prev = sqd;
    sqd = len2(nv);
    if (prev - sqd <= EPS * prev) { check_simplex = true; degenerate = 12; break; }
    axis = nv;
    if (iter++ > 1000) break;
    if (sx.n == 4) { degenerate = 13; break; }
  }
  if (check_simplex) {
    vs_update(sx);
    pointOnA = sx.cP1; pointOnB = sx.cP2;
    normalInB = axis;
    double lsq = len2(axis);
    if (lsq < 1e-4) degenerate = 5;
    if (lsq > EPS * EPS) {
      double rlen = 1.0 / sqrt(lsq);
      normalInB = normalInB * rlen;
      double s = sqrt(lsq);
      pointOnA = pointOnA - axis * (mA / s);
      pointOnB = pointOnB + axis * (mB / s);
      distance = (1.0 / rlen) - margin;
      is_valid = true;
      orgNormalInB = normalInB;
    }
  }
  bool catch_degenerate = degenerate && ((distance + margin) < PEN_TOL);
  if (!is_valid || catch_degenerate) {
    V3 tA, tB;
    axis = V3();
    bool valid2 = pen_depth(sA, xA, sB, xB, sc, axis, tA, tB);
    if (len2(axis) != 0) {
      if (valid2) {
        V3 tn = tB - tA; double lsq = len2(tn);
        if (lsq <= EPS * EPS) { tn = axis; lsq = len2(axis); }
        if (lsq > EPS * EPS) {
          tn = vdiv(tn, sqrt(lsq));
          double d2 = -len(tA - tB);
          if (!is_valid || d2 < distance) { distance = d2; pointOnA = tA; pointOnB = tB; normalInB = tn; is_valid = true; }
        }
      } else {
        double d2 = len(tA - tB) - margin;
        if (!is_valid || d2 < distance) {
          distance = d2; pointOnA = tA - axis * mA; pointOnB = tB + axis * mB;
          normalInB = vdiv(axis, len(axis)); is_valid = true;
        }
      }
    }
  }
  if (is_valid && (distance < 0 || distance * distance < max_dist2)) {
    V3 n = normalInB;
    V3 p = xf_apply(xA, support_core(sA, mulT(RA, -n))), q = xf_apply(xB, support_core(sB, mulT(RB, n)));
    double d0 = dot(n, p - q) - margin;
    V3 nn = -n;
    p = xf_apply(xA, support_core(sA, mulT(RA, -nn))); q = xf_apply(xB, support_core(sB, mulT(RB, nn)));
    double d1 = dot(nn, p - q) - margin;
    if (d1 > d0) normalInB = -normalInB;
    if (len2(orgNormalInB) != 0) {            // the detector's last check: the GJK normal wins when it separates more than the EPA result
      V3 no = orgNormalInB;
      p = xf_apply(xA, support_core(sA, mulT(RA, -no))); q = xf_apply(xB, support_core(sB, mulT(RB, no)));
      double d2 = dot(no, p - q) - margin;
      if (d2 > d0 && d2 > d1 && d2 > distance) { normalInB = orgNormalInB; distance = d2; }
    }
    normal_out = normalInB; point_b_out = pointOnB + offset; dist_out = distance;
    return true;
  }
  return false;
}

}  // namespace cvx
}  // namespace pg
