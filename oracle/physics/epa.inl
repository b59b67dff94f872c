// Penetration depth for overlapping convex pairs: a restatement, from the published algorithm, of Bullet's
// btGjkEpaSolver2::Penetration / ::Distance (btGjkEpa2.cpp: the solver's own GJK with projectorigin sub-simplex
// solvers, EncloseOrigin, and the EPA polytope expansion with horizon binding) and of
// btGjkEpaPenetrationDepthSolver::calcPenDepth (the list of nine guess vectors).  Double-precision build constants.
// Everything runs in shape 0's local frame on the MARGIN-INCLUSIVE shapes (vertex/corner + margin * direction).
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
namespace pgo {
namespace epa2 {

static const int GJK_MAX_ITERATIONS = 128;
static const double GJK_ACCURACY = 1e-12, GJK_MIN_DISTANCE = 1e-12, GJK_DUPLICATED_EPS = 1e-12;
static const int EPA_MAX_VERTICES = 128, EPA_MAX_ITERATIONS = 255, EPA_MAX_FACES = EPA_MAX_VERTICES * 2;
static double EPA_ACCURACY = 1e-12, EPA_PLANE_EPS = 1e-14;   // runtime for the what-if scans (Bullet, double precision: 1e-12 / 1e-14)

static inline V3 vdiv(V3 a, double s) { return a * (1.0 / s); }        // btVector3::operator/ multiplies by the reciprocal

struct MinkDiff {
  const Shape* s0; const Shape* s1;
  M3 toshape1, to0_basis; V3 to0_origin;
  bool margins;
  V3 ls(const Shape& s, V3 d) const {
    if (!margins) return support_core_local(s, d);
    V3 n = d;
    if (len2(n) < kEps * kEps) n = V3(-1, -1, -1);
    n = vdiv(n, len(n));
    return support_core_local(s, n) + n * s.margin;
  }
  V3 support0(V3 d) const { return ls(*s0, d); }
  V3 support1(V3 d) const { return mul(to0_basis, ls(*s1, mul(toshape1, d))) + to0_origin; }
  V3 support(V3 d) const { return support0(d) - support1(-d); }
};

static void mink_init(MinkDiff& md, const Shape& s0, const M3& R0, V3 o0, const Shape& s1, const M3& R1, V3 o1, bool margins) {
  md.s0 = &s0; md.s1 = &s1;
  md.toshape1 = mulTM(R1, R0);
  md.to0_basis = mulTM(R0, R1);
  md.to0_origin = mulT(R0, o1 - o0);
  md.margins = margins;
}

struct SV { V3 d, w; };
struct Simplex { SV* c[4]; double p[4]; int rank; };

static double det3(V3 a, V3 b, V3 c) {
  return a.y * b.z * c.x + a.z * b.x * c.y - a.x * b.z * c.y - a.y * b.x * c.z + a.x * b.y * c.z - a.z * b.y * c.x;
}

static double project2(V3 a, V3 b, double* w, unsigned& m) {
  V3 d = b - a; double l = len2(d);
  if (l > 0) {
    double t = l > 0 ? -dot(a, d) / l : 0;
    if (t >= 1) { w[0] = 0; w[1] = 1; m = 2; return len2(b); }
    else if (t <= 0) { w[0] = 1; w[1] = 0; m = 1; return len2(a); }
    else { w[0] = 1 - (w[1] = t); m = 3; return len2(a + d * t); }
  }
  return -1;
}

static double project3(V3 a, V3 b, V3 c, double* w, unsigned& m) {
  static const unsigned imd3[] = {1, 2, 0};
  const V3* vt[] = {&a, &b, &c};
  const V3 dl[] = {a - b, b - c, c - a};
  V3 n = cross(dl[0], dl[1]);
  double l = len2(n);
  if (l > 0) {
    double mindist = -1, subw[2] = {0, 0}; unsigned subm = 0;
    for (unsigned i = 0; i < 3; i++) {
      if (dot(*vt[i], cross(dl[i], n)) > 0) {
        unsigned j = imd3[i];
        double subd = project2(*vt[i], *vt[j], subw, subm);
        if (mindist < 0 || subd < mindist) {
          mindist = subd;
          m = ((subm & 1) ? 1u << i : 0) + ((subm & 2) ? 1u << j : 0);
          w[i] = subw[0]; w[j] = subw[1]; w[imd3[j]] = 0;
        }
      }
    }
    if (mindist < 0) {
      double d = dot(a, n), s = std::sqrt(l);
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

static double project4(V3 a, V3 b, V3 c, V3 d, double* w, unsigned& m) {
  static const unsigned imd3[] = {1, 2, 0};
  const V3* vt[] = {&a, &b, &c, &d};
  const V3 dl[] = {a - d, b - d, c - d};
  double vl = det3(dl[0], dl[1], dl[2]);
  bool ng = (vl * dot(a, cross(b - c, a - b))) <= 0;
  if (ng && std::fabs(vl) > 0) {
    double mindist = -1, subw[3] = {0, 0, 0}; unsigned subm = 0;
    for (unsigned i = 0; i < 3; i++) {
      unsigned j = imd3[i];
      double s = vl * dot(d, cross(dl[i], dl[j]));
      if (s > 0) {
        double subd = project3(*vt[i], *vt[j], d, subw, subm);
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

enum GjkStatus { GJK_VALID, GJK_INSIDE, GJK_FAILED };

struct Gjk {
  const MinkDiff* shape;
  V3 ray; double distance = 0;
  Simplex simplices[2]; SV store[4]; SV* free_[4]; int nfree = 0, current = 0;
  Simplex* simplex = nullptr;
  GjkStatus status = GJK_FAILED;

  void getsupport(V3 d, SV& sv) const { sv.d = vdiv(d, len(d)); sv.w = shape->support(sv.d); }
  void removevertice(Simplex& s) { free_[nfree++] = s.c[--s.rank]; }
  void appendvertice(Simplex& s, V3 v) { s.p[s.rank] = 0; s.c[s.rank] = free_[--nfree]; getsupport(v, *s.c[s.rank++]); }

  GjkStatus evaluate(const MinkDiff& sh, V3 guess) {
    int iterations = 0; double sqdist = 0, alpha = 0; V3 lastw[4]; int clastw = 0;
    for (int i = 0; i < 4; i++) free_[i] = &store[i];
    nfree = 4; current = 0; status = GJK_VALID; shape = &sh; distance = 0;
    simplices[0].rank = 0; ray = guess;
    double sqrl = len2(ray);
    appendvertice(simplices[0], sqrl > 0 ? -ray : V3(1, 0, 0));
    simplices[0].p[0] = 1; ray = simplices[0].c[0]->w; sqdist = sqrl;
    lastw[0] = lastw[1] = lastw[2] = lastw[3] = ray;
    do {
      int next = 1 - current;
      Simplex& cs = simplices[current]; Simplex& ns = simplices[next];
      double rl = len(ray);
      if (rl < GJK_MIN_DISTANCE) { status = GJK_INSIDE; break; }
      appendvertice(cs, -ray);
      V3 w = cs.c[cs.rank - 1]->w;
      bool found = false;
      for (int i = 0; i < 4; i++) if (len2(w - lastw[i]) < GJK_DUPLICATED_EPS) { found = true; break; }
      if (found) { removevertice(simplices[current]); break; }
      else lastw[clastw = (clastw + 1) & 3] = w;
      double omega = dot(ray, w) / rl;
      alpha = omega > alpha ? omega : alpha;
      if (((rl - alpha) - (GJK_ACCURACY * rl)) <= 0) { removevertice(simplices[current]); break; }
      double weights[4]; unsigned mask = 0;
      switch (cs.rank) {
        case 2: sqdist = project2(cs.c[0]->w, cs.c[1]->w, weights, mask); break;
        case 3: sqdist = project3(cs.c[0]->w, cs.c[1]->w, cs.c[2]->w, weights, mask); break;
        case 4: sqdist = project4(cs.c[0]->w, cs.c[1]->w, cs.c[2]->w, cs.c[3]->w, weights, mask); break;
      }
      if (sqdist >= 0) {
        ns.rank = 0; ray = V3(); current = next;
        for (int i = 0, ni = cs.rank; i < ni; i++) {
          if (mask & (1u << i)) { ns.c[ns.rank] = cs.c[i]; ns.p[ns.rank++] = weights[i]; ray += cs.c[i]->w * weights[i]; }
          else free_[nfree++] = cs.c[i];
        }
        if (mask == 15) status = GJK_INSIDE;
      } else { removevertice(simplices[current]); break; }
      status = ((++iterations) < GJK_MAX_ITERATIONS) ? status : GJK_FAILED;
    } while (status == GJK_VALID);
    simplex = &simplices[current];
    if (status == GJK_VALID) distance = len(ray); else if (status == GJK_INSIDE) distance = 0;
    return status;
  }

  bool enclose_origin() {
    switch (simplex->rank) {
      case 1:
        for (int i = 0; i < 3; i++) {
          V3 axis; axis[i] = 1;
          appendvertice(*simplex, axis); if (enclose_origin()) return true; removevertice(*simplex);
          appendvertice(*simplex, -axis); if (enclose_origin()) return true; removevertice(*simplex);
        }
        break;
      case 2: {
        V3 d = simplex->c[1]->w - simplex->c[0]->w;
        for (int i = 0; i < 3; i++) {
          V3 axis; axis[i] = 1;
          V3 p = cross(d, axis);
          if (len2(p) > 0) {
            appendvertice(*simplex, p); if (enclose_origin()) return true; removevertice(*simplex);
            appendvertice(*simplex, -p); if (enclose_origin()) return true; removevertice(*simplex);
          }
        }
        break;
      }
      case 3: {
        V3 n = cross(simplex->c[1]->w - simplex->c[0]->w, simplex->c[2]->w - simplex->c[0]->w);
        if (len2(n) > 0) {
          appendvertice(*simplex, n); if (enclose_origin()) return true; removevertice(*simplex);
          appendvertice(*simplex, -n); if (enclose_origin()) return true; removevertice(*simplex);
        }
        break;
      }
      case 4:
        if (std::fabs(det3(simplex->c[0]->w - simplex->c[3]->w, simplex->c[1]->w - simplex->c[3]->w, simplex->c[2]->w - simplex->c[3]->w)) > 0) return true;
        break;
    }
    return false;
  }
};

static int g_epa_trace = 0;
enum EpaStatus { EPA_VALID, EPA_TOUCHING, EPA_DEGENERATED, EPA_NONCONVEX, EPA_INVALIDHULL, EPA_OUTOFFACES, EPA_OUTOFVERTICES,
                 EPA_ACCURACYREACHED, EPA_FALLBACK, EPA_FAILED };

struct Face { V3 n; double d; SV* c[3]; Face* f[3]; Face* l[2]; unsigned char e[3]; unsigned char pass; };
struct FaceList { Face* root = nullptr; int count = 0; };
struct Horizon { Face* cf = nullptr; Face* ff = nullptr; int nf = 0; };

struct Epa {
  EpaStatus status = EPA_FAILED;
  Simplex result; V3 normal; double depth = 0;
  SV sv_store[EPA_MAX_VERTICES]; Face fc_store[EPA_MAX_FACES]; int nextsv = 0;
  FaceList hull, stock;

  Epa() { for (int i = 0; i < EPA_MAX_FACES; i++) append(stock, &fc_store[EPA_MAX_FACES - i - 1]); }
  static void bind(Face* fa, int ea, Face* fb, int eb) {
    fa->e[ea] = (unsigned char)eb; fa->f[ea] = fb; fb->e[eb] = (unsigned char)ea; fb->f[eb] = fa;
  }
  static void append(FaceList& list, Face* face) {
    face->l[0] = nullptr; face->l[1] = list.root;
    if (list.root) list.root->l[0] = face;
    list.root = face; ++list.count;
  }
  static void remove(FaceList& list, Face* face) {
    if (face->l[1]) face->l[1]->l[0] = face->l[0];
    if (face->l[0]) face->l[0]->l[1] = face->l[1];
    if (face == list.root) list.root = face->l[1];
    --list.count;
  }
  static bool getedgedist(Face* face, SV* a, SV* b, double& dist) {
    V3 ba = b->w - a->w, n_ab = cross(ba, face->n);
    double a_dot_nab = dot(a->w, n_ab);
    if (a_dot_nab < 0) {
      double ba_l2 = len2(ba), a_dot_ba = dot(a->w, ba), b_dot_ba = dot(b->w, ba);
      if (a_dot_ba > 0) dist = len(a->w);
      else if (b_dot_ba < 0) dist = len(b->w);
      else {
        double a_dot_b = dot(a->w, b->w), t = (len2(a->w) * len2(b->w) - a_dot_b * a_dot_b) / ba_l2;
        dist = std::sqrt(t > 0 ? t : 0);
      }
      return true;
    }
    return false;
  }
  Face* newface(SV* a, SV* b, SV* c, bool forced) {
    if (stock.root) {
      Face* face = stock.root;
      remove(stock, face); append(hull, face);
      face->pass = 0; face->c[0] = a; face->c[1] = b; face->c[2] = c;
      face->n = cross(b->w - a->w, c->w - a->w);
      double l = len(face->n);
      bool v = l > EPA_ACCURACY;
      if (v) {
        if (!(getedgedist(face, a, b, face->d) || getedgedist(face, b, c, face->d) || getedgedist(face, c, a, face->d)))
          face->d = dot(a->w, face->n) / l;
        face->n = vdiv(face->n, l);
        if (forced || face->d >= -EPA_PLANE_EPS) return face;
        else status = EPA_NONCONVEX;
      } else status = EPA_DEGENERATED;
      remove(hull, face); append(stock, face);
      return nullptr;
    }
    status = stock.root ? EPA_OUTOFVERTICES : EPA_OUTOFFACES;
    return nullptr;
  }
  Face* findbest() {
    Face* minf = hull.root; double mind = minf->d * minf->d;
    for (Face* f = minf->l[1]; f; f = f->l[1]) { double sqd = f->d * f->d; if (sqd < mind) { minf = f; mind = sqd; } }
    return minf;
  }
  bool expand(unsigned pass, SV* w, Face* f, unsigned e, Horizon& horizon) {
    static const unsigned i1m3[] = {1, 2, 0}, i2m3[] = {2, 0, 1};
    if (f->pass != pass) {
      unsigned e1 = i1m3[e];
      if ((dot(f->n, w->w) - f->d) < -EPA_PLANE_EPS) {
        Face* nf = newface(f->c[e1], f->c[e], w, false);
        if (nf) {
          bind(nf, 0, f, e);
          if (horizon.cf) bind(horizon.cf, 1, nf, 2); else horizon.ff = nf;
          horizon.cf = nf; ++horizon.nf;
          return true;
        }
      } else {
        unsigned e2 = i2m3[e];
        f->pass = (unsigned char)pass;
        if (expand(pass, w, f->f[e1], f->e[e1], horizon) && expand(pass, w, f->f[e2], f->e[e2], horizon)) {
          remove(hull, f); append(stock, f);
          return true;
        }
      }
    }
    return false;
  }
  EpaStatus evaluate(Gjk& gjk, V3 guess) {
    Simplex& simplex = *gjk.simplex;
    if (simplex.rank > 1 && gjk.enclose_origin()) {
      while (hull.root) { Face* f = hull.root; remove(hull, f); append(stock, f); }
      status = EPA_VALID; nextsv = 0;
      if (det3(simplex.c[0]->w - simplex.c[3]->w, simplex.c[1]->w - simplex.c[3]->w, simplex.c[2]->w - simplex.c[3]->w) < 0) {
        std::swap(simplex.c[0], simplex.c[1]); std::swap(simplex.p[0], simplex.p[1]);
      }
      Face* tetra[] = {newface(simplex.c[0], simplex.c[1], simplex.c[2], true), newface(simplex.c[1], simplex.c[0], simplex.c[3], true),
                       newface(simplex.c[2], simplex.c[1], simplex.c[3], true), newface(simplex.c[0], simplex.c[2], simplex.c[3], true)};
      if (hull.count == 4) {
        Face* best = findbest(); Face outer = *best; unsigned pass = 0; int iterations = 0;
        bind(tetra[0], 0, tetra[1], 0); bind(tetra[0], 1, tetra[2], 0); bind(tetra[0], 2, tetra[3], 0);
        bind(tetra[1], 1, tetra[3], 2); bind(tetra[1], 2, tetra[2], 1); bind(tetra[2], 2, tetra[3], 1);
        status = EPA_VALID;
        for (; iterations < EPA_MAX_ITERATIONS; ++iterations) {
          if (nextsv < EPA_MAX_VERTICES) {
            Horizon horizon; SV* w = &sv_store[nextsv++]; bool valid = true;
            best->pass = (unsigned char)(++pass);
            gjk.getsupport(best->n, *w);
            double wdist = dot(best->n, w->w) - best->d;
            if (g_epa_trace) fprintf(stderr, "    epa it %d best d %.15g wdist %.3e n %.9f %.9f %.9f\n", iterations, best->d, wdist, best->n.x, best->n.y, best->n.z);
            if (wdist > EPA_ACCURACY) {
              for (int j = 0; j < 3 && valid; ++j) valid &= expand(pass, w, best->f[j], best->e[j], horizon);
              if (valid && horizon.nf >= 3) {
                bind(horizon.cf, 1, horizon.ff, 2);
                remove(hull, best); append(stock, best);
                best = findbest(); outer = *best;
              } else { status = EPA_INVALIDHULL; break; }
            } else { status = EPA_ACCURACYREACHED; break; }
          } else { status = EPA_OUTOFVERTICES; break; }
        }
        V3 projection = outer.n * outer.d;
        if (g_epa_trace) fprintf(stderr, "    epa final face edges %.3e %.3e %.3e\n", len(outer.c[1]->w - outer.c[0]->w), len(outer.c[2]->w - outer.c[1]->w), len(outer.c[0]->w - outer.c[2]->w));
        normal = outer.n; depth = outer.d;
        result.rank = 3; result.c[0] = outer.c[0]; result.c[1] = outer.c[1]; result.c[2] = outer.c[2];
        result.p[0] = len(cross(outer.c[1]->w - projection, outer.c[2]->w - projection));
        result.p[1] = len(cross(outer.c[2]->w - projection, outer.c[0]->w - projection));
        result.p[2] = len(cross(outer.c[0]->w - projection, outer.c[1]->w - projection));
        double sum = result.p[0] + result.p[1] + result.p[2];
        result.p[0] /= sum; result.p[1] /= sum; result.p[2] /= sum;
        g_last_epa_iters = iterations;
        return status;
      }
    }
    status = EPA_FALLBACK;
    normal = -guess;
    double nl = len(normal);
    if (nl > 0) normal = vdiv(normal, nl); else normal = V3(1, 0, 0);
    depth = 0; result.rank = 1; result.c[0] = simplex.c[0]; result.p[0] = 1;
    return status;
  }
  static int g_last_epa_iters;
};
int Epa::g_last_epa_iters = 0;

static int g_epa_debug = 0;
static int g_epa_variant = 0;
struct Results { V3 witness0, witness1, normal; double distance; int status; };   // status 0 separated 1 penetrating 2 gjk failed 3 epa failed

static bool penetration(const Shape& s0, const M3& R0, V3 o0, const Shape& s1, const M3& R1, V3 o1, V3 guess, Results& res) {
  MinkDiff md; mink_init(md, s0, R0, o0, s1, R1, o1, true);
  Gjk gjk;
  GjkStatus st = gjk.evaluate(md, -guess);
  if (st == GJK_INSIDE) {
    static thread_local Epa* epa_mem = nullptr;
    if (!epa_mem) epa_mem = new Epa();
    Epa& epa = *epa_mem;
    EpaStatus es = epa.evaluate(gjk, -guess);
    if (es != EPA_FAILED) {
      V3 w0;
      for (int i = 0; i < epa.result.rank; i++) w0 += md.support0(epa.result.c[i]->d) * epa.result.p[i];
      res.status = 1;
      res.witness0 = mul(R0, w0) + o0;
      res.witness1 = mul(R0, w0 - epa.normal * epa.depth) + o0;
      res.normal = -epa.normal;
      res.distance = -epa.depth;
      res.status = 1 + 100 * (int)es;
      if (g_epa_debug) {
        fprintf(stderr, "  epa status %d iters %d nextsv %d depth %.12g normal %.9f %.9f %.9f\n", (int)es, Epa::g_last_epa_iters, epa.nextsv, epa.depth, epa.normal.x, epa.normal.y, epa.normal.z);
        for (int i = 0; i < epa.result.rank; i++) {
          V3 s0 = md.support0(epa.result.c[i]->d), s1 = md.support1(-epa.result.c[i]->d);
          fprintf(stderr, "    p %.12f d %.6f %.6f %.6f  s0 %.6f %.6f %.6f  s1 %.6f %.6f %.6f\n", epa.result.p[i], epa.result.c[i]->d.x, epa.result.c[i]->d.y, epa.result.c[i]->d.z, s0.x, s0.y, s0.z, s1.x, s1.y, s1.z);
        }
      }
      return true;
    } else res.status = 3;
  } else if (st == GJK_FAILED) res.status = 2;
  return false;
}

static bool distance(const Shape& s0, const M3& R0, V3 o0, const Shape& s1, const M3& R1, V3 o1, V3 guess, Results& res) {
  MinkDiff md; mink_init(md, s0, R0, o0, s1, R1, o1, false);
  Gjk gjk;
  GjkStatus st = gjk.evaluate(md, guess);
  if (st == GJK_VALID) {
    V3 w0, w1;
    for (int i = 0; i < gjk.simplex->rank; i++) {
      double p = gjk.simplex->p[i];
      w0 += md.support0(gjk.simplex->c[i]->d) * p;
      w1 += md.support1(-gjk.simplex->c[i]->d) * p;
    }
    res.witness0 = mul(R0, w0) + o0; res.witness1 = mul(R0, w1) + o0;
    res.normal = w0 - w1;
    res.distance = len(res.normal);
    res.normal = vdiv(res.normal, res.distance > GJK_MIN_DISTANCE ? res.distance : 1);
    return true;
  }
  res.status = st == GJK_INSIDE ? 1 : 2;
  return false;
}

static V3 safe_normalize(V3 v) {
  double l2 = len2(v);
  if (l2 >= kEps * kEps) return vdiv(v, std::sqrt(l2));
  return V3(1, 0, 0);
}

// btGjkEpaPenetrationDepthSolver::calcPenDepth
static bool calc_pen_depth(const Shape& sA, const M3& RA, V3 oA, const Shape& sB, const M3& RB, V3 oB, V3& v, V3& wA, V3& wB, int& info) {
  const double r2 = 0.7071067811865475244, r3 = 0.5773502691896257645;
  V3 guesses[] = {safe_normalize(oB - oA), safe_normalize(oA - oB), V3(0, 0, 1), V3(0, 1, 0), V3(1, 0, 0),
                  vdiv(V3(1, 1, 0), std::sqrt(2.0)), vdiv(V3(1, 1, 1), std::sqrt(3.0)), vdiv(V3(0, 1, 1), std::sqrt(2.0)), vdiv(V3(1, 0, 1), std::sqrt(2.0))};
  (void)r2; (void)r3;
  if (g_epa_variant == 1) guesses[0] = oB - oA;
  if (g_epa_variant == 2) { guesses[0] = safe_normalize(oA - oB); }
  if (g_epa_variant == 3) { guesses[0] = V3(0, 0, 1); }
  if (g_epa_variant == 4) { guesses[0] = V3(1, 0, 0); }
  for (int i = 0; i < 9; i++) {
    Results res;
    if (penetration(sA, RA, oA, sB, RB, oB, guesses[i], res)) { wA = res.witness0; wB = res.witness1; v = res.normal; info = res.status; return true; }
    else if (distance(sA, RA, oA, sB, RB, oB, guesses[i], res)) { wA = res.witness0; wB = res.witness1; v = res.normal; info = -1; return false; }
  }
  wA = V3(); wB = V3(); v = V3(); info = -2;
  return false;
}

}  // namespace epa2
}  // namespace pgo
