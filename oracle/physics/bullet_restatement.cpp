// CPU restatement (C++ fp64) of the rigid-body stepping the reference delegates to PyBullet.
// TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Nothing in the product links this file.
//
// The reference never implements physics itself: every `self._p.*` call in
// envs/plate_contact_graph_env.py:176-250 (reset), :454-482 (pre_simulate), :504-533 (servo loop)
// and :556-603 (settle) goes into the un-vendored third-party wheel `pybullet` (build of 2022-05-20,
// bullet3 ~3.2.4, double precision).  This file restates the PUBLISHED algorithm of that engine for
// exactly the feature set those call sites exercise:
//   * btMultiBodyDynamicsWorld single step with base-only multibodies: gravity + velocity-proportional
//     damping + gyroscopic term -> collision detection -> sequential-impulse solve -> semi-implicit Euler
//     with exponential-map orientation update;
//   * box-box narrowphase (the ODE dBoxBox2 routine Bullet ships as btBoxBoxDetector), convex-convex
//     narrowphase by GJK + EPA with margins (hull / cylinder objects), persistent 4-point manifolds;
//   * btMultiBodyConstraintSolver ordering: non-contact rows (alternating direction), normal rows,
//     paired cone-friction rows, 50 iterations, residual early-out;
//   * btMultiBodyFixedConstraint to the world (the fingertip servo), impulse limits +-maxForce*dt.
// Parity is pinned ONLY by the reference's recorded (u -> object pose) trajectories; every constant
// marked [H] is a hypothesis about PyBullet's server defaults that those trajectories test
// (SURVEY.md Appendix B).  Achieved tolerances are reported in DESIGN.md.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace pgo {

// ---------------------------------------------------------------------------------------------
// small linear algebra
struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(double s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
static inline V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double len2(V3 a) { return dot(a, a); }
static inline double len(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 mulc(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }

struct Q4 { double x, y, z, w; };
static inline Q4 qmul(Q4 a, Q4 b) {
  return Q4{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
static inline Q4 qnormalize(Q4 q) {
  double l = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Q4{q.x / l, q.y / l, q.z / l, q.w / l};
}
struct M3 {
  double m[3][3];
  V3 row(int i) const { return V3(m[i][0], m[i][1], m[i][2]); }
  V3 col(int j) const { return V3(m[0][j], m[1][j], m[2][j]); }
};
static inline M3 qmat(Q4 q) {  // btMatrix3x3::setRotation
  double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  double s = 2.0 / d;
  double xs = q.x * s, ys = q.y * s, zs = q.z * s;
  double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
  double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
  double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
  M3 r;
  r.m[0][0] = 1.0 - (yy + zz); r.m[0][1] = xy - wz; r.m[0][2] = xz + wy;
  r.m[1][0] = xy + wz; r.m[1][1] = 1.0 - (xx + zz); r.m[1][2] = yz - wx;
  r.m[2][0] = xz - wy; r.m[2][1] = yz + wx; r.m[2][2] = 1.0 - (xx + yy);
  return r;
}
static inline V3 mul(const M3& a, V3 v) { return V3(dot(a.row(0), v), dot(a.row(1), v), dot(a.row(2), v)); }
static inline V3 mulT(const M3& a, V3 v) { return V3(dot(a.col(0), v), dot(a.col(1), v), dot(a.col(2), v)); }
static inline M3 mulTM(const M3& a, const M3& b) {  // a^T b
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = dot(a.col(i), b.col(j));
  return r;
}

static const double kEps = 2.220446049250313e-16;  // SIMD_EPSILON in a double-precision build
static const double kLarge = 1e30;                  // BT_LARGE_FLOAT (double build)

static inline void plane_space1(V3 n, V3& p, V3& q) {  // btPlaneSpace1
  const double SQRT12 = 0.7071067811865475244008443621048490;
  if (std::fabs(n.z) > SQRT12) {
    double a = n.y * n.y + n.z * n.z;
    double k = 1.0 / std::sqrt(a);
    p = V3(0, -n.z * k, n.y * k);
    q = V3(a * k, -n.x * p.z, n.x * p.y);
  } else {
    double a = n.x * n.x + n.y * n.y;
    double k = 1.0 / std::sqrt(a);
    p = V3(-n.y * k, n.x * k, 0);
    q = V3(-n.z * p.y, n.z * p.x, a * k);
  }
}

// ---------------------------------------------------------------------------------------------
enum ShapeType { SH_BOX = 0, SH_HULL = 1, SH_CYL = 2 };

struct Shape {
  int type = SH_BOX;
  V3 half;                 // box: full half extents (margin included, as btBoxShape keeps them)
  std::vector<V3> verts;   // hull vertices (core shape; margin is added on top)
  bool in_compound = false; // hull child of a btCompoundShape (URDF mesh)
  double margin = 0.001;   // [H] gUrdfDefaultCollisionMargin / m_defaultCollisionMargin
  double radius = 0, halfh = 0;  // cylinder Z
};

struct Body {
  Shape shape;
  bool is_static = true;
  double mass = 0;
  V3 inertia;      // local diagonal
  V3 pos;
  Q4 q{0, 0, 0, 1};  // local -> world
  V3 v, w;
  double friction = 0.5;
  long uid = 0;    // broadphase proxy id (decides which body is A for non-compound pairs)
  int tag = -1;    // island tag
  bool compound = false;  // floor: btCompoundShape wrapper => always body A of its pairs
  M3 R;            // cached rotation
  V3 aabb_lo, aabb_hi;
  V3 proxy_lo, proxy_hi;   // tight AABB the broadphase proxy was (re)created with
  bool early = false;      // tips: the re-created object proxy found this tip at creation (tight AABBs overlapping)
};

struct ManifoldPoint {
  V3 localA, localB, normalB, posA, posB;
  double dist = 0;
  double impN = 0, impL1 = 0, impL2 = 0;
  V3 lat1, lat2;
  double friction = 0;
  int life = 0;
};

struct Manifold {
  int a = -1, b = -1;      // body indices (a = Bullet body0)
  ManifoldPoint p[4];
  int n = 0;
  double breaking = 0.02;
};

struct FixedConstraint {
  int body = -1;
  V3 pivotB;
  M3 frameB;
  double max_impulse = 500.0;
  bool active = false;
};

struct Params {
  double dt = 1.0 / 250.0;
  double gravity_z = -10.0;
  int iterations = 50;            // [H] PyBullet server default
  double erp = 0.2;               // [H] non-contact ERP
  double erp2 = 0.08;             // [H] contact ERP set by the PyBullet server
  double linear_slop = 1e-5;      // [H]
  double warmstart = 0.0;         // [H] multibody contacts are NOT warm started (goldens reject 0.1)
  double residual_threshold = 1e-7;  // [H] leastSquaresResidualThreshold
  double lin_damping = 0.04, ang_damping = 0.04;  // [H] btMultiBody defaults
  int gyro = 1;                   // [H] m_useGyroTerm
  int cone_friction = 1;          // [H] implicit cone on the two friction rows
  int two_dirs = 1;               // [H] SOLVER_USE_2_FRICTION_DIRECTIONS (btMultiBodyDynamicsWorld ctor)
  double breaking_factor = 0.02;  // gContactBreakingThreshold
  double aabb_margin = 0.02;      // contact threshold added to AABBs
  double max_coord_vel = 100.0;   // btMultiBody::m_maxCoordinateVelocity
  int pair_order_mode = -1;       // -1: by object type; 0: box-object order; 1: compound-object order (see collide)
  double tight_pad = 0.0;         // debugging: slack added to the tight-AABB test of compound children
  int order_step_min = 0;         // debugging: manifold_order3 applies from this step on
  int manifold_order2 = 0, order_step_min2 = 1 << 30;   // debugging: a second explicit order (same coding as manifold_order >= 10000) from step order_step_min2 on
  int manifold_reverse = 0;       // debugging: reverse the final manifold order
  int manifold_order3 = 0;        // debugging: explicit order (digits, creation indices) when exactly 3 manifolds are solved
  int manifold_order = 0;         // 0: emulate quickSort of dispatcher order; k>=1: (k-1)-th permutation of creation order
  int constraint_reverse = 1;     // [H] quickSort of equal island ids reverses the two servo constraints
  int residual_max = 1;           // 1: max over rows, 0: sum
  int merge_islands = 0;          // every island is its own solver call with its own early-out: PyBullet's server sets
                                  // minimumSolverBatchSize = 0.  Decided by cardboard_20.0, where the two parked fingertips collide
                                  // with each other: 1e-16 over 50 rows this way, 2e-7 growing to 1e-6 with one merged solve
                                  // (which would iterate the object's island until the fingertips' island converges);
                                  // every other recording is indifferent
  double default_max_impulse = 500.0;
  int pair_flip = 0;              // debugging: reverse the uid rule for non-compound pairs
  double fric_scale = 1.0;        // debugging: scales the floor friction at step dbg_step
  double pos_drop = 0;            // debugging: at step dbg_step, drop cached points farther apart than this
  int skip_body = -1, skip_other = -1;   // debugging: at step dbg_step the pair (skip_body, skip_other) is refreshed but not re-detected
  int marginal_rule = 0;          // 1: the round-1 rule for a merely touching fingertip (fitted without the idle fingertips; obsolete)
  int drop_scan_reverse = 0;      // see collide()
  int tipwall_order = 0;          // order of the (wall, fingertip) pairs, see collide()
  int compound_prerefresh = 1;    // see narrowphase()
  int tt_never_early = 1;         // [H] the fingertip-fingertip pair enters the cache after the object's pairs (plate_20.0, start of env step 2: row 51 6e-5 -> 4e-7); 0: round-1 rule (first when the tight proxies overlapped)
  int tip_idle_collide = 0;       // see collide(): parked active fingertips vs idle ones
  int tip_idle_order = 0;         // order of a parked fingertip's pairs with the idle ones: 0 ascending, 1 descending, -1: found last (generic)
  int idle_order = 0;             // creation order of the idle fingertips' pairs: 0: (2,3),(2,4),(3,4); 1: (2,3),(3,4),(2,4)
  int find_order = 4239;          // [H] order in which the re-created object proxy finds its pairs: digits 4 compound floor, 2 index tip, 3 thumb (tips whose tight AABB overlapped), 9 walls youngest first (1 / 5 / 6-8: round-1 and per-wall variants)
  int find_order_late = 0;        // what-if: find order of the object's re-created proxy in every env step after the first (world step > 52).
                                  // The held-out cardboard2_20.0 / cardboard3_20.0 (thumb on the plate in both env steps) and plate_no_df_20.0
                                  // (thumb new in env step 2) want the thumb BEFORE the floor there (3429: all 100 rows instead of 51),
                                  // keyboard_scale_20.0 and plate_20.0 want the floor first: the order follows the broadphase tree of the
                                  // moment, not a rule of the env
  int find_order2 = 0, find_order2_min = 1 << 30;   // what-if: another find order from world step find_order2_min on
  int fs_idx0 = -1, fs_idx1 = -1; double fs_val0 = 1, fs_val1 = 1;
  int tip_skip_n = 0;             // ... and for this many further steps
  int tip_skip = 0;               // debugging: no narrowphase for fingertip pairs at this step
  int cone_variant = 0;           // debugging: alternative friction-cone clips
  int pos_variant = 0;            // debugging: alternative right-hand sides for contacts with positive distance
  int dbg_skip_new = 0;           // debugging: no new compound-child manifold at this step
  int dbg_step = 0, dbg_idx = -1; // debugging: force the cache slot of contacts added at one step
  int walls_first = 1;            // [H] this many (object, wall) pairs enter the pair cache before the early fingertip pairs: 1 (or 2)
                                  // carries waterbottle_20.0 from row 2 to row 20 (7e-8) and improves wall_laptop_20.0; 0 and 3 do not
  int force_iters = 0, force_step = 0;  // debugging: run exactly force_iters iterations at substep #force_step
  int idle_rel01 = 0, idle_rel02 = 0, idle_rel12 = 0;   // debugging: world step at which the idle fingertips' pair (i, j) leaves the pair cache (0: when its AABBs separate)
  double perturb = 0;             // debugging: added to the object's velocity (v.x and w.y) at step dbg_step -- measures how fast the trajectory amplifies a rounding-level difference
  int debug = 0;
};

struct SolverRow {
  int bA = -1, bB = -1;     // dynamic body index or -1 (static / world)
  double jA[6], jB[6];      // [ang, lin] jacobians
  double dA[6], dB[6];      // unit-impulse velocity responses
  double rhs = 0, cfm = 0, lo = 0, hi = 0, inv = 0, imp = 0, friction = 0;
  int fric_of = -1;         // normal-row index for friction rows
  double prev_imp = 0;      // debugging what-if: impulse before this iteration's update
  ManifoldPoint* cp = nullptr;
};

struct World {
  Params P;
  std::vector<Body> bodies;
  std::vector<Manifold> manifolds;          // dispatcher order
  std::vector<std::pair<int, int>> disabled;  // collision filter (unordered body pairs)
  FixedConstraint cons[2];
  long next_uid = 1;
  int obj = -1, tip[2] = {-1, -1}, floor_body = -1;
  std::vector<int> idle;        // the env's inactive fingertips (plate_env:215-228: fingers 2-4, all created at (100,100,100))
  int last_iterations = 0;
  int step_count = 0;
  // scratch
  std::vector<double> dv;  // delta velocities per body [6] (warm start + solver deltas, as m_deltaVelocities)
  std::vector<double> ws;  // the part of dv already written into the body velocities by warm starting
};

static void update_cache(Body& b) { b.R = qmat(b.q); }

// Hull AABBs (btPolyhedralConvexAabbCachingShape): the cached local AABB is the vertex extent + margin, getAabb() adds
// the margin once more (child / tight AABB = extent + 2 margins); a btCompoundShape wrapper (URDF mesh) adds its own
// margin on top for the broadphase AABB, the bounding sphere and, through them, the contact breaking threshold.
static double g_hull_aabb_margins = 2.0, g_compound_margin = 0.001;
static void hull_extent(const Shape& s, double margins, bool with_compound, V3& c, V3& e) {
  V3 lo(1e300, 1e300, 1e300), hi(-1e300, -1e300, -1e300);
  for (auto& v : s.verts) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
  double m = margins * s.margin + ((with_compound && s.in_compound) ? g_compound_margin : 0.0);
  c = (hi + lo) * 0.5; e = (hi - lo) * 0.5 + V3(m, m, m);
}
static double bounding_radius(const Shape& s) {
  // btCollisionShape::getBoundingSphere via getAabb(identity): half diagonal, plus the centre offset (getAngularMotionDisc)
  if (s.type == SH_BOX) return len(s.half);
  if (s.type == SH_CYL) return len(V3(s.radius, s.radius, s.halfh));  // btCylinderShape aabb = half extents
  V3 c, e; hull_extent(s, g_hull_aabb_margins, true, c, e);
  return len(e) + len(c);
}

static void local_aabb(const Shape& s, V3& c, V3& e) {
  if (s.type == SH_BOX) { c = V3(); e = s.half; return; }
  if (s.type == SH_CYL) { c = V3(); e = V3(s.radius, s.radius, s.halfh); return; }
  hull_extent(s, g_hull_aabb_margins, true, c, e);
}

static void compute_aabb(Body& b, double extra) {
  V3 c, e; local_aabb(b.shape, c, e);
  V3 wc = mul(b.R, c) + b.pos, we;
  for (int i = 0; i < 3; i++) we[i] = std::fabs(b.R.m[i][0]) * e.x + std::fabs(b.R.m[i][1]) * e.y + std::fabs(b.R.m[i][2]) * e.z;
  b.aabb_lo = wc - we - V3(extra, extra, extra);
  b.aabb_hi = wc + we + V3(extra, extra, extra);
}

static double g_hull_inertia_pad = 0.002;   // [H] see shape_inertia
static V3 shape_inertia(const Shape& s, double mass) {
  double lx, ly, lz;
  if (s.type == SH_BOX) { lx = 2 * s.half.x; ly = 2 * s.half.y; lz = 2 * s.half.z; }
  else if (s.type == SH_CYL) {
    // btCylinderShape::calculateLocalInertia, Z-up cylinder, principal axis z
    double r2 = s.radius * s.radius, h2 = 4 * s.halfh * s.halfh;
    double t1 = mass / 12.0 * h2 + mass / 4.0 * r2, t2 = mass / 2.0 * r2;
    return V3(t1, t1, t2);
  } else {
    // btCompoundShape::calculateLocalInertia: box inertia of the compound's AABB.  The cached hull AABB carries the
    // margin three times (support with margin + recalcLocalAabb + btTransformAabb) and the compound adds its own
    // (btPolyhedralConvexShape::calculateLocalInertia for a bare hull gives the same: getAabb(identity) + one margin)
    V3 c, e; hull_extent(s, 1.0, false, c, e);
    e = e + V3(g_hull_inertia_pad, g_hull_inertia_pad, g_hull_inertia_pad);
    lx = 2 * e.x; ly = 2 * e.y; lz = 2 * e.z;
  }
  return V3(mass / 12.0 * (ly * ly + lz * lz), mass / 12.0 * (lx * lx + lz * lz), mass / 12.0 * (lx * lx + ly * ly));
}

// ---------------------------------------------------------------------------------------------
// persistent manifold (btPersistentManifold / btManifoldResult)
static int manifold_cache_entry(const Manifold& m, const ManifoldPoint& np) {
  double shortest = m.breaking * m.breaking;
  int nearest = -1;
  for (int i = 0; i < m.n; i++) {
    V3 d = m.p[i].localA - np.localA;
    double dd = dot(d, d);
    if (dd < shortest) { shortest = dd; nearest = i; }
  }
  return nearest;
}

static int manifold_sort_cached(const Manifold& m, const ManifoldPoint& pt) {
  int maxPenIdx = -1;
  double maxPen = pt.dist;
  for (int i = 0; i < 4; i++) if (m.p[i].dist < maxPen) { maxPenIdx = i; maxPen = m.p[i].dist; }
  double res[4] = {0, 0, 0, 0};
  if (maxPenIdx != 0) { V3 a = pt.localA - m.p[1].localA, b = m.p[3].localA - m.p[2].localA; res[0] = len2(cross(a, b)); }
  if (maxPenIdx != 1) { V3 a = pt.localA - m.p[0].localA, b = m.p[3].localA - m.p[2].localA; res[1] = len2(cross(a, b)); }
  if (maxPenIdx != 2) { V3 a = pt.localA - m.p[0].localA, b = m.p[3].localA - m.p[1].localA; res[2] = len2(cross(a, b)); }
  if (maxPenIdx != 3) { V3 a = pt.localA - m.p[0].localA, b = m.p[2].localA - m.p[1].localA; res[3] = len2(cross(a, b)); }
  int best = -1; double bv = -kLarge;   // btVector4::closestAxis4 = maxAxis4 of absolute values
  for (int i = 0; i < 4; i++) if (std::fabs(res[i]) > bv) { best = i; bv = std::fabs(res[i]); }
  return best;
}

struct ContactSink {
  World* w; Manifold* m; bool swapped;  // swapped: detector's (A,B) is the manifold's (B,A)
};

// normalOnB / pointOnB / depth in the DETECTOR's A/B convention
static void add_contact(ContactSink& s, V3 normalOnB, V3 pointOnB, double depth) {
  Manifold& m = *s.m;
  if (depth > m.breaking) return;
  const Body& bodyA = s.w->bodies[m.a];
  const Body& bodyB = s.w->bodies[m.b];
  V3 pointA = pointOnB + normalOnB * depth;
  ManifoldPoint np;
  if (s.swapped) {
    // btManifoldResult::addContactPoint, isSwapped branch: localA from result body1, localB from result body0
    np.localA = mulT(bodyA.R, pointA - bodyA.pos);
    np.localB = mulT(bodyB.R, pointOnB - bodyB.pos);
  } else {
    np.localA = mulT(bodyA.R, pointA - bodyA.pos);
    np.localB = mulT(bodyB.R, pointOnB - bodyB.pos);
  }
  np.normalB = normalOnB; np.dist = depth; np.posA = pointA; np.posB = pointOnB;
  double f = bodyA.friction * bodyB.friction;
  if (f < -10.0) f = -10.0; if (f > 10.0) f = 10.0;   // MAX_FRICTION
  np.friction = f;
  plane_space1(np.normalB, np.lat1, np.lat2);
  int idx = manifold_cache_entry(m, np);
  if (s.w->P.dbg_step > 0 && s.w->step_count == s.w->P.dbg_step && (m.n == 4 || s.w->P.dbg_idx >= 10) &&
      (m.a == s.w->tip[0] || m.a == s.w->tip[1] || m.b == s.w->tip[0] || m.b == s.w->tip[1])) {     // debugging: what-if on one insertion
    if (s.w->P.dbg_idx == 9) return;
    if (s.w->P.dbg_idx == 10) idx = -1;                       // add as a new point instead of replacing
    if (s.w->P.dbg_idx >= 11 && s.w->P.dbg_idx <= 14) idx = s.w->P.dbg_idx - 11 < m.n ? s.w->P.dbg_idx - 11 : idx;   // replace this slot
    if (s.w->P.dbg_idx == 19) return;                         // drop the point
    if (s.w->P.dbg_idx >= 0) { if (s.w->P.debug) fprintf(stderr, "dbg: natural idx %d -> %d\n", idx, s.w->P.dbg_idx); idx = s.w->P.dbg_idx; }
  }
  if (idx >= 0) {
    np.impN = m.p[idx].impN; np.impL1 = m.p[idx].impL1; np.impL2 = m.p[idx].impL2; np.life = m.p[idx].life;
    m.p[idx] = np;
  } else {
    idx = m.n;
    if (idx == 4) idx = manifold_sort_cached(m, np); else m.n++;
    if (idx < 0) idx = 0;
    m.p[idx] = np;
  }
}

static void refresh_manifold(World& w, Manifold& m) {
  const Body& A = w.bodies[m.a];
  const Body& B = w.bodies[m.b];
  for (int i = m.n - 1; i >= 0; i--) {
    ManifoldPoint& p = m.p[i];
    p.posA = mul(A.R, p.localA) + A.pos;
    p.posB = mul(B.R, p.localB) + B.pos;
    p.dist = dot(p.posA - p.posB, p.normalB);
    p.life++;
  }
  for (int i = m.n - 1; i >= 0; i--) {
    ManifoldPoint& p = m.p[i];
    bool remove = false;
    if (!(p.dist <= m.breaking)) remove = true;
    else if (w.P.pos_drop > 0 && w.step_count == w.P.dbg_step && p.dist > w.P.pos_drop) remove = true;      // debugging what-if
    else {
      V3 proj = p.posA - p.normalB * p.dist;
      V3 d = p.posB - proj;
      if (dot(d, d) > m.breaking * m.breaking) remove = true;
      if (w.P.debug == 5) fprintf(stderr, "step %d refresh %d-%d slot %d dist %+.4e drift %.4e breaking %.4e life %d%s\n", w.step_count, m.a, m.b, i, p.dist, std::sqrt(dot(d, d)), m.breaking, p.life, remove ? " REMOVED" : "");
    }
    if (remove) {
      int last = m.n - 1;
      if (i != last) m.p[i] = m.p[last];
      m.n--;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// box-box (btBoxBoxDetector = ODE dBoxBox2).  R are row-major 3x3, columns = box axes.
static void line_closest_approach(V3 pa, V3 ua, V3 pb, V3 ub, double* alpha, double* beta) {
  V3 p = pb - pa;
  double uaub = dot(ua, ub), q1 = dot(ua, p), q2 = -dot(ub, p);
  double d = 1 - uaub * uaub;
  if (d <= (double)0.0001f) { *alpha = 0; *beta = 0; }
  else { d = 1.0 / d; *alpha = (q1 + uaub * q2) * d; *beta = (uaub * q1 + q2) * d; }
}

static int intersect_rect_quad2(double h[2], double p[8], double ret[16]) {
  int nq = 4, nr = 0;
  double buffer[16];
  double* q = p;
  double* r = ret;
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      double* pq = q;
      double* pr = r;
      nr = 0;
      for (int i = nq; i > 0; i--) {
        if (sign * pq[dir] < h[dir]) {
          pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++;
          if (nr & 8) { q = r; goto done; }
        }
        double* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          pr[dir] = sign * h[dir];
          pr += 2; nr++;
          if (nr & 8) { q = r; goto done; }
        }
        pq += 2;
      }
      q = r;
      r = (q == ret) ? buffer : ret;
      nq = nr;
    }
  }
done:
  if (q != ret) memcpy(ret, q, nr * 2 * sizeof(double));
  return nr;
}

static const double M__PI = (double)3.14159265f;

static void cull_points2(int n, double p[], int m, int i0, int iret[]) {
  int i, j;
  double a, cx, cy, q;
  if (n == 1) { cx = p[0]; cy = p[1]; }
  else if (n == 2) { cx = 0.5 * (p[0] + p[2]); cy = 0.5 * (p[1] + p[3]); }
  else {
    a = 0; cx = 0; cy = 0;
    for (i = 0; i < (n - 1); i++) {
      q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
      a += q;
      cx += q * (p[i * 2] + p[i * 2 + 2]);
      cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
    }
    q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
    if (std::fabs(a + q) > kEps) a = 1.0 / (3.0 * (a + q)); else a = kLarge;
    cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
    cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
  }
  double A[8];
  for (i = 0; i < n; i++) A[i] = std::atan2(p[i * 2 + 1] - cy, p[i * 2] - cx);
  int avail[8];
  for (i = 0; i < n; i++) avail[i] = 1;
  avail[i0] = 0;
  iret[0] = i0;
  iret++;
  for (j = 1; j < m; j++) {
    a = double(j) * (2 * M__PI / m) + A[i0];
    if (a > M__PI) a -= 2 * M__PI;
    double maxdiff = 1e9, diff;
    *iret = i0;
    for (i = 0; i < n; i++) {
      if (avail[i]) {
        diff = std::fabs(A[i] - a);
        if (diff > M__PI) diff = 2 * M__PI - diff;
        if (diff < maxdiff) { maxdiff = diff; *iret = i; }
      }
    }
    avail[*iret] = 0;
    iret++;
  }
}

static int g_last_code = 0, g_last_cnum = 0;
static int g_bb_edge_point = 0, g_bb_edge_flip = 0;
static double g_bb_fudge = 1.05;   // debugging: dBoxBox2's edge-axis preference factor
static const bool g_bb_debug = getenv("PGO_BB_DEBUG") != nullptr;   // debugging: print every separating-axis value
static int box_box(V3 p1, const M3& R1, V3 half1, V3 p2, const M3& R2, V3 half2, ContactSink& out) {
  g_last_code = 0; g_last_cnum = 0;
  const double fudge_factor = g_bb_fudge;
  V3 p, pp, normalC;
  int normalR_box = 0, normalR_col = -1;  // column of R1 (box 1) or R2 (box 2)
  double A[3], B[3], R[3][3], Q[3][3], s, s2, l;
  int i, j, invert_normal, code;
  p = p2 - p1;
  pp = mulT(R1, p);
  // side = 2*half, A = side*0.5
  for (i = 0; i < 3; i++) { A[i] = (2.0 * half1[i]) * 0.5; B[i] = (2.0 * half2[i]) * 0.5; }
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) { R[i][j] = dot(R1.col(i), R2.col(j)); Q[i][j] = std::fabs(R[i][j]); }
  s = -INFINITY; invert_normal = 0; code = 0;
#define TST1(expr1, expr2, box, col, cc) \
  { double e1 = (expr1); s2 = std::fabs(e1) - (expr2); if (s2 > 0) return 0; \
    if (g_bb_debug) fprintf(stderr, "    bb face %d s2 %.12e\n", (cc), s2); \
    if (s2 > s) { s = s2; normalR_box = box; normalR_col = col; invert_normal = (e1 < 0); code = (cc); } }
  TST1(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), 1, 0, 1);
  TST1(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), 1, 1, 2);
  TST1(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), 1, 2, 3);
  TST1(dot(R2.col(0), p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), 2, 0, 4);
  TST1(dot(R2.col(1), p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), 2, 1, 5);
  TST1(dot(R2.col(2), p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), 2, 2, 6);
#undef TST1
#define TST2(expr1, expr2, n1, n2, n3, cc) \
  { double e1 = (expr1); s2 = std::fabs(e1) - (expr2); if (s2 > kEps) return 0; \
    l = std::sqrt((n1) * (n1) + (n2) * (n2) + (n3) * (n3)); \
    if (l > kEps) { s2 /= l; if (g_bb_debug) fprintf(stderr, "    bb edge %d s2 %.12e (x1.05 %.12e)\n", (cc), s2, s2 * fudge_factor); if (s2 * fudge_factor > s) { s = s2; normalR_col = -1; \
      normalC = V3((n1) / l, (n2) / l, (n3) / l); invert_normal = (e1 < 0); code = (cc); } } }
  const double fudge2 = (double)1.0e-5f;
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) Q[i][j] += fudge2;
  TST2(pp[2] * R[1][0] - pp[1] * R[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -R[2][0], R[1][0], 7);
  TST2(pp[2] * R[1][1] - pp[1] * R[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -R[2][1], R[1][1], 8);
  TST2(pp[2] * R[1][2] - pp[1] * R[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -R[2][2], R[1][2], 9);
  TST2(pp[0] * R[2][0] - pp[2] * R[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), R[2][0], 0, -R[0][0], 10);
  TST2(pp[0] * R[2][1] - pp[2] * R[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), R[2][1], 0, -R[0][1], 11);
  TST2(pp[0] * R[2][2] - pp[2] * R[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), R[2][2], 0, -R[0][2], 12);
  TST2(pp[1] * R[0][0] - pp[0] * R[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -R[1][0], R[0][0], 0, 13);
  TST2(pp[1] * R[0][1] - pp[0] * R[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -R[1][1], R[0][1], 0, 14);
  TST2(pp[1] * R[0][2] - pp[0] * R[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -R[1][2], R[0][2], 0, 15);
#undef TST2
  if (!code) return 0;
  g_last_code = code;
  if (g_bb_debug) fprintf(stderr, "    bb: code %d s %.12e normalC %.6f %.6f %.6f\n", code, s, normalC.x, normalC.y, normalC.z);
  V3 normal;
  if (normalR_col >= 0) normal = (normalR_box == 1 ? R1 : R2).col(normalR_col);
  else normal = mul(R1, normalC);
  if (invert_normal) normal = -normal;
  double depth = -s;

  if (code > 6) {
    V3 pa = p1;
    for (j = 0; j < 3; j++) {
      double sign = (dot(normal, R1.col(j)) > 0) ? 1.0 : -1.0;
      if ((g_bb_edge_flip >> j) & 1) sign = -sign;           // debugging what-if
      pa += R1.col(j) * (sign * A[j]);
    }
    V3 pb = p2;
    for (j = 0; j < 3; j++) {
      double sign = (dot(normal, R2.col(j)) > 0) ? -1.0 : 1.0;
      if ((g_bb_edge_flip >> (3 + j)) & 1) sign = -sign;     // debugging what-if
      pb += R2.col(j) * (sign * B[j]);
    }
    double alpha, beta;
    V3 ua = R1.col((code - 7) / 3), ub = R2.col((code - 7) % 3);
    line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
    pa += ua * alpha;
    pb += ub * beta;
    if (g_bb_edge_point == 1) add_contact(out, -normal, (pa + pb) * 0.5, -depth);     // debugging what-ifs (USE_CENTER_POINT)
    else if (g_bb_edge_point == 2) add_contact(out, -normal, pa, -depth);
    else add_contact(out, -normal, pb, -depth);
    return 1;
  }

  const M3 *Ra, *Rb;
  V3 pa, pb;
  const double *Sa, *Sb;
  if (code <= 3) { Ra = &R1; Rb = &R2; pa = p1; pb = p2; Sa = A; Sb = B; }
  else { Ra = &R2; Rb = &R1; pa = p2; pb = p1; Sa = B; Sb = A; }
  V3 normal2 = (code <= 3) ? normal : -normal, nr, anr;
  nr = mulT(*Rb, normal2);
  anr = V3(std::fabs(nr.x), std::fabs(nr.y), std::fabs(nr.z));
  int lanr, a1, a2;
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  V3 center;
  if (nr[lanr] < 0) center = pb - pa + Rb->col(lanr) * Sb[lanr];
  else center = pb - pa - Rb->col(lanr) * Sb[lanr];
  int codeN, code1, code2;
  codeN = (code <= 3) ? code - 1 : code - 4;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double quad[8];
  double c1, c2, m11, m12, m21, m22;
  c1 = dot(center, Ra->col(code1));
  c2 = dot(center, Ra->col(code2));
  m11 = dot(Ra->col(code1), Rb->col(a1));
  m12 = dot(Ra->col(code1), Rb->col(a2));
  m21 = dot(Ra->col(code2), Rb->col(a1));
  m22 = dot(Ra->col(code2), Rb->col(a2));
  {
    double k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = {Sa[code1], Sa[code2]};
  double ret[16];
  int n = intersect_rect_quad2(rect, quad, ret);
  if (n < 1) return 0;
  double point[3 * 8], dep[8];
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (j = 0; j < n; j++) {
    double k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    double k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    V3 pt = center + Rb->col(a1) * k1 + Rb->col(a2) * k2;
    point[cnum * 3] = pt.x; point[cnum * 3 + 1] = pt.y; point[cnum * 3 + 2] = pt.z;
    dep[cnum] = Sa[codeN] - dot(normal2, pt);
    if (g_bb_debug) fprintf(stderr, "    bb poly vertex %d/%d ret %.9f %.9f dep %.6e\n", j, n, ret[j * 2], ret[j * 2 + 1], dep[cnum]);
    if (dep[cnum] >= 0) { ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1]; cnum++; }
  }
  if (cnum < 1) return 0;
  g_last_cnum = cnum;
  int maxc = 4;
  if (maxc > cnum) maxc = cnum;
  if (maxc < 1) maxc = 1;
  if (cnum <= maxc) {
    for (j = 0; j < cnum; j++) {
      V3 pw = V3(point[j * 3], point[j * 3 + 1], point[j * 3 + 2]) + pa;
      if (code >= 4) pw = pw - normal * dep[j];
      add_contact(out, -normal, pw, -dep[j]);
    }
  } else {
    int i1 = 0;
    double maxdepth = dep[0];
    for (i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
    int iret[8];
    cull_points2(cnum, ret, maxc, i1, iret);
    for (j = 0; j < maxc; j++) {
      V3 pw = V3(point[iret[j] * 3], point[iret[j] * 3 + 1], point[iret[j] * 3 + 2]) + pa;
      if (code >= 4) pw = pw - normal * dep[iret[j]];
      add_contact(out, -normal, pw, -dep[iret[j]]);
    }
    cnum = maxc;
  }
  return cnum;
}

}  // namespace pgo

#include "gjk_epa.inl"

namespace pgo {

// ---------------------------------------------------------------------------------------------
// collision pipeline
static bool filter_disabled(const World& w, int a, int b) {
  for (auto& d : w.disabled) if ((d.first == a && d.second == b) || (d.first == b && d.second == a)) return true;
  return false;
}

static bool aabb_overlap(const Body& a, const Body& b) {
  for (int k = 0; k < 3; k++) if (a.aabb_lo[k] > b.aabb_hi[k] || a.aabb_hi[k] < b.aabb_lo[k]) return false;
  return true;
}

// pair-level test.  A compound body (the floor: its URDF collision box has an origin offset, so Bullet keeps
// it wrapped in a btCompoundShape) goes through btCompoundCollisionAlgorithm, which creates the child
// algorithm + manifold only while the child's TIGHT aabb overlaps the other shape's aabb and destroys it
// (contact cache and all) as soon as they separate.
static bool pair_overlap(const World& w, const Body& a, const Body& b) {
  if (!aabb_overlap(a, b)) return false;
  if (a.compound || b.compound) {
    double ea = w.P.aabb_margin + (a.shape.in_compound ? g_compound_margin : 0.0) - 0.5 * w.P.tight_pad;
    double eb = w.P.aabb_margin + (b.shape.in_compound ? g_compound_margin : 0.0) - 0.5 * w.P.tight_pad;
    for (int k = 0; k < 3; k++)
      if (a.aabb_lo[k] + ea > b.aabb_hi[k] - eb || a.aabb_hi[k] - ea < b.aabb_lo[k] + eb) return false;
  }
  return true;
}

static Manifold* find_manifold(World& w, int a, int b) {
  for (auto& m : w.manifolds) if ((m.a == a && m.b == b) || (m.a == b && m.b == a)) return &m;
  return nullptr;
}

static void destroy_manifolds_of(World& w, int body) {
  for (size_t i = 0; i < w.manifolds.size();) {
    if (w.manifolds[i].a == body || w.manifolds[i].b == body) {
      w.manifolds[i] = w.manifolds.back();   // dispatcher swap-remove
      w.manifolds.pop_back();
    } else i++;
  }
}

// PyBullet's setCollisionFilterPair refreshes the broadphase proxies of both bodies (new ids, all
// their pairs and manifolds dropped)
static void refresh_proxy(World& w, int body) {
  destroy_manifolds_of(w, body);
  Body& b = w.bodies[body];
  b.uid = w.next_uid++;
  // btDbvtBroadphase::createProxy collides the NEW proxy's tight AABB against both trees at once (no deferred
  // collide): pairs found here enter the pair cache before the ones the next updateAabbs finds with fat AABBs
  update_cache(b); compute_aabb(b, 0.0);
  b.proxy_lo = b.aabb_lo; b.proxy_hi = b.aabb_hi;
  if (body == w.obj) {
    double depth[2] = {-1, -1};
    for (int t = 0; t < 2; t++) if (w.tip[t] >= 0) {
      Body& tb = w.bodies[w.tip[t]];
      double d = 1e300;                       // smallest per-axis overlap of the two tight AABBs (< 0: separated)
      for (int k = 0; k < 3; k++) d = std::min(d, std::min(tb.proxy_hi[k] - b.proxy_lo[k], b.proxy_hi[k] - tb.proxy_lo[k]));
      depth[t] = filter_disabled(w, body, w.tip[t]) ? -1.0 : d;
    }
    for (int t = 0; t < 2; t++) if (w.tip[t] >= 0) {
      // [H] a tip that merely touches the object (overlap below 1e-6: placed exactly on a face) is found at creation
      // only when it is the single active tip; with the other tip clearly overlapping it enters the cache later
      bool marginal = depth[t] < 1e-6, other_clear = depth[1 - t] >= 1e-6;
      w.bodies[w.tip[t]].early = depth[t] >= 0 && !(marginal && other_clear && w.P.marginal_rule);
      if (w.P.debug == 1) fprintf(stderr, "refresh obj: tip %d overlap %.3g early %d\n", t, depth[t], (int)w.bodies[w.tip[t]].early);
    }
  }
}

static void set_collision_filter(World& w, int a, int b, bool enable) {
  bool dis = filter_disabled(w, a, b);
  if (!enable && !dis) w.disabled.push_back({a, b});
  if (enable && dis) {
    for (size_t i = 0; i < w.disabled.size(); i++) {
      auto& d = w.disabled[i];
      if ((d.first == a && d.second == b) || (d.first == b && d.second == a)) { w.disabled.erase(w.disabled.begin() + i); break; }
    }
  }
  refresh_proxy(w, a);
  refresh_proxy(w, b);
}

static void narrowphase(World& w, Manifold& m) {
  Body& A = w.bodies[m.a];
  Body& B = w.bodies[m.b];
  ContactSink sink{&w, &m, false};
  if (w.P.tip_skip > 0 && w.step_count >= w.P.tip_skip && w.step_count <= w.P.tip_skip + w.P.tip_skip_n && (m.a == w.tip[0] || m.b == w.tip[0] || m.a == w.tip[1] || m.b == w.tip[1])) return;   // debugging what-if
  const bool dbg_no_detect = w.P.skip_body >= 0 && w.step_count == w.P.dbg_step && (m.a == w.P.skip_body || m.b == w.P.skip_body) && (w.P.skip_other < 0 || m.a == w.P.skip_other || m.b == w.P.skip_other);   // debugging what-if: refresh only
  if (dbg_no_detect) { refresh_manifold(w, m); return; }
  // btCompoundCollisionAlgorithm::processCollision refreshes the contact points of its child manifolds BEFORE it runs the child
  // algorithm (which refreshes them again afterwards): cached points that have drifted away are gone before the new ones arrive,
  // and a new point that has to displace one of four cached points (sortCachedPoints) is compared with their CURRENT depths
  if ((A.compound || B.compound) && w.P.compound_prerefresh) refresh_manifold(w, m);
  if (A.shape.type == SH_BOX && B.shape.type == SH_BOX) {
    int nbefore = m.n;
    box_box(A.pos, A.R, A.shape.half, B.pos, B.R, B.shape.half, sink);
    if (w.P.debug > 2) fprintf(stderr, "step %d pair %d %d code %d cnum %d n %d->%d\n", w.step_count, m.a, m.b, g_last_code, g_last_cnum, nbefore, m.n);
  } else {
    convex_convex(w, A, B, sink);
  }
  refresh_manifold(w, m);
}

static void collide(World& w) {
  for (auto& b : w.bodies) { update_cache(b); compute_aabb(b, w.P.aabb_margin); }
  int nb = (int)w.bodies.size();
  if (w.P.debug == 7 && nb > 6) fprintf(stderr, "step %d aabb: body4 hi.x %.12e  body6 lo.x %.12e  gap %.3e\n", w.step_count, w.bodies[4].aabb_hi.x, w.bodies[6].aabb_lo.x, w.bodies[6].aabb_lo.x - w.bodies[4].aabb_hi.x);
  // drop pairs whose AABBs separated
  if (w.P.drop_scan_reverse) {                 // debugging what-if: the broadphase's clean-up scan starts at a rotating index
    for (int i = (int)w.manifolds.size() - 1; i >= 0; i--) {
      Manifold& m = w.manifolds[i];
      if (!pair_overlap(w, w.bodies[m.a], w.bodies[m.b])) { w.manifolds[i] = w.manifolds.back(); w.manifolds.pop_back(); }
    }
  } else
  for (size_t i = 0; i < w.manifolds.size();) {
    Manifold& m = w.manifolds[i];
    bool gone = !pair_overlap(w, w.bodies[m.a], w.bodies[m.b]);
    if (w.idle.size() == 3) {                    // debugging what-if: explicit release steps of the idle fingertips' pairs
      auto ix = [&](int b) { for (int q = 0; q < 3; q++) if (w.idle[q] == b) return q; return -1; };
      int qa = ix(m.a), qb = ix(m.b);
      if (qa >= 0 && qb >= 0) {
        int lo = std::min(qa, qb), hi = std::max(qa, qb);
        int rel = (lo == 0 && hi == 1) ? w.P.idle_rel01 : (lo == 0 && hi == 2) ? w.P.idle_rel02 : w.P.idle_rel12;
        if (rel > 0) gone = w.step_count >= rel;
      }
    }
    if (gone) { w.manifolds[i] = w.manifolds.back(); w.manifolds.pop_back(); }
    else i++;
  }
  // new pairs.  Creation order emulates the order in which Bullet's broadphase reports the object's
  // pairs after PyBullet re-creates its proxy (fitted to the recorded trajectories): thumb, floor,
  // index tip, then the walls; pairs that do not involve the object come last.
  std::vector<std::pair<int, int>> cand;
  auto push = [&](int i, int j) { if (i >= 0 && j >= 0 && i != j) cand.push_back({std::min(i, j), std::max(i, j)}); };
  int mode = w.P.pair_order_mode;
  if (mode < 0) mode = 1;
  if (w.obj >= 0 && mode == 0) {
    push(w.obj, w.tip[0]);
    for (int i = 0; i < nb; i++) if (w.bodies[i].compound) push(w.obj, i);
    push(w.obj, w.tip[1]);
    for (int i = 0; i < nb; i++) if (w.bodies[i].is_static && !w.bodies[i].compound) push(w.obj, i);
  } else if (w.obj >= 0) {
    // pair-cache order after PyBullet's setCollisionFilterPair sequence (every call re-creates both proxies): the
    // surviving tip-tip / wall-tip pairs, then what the re-created object proxy found at creation -- dynamic tree
    // first (index tip, thumb: only tips whose tight AABB overlapped), then the static tree -- then the tip pairs
    // the first updateAabbs finds with fat AABBs.  Tree traversal orders are fitted to the recorded trajectories.
    // the idle fingertips overlap each other from the moment they are created: their pairs are the oldest of the cache
    if (w.idle.size() == 3) {
      push(w.idle[0], w.idle[1]);
      if (w.P.idle_order == 0) { push(w.idle[0], w.idle[2]); push(w.idle[1], w.idle[2]); }
      else { push(w.idle[1], w.idle[2]); push(w.idle[0], w.idle[2]); }
    }
    // a parked active fingertip is teleported to (100, 100, 100) -- where the idle ones were created -- BEFORE the object's proxy
    // is re-created for the last time: its pairs with the idle fingertips are older than all of the object's
    if (w.idle.size() == 3 && w.P.tip_idle_order >= 0) {
      for (int t = 0; t < 2; t++) if (w.tip[t] >= 0)
        for (int q = 0; q < 3; q++) push(w.tip[t], w.idle[w.P.tip_idle_order == 0 ? q : 2 - q]);
    }
    bool tt_early = false;
    if (w.tip[0] >= 0 && w.tip[1] >= 0) {
      const Body& a = w.bodies[w.tip[0]]; const Body& b = w.bodies[w.tip[1]];
      tt_early = true;
      for (int k = 0; k < 3; k++) if (a.proxy_lo[k] > b.proxy_hi[k] || a.proxy_hi[k] < b.proxy_lo[k]) tt_early = false;
    }
    if (w.P.tt_never_early) tt_early = false;          // debugging what-if
    if (tt_early) push(w.tip[0], w.tip[1]);
    switch (w.P.tipwall_order) {                 // [H] 0; the others are what-ifs
      case 0: for (int i = 0; i < nb; i++) if (w.bodies[i].is_static && !w.bodies[i].compound) { push(i, w.tip[1]); push(i, w.tip[0]); } break;
      case 1: for (int i = nb - 1; i >= 0; i--) if (w.bodies[i].is_static && !w.bodies[i].compound) { push(i, w.tip[1]); push(i, w.tip[0]); } break;
      case 2: for (int t = 0; t < 2; t++) for (int i = 0; i < nb; i++) if (w.bodies[i].is_static && !w.bodies[i].compound) push(i, w.tip[t]); break;
      case 3: for (int t = 0; t < 2; t++) for (int i = nb - 1; i >= 0; i--) if (w.bodies[i].is_static && !w.bodies[i].compound) push(i, w.tip[t]); break;
      case 4: for (int t = 1; t >= 0; t--) for (int i = nb - 1; i >= 0; i--) if (w.bodies[i].is_static && !w.bodies[i].compound) push(i, w.tip[t]); break;
      case 5: break;                              // found last (generic order)
    }
    {
      int digits[8], nd = 0;
      int fo = w.P.find_order;
      if (w.P.find_order_late > 0 && w.step_count > 52) fo = w.P.find_order_late;
      if (w.P.find_order2 > 0 && w.step_count >= w.P.find_order2_min) fo = w.P.find_order2;
      for (int c = fo; c > 0; c /= 10) digits[nd++] = c % 10;
      for (int q = nd - 1; q >= 0; q--) {
        switch (digits[q]) {
          case 1: { int k = 0; for (int i = 0; i < nb && k < w.P.walls_first; i++) if (w.bodies[i].is_static && !w.bodies[i].compound) { push(w.obj, i); k++; } } break;
          case 2: if (w.tip[1] >= 0 && w.bodies[w.tip[1]].early) push(w.obj, w.tip[1]); break;
          case 3: if (w.tip[0] >= 0 && w.bodies[w.tip[0]].early) push(w.obj, w.tip[0]); break;
          case 4: for (int i = 0; i < nb; i++) if (w.bodies[i].compound && i != w.obj) push(w.obj, i); break;
          case 5: for (int i = 0; i < nb; i++) if (w.bodies[i].is_static && !w.bodies[i].compound) push(w.obj, i); break;
          case 6: case 7: case 8: { int k = 0; for (int i = 0; i < nb; i++) if (w.bodies[i].is_static && !w.bodies[i].compound) { if (k == digits[q] - 6) push(w.obj, i); k++; } } break;   // wall #0, #1, #2 on its own
          case 9: for (int i = nb - 1; i >= 0; i--) if (w.bodies[i].is_static && !w.bodies[i].compound) push(w.obj, i); break;                                       // all walls, youngest first
        }
      }
    }
    for (int t = 1; t >= 0; t--) if (w.tip[t] >= 0 && !w.bodies[w.tip[t]].early) push(w.obj, w.tip[t]);
  }
  for (int i = 0; i < nb; i++) for (int j = i + 1; j < nb; j++) {
    bool seen = false;
    for (auto& c : cand) if (c.first == i && c.second == j) seen = true;
    if (!seen) cand.push_back({i, j});
  }
  for (auto& c : cand) {
    int i = c.first, j = c.second;
    Body& bi = w.bodies[i];
    Body& bj = w.bodies[j];
    if (bi.is_static && bj.is_static) continue;
    if (filter_disabled(w, i, j)) continue;
    if (!w.P.tip_idle_collide) {
      // [simplification shared with the CUDA kernel, which carries the idle fingertips only as entries of the dispatcher array]
      // an ACTIVE fingertip parked at (100, 100, 100) does collide with the idle ones in PyBullet during the first env step;
      // with tip_idle_collide = 1 (and tip_idle_order = 0) wall_laptop_20.0 is exact for three more rows (7-9)
      bool ii = std::find(w.idle.begin(), w.idle.end(), i) != w.idle.end(), ij = std::find(w.idle.begin(), w.idle.end(), j) != w.idle.end();
      if (ii != ij) continue;
    }
    if (!pair_overlap(w, bi, bj)) continue;
    if (find_manifold(w, i, j)) continue;
    if (w.idle.size() == 3 && (w.P.idle_rel01 || w.P.idle_rel02 || w.P.idle_rel12)) {
      auto ix = [&](int b) { for (int q = 0; q < 3; q++) if (w.idle[q] == b) return q; return -1; };
      int qa = ix(i), qb = ix(j);
      if (qa >= 0 && qb >= 0) {
        int lo = std::min(qa, qb), hi = std::max(qa, qb);
        int rel = (lo == 0 && hi == 1) ? w.P.idle_rel01 : (lo == 0 && hi == 2) ? w.P.idle_rel02 : w.P.idle_rel12;
        if (rel > 0 && w.step_count >= rel) continue;
      }
    }
    if (w.P.dbg_skip_new > 0 && w.step_count == w.P.dbg_skip_new && (bi.compound || bj.compound)) continue;   // debugging what-if
    Manifold m;
    if (bi.compound) { m.a = i; m.b = j; }
    else if (bj.compound) { m.a = j; m.b = i; }
    else if ((bi.uid < bj.uid) != (w.P.pair_flip != 0)) { m.a = i; m.b = j; }
    else { m.a = j; m.b = i; }
    if (w.P.debug == 1) fprintf(stderr, "oracle new manifold %d-%d  lo_i %.5f %.5f %.5f hi_i %.5f %.5f %.5f | lo_j %.5f %.5f %.5f hi_j %.5f %.5f %.5f\n", i, j, bi.aabb_lo.x, bi.aabb_lo.y, bi.aabb_lo.z, bi.aabb_hi.x, bi.aabb_hi.y, bi.aabb_hi.z, bj.aabb_lo.x, bj.aabb_lo.y, bj.aabb_lo.z, bj.aabb_hi.x, bj.aabb_hi.y, bj.aabb_hi.z);
    double ta = bounding_radius(w.bodies[m.a].shape) * w.P.breaking_factor;
    double tb = bounding_radius(w.bodies[m.b].shape) * w.P.breaking_factor;
    m.breaking = std::min(ta, tb);
    w.manifolds.push_back(m);
  }
  for (auto& m : w.manifolds) narrowphase(w, m);
}

// ---------------------------------------------------------------------------------------------
// dynamics
static void apply_external(World& w) {
  const Params& P = w.P;
  for (auto& b : w.bodies) {
    if (b.is_static) continue;
    // btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof, base only, body frame
    V3 wl = mulT(b.R, b.w), vl = mulT(b.R, b.v);
    V3 force_l = mulT(b.R, V3(0, 0, P.gravity_z * b.mass));
    V3 zero_ang = V3(), zero_lin = -force_l;
    double wn = len(wl), vn = len(vl);
    zero_ang += mulc(b.inertia, wl) * (P.ang_damping + P.ang_damping * wn);
    zero_lin += vl * (b.mass * (P.lin_damping + P.lin_damping * vn));
    if (P.gyro) zero_ang += cross(wl, mulc(b.inertia, wl));
    zero_lin += cross(wl, vl) * b.mass;
    V3 acc_ang = V3(-zero_ang.x / b.inertia.x, -zero_ang.y / b.inertia.y, -zero_ang.z / b.inertia.z);
    V3 acc_lin = V3(-zero_lin.x / b.mass, -zero_lin.y / b.mass, -zero_lin.z / b.mass);
    V3 wdot = mul(b.R, acc_ang);
    V3 vdot = mul(b.R, acc_lin + cross(wl, vl));
    for (int k = 0; k < 3; k++) {
      b.w[k] += wdot[k] * P.dt; b.w[k] = std::max(-P.max_coord_vel, std::min(P.max_coord_vel, b.w[k]));
    }
    for (int k = 0; k < 3; k++) {
      b.v[k] += vdot[k] * P.dt; b.v[k] = std::max(-P.max_coord_vel, std::min(P.max_coord_vel, b.v[k]));
    }
  }
}

static void unit_response(const Body& b, const double j[6], double d[6]) {
  if (b.is_static) { for (int k = 0; k < 6; k++) d[k] = 0; return; }
  V3 tl = mulT(b.R, V3(j[0], j[1], j[2]));
  V3 al(tl.x / b.inertia.x, tl.y / b.inertia.y, tl.z / b.inertia.z);
  V3 aw = mul(b.R, al);
  d[0] = aw.x; d[1] = aw.y; d[2] = aw.z;
  d[3] = j[3] / b.mass; d[4] = j[4] / b.mass; d[5] = j[5] / b.mass;
}

static inline double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}

static double body_vel_dot(const Body& b, const double j[6]) {
  // rel_vel += velocityVector[i]*jac[i], velocityVector = [omega, v]
  double s = 0;
  s += b.w.x * j[0]; s += b.w.y * j[1]; s += b.w.z * j[2];
  s += b.v.x * j[3]; s += b.v.y * j[4]; s += b.v.z * j[5];
  return s;
}

static void setup_contact_row(World& w, SolverRow& r, const Manifold& m, ManifoldPoint& cp, V3 axis, bool is_friction) {
  const Params& P = w.P;
  Body& A = w.bodies[m.a];
  Body& B = w.bodies[m.b];
  V3 rel1 = cp.posA - A.pos, rel2 = cp.posB - B.pos;
  r.bA = m.a; r.bB = m.b; r.cp = &cp;
  V3 t1 = cross(rel1, axis);
  r.jA[0] = t1.x; r.jA[1] = t1.y; r.jA[2] = t1.z; r.jA[3] = axis.x; r.jA[4] = axis.y; r.jA[5] = axis.z;
  V3 t2 = cross(rel2, -axis);
  r.jB[0] = t2.x; r.jB[1] = t2.y; r.jB[2] = t2.z; r.jB[3] = -axis.x; r.jB[4] = -axis.y; r.jB[5] = -axis.z;
  unit_response(A, r.jA, r.dA);
  unit_response(B, r.jB, r.dB);
  double denom0 = 0, denom1 = 0;
  for (int i = 0; i < 6; i++) denom0 += r.jA[i] * r.dA[i];
  for (int i = 0; i < 6; i++) denom1 += r.jB[i] * r.dB[i];
  double d = denom0 + denom1;
  r.inv = (d > kEps) ? 1.0 / d : 0.0;
  double rel_vel = 0;
  rel_vel += body_vel_dot(A, r.jA);   // static bodies have zero velocity
  rel_vel += body_vel_dot(B, r.jB);
  r.friction = cp.friction;
  if (P.fric_scale != 1.0 && w.step_count == P.dbg_step && A.compound) r.friction *= P.fric_scale;   // debugging what-if
  double distance = is_friction ? 0.0 : cp.dist + P.linear_slop;
  double positionalError = 0, velocityError = 0 - rel_vel;
  double erp = P.erp2;
  if (is_friction) positionalError = -distance * erp / P.dt;
  else {
    if (distance > 0) {
      if (P.pos_variant == 1) { positionalError = -distance * erp / P.dt; }                      // debugging what-ifs
      else if (P.pos_variant == 2) { positionalError = 0; velocityError -= cp.dist / P.dt; }
      else if (P.pos_variant == 3) { positionalError = 0; }
      else { positionalError = 0; velocityError -= distance / P.dt; }
    }
    else positionalError = -distance * erp / P.dt;
  }
  r.rhs = positionalError * r.inv + velocityError * r.inv;
  r.cfm = 0;
  if (is_friction) { r.lo = -r.friction; r.hi = r.friction; }
  else { r.lo = 0; r.hi = 1e10; }
  r.imp = 0;
}

static void apply_impulse(World& w, const SolverRow& r, double dimp) {
  if (r.bA >= 0 && !w.bodies[r.bA].is_static) { double* dv = &w.dv[6 * r.bA]; for (int i = 0; i < 6; i++) dv[i] += r.dA[i] * dimp; }
  if (r.bB >= 0 && !w.bodies[r.bB].is_static) { double* dv = &w.dv[6 * r.bB]; for (int i = 0; i < 6; i++) dv[i] += r.dB[i] * dimp; }
}

static double row_delta(World& w, const SolverRow& r) {
  double delta = r.rhs - r.imp * r.cfm;
  double a = 0, b = 0;
  if (r.bA >= 0) a = dot6(r.jA, &w.dv[6 * r.bA]);
  if (r.bB >= 0) b = dot6(r.jB, &w.dv[6 * r.bB]);
  delta -= a * r.inv;
  delta -= b * r.inv;
  return delta;
}

static double resolve_row(World& w, SolverRow& r) {
  double delta = row_delta(w, r);
  double sum = r.imp + delta;
  if (sum < r.lo) { delta = r.lo - r.imp; r.imp = r.lo; }
  else if (sum > r.hi) { delta = r.hi - r.imp; r.imp = r.hi; }
  else r.imp = sum;
  apply_impulse(w, r, delta);
  return r.inv != 0 ? delta / r.inv : 0.0 * (delta / r.inv);
}

static double resolve_cone(World& w, SolverRow& cA, SolverRow& cB) {
  double dB = row_delta(w, cB), sumB = cB.imp + dB;
  double dA = row_delta(w, cA), sumA = cA.imp + dA;
  const int cv = w.P.cone_variant;             // debugging what-ifs of the clip
  if (cv == 1) {                               // per-axis box clip, Jacobi inside the pair
    if (sumA < cA.lo) { dA = cA.lo - cA.imp; cA.imp = cA.lo; } else if (sumA > cA.hi) { dA = cA.hi - cA.imp; cA.imp = cA.hi; } else cA.imp = sumA;
    if (sumB < cB.lo) { dB = cB.lo - cB.imp; cB.imp = cB.lo; } else if (sumB > cB.hi) { dB = cB.hi - cB.imp; cB.imp = cB.hi; } else cB.imp = sumB;
  } else
  if (sumA * sumA + sumB * sumB >= cA.lo * cB.lo) {
    double angle = std::atan2(sumA, sumB);
    double sumAclipped = std::fabs(cA.lo * std::sin(angle));
    double sumBclipped = std::fabs(cB.lo * std::cos(angle));
    if (sumA < -sumAclipped) { dA = -sumAclipped - cA.imp; cA.imp = -sumAclipped; }
    else if (sumA > sumAclipped) { dA = sumAclipped - cA.imp; cA.imp = sumAclipped; }
    else cA.imp = sumA;
    if (sumB < -sumBclipped) { dB = -sumBclipped - cB.imp; cB.imp = -sumBclipped; }
    else if (sumB > sumBclipped) { dB = sumBclipped - cB.imp; cB.imp = sumBclipped; }
    else cB.imp = sumB;
  } else { cA.imp = sumA; cB.imp = sumB; }
  apply_impulse(w, cA, dA);
  apply_impulse(w, cB, dB);
  return dA / cA.inv + dB / cB.inv;
}

static void matrix_to_euler_xyz(const M3& mat, V3& xyz) {
  // btGeneric6DofSpring2Constraint::matrixToEulerXYZ with btGetMatrixElem(mat, idx) = mat[idx%3][idx/3]
  auto el = [&](int idx) { return mat.m[idx % 3][idx / 3]; };
  double fi = el(2);
  if (fi < 1.0) {
    if (fi > -1.0) {
      xyz.x = std::atan2(-el(5), el(8)); xyz.y = std::asin(el(2)); xyz.z = std::atan2(-el(1), el(0));
    } else { xyz.x = -std::atan2(el(3), el(4)); xyz.y = -M_PI / 2; xyz.z = 0; }
  } else { xyz.x = std::atan2(el(3), el(4)); xyz.y = M_PI / 2; xyz.z = 0; }
}

static void setup_fixed_rows(World& w, const FixedConstraint& c, std::vector<SolverRow>& rows) {
  const Params& P = w.P;
  Body& A = w.bodies[c.body];
  V3 pivotA = A.pos;          // pivotInA = 0
  const M3& frameA = A.R;     // frameInA = identity
  M3 relRot = mulTM(frameA, c.frameB);   // frameAworld.inverse() * frameBworld
  V3 angleDiff;
  matrix_to_euler_xyz(relRot, angleDiff);
  for (int i = 0; i < 6; i++) {
    SolverRow r;
    r.bA = c.body; r.bB = -1;
    V3 nl, na;
    double posError;
    if (i < 3) { nl[i] = 1; posError = dot(pivotA - c.pivotB, nl); }
    else { na = frameA.col(i % 3); posError = angleDiff[i % 3]; }
    V3 oc = cross(pivotA - A.pos, nl);
    r.jA[0] = oc.x + na.x; r.jA[1] = oc.y + na.y; r.jA[2] = oc.z + na.z; r.jA[3] = nl.x; r.jA[4] = nl.y; r.jA[5] = nl.z;
    for (int k = 0; k < 6; k++) { r.jB[k] = 0; r.dB[k] = 0; }
    unit_response(A, r.jA, r.dA);
    double d = 0;
    for (int k = 0; k < 6; k++) d += r.jA[k] * r.dA[k];
    r.inv = (d > kEps) ? 1.0 / d : 0.0;
    double rel_vel = body_vel_dot(A, r.jA);
    double velocityError = (0 - rel_vel) * 1.0;
    double positionalError = -posError * P.erp / P.dt;
    r.rhs = positionalError * r.inv + velocityError * r.inv;
    r.cfm = 0; r.lo = -c.max_impulse; r.hi = c.max_impulse; r.imp = 0;
    rows.push_back(r);
  }
}

// btAlignedObjectArray::quickSortInternal on integer keys carrying payload indices
static void bt_quicksort(std::vector<std::pair<int, int>>& a, int lo, int hi) {
  int i = lo, j = hi;
  std::pair<int, int> x = a[(lo + hi) / 2];
  do {
    while (a[i].first < x.first) i++;
    while (x.first < a[j].first) j--;
    if (i <= j) { std::swap(a[i], a[j]); i++; j--; }
  } while (i <= j);
  if (lo < j) bt_quicksort(a, lo, j);
  if (i < hi) bt_quicksort(a, i, hi);
}

static int uf_find(std::vector<int>& id, int x) {
  while (x != id[x]) { id[x] = id[id[x]]; x = id[x]; }
  return x;
}

static void solve(World& w) {
  const Params& P = w.P;
  int nb = (int)w.bodies.size();
  // islands: union over overlapping (AABB) pairs = our manifold list; constraints with the world do not merge
  std::vector<int> tagof(nb, -1), uf;
  for (int i = 0; i < nb; i++) if (!w.bodies[i].is_static) { tagof[i] = (int)uf.size(); uf.push_back((int)uf.size()); }
  for (auto& m : w.manifolds) {
    if (tagof[m.a] >= 0 && tagof[m.b] >= 0) {
      int i = uf_find(uf, tagof[m.a]), j = uf_find(uf, tagof[m.b]);
      if (i != j) uf[i] = j;
    }
  }
  for (int i = 0; i < nb; i++) w.bodies[i].tag = tagof[i] >= 0 ? uf_find(uf, tagof[i]) : -1;
  auto island_of = [&](const Manifold& m) { return w.bodies[m.a].tag >= 0 ? w.bodies[m.a].tag : w.bodies[m.b].tag; };

  std::vector<std::pair<int, int>> order;  // (island, manifold index)
  for (int i = 0; i < (int)w.manifolds.size(); i++) if (w.manifolds[i].n > 0 || true) order.push_back({island_of(w.manifolds[i]), i});
  if (order.size() > 1) bt_quicksort(order, 0, (int)order.size() - 1);
  std::vector<std::pair<int, int>> corder;
  for (int c = 0; c < 2; c++) if (w.cons[c].active) corder.push_back({w.bodies[w.cons[c].body].tag, c});
  if (corder.size() > 1) bt_quicksort(corder, 0, (int)corder.size() - 1);

  std::vector<int> islands;
  for (int i = 0; i < nb; i++) if (w.bodies[i].tag >= 0 && std::find(islands.begin(), islands.end(), w.bodies[i].tag) == islands.end()) islands.push_back(w.bodies[i].tag);
  std::sort(islands.begin(), islands.end());
  if (P.merge_islands) islands.assign(1, -2);

  w.last_iterations = 0;
  for (int isl : islands) {
    std::vector<int> mlist;
    for (auto& o : order) if (o.first == isl || isl == -2) mlist.push_back(o.second);
    int nd3 = 0; for (int c = P.manifold_order3; c > 0; c /= 10) nd3++;
    if (P.manifold_order3 > 0 && (int)mlist.size() == nd3 - 1 && w.step_count >= P.order_step_min) {
      std::sort(mlist.begin(), mlist.end());
      std::vector<int> src = mlist; int code = P.manifold_order3;
      for (int k = nd3 - 2; k >= 0; k--) { mlist[k] = src[code % 10]; code /= 10; }
    } else if (P.manifold_order >= 10000 && mlist.size() > 1 && w.step_count >= P.order_step_min) {
      // explicit order: digits of (manifold_order - 10000), zero padded to mlist.size(), index creation order
      std::sort(mlist.begin(), mlist.end());
      std::vector<int> src = mlist; int code = (w.step_count >= P.order_step_min2 ? P.manifold_order2 : P.manifold_order) - 10000; int n = (int)src.size();
      for (int k = n - 1; k >= 0; k--) { int dgt = code % 10; code /= 10; if (dgt < n) mlist[k] = src[dgt]; }
    } else if (P.manifold_order > 0 && P.manifold_order < 10000 && mlist.size() > 1) {
      std::sort(mlist.begin(), mlist.end());
      for (int k = 0; k < P.manifold_order - 1; k++) std::next_permutation(mlist.begin(), mlist.end());
    }
    if (P.debug == 1) { fprintf(stderr, "step %d solve order:", w.step_count); for (int mi : mlist) fprintf(stderr, " [%d](%d-%d n%d)", mi, w.manifolds[mi].a, w.manifolds[mi].b, w.manifolds[mi].n); fprintf(stderr, "\n"); }
    if (P.manifold_reverse) std::reverse(mlist.begin(), mlist.end());
    std::vector<SolverRow> nc, normal, fric;
    w.dv.assign(6 * nb, 0.0);
    w.ws.assign(6 * nb, 0.0);
    // contacts first (convertContacts), then constraint rows
    for (int mi : mlist) {
      Manifold& m = w.manifolds[mi];
      for (int k = 0; k < m.n; k++) {
        ManifoldPoint& cp = m.p[k];
        SolverRow r;
        setup_contact_row(w, r, m, cp, cp.normalB, false);
        if (P.warmstart > 0) {
          r.imp = cp.impN * P.warmstart;
          if (r.imp != 0) {
            // applied directly to the body velocities AND to the delta-velocity accumulator (as Bullet does)
            for (int side = 0; side < 2; side++) {
              int bi = side ? r.bB : r.bA;
              const double* d = side ? r.dB : r.dA;
              Body& b = w.bodies[bi];
              if (b.is_static) continue;
              b.w.x += d[0] * r.imp; b.w.y += d[1] * r.imp; b.w.z += d[2] * r.imp;
              b.v.x += d[3] * r.imp; b.v.y += d[4] * r.imp; b.v.z += d[5] * r.imp;
              for (int q = 0; q < 6; q++) { w.dv[6 * bi + q] += d[q] * r.imp; w.ws[6 * bi + q] += d[q] * r.imp; }
            }
          }
        }
        int nidx = (int)normal.size();
        normal.push_back(r);
        plane_space1(cp.normalB, cp.lat1, cp.lat2);
        cp.lat1 = cp.lat1 / len(cp.lat1);
        cp.lat2 = cp.lat2 / len(cp.lat2);
        SolverRow f1; setup_contact_row(w, f1, m, cp, cp.lat1, true); f1.fric_of = nidx; fric.push_back(f1);
        if (P.two_dirs) { SolverRow f2; setup_contact_row(w, f2, m, cp, cp.lat2, true); f2.fric_of = nidx; fric.push_back(f2); }
      }
    }
    for (auto& co : corder) if (co.first == isl || isl == -2) setup_fixed_rows(w, w.cons[co.second], nc);
    if (P.constraint_reverse == 0 && nc.size() == 12) std::rotate(nc.begin(), nc.begin() + 6, nc.end());
    if (nc.empty() && normal.empty()) continue;

    int it;
    for (it = 0; it < P.iterations; it++) {
      double lsr = 0;
      int row_kind = 0; double kind_max[3] = {0, 0, 0};
      auto acc = [&](double res) { kind_max[row_kind] = std::max(kind_max[row_kind], res * res);
                                   if (P.residual_max) lsr = std::max(lsr, res * res); else lsr += res * res; };
      int nn = (int)nc.size();
      for (int j = 0; j < nn; j++) {
        int index = (it & 1) ? j : nn - 1 - j;
        double rr = resolve_row(w, nc[index]);
        if (P.debug > 1 && w.step_count == 3) fprintf(stderr, "  nc row %d res %.3e imp %.6e rhs %.3e inv %.3e\n", index, rr, nc[index].imp, nc[index].rhs, nc[index].inv);
        acc(rr);
      }
      row_kind = 1;
      for (auto& r : normal) { r.prev_imp = r.imp; double rr = resolve_row(w, r); if (P.debug > 1 && w.step_count == 3) fprintf(stderr, "  normal bA %d bB %d res %.3e imp %.6e rhs %.3e\n", r.bA, r.bB, rr, r.imp, r.rhs); acc(rr); }
      row_kind = 2;
      if (P.two_dirs && P.cone_friction) {
        for (size_t j = 0; j + 1 < fric.size(); j += 2) {
          SolverRow& fa = fric[j];
          SolverRow& fb = fric[j + 1];
          double total = P.cone_variant == 2 ? normal[fa.fric_of].prev_imp : normal[fa.fric_of].imp;
          if (w.step_count == P.dbg_step && fa.fric_of == P.fs_idx0) total *= P.fs_val0;       // debugging what-ifs: per-contact scale of the cone
          if (w.step_count == P.dbg_step && fa.fric_of == P.fs_idx1) total *= P.fs_val1;
          fa.lo = -(fa.friction * total); fa.hi = fa.friction * total;
          fb.lo = -(fb.friction * total); fb.hi = fb.friction * total;
          if (P.cone_variant == 3 && !(total > 0)) continue;      // debugging what-if: no friction update while the normal impulse is zero
          acc(resolve_cone(w, fa, fb));
        }
      } else {
        for (auto& f : fric) {
          double total = normal[f.fric_of].imp;
          if (total > 0) { f.lo = -(f.friction * total); f.hi = f.friction * total; acc(resolve_row(w, f)); }
        }
      }
      if (P.debug == 6 && w.step_count >= P.dbg_step && w.step_count <= P.dbg_step + 1) {
        fprintf(stderr, "step %d it %2d:", w.step_count, it);
        for (size_t k = 0; k < normal.size(); k++) { double f1 = fric[2 * k].imp, f2 = fric[2 * k + 1].imp, lim = normal[k].friction * normal[k].imp;
          fprintf(stderr, " [N %.5e r %.6f]", normal[k].imp, lim > 0 ? std::sqrt(f1 * f1 + f2 * f2) / lim : 0.0); }
        fprintf(stderr, " servo");
        for (size_t k = 0; k < nc.size(); k++) fprintf(stderr, " %+.3e", nc[k].imp);
        fprintf(stderr, "\n");
      }
      if (P.debug) fprintf(stderr, "step %d isl %d it %d residual^2 nc %.3e normal %.3e fric %.3e\n", w.step_count, isl, it, kind_max[0], kind_max[1], kind_max[2]);
      if (P.force_iters > 0 && w.step_count == P.force_step) { if (it + 1 >= P.force_iters) { it++; break; } else continue; }
      if (lsr <= P.residual_threshold || it >= P.iterations - 1) { it++; break; }
    }
    w.last_iterations = std::max(w.last_iterations, it);
    if (P.debug == 4 && w.step_count >= P.dbg_step && w.step_count <= P.dbg_step + 2) {      // debugging: the solved impulses of one step
      for (size_t k = 0; k < normal.size(); k++) {
        const SolverRow& r = normal[k];
        double f1 = 2 * k < fric.size() ? fric[2 * k].imp : 0, f2 = 2 * k + 1 < fric.size() ? fric[2 * k + 1].imp : 0;
        fprintf(stderr, "step %d isl %d contact %zu (%d-%d) dist %+.3e impN %.6e fric %+.6e %+.6e |f|/(mu N) %.4f mu %.2f\n", w.step_count, isl, k, r.bA, r.bB,
                r.cp->dist, r.imp, f1, f2, r.imp > 0 ? std::sqrt(f1 * f1 + f2 * f2) / (r.friction * r.imp) : 0.0, r.friction);
      }
      for (size_t k = 0; k < nc.size(); k++) fprintf(stderr, "step %d servo row %zu imp %+.6e lim %.3e\n", w.step_count, k, nc[k].imp, nc[k].hi);
    }
    // write back
    for (size_t k = 0; k < normal.size(); k++) {
      ManifoldPoint* cp = normal[k].cp;
      cp->impN = normal[k].imp;
    }
    for (auto& f : fric) { /* lateral impulses are stored but friction is never warm-started */ (void)f; }
    for (int i = 0; i < nb; i++) {
      Body& b = w.bodies[i];
      if (b.is_static || (b.tag != isl && isl != -2)) continue;
      double d[6];
      // warm-start part was already written into the velocities; add only the solver's own deltas
      for (int q = 0; q < 6; q++) d[q] = w.dv[6 * i + q] - w.ws[6 * i + q];
      b.w.x += d[0]; b.w.y += d[1]; b.w.z += d[2]; b.v.x += d[3]; b.v.y += d[4]; b.v.z += d[5];
    }
  }
}

static void integrate(World& w) {
  const Params& P = w.P;
  for (auto& b : w.bodies) {
    if (b.is_static) continue;
    b.pos += b.v * P.dt;
    // btMultiBody::stepPositionsMultiDof pQuatUpdateFun, base body
    V3 angvel = b.w;
    double fAngle = len(angvel);
    const double ANGULAR_MOTION_THRESHOLD = 0.5 * 1.57079632679489661923;
    if (fAngle * P.dt > ANGULAR_MOTION_THRESHOLD) fAngle = 0.5 * 1.57079632679489661923 / P.dt;
    V3 axis;
    if (fAngle < 0.001) axis = angvel * (0.5 * P.dt - (P.dt * P.dt * P.dt) * 0.020833333333 * fAngle * fAngle);
    else axis = angvel * (std::sin(0.5 * fAngle * P.dt) / fAngle);
    // world->local quat' = quat * (-axis, cos)  <=>  local->world q' = (axis, cos) * q
    Q4 dq{axis.x, axis.y, axis.z, std::cos(fAngle * P.dt * 0.5)};
    Q4 inv{-b.q.x, -b.q.y, -b.q.z, b.q.w};
    Q4 dqi{-dq.x, -dq.y, -dq.z, dq.w};
    Q4 ninv = qnormalize(qmul(inv, dqi));
    b.q = Q4{-ninv.x, -ninv.y, -ninv.z, ninv.w};
    update_cache(b);
  }
}

static void step(World& w) {
  w.step_count++;
  if (w.P.perturb != 0 && w.step_count == w.P.dbg_step && w.obj >= 0) { w.bodies[w.obj].v.x += w.P.perturb; w.bodies[w.obj].w.y += w.P.perturb; }
  collide(w);
  apply_external(w);
  solve(w);
  integrate(w);
}

}  // namespace pgo

// =============================================================================================
// C API (ctypes)
using namespace pgo;

extern "C" {

void* pgo_create() { return new World(); }
void pgo_destroy(void* h) { delete (World*)h; }
double* pgo_params(void* h) { return (double*)&((World*)h)->P; }

// Param setters by name keep the Python side independent of struct layout
void pgo_set_param(void* h, const char* name, double v) {
  Params& P = ((World*)h)->P;
#define S(n) if (!strcmp(name, #n)) { P.n = (decltype(P.n))v; return; }
  S(dt) S(gravity_z) S(iterations) S(erp) S(erp2) S(linear_slop) S(warmstart) S(residual_threshold)
  S(lin_damping) S(ang_damping) S(gyro) S(cone_friction) S(two_dirs) S(breaking_factor) S(aabb_margin) S(tight_pad)
  S(max_coord_vel) S(manifold_order) S(manifold_order3) S(manifold_reverse) S(order_step_min) S(manifold_order2) S(order_step_min2) S(pair_order_mode) S(constraint_reverse) S(residual_max) S(merge_islands) S(force_iters) S(force_step) S(pair_flip) S(dbg_skip_new) S(pos_variant) S(cone_variant) S(tip_skip) S(tip_skip_n) S(idle_order) S(tip_idle_order) S(tip_idle_collide) S(tt_never_early) S(compound_prerefresh) S(tipwall_order) S(drop_scan_reverse) S(marginal_rule) S(find_order) S(find_order_late) S(find_order2) S(find_order2_min) S(fs_idx0) S(fs_idx1) S(fs_val0) S(fs_val1) S(skip_body) S(skip_other) S(pos_drop) S(fric_scale) S(dbg_step) S(dbg_idx) S(walls_first) S(default_max_impulse) S(perturb) S(idle_rel01) S(idle_rel02) S(idle_rel12) S(debug)
#undef S
  if (!strcmp(name, "hull_aabb_margins")) { g_hull_aabb_margins = v; return; }
  if (!strcmp(name, "compound_margin")) { g_compound_margin = v; return; }
  if (!strcmp(name, "hull_inertia_pad")) { g_hull_inertia_pad = v; return; }
  if (!strcmp(name, "pen_tie")) { g_pen_tie = v; return; }
  if (!strcmp(name, "gjk_rel_error2")) { g_gjk_rel_error2 = v; return; }
  if (!strcmp(name, "gjk_pen_tol")) { g_gjk_pen_tol = v; return; }
  if (!strcmp(name, "gjk_org_rule")) { g_gjk_org_rule = (int)v; return; }
  if (!strncmp(name, "cc_nudge", 8) && name[8] >= '0' && name[8] <= '6') { g_cc_nudge[name[8] - '0'] = v; return; }
  if (!strcmp(name, "bb_fudge")) { g_bb_fudge = v; return; }
  if (!strcmp(name, "bb_edge_point")) { g_bb_edge_point = (int)v; return; }
  if (!strcmp(name, "bb_edge_flip")) { g_bb_edge_flip = (int)v; return; }
  if (!strcmp(name, "epa_accuracy")) { epa2::EPA_ACCURACY = v; return; }
  if (!strcmp(name, "epa_plane_eps")) { epa2::EPA_PLANE_EPS = v; return; }
  if (!strcmp(name, "epa_variant")) { epa2::g_epa_variant = (int)v; return; }
  fprintf(stderr, "pgo_set_param: unknown %s\n", name);
}

// pybullet.multiplyTransforms(posA, ornA, posB, identity)[0] in SINGLE precision, as b3MultiplyTransforms computes it (b3Transform with
// float b3Scalar): b3Quaternion::length2, b3Matrix3x3::setRotation, b3Transform::operator() -- every operation one float operation in
// b3's order (this file is compiled with -ffp-contract=off).  See oracle/envsim.py b3_multiply_transforms (the numpy statement of the
// same arithmetic, kept as the documented definition and cross-checked against this one by tests/test_oracle_cpu.py).
void pgo_b3_multiply_transforms(const double* posA, const double* ornA, const double* posB, double* out) {
  const float x = (float)ornA[0], y = (float)ornA[1], z = (float)ornA[2], w = (float)ornA[3];
  const float xx0 = x * x, yy0 = y * y, zz0 = z * z, ww0 = w * w;
  float d = xx0 + yy0; d = d + zz0; d = d + ww0;
  const float s = 2.0f / d;
  const float xs = x * s, ys = y * s, zs = z * s;
  const float wx = w * xs, wy = w * ys, wz = w * zs;
  const float xx = x * xs, xy = x * ys, xz = x * zs;
  const float yy = y * ys, yz = y * zs, zz = z * zs;
  const float t0 = yy + zz, t1 = xx + zz, t2 = xx + yy;
  const float R[3][3] = {{1.0f - t0, xy - wz, xz + wy}, {xy + wz, 1.0f - t1, yz - wx}, {xz - wy, yz + wx, 1.0f - t2}};
  const float v0 = (float)posB[0], v1 = (float)posB[1], v2 = (float)posB[2];
  for (int i = 0; i < 3; i++) {
    const float a = v0 * R[i][0], b = v1 * R[i][1], c = v2 * R[i][2];
    float acc = a + b; acc = acc + c; acc = acc + (float)posA[i];
    out[i] = (double)acc;
  }
}

// plate_env:526-530 for both fingers in one call: target = multiplyTransforms(pos, quat, cur_fin) + (multiplyTransforms(0, quat, f) + vel * dt),
// the transforms in single precision (above), the sums in double in the env's order
void pgo_servo_targets(const double* pose7, const double* cur_pos6, const double* f6, const double* vel3, double dt, double* out6) {
  const double zero[3] = {0, 0, 0};
  for (int i = 0; i < 2; i++) {
    double v32[3], t32[3];
    pgo_b3_multiply_transforms(zero, pose7 + 3, f6 + 3 * i, v32);
    pgo_b3_multiply_transforms(pose7, pose7 + 3, cur_pos6 + 3 * i, t32);
    for (int k = 0; k < 3; k++) { double vg = v32[k] + vel3[k] * dt; out6[3 * i + k] = t32[k] + vg; }
  }
}

// add a body; returns index.  type 0 box (half), 1 hull (verts), 2 cylinder (half = r, r, halfh)
int pgo_add_body(void* h, int type, const double* half, const double* verts, int nverts, double mass,
                 const double* pos, const double* quat, double friction, int compound, double margin) {
  World& w = *(World*)h;
  Body b;
  b.shape.type = type;
  b.shape.half = V3(half[0], half[1], half[2]);
  b.shape.margin = margin;
  if (type == SH_CYL) { b.shape.radius = half[0]; b.shape.halfh = half[2]; }
  for (int i = 0; i < nverts; i++) b.shape.verts.push_back(V3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
  b.mass = mass; b.is_static = (mass == 0);
  if (!b.is_static) b.inertia = shape_inertia(b.shape, mass);
  b.pos = V3(pos[0], pos[1], pos[2]);
  b.q = qnormalize(Q4{quat[0], quat[1], quat[2], quat[3]});
  b.friction = friction;
  b.compound = compound != 0;
  b.shape.in_compound = b.compound && type == SH_HULL;
  b.uid = w.next_uid++;
  update_cache(b);
  w.bodies.push_back(b);
  return (int)w.bodies.size() - 1;
}

void pgo_set_roles(void* h, int obj, int tip0, int tip1) {
  World& w = *(World*)h; w.obj = obj; w.tip[0] = tip0; w.tip[1] = tip1;
}

void pgo_add_idle(void* h, int body) { ((World*)h)->idle.push_back(body); }

void pgo_set_filter(void* h, int a, int b, int enable) { set_collision_filter(*(World*)h, a, b, enable != 0); }

void pgo_reset_pose(void* h, int body, const double* pos, const double* quat) {
  World& w = *(World*)h; Body& b = w.bodies[body];
  b.pos = V3(pos[0], pos[1], pos[2]);
  b.q = qnormalize(Q4{quat[0], quat[1], quat[2], quat[3]});
  b.v = V3(); b.w = V3();
  update_cache(b);
}

void pgo_get_pose(void* h, int body, double* out7) {
  Body& b = ((World*)h)->bodies[body];
  out7[0] = b.pos.x; out7[1] = b.pos.y; out7[2] = b.pos.z; out7[3] = b.q.x; out7[4] = b.q.y; out7[5] = b.q.z; out7[6] = b.q.w;
}
// the body's AABB as btCollisionWorld::updateSingleAabb would pass it to the broadphase (extra = 0.02) or as refreshBroadphaseProxy does
// (extra = 0): used by tools/whatif/dbvt_find_order.py
void pgo_get_aabb(void* h, int body, double extra, double* out6) {
  World& w = *(World*)h; Body b = w.bodies[body];
  update_cache(b); compute_aabb(b, extra);
  out6[0] = b.aabb_lo.x; out6[1] = b.aabb_lo.y; out6[2] = b.aabb_lo.z; out6[3] = b.aabb_hi.x; out6[4] = b.aabb_hi.y; out6[5] = b.aabb_hi.z;
}
void pgo_get_vel(void* h, int body, double* out6) {
  Body& b = ((World*)h)->bodies[body];
  out6[0] = b.v.x; out6[1] = b.v.y; out6[2] = b.v.z; out6[3] = b.w.x; out6[4] = b.w.y; out6[5] = b.w.z;
}

void pgo_create_constraint(void* h, int slot, int body, const double* pivotB, const double* quatB) {
  World& w = *(World*)h; FixedConstraint& c = w.cons[slot];
  c.body = body; c.pivotB = V3(pivotB[0], pivotB[1], pivotB[2]);
  c.frameB = qmat(Q4{quatB[0], quatB[1], quatB[2], quatB[3]});
  c.max_impulse = w.P.default_max_impulse; c.active = true;
}
void pgo_change_constraint(void* h, int slot, const double* pivotB, const double* quatB, double max_force) {
  World& w = *(World*)h; FixedConstraint& c = w.cons[slot];
  c.pivotB = V3(pivotB[0], pivotB[1], pivotB[2]);
  c.frameB = qmat(Q4{quatB[0], quatB[1], quatB[2], quatB[3]});
  c.max_impulse = max_force * w.P.dt;
}
void pgo_remove_constraint(void* h, int slot) { ((World*)h)->cons[slot].active = false; }

void pgo_step(void* h) { step(*(World*)h); }
int pgo_last_iterations(void* h) { return ((World*)h)->last_iterations; }

int pgo_num_manifolds(void* h) { return (int)((World*)h)->manifolds.size(); }
// out: per manifold [a, b, n, then 4 x (posB xyz, normal xyz, dist, impN)]
int pgo_dump_manifolds(void* h, double* out, int cap) {
  World& w = *(World*)h; int k = 0;
  for (auto& m : w.manifolds) {
    if (k + 35 > cap) break;
    out[k++] = m.a; out[k++] = m.b; out[k++] = m.n;
    for (int i = 0; i < 4; i++) {
      const ManifoldPoint& p = m.p[i];
      out[k++] = p.posB.x; out[k++] = p.posB.y; out[k++] = p.posB.z;
      out[k++] = p.normalB.x; out[k++] = p.normalB.y; out[k++] = p.normalB.z;
      out[k++] = p.dist; out[k++] = p.impN;
    }
  }
  return k;
}

}  // extern "C"
