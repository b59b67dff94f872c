// One rollout = reset + N env steps (action decode, pre-simulate, 50 servo substeps, settle on the last).
// Restates the control flow of envs/plate_contact_graph_env.py:253-632 and the worker loop of
// stoch_traj_opt.py:124-136 for ONE sample; the kernel in pg_rollout.cu runs K of them, one per thread.
#pragma once
#include "pg_physics.cuh"

namespace pg {

struct Carry {            // last_fins / last_surface_norm carried across env steps (plate_env:137, :112-114)
  V3 pos[NFING], norm[NFING];
  int on[NFING];
};

PG_HD void region_parse(const EnvDev& E, int state, int f, double s0, double s1, V3& pos, V3& nrm) {
  int rid = E.csg[state * NFING + f] - 1;                    // contact_state_graph.py:37-38, mesh_region.py:37
  double sa0 = (s0 + 1) * 0.5, sa1 = (s1 + 1) * 0.5;
  if (E.region_type == REGION_MESH) {
    if (rid < 0) rid += E.n_regions;
    const double* v = &E.tri_v[rid * 9];
    double sx = sqrt(sa0);
    double w0 = 1 - sx, w1 = sx * (1 - sa1), w2 = sx * sa1;      // mesh_region.py:19-23
    pos = V3();
    pos = pos + V3(v[0], v[1], v[2]) * w0;
    pos = pos + V3(v[3], v[4], v[5]) * w1;
    pos = pos + V3(v[6], v[7], v[8]) * w2;
    nrm = V3(E.region_n[rid * 3], E.region_n[rid * 3 + 1], E.region_n[rid * 3 + 2]);
  } else if (E.region_type == REGION_BLOCK) {
    if (rid < 0) rid += E.n_regions;
    const double* r = &E.block_r[rid * 6];
    int fa = E.block_axis[rid];
    if (fa == 0) pos = V3(r[0], (r[3] - r[2]) * sa0 + r[2], (r[5] - r[4]) * sa1 + r[4]);       // small_block_region.py:24-31
    else if (fa == 1) pos = V3((r[1] - r[0]) * sa0 + r[0], r[2], (r[5] - r[4]) * sa1 + r[4]);
    else pos = V3((r[1] - r[0]) * sa0 + r[0], (r[3] - r[2]) * sa1 + r[2], r[4]);
    nrm = V3(E.region_n[rid * 3], E.region_n[rid * 3 + 1], E.region_n[rid * 3 + 2]);
  } else {
    const double PI = 3.141592653589793;
    double radius = E.bottle_radius, length = E.bottle_length;   // waterbottle_region.py:9-40
    if (rid == 0 || rid == 3) {
      double rr = radius * sqrt(sa0), th = sa1 * 2 * PI;
      pos = V3(-rr * sin(th), -rr * cos(th), rid == 0 ? length / 2 : -length / 2);
      nrm = V3(0, 0, rid == 0 ? 1.0 : -1.0);
    } else {
      double z = sa0 * length - length / 2;
      double th = sa1 * PI + (rid == 2 ? PI : 0.0);
      double x = -radius * sin(th), y = -radius * cos(th);
      pos = V3(x, y, z);
      double l = sqrt(x * x + y * y);
      nrm = V3(x / l, y / l, 0 / l);
    }
  }
}

PG_HD M3 rot_from_z(V3 n) {      // utils/math_utils.py:457-474 with vec1 = (0,0,1)
  double ln = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
  V3 b(n.x / ln, n.y / ln, n.z / ln);
  V3 v(0 * b.z - 1 * b.y, 1 * b.x - 0 * b.z, 0);
  double c = b.z, s = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  M3 R;
  for (int i = 0; i < 9; i++) R.m[i] = 0;
  if (s == 0) {
    if (c > 0) { R.m[0] = 1; R.m[4] = 1; R.m[8] = 1; }
    else { R.m[0] = -1; R.m[4] = 1; R.m[8] = -1; }
    return R;
  }
  double k[9] = {0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0};
  double f = (1 - c) / (s * s);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double kk = 0;
    for (int q = 0; q < 3; q++) kk += k[3 * i + q] * k[3 * q + j];
    R.m[3 * i + j] = ((i == j) ? 1.0 : 0.0) + k[3 * i + j] + kk * f;
  }
  return R;
}

// np.interp(t, linspace(0,50,7), y) for integer t in [0,50)
PG_HD double interp7(const EnvDev& E, const double* y, int stride, int t) {
  double x = (double)t;
  int j = 0;
  while (j < NKNOT - 2 && E.interp_xp[j + 1] <= x) j++;
  if (E.interp_xp[j] == x) return y[j * stride];
  double slope = (y[(j + 1) * stride] - y[j * stride]) / (E.interp_xp[j + 1] - E.interp_xp[j]);
  return slope * (x - E.interp_xp[j]) + y[j * stride];
}

PG_HD void decode_step(const EnvDev& E, const double* a_raw, Carry& carry, int state, int prev_state,
                       V3* cur_pos, int* cur_on, double knots[NFING][NKNOT][3]) {
  V3 new_norm[NFING];
  for (int f = 0; f < NFING; f++) {
    double sub[SINGLE];
    for (int i = 0; i < SINGLE; i++) sub[i] = tanh(a_raw[f * SINGLE + i]);          // plate_env:487
    V3 pos, nrm;
    new_norm[f] = carry.norm[f];
    if (carry.on[f]) { pos = carry.pos[f]; nrm = carry.norm[f]; }                 // :290-305
    else {
      region_parse(E, state, f, sub[SINGLE - 3], sub[SINGLE - 2], pos, nrm);      // :307
      pos = pos + nrm * 0.01;                                                     // :279
      new_norm[f] = nrm;
    }
    M3 R = rot_from_z(nrm);
    for (int i = 0; i < NKNOT; i++) {                                             // :313-330
      double vn = 0, vt1 = 0, vt2 = 0;
      if (sub[3 * i] > 0) {
        vn = sub[3 * i] * 5.0 * E.dt;
        vt1 = sub[3 * i + 1] * 1.0 * vn;
        vt2 = sub[3 * i + 2] * 1.0 * vn;
      }
      V3 t = mul(R, V3(vt1, vt2, 0));
      knots[f][i][0] = -vn * nrm.x + t.x; knots[f][i][1] = -vn * nrm.y + t.y; knots[f][i][2] = -vn * nrm.z + t.z;
    }
    bool broke = (prev_state >= 0 && E.csg[state * NFING + f] != E.csg[prev_state * NFING + f]) && carry.on[f];
    int flag = broke ? 0 : (sub[SINGLE - 1] > 0 ? 1 : 0);                           // :336-339
    if (!flag) {                                                                  // :380-382
      for (int i = 0; i < NKNOT; i++) for (int c = 0; c < 3; c++) knots[f][i][c] = 0.0;
      pos = V3(100.0, 100.0, 100.0);
    }
    if (knots[f][0][0] == 0) for (int i = 0; i < NKNOT; i++) for (int c = 0; c < 3; c++) knots[f][i][c] = 0.0;  // :384
    cur_pos[f] = pos; cur_on[f] = flag;
  }
  for (int f = 0; f < NFING; f++) { carry.pos[f] = cur_pos[f]; carry.on[f] = cur_on[f]; carry.norm[f] = new_norm[f]; }
}

PG_HD double min_dist_box(const EnvDev& E, int wall, V3 p) {     // utils/math_utils.py:476-490
  const BoxBody& s = E.statics[wall];
  Q4 q; q.x = s.quat[0]; q.y = s.quat[1]; q.z = s.quat[2]; q.w = s.quat[3];
  // pytorch_kinematics.quaternion_to_matrix: two_s = 2/|q|^2
  double two_s = 2.0 / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  M3 R;
  R.m[0] = 1 - two_s * (q.y * q.y + q.z * q.z); R.m[1] = two_s * (q.x * q.y - q.z * q.w); R.m[2] = two_s * (q.x * q.z + q.y * q.w);
  R.m[3] = two_s * (q.x * q.y + q.z * q.w); R.m[4] = 1 - two_s * (q.x * q.x + q.z * q.z); R.m[5] = two_s * (q.y * q.z - q.x * q.w);
  R.m[6] = two_s * (q.x * q.z - q.y * q.w); R.m[7] = two_s * (q.y * q.z + q.x * q.w); R.m[8] = 1 - two_s * (q.x * q.x + q.y * q.y);
  V3 l = mulT(R, p - V3(s.pos[0], s.pos[1], s.pos[2]));
  double dx = fmax(fabs(l.x) - E.clearance_wall_size[0], 0.0), dy = fmax(fabs(l.y) - E.clearance_wall_size[1], 0.0),
         dz = fmax(fabs(l.z) - E.clearance_wall_size[2], 0.0);
  return sqrt(dx * dx + dy * dy + dz * dz);
}

PG_HD bool check_clearance(const EnvDev& E, V3 p) {             // plate_env:463 / bookshelf_graph_env.py:645-654
  if (p.z < E.clearance_h) return false;
  for (int i = 0; i < E.n_clear_walls; i++) if (min_dist_box(E, E.clear_walls[i], p) < E.clearance_h) return false;
  return true;
}

struct RolloutOut {
  double* pose;       // [N][7]  object pose (xyz + quat xyzw) at each env step's cost evaluation
  double* fins;       // [N][NFING][3] cur_fins local positions at each env step
  double* metric;     // scalar
  double* trace;      // optional [N*SKIP][7]: pose at the START of every control substep (model_test.py:124-135)
  int* iters;         // optional [N*SKIP]
  int* settle_iters;  // optional [settle_substeps] (host diagnostics only)
  int sync_cta;       // device: every thread of this CTA runs a rollout in this pass (see substep)
  void* regroup;      // device: [K] RegroupSlot exchange buffer, or null (see settle_regroup)
  int k;              // device: index of the rollout this thread currently runs
  int k0;             // device: index of the CTA's first rollout in this pass (regroup exchanges stay inside [k0, k0 + blockDim.x))
};

// Settle-phase regrouping (device only).  The warp pays for its slowest lane: a rollout whose contact solver stops at the
// residual threshold after ~7 sweeps waits for lanes that run all 50, and which kind a rollout is hardly changes once the
// fingers hold still.  Before the settle substeps (49-74 % of a rollout) the rollouts of a CTA are therefore re-dealt to
// its threads in order of the solver work they needed during the last control phase, so that warps hold rollouts of one
// kind.  The state travels through a global exchange buffer (5 KB per rollout, once); results are unchanged bit for bit
// because a rollout's arithmetic does not depend on the thread that runs it.
struct RegroupSlot {
  Sim S;
  double m0;
  double* pose;
  double* metric;
  int k;
};

#if defined(__CUDA_ARCH__)
__device__ __noinline__ void settle_regroup(Sim& S, double& m0, RolloutOut& out, int key) {
  RegroupSlot* slots = (RegroupSlot*)out.regroup;
  int* keys = (int*)pg_solver_sm;                 // the solver slices are dead between substeps
  int* src = keys + blockDim.x;
  const int me = threadIdx.x, n = blockDim.x;
  __syncthreads();                                // every thread has left its last solve
  RegroupSlot& mine = slots[out.k0 + me];      // slot of this THREAD (the rollout it currently runs may come from an earlier re-deal)
  mine.S = S; mine.m0 = m0; mine.pose = out.pose; mine.metric = out.metric; mine.k = out.k;
  keys[me] = key;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < n; j++) { const int kj = keys[j]; rank += (kj < key || (kj == key && j < me)) ? 1 : 0; }
  src[rank] = me;
  __syncthreads();
  const int from = src[me];
  const RegroupSlot& theirs = slots[out.k0 + from];
  S = theirs.S; m0 = theirs.m0; out.pose = theirs.pose; out.metric = theirs.metric;
  out.k = theirs.k;
  __syncthreads();                                // keys / src are overwritten by the next solve
}
#endif

// ---- reset(): plate_env:176-250.  Bodies at their spawn poses, tips parked at (f+2, f+2, f+2), one substep (:237).
template <int SHAPE>
PG_HD void env_reset(const EnvDev& E, Sim& S, const M3* statR, RowSet& rs, Carry& carry, int sync_cta, double* sm, int ss) {
  {
    DynBody& o = S.dyn[0];
    o.pos = V3(E.obj_pos[0], E.obj_pos[1], E.obj_pos[2]);
    Q4 q; q.x = E.obj_quat[0]; q.y = E.obj_quat[1]; q.z = E.obj_quat[2]; q.w = E.obj_quat[3];
    o.q = qnormalize(q); o.R = qmat(o.q);
    o.mass = E.obj_mass; o.inertia = V3(E.obj_inertia[0], E.obj_inertia[1], E.obj_inertia[2]);
    o.half = V3(E.obj_half[0], E.obj_half[1], E.obj_half[2]); o.friction = E.obj_friction;
    for (int f = 0; f < NFING; f++) {
      DynBody& t = S.dyn[1 + f];
      t.pos = V3(f + 2.0, f + 2.0, f + 2.0);
      t.q.x = 0; t.q.y = 0; t.q.z = 0; t.q.w = 1; t.R = qmat(t.q);
      t.mass = E.tip_mass; t.inertia = V3(E.tip_inertia[f][0], E.tip_inertia[f][1], E.tip_inertia[f][2]);
      t.half = V3(E.tip_half[f][0], E.tip_half[f][1], E.tip_half[f][2]); t.friction = E.tip_friction;
    }
  }
  S.nman = 0; S.step = 0;
  S.ndisp = 3; S.disp[0] = DISP_IDLE; S.disp[1] = DISP_IDLE + 1; S.disp[2] = DISP_IDLE + 2;     // the idle fingertips' pairs are the oldest (pg_physics.cuh)
  S.filter_off = 0; S.obj_first = 1; S.last_iterations = 0; S.early = 0; S.sync_cta = sync_cta;
  for (int f = 0; f < 2; f++) { S.plo[f] = V3(); S.phi[f] = V3(); }
  S.cons[0].active = 0; S.cons[1].active = 0;
  substep<SHAPE>(E, S, statR, rs, sm, ss);                                                       // :237
  S.obj_first = 0;
  for (int f = 0; f < NFING; f++) { carry.on[f] = 0; carry.pos[f] = V3(); carry.norm[f] = V3(); }
}

// ---- step(a), control part: plate_env:253-555 for env step j of the episode (a = the raw action, before tanh): decode,
// pre_simulate, the 50 servo substeps.  Writes out.fins row j (+ trace / iters rows of this step); returns the solver work of
// the servo phase (the key the rollouts are re-dealt / sorted by).
template <int SHAPE>
PG_HD int env_control(const EnvDev& E, Sim& S, const M3* statR, RowSet& rs, Carry& carry, const double* a, int state, int prev_state,
                      double max_force, int j, RolloutOut& out, double* sm, int ss) {
  V3 cur_pos[NFING];
  int cur_on[NFING];
  double knots[NFING][NKNOT][3];
  decode_step(E, a, carry, state, prev_state, cur_pos, cur_on, knots);
  // ---- pre_simulate :454-482
  {
    DynBody& o = S.dyn[0];
    M3 R = qmat(o.q);
    for (int i = 0; i < NFING; i++) {
      V3 g = b3_multiply_transforms(o.pos, o.q, cur_pos[i]);       // float32, as pybullet.multiplyTransforms returns it
      if (!check_clearance(E, g)) g = V3(100.0, 100.0, 100.0);
      S.filter_off |= (1u << i);
      refresh_proxy<SHAPE>(E, S, B_TIP0 + i);
      refresh_proxy<SHAPE>(E, S, B_OBJ);
      DynBody& t = S.dyn[1 + i];
      t.pos = g; t.q = qnormalize(o.q); t.v = V3(); t.w = V3(); t.R = qmat(t.q);
      S.cons[i].pivotB = g; S.cons[i].frameB = R; S.cons[i].max_impulse = E.default_max_impulse; S.cons[i].active = 1;
    }
    substep<SHAPE>(E, S, statR, rs, sm, ss);
    for (int i = 0; i < NFING; i++) {
      S.filter_off &= ~(1u << i);
      refresh_proxy<SHAPE>(E, S, B_TIP0 + i);
      refresh_proxy<SHAPE>(E, S, B_OBJ);
    }
  }
  // ---- servo loop :504-533
  int solver_work = 0;                            // regroup key: solver work of this control phase
  for (int t = 0; t < SKIP; t++) {
    DynBody& o = S.dyn[0];
    if (out.trace) {
      double* p = out.trace + (j * SKIP + t) * 7;
      p[0] = o.pos.x; p[1] = o.pos.y; p[2] = o.pos.z; p[3] = o.q.x; p[4] = o.q.y; p[5] = o.q.z; p[6] = o.q.w;
    }
    M3 R = qmat(o.q);
    V3 vel = o.v;
    for (int i = 0; i < NFING; i++) {
      V3 f(interp7(E, &knots[i][0][0], 3, t), interp7(E, &knots[i][0][1], 3, t), interp7(E, &knots[i][0][2], 3, t));
      V3 v_g = b3_multiply_transforms(V3(), o.q, f) + vel * E.dt;       // the two transforms in float32 (pg_math.cuh), the sums in double
      V3 tar = b3_multiply_transforms(o.pos, o.q, cur_pos[i]) + v_g;
      S.cons[i].pivotB = tar; S.cons[i].frameB = R; S.cons[i].max_impulse = max_force * E.dt;
    }
    substep<SHAPE>(E, S, statR, rs, sm, ss);
#if !defined(PG_PHASE_CLOCKS)
    if (out.iters) out.iters[j * SKIP + t] = S.last_iterations;
#endif
    solver_work += S.last_iterations * (12 + 2 * rs.n_pts);    // sweeps x rows
  }
  for (int i = 0; i < NFING; i++) {
    out.fins[(j * NFING + i) * 3 + 0] = cur_pos[i].x;
    out.fins[(j * NFING + i) * 3 + 1] = cur_pos[i].y;
    out.fins[(j * NFING + i) * 3 + 2] = cur_pos[i].z;
  }
  return solver_work;
}

PG_HD void write_pose(const Sim& S, double* p) {
  const DynBody& o = S.dyn[0];
  p[0] = o.pos.x; p[1] = o.pos.y; p[2] = o.pos.z; p[3] = o.q.x; p[4] = o.q.y; p[5] = o.q.z; p[6] = o.q.w;
}

// first half of the settle metric (plate_env:558): the object's height / depth before the settle substeps
PG_HD double settle_metric_half(const EnvDev& E, const Sim& S) { return (S.dyn[0].pos[E.metric_axis] + E.metric_offset) * 0.5; }

// ---- step(a) = control part + (last step, train mode) settle + metric :556-603, then the pose the cost is evaluated at
template <int SHAPE>
PG_HD void env_step(const EnvDev& E, Sim& S, const M3* statR, RowSet& rs, Carry& carry, const double* a, int state, int prev_state,
                    double max_force, bool settle, int j, RolloutOut& out, double* sm, int ss) {
  int solver_work = env_control<SHAPE>(E, S, statR, rs, carry, a, state, prev_state, max_force, j, out, sm, ss);
  if (settle) {
    double m0 = settle_metric_half(E, S);
#if defined(__CUDA_ARCH__)
    if (out.regroup) settle_regroup(S, m0, out, solver_work);     // from here on this thread may run another rollout of its CTA
#endif
    int recent_work = 0;                              // solver work since the last re-deal
    for (int s = 0; s < E.settle_substeps; s++) {
#if defined(__CUDA_ARCH__) && defined(PG_REGROUP_EVERY)
      // the kind of a rollout drifts while it settles: re-deal again every PG_REGROUP_EVERY substeps by the work just seen
      if (out.regroup && s > 0 && s % PG_REGROUP_EVERY == 0) { settle_regroup(S, m0, out, recent_work); recent_work = 0; }
#endif
      substep<SHAPE>(E, S, statR, rs, sm, ss);
      recent_work += S.last_iterations * (12 + 2 * rs.n_pts);
      if (out.settle_iters) out.settle_iters[s] = S.last_iterations;
    }
    if (out.metric) *out.metric = m0 + settle_metric_half(E, S);
  }
  write_pose(S, out.pose + j * 7);
}

PG_HD void static_rotations(const EnvDev& E, M3* statR) {
  for (int s = 0; s < E.n_static; s++) {
    Q4 q; q.x = E.statics[s].quat[0]; q.y = E.statics[s].quat[1]; q.z = E.statics[s].quat[2]; q.w = E.statics[s].quat[3];
    statR[s] = qmat(qnormalize(q));
  }
}

// sm / ss: the thread's on-chip solver slice (element i at sm[i*ss]); SM_DOUBLES elements
template <int SHAPE>
PG_HDN void rollout_one(const EnvDev& E, const double* u, const float* noise, int N, const int* path,
                        double max_force, int train, RolloutOut out, double* sm, int ss) {
  Sim S;
  RowSet rs;
  M3 statR[MAX_STATICS];
  static_rotations(E, statR);
#if defined(PG_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
  for (int i = 0; i < 6; i++) S.clk[i] = 0;
  PG_CLK0(roll_t0);
#endif
  Carry carry;
  env_reset<SHAPE>(E, S, statR, rs, carry, out.sync_cta, sm, ss);
  for (int j = 0; j < N; j++) {
    double a[ADIM];
    for (int d = 0; d < ADIM; d++) a[d] = u[j * ADIM + d] + (noise ? (double)noise[j * ADIM + d] : 0.0);
    env_step<SHAPE>(E, S, statR, rs, carry, a, path[j], j == 0 ? -1 : path[j - 1], max_force, j == N - 1 && train, j, out, sm, ss);
  }
#if defined(PG_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
  // profiling builds: lane 0 of every warp adds its phase cycles to the counters behind out.iters (6 x unsigned long long)
  if (out.iters && (threadIdx.x & 31) == 0) {
    long long total = clock64() - roll_t0, sum = 0;
    unsigned long long* g = (unsigned long long*)out.iters;
    for (int i = 0; i < 5; i++) { atomicAdd(g + i, (unsigned long long)S.clk[i]); sum += S.clk[i]; }
    atomicAdd(g + 5, (unsigned long long)(total - sum));
  }
#endif
}

// The state one episode carries from env step to env step when it is advanced one step at a time (pg_env_step_single): the
// rigid-body world (bodies, persistent manifolds, servo constraints, broadphase bookkeeping) and the decode carry
// (last_fins / last_surface_norm, plate_env:112-114, :137).  Opaque to the ABI (PgCarry).
struct EpisodeState {
  Sim S;
  Carry carry;
  int step, prev_state;
};

}  // namespace pg
