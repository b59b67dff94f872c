// Host-side construction of the device constant block from the ABI spec (no CUDA calls: also used by the
// CPU logic tests, which compile pg_rollout.cuh for the host).
#pragma once
#include <math.h>
#include <string.h>
#include <string>
#include "../../include/pregrasp_b200.h"
#include "pg_env.cuh"

namespace pg {

static inline double box_radius(const double* h) { return sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]); }

static inline int build_envdev(const PgEnvSpec* sp, EnvDev& E, std::string& err) {
  if (sp->obj_shape != SHAPE_BOX && sp->obj_shape != SHAPE_HULL && sp->obj_shape != SHAPE_CYL) { err = "unknown object shape"; return PG_ERR_INVALID; }
  if (sp->obj_shape != SHAPE_BOX && (sp->n_hull < 4 || sp->n_hull > MAX_HULL || !sp->hull)) { err = "hull needs 4..64 vertices"; return PG_ERR_INVALID; }
  if (sp->n_static < 1 || sp->n_static > MAX_STATICS || sp->n_regions > MAX_REGIONS || sp->n_states > MAX_STATES || !sp->csg) {
    err = "spec out of supported bounds"; return PG_ERR_INVALID;
  }
  if (sp->n_regions < 0 || sp->n_states < 1 || sp->n_clear_walls < 0 || sp->n_clear_walls > 2 || sp->floor_index < 0 || sp->floor_index >= sp->n_static ||
      sp->settle_substeps < 0 || sp->metric_axis < 0 || sp->metric_axis > 2) { err = "negative / out-of-range count or index in the spec"; return PG_ERR_INVALID; }
  for (int w = 0; w < sp->n_clear_walls; w++)
    if (sp->clear_walls[w] < 0 || sp->clear_walls[w] >= sp->n_static) { err = "clear_walls index out of range"; return PG_ERR_INVALID; }
  if (sp->region_type != REGION_MESH && sp->region_type != REGION_BLOCK && sp->region_type != REGION_BOTTLE) { err = "unknown region type"; return PG_ERR_INVALID; }
  for (int i = 0; i < sp->n_states * NFING; i++) {
    // region ids are 1-based; 0 / negative ids wrap like the reference's python indexing (region_parse)
    const int rid = sp->csg[i] - 1, nreg = sp->region_type == REGION_BOTTLE ? 4 : sp->n_regions;
    if (rid >= nreg || rid < -nreg) { err = "contact-state graph names a region outside the region table"; return PG_ERR_INVALID; }
  }
  if (sp->region_type == REGION_MESH && (!sp->tri_v || !sp->region_n)) { err = "mesh regions need tri_v and region_n"; return PG_ERR_INVALID; }
  if (sp->region_type == REGION_BLOCK && (!sp->block_r || !sp->block_axis || !sp->region_n)) { err = "block regions need block_r, block_axis and region_n"; return PG_ERR_INVALID; }
  memset(&E, 0, sizeof(EnvDev));
  // ---- PyBullet server defaults confirmed by the recorded trajectories
  E.dt = 1.0 / 250.0; E.gravity_z = -10.0; E.erp = 0.2; E.erp2 = 0.08; E.slop = 1e-5; E.residual_threshold = 1e-7;
  E.lin_damping = 0.04; E.ang_damping = 0.04; E.breaking_factor = 0.02; E.aabb_margin = 0.02; E.max_coord_vel = 100.0;
  E.default_max_impulse = 500.0; E.iterations = 50;
  E.obj_shape = sp->obj_shape; E.n_hull = 0;
  E.obj_compound = (sp->obj_shape == SHAPE_HULL);      // URDF mesh: hull wrapped in a btCompoundShape; URDF cylinder: bare hull
  for (int i = 0; i < 3; i++) { E.obj_half[i] = sp->obj_half[i]; E.obj_pos[i] = sp->obj_pos[i]; }
  for (int i = 0; i < 4; i++) E.obj_quat[i] = sp->obj_quat[i];
  E.obj_mass = sp->obj_mass; E.obj_friction = sp->obj_friction; E.obj_margin = 0.001;
  if (sp->obj_shape == SHAPE_BOX) {
    double lx = 2 * E.obj_half[0], ly = 2 * E.obj_half[1], lz = 2 * E.obj_half[2], m = E.obj_mass;   // btBoxShape::calculateLocalInertia
    E.obj_inertia[0] = m / 12.0 * (ly * ly + lz * lz); E.obj_inertia[1] = m / 12.0 * (lx * lx + lz * lz); E.obj_inertia[2] = m / 12.0 * (lx * lx + ly * ly);
    E.obj_radius_bs = box_radius(E.obj_half);
    for (int i = 0; i < 3; i++) { E.obj_aabb_c[i] = 0; E.obj_aabb_e[i] = E.obj_half[i]; }
  } else {
    // URDF mesh -> btCompoundShape(btConvexHullShape); URDF <cylinder> -> bare 64-vertex btConvexHullShape
    E.n_hull = sp->n_hull;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int v = 0; v < sp->n_hull; v++) for (int i = 0; i < 3; i++) {
      double x = sp->hull[3 * v + i]; E.hull[3 * v + i] = x;
      if (x < lo[i]) lo[i] = x;
      if (x > hi[i]) hi[i] = x;
    }
    // btPolyhedralConvexAabbCachingShape: cached local AABB = vertex extent + margin, getAabb() adds the margin again
    // (child / tight AABB: extent + 2 margins); the btCompoundShape of a URDF mesh adds its own margin for the
    // broadphase AABB and the bounding sphere; calculateLocalInertia uses the box of extent + 3 margins either way
    double l[3], r2 = 0, c2 = 0;
    E.obj_compound_margin = E.obj_compound ? 0.001 : 0.0;
    for (int i = 0; i < 3; i++) {
      double he = 0.5 * (hi[i] - lo[i]);
      E.obj_aabb_c[i] = 0.5 * (hi[i] + lo[i]); E.obj_aabb_e[i] = he + 2 * E.obj_margin + E.obj_compound_margin;
      l[i] = 2 * ((he + E.obj_margin) + 0.002);
      r2 += E.obj_aabb_e[i] * E.obj_aabb_e[i]; c2 += E.obj_aabb_c[i] * E.obj_aabb_c[i];
    }
    double m = E.obj_mass;
    E.obj_inertia[0] = m / 12.0 * (l[1] * l[1] + l[2] * l[2]); E.obj_inertia[1] = m / 12.0 * (l[0] * l[0] + l[2] * l[2]); E.obj_inertia[2] = m / 12.0 * (l[0] * l[0] + l[1] * l[1]);
    E.obj_radius_bs = sqrt(r2) + sqrt(c2);
  }
  E.n_static = sp->n_static; E.floor_index = sp->floor_index;
  for (int s = 0; s < sp->n_static; s++) {
    memcpy(&E.statics[s], &sp->statics[s], sizeof(PgBox));
    E.static_radius_bs[s] = box_radius(sp->statics[s].half);
  }
  E.tip_mass = 0.01; E.tip_friction = sp->tip_friction;
  for (int f = 0; f < 2; f++) {
    for (int i = 0; i < 3; i++) E.tip_half[f][i] = sp->tip_half[f][i];
    double lx = 2 * E.tip_half[f][0], ly = 2 * E.tip_half[f][1], lz = 2 * E.tip_half[f][2], m = E.tip_mass;
    E.tip_inertia[f][0] = m / 12.0 * (ly * ly + lz * lz); E.tip_inertia[f][1] = m / 12.0 * (lx * lx + lz * lz); E.tip_inertia[f][2] = m / 12.0 * (lx * lx + ly * ly);
    E.tip_radius_bs[f] = box_radius(E.tip_half[f]);
  }
  E.clearance_h = sp->clearance_h;
  for (int i = 0; i < 3; i++) E.clearance_wall_size[i] = sp->clearance_wall_size[i];
  E.n_clear_walls = sp->n_clear_walls; E.clear_walls[0] = sp->clear_walls[0]; E.clear_walls[1] = sp->clear_walls[1];
  E.settle_substeps = sp->settle_substeps; E.metric_axis = sp->metric_axis; E.metric_offset = sp->metric_offset;
  E.region_type = sp->region_type; E.n_regions = sp->n_regions; E.n_states = sp->n_states;
  if (sp->tri_v) memcpy(E.tri_v, sp->tri_v, sizeof(double) * 9 * sp->n_regions);
  if (sp->region_n) memcpy(E.region_n, sp->region_n, sizeof(double) * 3 * sp->n_regions);
  if (sp->block_r) memcpy(E.block_r, sp->block_r, sizeof(double) * 6 * sp->n_regions);
  if (sp->block_axis) for (int i = 0; i < sp->n_regions; i++) E.block_axis[i] = sp->block_axis[i];
  for (int i = 0; i < sp->n_states * 2; i++) E.csg[i] = sp->csg[i];
  E.bottle_radius = 0.045; E.bottle_length = 0.22;
  for (int i = 0; i < NKNOT; i++) E.interp_xp[i] = (i == NKNOT - 1) ? 50.0 : (double)i * (50.0 / 6.0);   // np.linspace(0,50,7)

  return PG_OK;
}

}  // namespace pg
