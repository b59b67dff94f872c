// Device-resident environment constants (one per pg_env handle).  Mirrors include/pregrasp_b200.h::PgEnvSpec
// plus derived quantities (inertias, breaking thresholds) computed once at pg_env_create.
#pragma once
#include "pg_math.cuh"

namespace pg {

enum { SHAPE_BOX = 0, SHAPE_HULL = 1, SHAPE_CYL = 2 };
enum { REGION_MESH = 0, REGION_BLOCK = 1, REGION_BOTTLE = 2 };
enum { MAX_STATICS = 4, MAX_REGIONS = 64, MAX_STATES = 16, MAX_HULL = 64, MAX_CROP = 2, MAX_DBOX = 4, NFING = 2,
       NKNOT = 7, SKIP = 50, SINGLE = 24, ADIM = 48 };

struct BoxBody {        // static box (floor / walls)
  double half[3], pos[3], quat[4], friction;
  int compound;         // floor: wrapped in a btCompoundShape by the URDF importer (always body A of its pairs)
  int pad;
};

struct EnvDev {
  // physics parameters (the PyBullet server defaults the recorded trajectories confirm)
  double dt, gravity_z, erp, erp2, slop, residual_threshold, lin_damping, ang_damping, breaking_factor,
      aabb_margin, max_coord_vel, default_max_impulse;
  int iterations, pad0;
  // object
  int obj_shape, n_hull, obj_compound;
  double obj_half[3], obj_mass, obj_inertia[3], obj_pos[3], obj_quat[4], obj_friction, obj_margin, obj_radius_bs;
  double hull[MAX_HULL * 3];
  double obj_aabb_c[3], obj_aabb_e[3];   // local AABB centre / half extent of the object's broadphase AABB (hull: extent + 2 margins + compound margin)
  double obj_compound_margin;            // what the compound wrapper adds on top of the child (tight) AABB
  // static scene
  int n_static, floor_index;
  BoxBody statics[MAX_STATICS];
  double static_radius_bs[MAX_STATICS];
  // tips
  double tip_half[NFING][3], tip_mass, tip_inertia[NFING][3], tip_friction, tip_radius_bs[NFING];
  // env logic
  double clearance_h, clearance_wall_size[3];
  int n_clear_walls, clear_walls[2];
  int settle_substeps, metric_axis;
  double metric_offset;
  // decode tables
  int region_type, n_regions, n_states, pad1;
  double tri_v[MAX_REGIONS * 9];     // mesh region: 3 vertices per triangle (already 0.9-scaled)
  double region_n[MAX_REGIONS * 3];  // triangle normal | block surface normal
  double block_r[MAX_REGIONS * 6];   // small-block rectangles
  int block_axis[MAX_REGIONS];
  int csg[MAX_STATES * NFING];       // region id (1-based) per state and finger
  double bottle_radius, bottle_length;
  double interp_xp[NKNOT];           // np.linspace(0, 50, 7)
};

}  // namespace pg
