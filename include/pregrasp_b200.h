/* pregrasp_b200 -- C ABI of the B200-native MPPI rollout path of synthesize_pregrasp.
 *
 * The reference (Python) has no FFI seam; these entry points are what a binding for its hot path binds.
 * Each one names the reference interface it replaces (paths relative to the reference root).
 * Conventions: every function returns 0 on success, a PG_ERR_* code otherwise; pg_last_error() gives the
 * text of the last failure on the calling thread.  No exceptions cross the ABI.  Pointers named *_dev are
 * device pointers on the handle's GPU, *_host are host pointers.  `stream` is a cudaStream_t passed as
 * void* (NULL = default stream); *_dev calls are asynchronous on that stream.  Handles are not thread-safe;
 * independent handles are.  Quaternions are xyzw.
 */
#ifndef PREGRASP_B200_H
#define PREGRASP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_INVALID 1      /* bad argument */
#define PG_ERR_CUDA 2         /* CUDA runtime error (text in pg_last_error) */
#define PG_ERR_UNSUPPORTED 3  /* valid reference feature not built */

#define PG_ACTION_DIM 48      /* 2 fingers x (7 knots x 3 + 2 position params + on/off), plate_env:145-147 */
#define PG_CONTROL_SKIP 50    /* plate_env:51 */
#define PG_MAX_STATICS 4

typedef struct PgBox {        /* static box: setup_pybullet.py create_* wall / plane.urdf floor */
  double half[3], pos[3], quat[4], friction;
  int32_t compound;           /* 1 for the URDF floor (collision origin offset keeps it a compound shape) */
  int32_t pad;
} PgBox;

typedef struct PgScoreWeights { /* neurals/network.py:434-441 LargeScoreFunction(num_fingers=2, latent_dim=20, df) */
  const float *conv1_w, *conv1_b; /* [32,4], [32]   */
  const float *conv2_w, *conv2_b; /* [64,32], [64]  */
  const float *conv3_w, *conv3_b; /* [256,128], [256] */
  const float *conv4_w, *conv4_b; /* [20,256], [20] */
  const float *fc1_w, *fc1_b;     /* [512,26], [512] */
  const float *fc2_w, *fc2_b;     /* [512,512], [512] */
  const float *fc3_w, *fc3_b;     /* [256,512], [256] */
  const float *out_w, *out_b;     /* [1,256], [1] */
} PgScoreWeights;

typedef struct PgEnvSpec {     /* the constructor kwargs + literals of envs/<env>_contact_graph_env.py */
  /* dynamic object */
  int32_t obj_shape;           /* 0 box, 1 convex hull of a URDF mesh (child of a compound shape), 2 URDF <cylinder>:
                                  the 64-vertex prism hull PyBullet builds for it (hull = its vertices) */
  int32_t n_hull;
  double obj_half[3];
  const double* hull;          /* [n_hull,3] or NULL */
  double obj_mass, obj_pos[3], obj_quat[4], obj_friction;
  /* static scene, creation order */
  int32_t n_static, floor_index;
  PgBox statics[PG_MAX_STATICS];
  /* fingertips (thumb, index) */
  double tip_half[2][3], tip_friction;
  /* env logic */
  double clearance_h, clearance_wall_size[3];
  int32_t n_clear_walls, clear_walls[2];
  int32_t project_uses_clearance; /* 0: plate rule point.z > CLEARANCE_H (plate_env:669); 1: checkClearance(point) */
  int32_t settle_substeps, metric_axis;
  double metric_offset;
  /* contact regions + contact state graph */
  int32_t region_type, n_regions, n_states, pad0;
  const double* tri_v;         /* mesh regions [n_regions,3,3] */
  const double* region_n;      /* [n_regions,3] */
  const double* block_r;       /* small-block regions [n_regions,6] */
  const int32_t* block_axis;   /* [n_regions] */
  const int32_t* csg;          /* [n_states,2] 1-based region ids */
  /* cost: object cloud, de-occlusion boxes, distance-field meshes, score network */
  int32_t n_cloud, n_crop;
  const double* cloud;         /* [n_cloud,3] object frame */
  double crop[2][2][3];        /* (min,max) world AABBs, union */
  int32_t n_dist_box, n_dist_tri;
  double dist_box[4][2][3];    /* (min,max) corners of the distance-field boxes */
  const double* dist_tri;      /* [n_dist_tri,3,3] or NULL */
  /* right prisms with a regular polygon cross-section, axis z: what Open3D's create_cylinder tessellates a cylinder into
     (waterbottle holders, distancefield_utils.py:62-68).  Exactly the same surface as its triangle list, evaluated in closed
     form: (cx, cy, cz, radius of the vertex circle, half height, number of sides <= 32); first vertex at angle 0 */
  int32_t n_dist_prism, pad1;
  double dist_prism[2][6];
  PgScoreWeights weights;
} PgEnvSpec;

typedef struct PgEnv PgEnv;

const char* pg_last_error(void);
const char* pg_version(void);

/* replaces env_fn(render=False, device=..., **kwargs) in the worker, stoch_traj_opt.py:113 */
int pg_env_create(const PgEnvSpec* spec, int device, PgEnv** out);
int pg_env_destroy(PgEnv* env);

/* replaces the worker's "control" branch, stoch_traj_opt.py:119-136: K rollouts of N env steps from the
 * post-reset state with actions u[j] + noise[k,j], cost J[k] = sum_j -reward (reward = 100 x score logit).
 * u_dev [N,48] f64; noise_dev [K,N,48] f32 or NULL (=> K copies of u, replay_traj :233-247);
 * path_host [N] contact-state ids; J_dev [K] f64; metric_dev [K] f64 or NULL (info["metric"]);
 * trace_dev [K,N*50,7] f64 or NULL (object pose at the start of every control substep, model_test.py:124-135). */
int pg_rollout(PgEnv* env, const double* u_dev, const float* noise_dev, int K, int N, const int32_t* path_host,
               double max_force, double* J_dev, double* metric_dev, double* trace_dev, void* stream);

/* same call with HOST buffers: copies in, runs, copies J/metric back, synchronises (the plugin-facing form) */
int pg_rollout_host(PgEnv* env, const double* u_host, const float* noise_host, int K, int N,
                    const int32_t* path_host, double max_force, double* J_host, double* metric_host,
                    double* trace_host);

/* Q independent MPPI problems in one launch.  Replaces the serial outer loops of model_optimize.py:74-117 (one optimiser per
   contact-state path and per repetition): rollout k uses u[prob[k]] and paths[prob[k]].  Device: u [Q,N,48] f64, noise [K,N,48] f32,
   prob [K] i32, J [K], metric [K] or NULL; host: paths [Q,N] i32.  The prob values must lie in [0,Q) (not checked on the device). */
int pg_rollout_multi(PgEnv* env, const double* u_dev, const float* noise_dev, int K, int N, int Q, const int32_t* prob_dev,
                     const int32_t* paths_host, double max_force, double* J_dev, double* metric_dev, void* stream);

/* the two stages of pg_rollout, exposed for tests and profiling */
int pg_physics(PgEnv* env, const double* u_dev, const float* noise_dev, int K, int N, const int32_t* path_host,
               double max_force, double* pose_dev /*[K,N,7]*/, double* fins_dev /*[K,N,2,3]*/,
               double* metric_dev, double* trace_dev, int32_t* iters_dev, void* stream);
/* calc_reward_value_stage_one, plate_env:393-437: B evaluations -> logits */
int pg_cost(PgEnv* env, const double* pose_dev /*[B,7]*/, const double* fins_dev /*[B,2,3]*/, int B,
            float* logit_dev /*[B]*/, void* stream);
/* LargeScoreFunction.pred_score on explicit inputs (neurals/network.py:443-455): pts [B,P,3] f32, df [B,P] f32,
 * mask [B,P] u8 (0 = cropped away), grasp [B,6] f32 -> logit [B]; what generate_csg.py:108-130 batches */
int pg_score(PgEnv* env, const float* pts_dev, const float* df_dev, const uint8_t* mask_dev, const float* grasp_dev,
             int B, int P, float* logit_dev, void* stream);

/* replaces stoch_traj_opt.py:195-204.  J [K] f64, noise [K,N,D] f32, u_inout [N,D] f64.
 * partial_dev (nullable) receives (min J, sum w, sum J, sum w*noise[N*D]) = 3+N*D doubles for the multi-GPU merge;
 * when apply != 0 the update u += alpha * sum(w noise)/sum(w) is applied in place. */
int pg_mppi_update(const double* J_dev, const float* noise_dev, int K, int N, int D, double kappa, double alpha,
                   double* u_inout_dev, double* partial_dev, int apply, void* stream);
/* merge G partials (host side of the per-iteration exchange) and apply: u += alpha * V/S */
int pg_mppi_merge_apply(const double* partials_host /*[G,3+N*D]*/, int G, int N, int D, double kappa, double alpha,
                        double* u_inout_host);

/* utils/dyn_feasibility_check.py:153-235 simple_check_dyn_feasible: pts/nrm [B,M,3] f64, count [B] (<= M) or NULL */
int pg_dyn_feasible_simple(const double* pts_dev, const double* nrm_dev, const int32_t* count_dev, int B, int M,
                           uint8_t* out_dev, void* stream);
/* all C(P,3) triples of a cloud (find_dynamically_feasible_contacts :239-303 with the simple criterion and the
 * pairwise min-distance prefilter :260-265); triple index range [first,last) for multi-GPU partitioning.
 * triples_out_dev [cap,3] i32 (nullable), count_dev [1] u64 */
int pg_dyn_sweep_simple(const double* cloud_pts_dev, const double* cloud_nrm_dev, int P, double min_dist,
                        uint64_t first, uint64_t last, int32_t* triples_out_dev, uint64_t cap,
                        unsigned long long* count_dev, void* stream);

/* utils/dyn_feasibility_check.py:28-112 check_dyn_feasible: the wrench-closure QP (min 1/2 |net wrench|^2 over friction pyramids
 * mu, normal forces >= f_min; reference values mu = 5, f_min = MIN_NORMAL_FORCE[3] = 1) solved exactly per contact triple; a grasp
 * is feasible when some triple reaches objective <= tol (2e-3 at :77, 1e-4 in the sweep).  obj_out_dev [B] (nullable): smallest
 * objective seen.  The reference solves the same QP with cvxpy/OSQP to solver tolerance: decisions can differ for objectives
 * within that tolerance of tol (parity unpinned). */
int pg_dyn_feasible_qp(const double* pts_dev, const double* nrm_dev, const int32_t* count_dev, int B, int M, double mu, double f_min,
                       double tol, uint8_t* out_dev, double* obj_out_dev, void* stream);
/* find_dynamically_feasible_contacts :239-303 as written (QP criterion, min-distance prefilter), triple range [first,last) */
int pg_dyn_sweep_qp(const double* cloud_pts_dev, const double* cloud_nrm_dev, int P, double min_dist, double mu, double f_min,
                    double tol, uint64_t first, uint64_t last, int32_t* triples_out_dev, uint64_t cap,
                    unsigned long long* count_dev, void* stream);

/* Single-sample gym-style stepping: replaces env.reset() / env.step(a) of envs/<env>_contact_graph_env.py (plate_env:176-250,
 * :253-632) for ONE episode advanced one env step per call.  PgCarry is the state an episode carries between env steps: the
 * rigid-body world (bodies, persistent manifolds, servo constraints) and the decode carry (last_fins / last_surface_norm,
 * plate_env:112-114, :137); it lives on the env's GPU.  All pointers are HOST pointers; the calls are synchronous.
 *   pg_carry_create      a new episode in its post-reset state (runs reset())
 *   pg_carry_reset       reset(); pose7_host (nullable) receives the object pose
 *   pg_env_step_single   step(a): a_host [48] raw action (tanh is applied inside, :487); state = path[step_cnt] (:269-271);
 *                        last_step != 0 runs the train-mode settle substeps and the metric (:556-603);
 *                        reward = 100 x score logit (:611-612); trace_host [50,7] (nullable) = object pose at the start of each
 *                        control substep (what model_test.py:124-135 records) */
typedef struct PgCarry PgCarry;
int pg_carry_create(PgEnv* env, PgCarry** out);
int pg_carry_reset(PgCarry* carry, double* pose7_host);
int pg_carry_destroy(PgCarry* carry);
int pg_env_step_single(PgEnv* env, PgCarry* carry, const double* a_host, int32_t state, double max_force, int32_t last_step,
                       double* reward_host, double* metric_host, double* pose7_host, double* trace_host);

/* diagnostics / tuning: how pg_rollout, pg_rollout_multi and pg_physics run the rigid-body part.  0 / 1 (default): one persistent
 * launch for the whole episode; 2: one launch per episode phase with the rollouts re-sorted by solver work between phases (the
 * measured alternative decomposition; slower on the B200).  Results are bit-identical either way (a rollout's
 * arithmetic does not depend on the thread that runs it). */
int pg_set_rollout_mode(PgEnv* env, int mode);

/* diagnostics: number of kernel launches issued through this library since load */
uint64_t pg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
