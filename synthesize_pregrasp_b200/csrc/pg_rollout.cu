// Kernel 1: batched rigid-body rollouts, one rollout per thread (see pg_physics.cuh / pg_rollout.cuh).
//
// Why one THREAD per rollout rather than one warp: the dominant cost is the projected Gauss-Seidel sweep,
// a strictly sequential chain over ~40-100 rows x up to 50 iterations whose per-row work is two 6-wide dot
// products and two 6-wide axpys.  A warp could put at most 6-12 lanes on one row and would pay shuffle
// latency on every row; a thread per rollout keeps all 32 lanes busy with independent rollouts and the only
// divergence is the per-rollout row count.  HBM sees u + noise in and pose/fins/metric out; everything else
// (bodies, manifolds, solver rows) lives in registers and thread-local memory for the whole rollout.
//
// CTA shape (pg_rollout_kernel.cuh): ONE 448-thread CTA per SM.  Its warps meet on CTA barriers inside every substep so that
// they walk the collision code together (instruction fetch), and before the settle phase the CTA re-deals its rollouts to
// its threads by solver work (settle_regroup, pg_rollout.cuh) so that a warp holds rollouts of one kind: bookshelf physics
// 1425 -> 1268 ms with the regrouping, 1210 ms with the barriers as well (64-thread CTAs, 7 per SM, before).
#include "pg_rollout_kernel.cuh"

namespace pg {

// defined in pg_rollout_hull.cu (compiled without FMA contraction: GJK / polytope decisions are rounding-sensitive)
int launch_rollout_hull(cudaStream_t st, const EnvDev* env, const double* u, const float* noise, int K, int N, const int* path, double max_force,
                        double* pose, double* fins, double* metric, double* trace, int* iters, const int* prob, void* regroup);

int launch_episode_hull(cudaStream_t st, const EnvDev* env, void* state, const double* a, int cstate, double max_force, int settle,
                        double* pose7, double* fins6, double* metric, double* trace);

size_t episode_state_bytes() { return sizeof(EpisodeState); }

// one env step (or the reset, a == nullptr) of the single episode behind `state` (pg_env_step_single)
int launch_episode(PgEnv* env, void* state, const double* a, int cstate, double max_force, int settle, double* pose7, double* fins6,
                   double* metric, double* trace, cudaStream_t st) {
  int rc;
  if (env->host_env.obj_shape != SHAPE_BOX) rc = launch_episode_hull(st, env->dev_env, state, a, cstate, max_force, settle, pose7, fins6, metric, trace);
  else rc = launch_episode_sized<SHAPE_BOX>(st, env->dev_env, state, a, cstate, max_force, settle, pose7, fins6, metric, trace);
  if (rc) return rc;
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

int launch_physics(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                   double* pose, double* fins, double* metric, double* trace, int* iters, cudaStream_t st, const int* prob_dev) {
  // The measured alternative (pg_rollout_phase.cuh): one launch per episode phase with the rollouts re-sorted by solver work in
  // between.  Bit-identical results; on the B200 it is 11-40 % SLOWER than the persistent kernel below (profiles/README.md), so it
  // runs only when asked for (pg_set_rollout_mode(env, 2)).
  const bool phased = env->rollout_mode == 2;
  if (phased && !trace && !iters)
    return launch_physics_phased(env, u, noise, K, N, path_dev, max_force, pose, fins, metric, st, prob_dev);
  // exchange buffer of the settle-phase regrouping (pg_rollout.cuh), one slot per rollout
  const size_t need = (size_t)K * sizeof(RegroupSlot);
  if (env->cap_regroup < need) {
    cudaFree(env->regroup); env->regroup = nullptr; env->cap_regroup = 0;
    PG_CUDA(cudaMalloc(&env->regroup, need));
    env->cap_regroup = need;
  }
  void* regroup = env->regroup;
  int rc;
  if (env->host_env.obj_shape != SHAPE_BOX)
    rc = launch_rollout_hull(st, env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters, prob_dev, regroup);
  else
    rc = launch_rollout_sized<SHAPE_BOX>(st, env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters, prob_dev, regroup);
  if (rc) return rc;
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
