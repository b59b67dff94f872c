// Phased rollouts, box objects (see pg_rollout_phase.cuh).  Own translation unit: PG_BLOCK_THREADS = 128.
#define PG_BLOCK_THREADS 128
#include "pg_rollout_phase.cuh"

namespace pg {

int launch_rollout_phased_hull(cudaStream_t st, const EnvDev* env, int settle_substeps, void* states, int* hist, int* order, const double* u,
                               const float* noise, int K, int N, const int* path, double max_force, double* pose, double* fins,
                               double* metric, const int* prob);

// K rollouts phase by phase with the rollouts re-sorted by solver work between phases; scratch grown on demand
int launch_physics_phased(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                          double* pose, double* fins, double* metric, cudaStream_t st, const int* prob_dev) {
  const size_t need = (size_t)K * rollout_state_bytes();
  if (env->cap_ph < need) {
    PG_CUDA(cudaStreamSynchronize(st));
    cudaFree(env->ph_states); cudaFree(env->ph_order); env->ph_states = nullptr; env->ph_order = nullptr; env->cap_ph = 0;
    PG_CUDA(cudaMalloc(&env->ph_states, need));
    PG_CUDA(cudaMalloc(&env->ph_order, (size_t)K * sizeof(int)));
    env->cap_ph = need;
  }
  if (!env->ph_hist) PG_CUDA(cudaMalloc(&env->ph_hist, SORT_BUCKETS * sizeof(int)));
  const int settle = env->host_env.settle_substeps;
  if (env->host_env.obj_shape != SHAPE_BOX)
    return launch_rollout_phased_hull(st, env->dev_env, settle, env->ph_states, env->ph_hist, env->ph_order, u, noise, K, N, path_dev, max_force,
                                      pose, fins, metric, prob_dev);
  return launch_rollout_phased<SHAPE_BOX>(st, env->dev_env, settle, env->ph_states, env->ph_hist, env->ph_order, u, noise, K, N, path_dev, max_force,
                                          pose, fins, metric, prob_dev);
}

}  // namespace pg
