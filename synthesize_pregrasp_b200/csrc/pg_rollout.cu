// Kernel 1: batched rigid-body rollouts, one rollout per thread (see pg_physics.cuh / pg_rollout.cuh).
//
// Why one THREAD per rollout rather than one warp: the dominant cost is the projected Gauss-Seidel sweep,
// a strictly sequential chain over ~40-100 rows x up to 50 iterations whose per-row work is two 6-wide dot
// products and two 6-wide axpys.  A warp could put at most 6-12 lanes on one row and would pay shuffle
// latency on every row; a thread per rollout keeps all 32 lanes busy with independent rollouts and the only
// divergence is the per-rollout row count.  HBM sees u + noise in and pose/fins/metric out; everything else
// (bodies, manifolds, solver rows) lives in registers and thread-local memory for the whole rollout.
#include "pg_rollout_kernel.cuh"

namespace pg {

// defined in pg_rollout_hull.cu (compiled without FMA contraction: GJK / polytope decisions are rounding-sensitive)
void launch_rollout_hull(int blocks, int threads, size_t smem, cudaStream_t st, const EnvDev* env, const double* u, const float* noise,
                         int K, int N, const int* path, double max_force, double* pose, double* fins, double* metric, double* trace, int* iters,
                         const int* prob);

int launch_physics(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                   double* pose, double* fins, double* metric, double* trace, int* iters, cudaStream_t st, const int* prob_dev) {
  const int threads = RK_THREADS;
  int blocks = (K + threads - 1) / threads;
  // Resident CTAs per SM: the solver rows of all resident rollouts (2-3 KB each) are re-read every Gauss-Seidel
  // iteration; the grid can be capped below K/64 (grid-stride kernel) to keep that working set nearer the 126 MB L2.
  // PG_ROLLOUT_CTAS overrides (tuning only).
  static int ctas_per_sm = -1, n_sm = 0;
  if (ctas_per_sm < 0) {
    const char* e = getenv("PG_ROLLOUT_CTAS"); ctas_per_sm = e ? atoi(e) : RK_MINBLOCKS;
    if (ctas_per_sm < 1 || ctas_per_sm > RK_MINBLOCKS) ctas_per_sm = RK_MINBLOCKS;
    int dev = 0; PG_CUDA(cudaGetDevice(&dev)); PG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t smem = SM_DOUBLES * RK_THREADS * sizeof(double);
  if (blocks > ctas_per_sm * n_sm) blocks = ctas_per_sm * n_sm;
  static bool optin = false;                                   // CTA shapes whose slice exceeds the 48 KB default (tuning builds)
  if (smem > 48 * 1024 && !optin) { PG_CUDA(cudaFuncSetAttribute(rollout_kernel<SHAPE_BOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); optin = true; }
  if (env->host_env.obj_shape != SHAPE_BOX)
    launch_rollout_hull(blocks, threads, smem, st, env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters, prob_dev);
  else
    rollout_kernel<SHAPE_BOX><<<blocks, threads, smem, st>>>(env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters, prob_dev);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
