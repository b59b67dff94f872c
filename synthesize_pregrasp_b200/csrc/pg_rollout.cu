// Kernel 1: batched rigid-body rollouts, one rollout per thread (see pg_physics.cuh / pg_rollout.cuh).
//
// Why one THREAD per rollout rather than one warp: the dominant cost is the projected Gauss-Seidel sweep,
// a strictly sequential chain over ~40-100 rows x up to 50 iterations whose per-row work is two 6-wide dot
// products and two 6-wide axpys.  A warp could put at most 6-12 lanes on one row and would pay shuffle
// latency on every row; a thread per rollout keeps all 32 lanes busy with independent rollouts and the only
// divergence is the per-rollout row count.  HBM sees u + noise in and pose/fins/metric out; everything else
// (bodies, manifolds, solver rows) lives in registers and thread-local memory for the whole rollout.
#define PG_BLOCK_THREADS 64
#include "pg_internal.h"
#include "pg_rollout.cuh"

namespace pg {

#define RK_THREADS PG_BLOCK_THREADS
// SHAPE: the object's collision shape; box objects never instantiate the GJK / polytope code (and its ~21 KB of
// per-thread scratch), hull objects (plate) get their own kernel
template <int SHAPE>
__global__ void __launch_bounds__(RK_THREADS, 7) rollout_kernel(const EnvDev* __restrict__ env, const double* __restrict__ u,
                                                      const float* __restrict__ noise, int K, int N,
                                                      const int* __restrict__ path, double max_force,
                                                      double* __restrict__ pose, double* __restrict__ fins,
                                                      double* __restrict__ metric, double* __restrict__ trace,
                                                      int* __restrict__ iters) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  RolloutOut out;
  out.pose = pose + (size_t)k * N * 7;
  out.fins = fins + (size_t)k * N * NFING * 3;
  out.metric = metric ? metric + k : nullptr;
  out.trace = trace ? trace + (size_t)k * N * SKIP * 7 : nullptr;
  out.iters = iters ? iters + (size_t)k * N * SKIP : nullptr;
  out.settle_iters = nullptr;
  rollout_one<SHAPE>(*env, u, noise ? noise + (size_t)k * N * ADIM : nullptr, N, path, max_force, 1, out, nullptr, 0);   // solver slice: dynamic shared memory, see PG_SM
}

int launch_physics(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                   double* pose, double* fins, double* metric, double* trace, int* iters, cudaStream_t st) {
  const int threads = RK_THREADS;
  int blocks = (K + threads - 1) / threads;
  const size_t smem = SM_DOUBLES * RK_THREADS * sizeof(double);
  if (env->host_env.obj_shape != SHAPE_BOX)
    rollout_kernel<SHAPE_HULL><<<blocks, threads, smem, st>>>(env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters);
  else
    rollout_kernel<SHAPE_BOX><<<blocks, threads, smem, st>>>(env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
