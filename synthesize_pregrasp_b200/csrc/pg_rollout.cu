// Kernel 1: batched rigid-body rollouts, one rollout per thread (see pg_physics.cuh / pg_rollout.cuh).
//
// Why one THREAD per rollout rather than one warp: the dominant cost is the projected Gauss-Seidel sweep,
// a strictly sequential chain over ~40-100 rows x up to 50 iterations whose per-row work is two 6-wide dot
// products and two 6-wide axpys.  A warp could put at most 6-12 lanes on one row and would pay shuffle
// latency on every row; a thread per rollout keeps all 32 lanes busy with independent rollouts and the only
// divergence is the per-rollout row count.  HBM sees u + noise in and pose/fins/metric out; everything else
// (bodies, manifolds, solver rows) lives in registers and thread-local memory for the whole rollout.
#define PG_BLOCK_THREADS 64
#include <cstdlib>
#include "pg_internal.h"
#include "pg_rollout.cuh"

namespace pg {

#define RK_THREADS PG_BLOCK_THREADS
#ifndef RK_MINBLOCKS
#define RK_MINBLOCKS 7
#endif
// SHAPE: the object's collision shape; box objects never instantiate the GJK / polytope code (and its ~21 KB of
// per-thread scratch), hull objects (plate) get their own kernel
template <int SHAPE>
__global__ void __launch_bounds__(RK_THREADS, RK_MINBLOCKS) rollout_kernel(const EnvDev* __restrict__ env, const double* __restrict__ u,
                                                      const float* __restrict__ noise, int K, int N,
                                                      const int* __restrict__ path, double max_force,
                                                      double* __restrict__ pose, double* __restrict__ fins,
                                                      double* __restrict__ metric, double* __restrict__ trace,
                                                      int* __restrict__ iters) {
  // grid-stride over the batch: the host may launch fewer CTAs than K/64 to cap how many rollouts are resident
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    RolloutOut out;
    out.pose = pose + (size_t)k * N * 7;
    out.fins = fins + (size_t)k * N * NFING * 3;
    out.metric = metric ? metric + k : nullptr;
    out.trace = trace ? trace + (size_t)k * N * SKIP * 7 : nullptr;
    out.iters = iters ? iters + (size_t)k * N * SKIP : nullptr;
    out.settle_iters = nullptr;
    rollout_one<SHAPE>(*env, u, noise ? noise + (size_t)k * N * ADIM : nullptr, N, path, max_force, 1, out, nullptr, 0);   // solver slice: dynamic shared memory, see PG_SM
  }
}

int launch_physics(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                   double* pose, double* fins, double* metric, double* trace, int* iters, cudaStream_t st) {
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
  if (env->host_env.obj_shape != SHAPE_BOX)
    rollout_kernel<SHAPE_HULL><<<blocks, threads, smem, st>>>(env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters);
  else
    rollout_kernel<SHAPE_BOX><<<blocks, threads, smem, st>>>(env->dev_env, u, noise, K, N, path_dev, max_force, pose, fins, metric, trace, iters);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
