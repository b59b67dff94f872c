// The rollout kernel template (one thread per rollout); instantiated for box objects in pg_rollout.cu and for hull objects
// in pg_rollout_hull.cu -- two translation units because they are compiled with different floating-point contraction.
#pragma once
#ifndef PG_BLOCK_THREADS
#define PG_BLOCK_THREADS 448
#endif
#define PG_SUBSTEP_BARRIER 1     // CTA barriers inside substep() (pg_physics.cuh): the warps of a CTA walk the collision code together
#include <cstdlib>
#include "pg_internal.h"
#include "pg_rollout.cuh"

namespace pg {

#define RK_THREADS PG_BLOCK_THREADS
#ifndef RK_MINBLOCKS
#define RK_MINBLOCKS 1
#endif
// SHAPE: the object's collision shape; box objects never instantiate the GJK / polytope code (and its ~21 KB of
// per-thread scratch), hull objects (plate) get their own kernel
template <int SHAPE>
__global__ void __launch_bounds__(RK_THREADS, RK_MINBLOCKS) rollout_kernel(const EnvDev* __restrict__ env, const double* __restrict__ u,
                                                      const float* __restrict__ noise, int K, int N,
                                                      const int* __restrict__ path, double max_force,
                                                      double* __restrict__ pose, double* __restrict__ fins,
                                                      double* __restrict__ metric, double* __restrict__ trace,
                                                      int* __restrict__ iters, const int* __restrict__ prob, void* regroup) {
  if (SHAPE == SHAPE_HULL) {
    for (int i = threadIdx.x; i < 3 * env->n_hull; i += blockDim.x) pg_hull_sm[i] = env->hull[i];
    __syncthreads();
  }
  // grid-stride over the batch: the host may launch fewer CTAs than K/64 to cap how many rollouts are resident
  for (int k0 = blockIdx.x * blockDim.x; k0 < K; k0 += gridDim.x * blockDim.x) {
    const int k = k0 + threadIdx.x;
    const int cta_full = (k0 + (int)blockDim.x <= K);          // uniform over the CTA
    if (k >= K) continue;
    RolloutOut out;
    out.pose = pose + (size_t)k * N * 7;
    out.fins = fins + (size_t)k * N * NFING * 3;
    out.metric = metric ? metric + k : nullptr;
    out.trace = trace ? trace + (size_t)k * N * SKIP * 7 : nullptr;
    out.iters = iters ? iters + (size_t)k * N * SKIP : nullptr;
    out.settle_iters = nullptr;
    out.sync_cta = cta_full;
    out.regroup = (cta_full && blockDim.x > 32) ? regroup : nullptr;
    out.k = k;
    // several independent MPPI problems may share a launch: rollout k then uses the controls and the path of problem prob[k]
    const int q = prob ? prob[k] : 0;
    rollout_one<SHAPE>(*env, u + (size_t)q * N * ADIM, noise ? noise + (size_t)k * N * ADIM : nullptr, N, path + (size_t)q * N, max_force, 1, out, nullptr, 0);   // solver slice: dynamic shared memory, see PG_SM
  }
}

// One CTA per SM, sized to the batch: the shared-memory slices are strided by the compile-time CTA size, so any launch size up
// to it is valid; small batches get smaller CTAs so that every SM has work (K = 360: 12 CTAs of 32; 30 x 360: 113 CTAs of 96;
// K = 65 536: 147 CTAs of 448).
template <int SHAPE>
static int launch_rollout_sized(cudaStream_t st, const EnvDev* env, const double* u, const float* noise, int K, int N, const int* path,
                                double max_force, double* pose, double* fins, double* metric, double* trace, int* iters, const int* prob,
                                void* regroup) {
  const size_t smem = SM_DOUBLES * RK_THREADS * sizeof(double);
  static int sm_of[64] = {0};                                  // per device: SM count, and the one-time shared-memory opt-in
  int dev = 0; PG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sm_of[dev]) {
    PG_CUDA(cudaFuncSetAttribute(rollout_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(cudaDeviceGetAttribute(&sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_sm = sm_of[dev];
  int threads = ((K + n_sm - 1) / n_sm + 31) / 32 * 32;
  if (threads < 32) threads = 32;
  if (threads > RK_THREADS) threads = RK_THREADS;
  int blocks = (K + threads - 1) / threads;
  if (blocks > RK_MINBLOCKS * n_sm) blocks = RK_MINBLOCKS * n_sm;
  rollout_kernel<SHAPE><<<blocks, threads, smem, st>>>(env, u, noise, K, N, path, max_force, pose, fins, metric, trace, iters, prob, regroup);
  return PG_OK;
}

}  // namespace pg
