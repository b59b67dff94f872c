// The rollout kernel template (one thread per rollout); instantiated for box objects in pg_rollout.cu and for hull objects
// in pg_rollout_hull.cu -- two translation units because they are compiled with different floating-point contraction.
#pragma once
#ifndef PG_BLOCK_THREADS
#define PG_BLOCK_THREADS 448
#endif
#ifndef PG_NO_SUBSTEP_BARRIER
#define PG_SUBSTEP_BARRIER 1     // CTA barriers inside substep() (pg_physics.cuh): the warps of a CTA walk the collision code together
#endif
// settle phase: the CTA re-deals its rollouts to its threads again every PG_REGROUP_EVERY substeps (pg_rollout.cuh); measured on
// bookshelf K = 65536: 1485-1519 ms without, 1416-1438 ms every 100, 1405-1424 ms every 50 (results bit-identical)
#ifndef PG_REGROUP_EVERY
#define PG_REGROUP_EVERY 50
#endif
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
                                                      int* __restrict__ iters, const int* __restrict__ prob, void* regroup, int lanes) {
  if (SHAPE == SHAPE_HULL) {
    for (int i = threadIdx.x; i < 3 * env->n_hull; i += blockDim.x) pg_hull_sm[i] = env->hull[i];
    __syncthreads();
  }
  // `lanes` rollouts per warp (small batches leave lanes empty so that every SM has a warp); the other lanes leave at once
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, per_cta = (blockDim.x >> 5) * lanes;
  if (lane >= lanes) return;
  // grid-stride over the batch
  for (int k0 = blockIdx.x * per_cta; k0 < K; k0 += gridDim.x * per_cta) {
    const int k = k0 + warp * lanes + lane;
    const int cta_full = (k0 + per_cta <= K);                  // uniform over the CTA: all its remaining threads run a rollout
    if (k >= K) continue;
    RolloutOut out;
    out.pose = pose + (size_t)k * N * 7;
    out.fins = fins + (size_t)k * N * NFING * 3;
    out.metric = metric ? metric + k : nullptr;
    out.trace = trace ? trace + (size_t)k * N * SKIP * 7 : nullptr;
#if defined(PG_PHASE_CLOCKS)
    out.iters = iters;                                         // profiling builds: 6 global cycle counters (tools/gpu_phase_clocks.py)
#else
    out.iters = iters ? iters + (size_t)k * N * SKIP : nullptr;
#endif
    out.settle_iters = nullptr;
    out.sync_cta = cta_full;
    out.regroup = (cta_full && lanes == 32 && blockDim.x > 32) ? regroup : nullptr;
    out.k = k; out.k0 = k0;
    // several independent MPPI problems may share a launch: rollout k then uses the controls and the path of problem prob[k]
    const int q = prob ? prob[k] : 0;
    rollout_one<SHAPE>(*env, u + (size_t)q * N * ADIM, noise ? noise + (size_t)k * N * ADIM : nullptr, N, path + (size_t)q * N, max_force, 1, out, nullptr, 0);   // solver slice: dynamic shared memory, see PG_SM
  }
}

// One CTA per SM, sized to the batch: the shared-memory slices are strided by the compile-time CTA size, so any launch size up
// to it is valid.  A warp takes about as long with 1 rollout as with 32 (the sequential chain of one rollout; measured on the
// plate: 297 vs 381 ms), and more warps per SM cost instruction fetch, so: one warp per SM first (K = 360: 120 warps of 3
// rollouts, 300 ms instead of 381 ms with 12 full warps), then full lanes, then more warps per SM (30 x 360: 338 full warps;
// K = 65 536: 147 CTAs of 448 threads).  Spreading 30 x 360 rollouts over 1800 warps of 6 lanes: 779 ms vs 553 ms.
template <int SHAPE>
static int launch_rollout_sized(cudaStream_t st, const EnvDev* env, const double* u, const float* noise, int K, int N, const int* path,
                                double max_force, double* pose, double* fins, double* metric, double* trace, int* iters, const int* prob,
                                void* regroup) {
  const size_t smem = SM_DOUBLES * RK_THREADS * sizeof(double);
  int dev = 0, n_sm = 0;
  PG_CUDA(cudaGetDevice(&dev));
  // per-device opt-in and SM count, queried on every launch (host-side, cheap; no process-global state)
  PG_CUDA(cudaFuncSetAttribute(rollout_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int max_warps = RK_THREADS / 32;
  int lanes = (K + n_sm - 1) / n_sm;                           // one warp per SM first, then fill its lanes, then add warps
  if (lanes < 1) lanes = 1;
  if (lanes > 32) lanes = 32;
  const int warps = (K + lanes - 1) / lanes;
  int wpc = (warps + n_sm - 1) / n_sm;
  if (wpc > max_warps) wpc = max_warps;
  int blocks = (warps + wpc - 1) / wpc;
  if (blocks > RK_MINBLOCKS * n_sm) blocks = RK_MINBLOCKS * n_sm;
#if defined(PG_RUNTIME_STRIDE)
  const size_t smem_launch = SM_DOUBLES * (size_t)(32 * wpc) * sizeof(double);
#else
  const size_t smem_launch = smem;
#endif
  rollout_kernel<SHAPE><<<blocks, 32 * wpc, smem_launch, st>>>(env, u, noise, K, N, path, max_force, pose, fins, metric, trace, iters, prob, regroup, lanes);
  return PG_OK;
}


// One episode advanced ONE env step per launch (pg_env_step_single: the gym-style env.step of the reference, plate_env:253-632):
// the state between steps lives in an EpisodeState in device memory.  a == nullptr: reset() (plate_env:176-250).
// Same per-thread code as the batch kernel (env_reset / env_step), run by one thread.
template <int SHAPE>
__global__ void __launch_bounds__(32, 1) episode_kernel(const EnvDev* __restrict__ env, EpisodeState* __restrict__ st, const double* __restrict__ a,
                                                        int state, double max_force, int settle, double* __restrict__ pose7,
                                                        double* __restrict__ fins6, double* __restrict__ metric, double* __restrict__ trace) {
  if (SHAPE == SHAPE_HULL) {
    for (int i = threadIdx.x; i < 3 * env->n_hull; i += blockDim.x) pg_hull_sm[i] = env->hull[i];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  RowSet rs;
  M3 statR[MAX_STATICS];
  static_rotations(*env, statR);
  Sim S;
  Carry carry;
#if defined(PG_PHASE_CLOCKS)
  for (int i = 0; i < 6; i++) S.clk[i] = 0;
#endif
  if (!a) {
    env_reset<SHAPE>(*env, S, statR, rs, carry, 0, nullptr, 0);
    st->step = 0; st->prev_state = -1;
  } else {
    S = st->S; carry = st->carry;
    S.sync_cta = 0;
    RolloutOut out;
    out.pose = pose7; out.fins = fins6; out.metric = metric; out.trace = trace; out.iters = nullptr; out.settle_iters = nullptr;
    out.sync_cta = 0; out.regroup = nullptr; out.k = 0; out.k0 = 0;
    double act[ADIM];
    for (int d = 0; d < ADIM; d++) act[d] = a[d];
    env_step<SHAPE>(*env, S, statR, rs, carry, act, state, st->prev_state, max_force, settle != 0, 0, out, nullptr, 0);
    st->step += 1; st->prev_state = state;
  }
  st->S = S; st->carry = carry;
  if (!a && pose7) {
    const DynBody& o = S.dyn[0];
    pose7[0] = o.pos.x; pose7[1] = o.pos.y; pose7[2] = o.pos.z; pose7[3] = o.q.x; pose7[4] = o.q.y; pose7[5] = o.q.z; pose7[6] = o.q.w;
  }
}

template <int SHAPE>
static int launch_episode_sized(cudaStream_t st, const EnvDev* env, void* state, const double* a, int cstate, double max_force, int settle,
                                double* pose7, double* fins6, double* metric, double* trace) {
  const size_t smem = SM_DOUBLES * RK_THREADS * sizeof(double);       // the solver slices are strided by the compile-time CTA size
  PG_CUDA(cudaFuncSetAttribute(episode_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  episode_kernel<SHAPE><<<1, 32, smem, st>>>(env, (EpisodeState*)state, a, cstate, max_force, settle, pose7, fins6, metric, trace);
  return PG_OK;
}

}  // namespace pg
