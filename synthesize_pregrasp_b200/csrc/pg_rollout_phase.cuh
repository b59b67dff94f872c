// Phased rollout kernels: the same per-rollout code as rollout_kernel (env_reset / env_control / substep, pg_rollout.cuh), but a
// launch advances every rollout of the batch by ONE phase of its episode -- reset + first env step, a later env step, or a chunk
// of settle substeps -- and the rollouts' state lives in global memory between launches (RolloutState, ~5 KB each: read and
// written once per phase, i.e. ~10 KB per rollout per 50 substeps -- nothing next to the kernel's time).
//
// Why it exists (profiles/README.md, round 2): with one persistent launch a rollout is bound to a lane for its whole episode.  ncu
// on the bench size shows 13.6 of 32 lanes active (rollouts of one warp need 7...50 solver sweeps over 4...20 contact points) and
// 45 % of all warp stall time in the CTA barriers; a batch of IDENTICAL rollouts runs 3.6x faster than a random one, a batch
// whose warps hold 32 copies of one rollout each 1.65x faster.  Here the rollouts are SORTED between two phases by the solver work
// they just needed (descending counting sort on the device) and the next launch deals them to threads in that order: the lanes of a
// warp run rollouts of one kind, the warps are dealt to the CTAs round-robin so that every SM gets the same mix.
// MEASURED (bookshelf / plate, K = 65536): 1572-1600 / 1743-2350 ms against 1411 / 1277 ms of the persistent kernel with its
// CTA-level re-deals -- the work of the previous 50 substeps predicts the next 50 too loosely to buy more than the re-deal inside
// a CTA already does, and what it buys is lost to register spills at 128-thread CTAs, the tail of eight launches and (hull objects)
// the instruction-cache alignment the CTA barriers of the persistent kernel provide.  Kept as the measured alternative
// (pg_set_rollout_mode(env, 2)); results are bit-identical (GPU test test_phased_sorted_rollouts_equal_persistent_kernel).
//
// Compiled in its own translation units with PG_BLOCK_THREADS = 128: the solver slices in shared memory are strided by that
// compile-time CTA size.
#pragma once
#define PG_SUBSTEP_BARRIER 1     // CTA barriers inside substep(): the (similar) warps of a CTA walk the collision code together
#include "pg_internal.h"
#include "pg_rollout.cuh"

namespace pg {

#ifndef PH_MINBLOCKS
#define PH_MINBLOCKS 4
#endif

struct RolloutState {            // what a rollout carries from phase to phase
  Sim S;
  Carry carry;
  double m0;                     // first half of the settle metric (plate_env:558)
  int work;                      // solver work of the phase just run: sweeps x rows, the sort key
  int pad;
};

enum { PH_FIRST = 0, PH_CONTROL = 1, PH_SETTLE = 2 };

struct PhaseArgs {
  int kind;                      // PH_FIRST: reset() + control part of env step 0; PH_CONTROL: control part of env step j; PH_SETTLE: settle substeps [s0, s1)
  int j, s0, s1;
  int settle_after;              // control phase of the LAST env step of a train rollout: the settle follows (metric half recorded, pose written later)
  int last;                      // settle chunk that ends the episode: metric + pose
  int K, N;
};

template <int SHAPE>
__global__ void __launch_bounds__(PG_BLOCK_THREADS, PH_MINBLOCKS)
phase_kernel(const EnvDev* __restrict__ env, RolloutState* __restrict__ states, const int* __restrict__ order, PhaseArgs A,
             const double* __restrict__ u, const float* __restrict__ noise, const int* __restrict__ path, const int* __restrict__ prob,
             double max_force, double* __restrict__ pose, double* __restrict__ fins, double* __restrict__ metric) {
  if (SHAPE == SHAPE_HULL) {
    for (int i = threadIdx.x; i < 3 * env->n_hull; i += blockDim.x) pg_hull_sm[i] = env->hull[i];
    __syncthreads();
  }
  // Dealing: the sorted list is cut into warps of 32 neighbours (one kind of rollout per warp) and the warps are dealt to the
  // CTAs round-robin -- CTA b runs sorted warps b, b + gridDim, b + 2 gridDim, ... -- so that every CTA (every SM) gets the same
  // mix of heavy and light WARPS.  (Dealt in blocks, the 128 heaviest rollouts of the batch ended up in one CTA and set the
  // makespan of the phase: measured 13 % slower than the persistent kernel.)
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = (wi * (int)gridDim.x + (int)blockIdx.x) * 32 + lane;
  // the warps of a CTA are of different kinds by construction, so there are no CTA barriers inside substep() here
  const int cta_full = 0;
  if (i >= A.K) return;
  const int k = order ? order[i] : i;
  const int N = A.N;
  RolloutOut out;
  out.pose = pose + (size_t)k * N * 7;
  out.fins = fins + (size_t)k * N * NFING * 3;
  out.metric = metric ? metric + k : nullptr;
  out.trace = nullptr; out.iters = nullptr; out.settle_iters = nullptr;
  out.sync_cta = cta_full; out.regroup = nullptr; out.k = k; out.k0 = 0;
  const EnvDev& E = *env;
  Sim S;
  RowSet rs;
  Carry carry;
  M3 statR[MAX_STATICS];
  static_rotations(E, statR);
  RolloutState& st = states[k];
  double m0 = 0;
  if (A.kind == PH_FIRST) env_reset<SHAPE>(E, S, statR, rs, carry, cta_full, nullptr, 0);
  else { S = st.S; carry = st.carry; m0 = st.m0; S.sync_cta = cta_full; }
  int work = 0;
  if (A.kind != PH_SETTLE) {
    const int j = A.j;
    const int q = prob ? prob[k] : 0;                      // several MPPI problems may share a launch (pg_rollout_multi)
    const double* uq = u + (size_t)q * N * ADIM;
    const int* pq = path + (size_t)q * N;
    const float* nz = noise ? noise + (size_t)k * N * ADIM : nullptr;
    double a[ADIM];
    for (int d = 0; d < ADIM; d++) a[d] = uq[j * ADIM + d] + (nz ? (double)nz[j * ADIM + d] : 0.0);
    work = env_control<SHAPE>(E, S, statR, rs, carry, a, pq[j], j == 0 ? -1 : pq[j - 1], max_force, j, out, nullptr, 0);
    if (A.settle_after) m0 = settle_metric_half(E, S);
    else {
      write_pose(S, out.pose + j * 7);
      if (j == N - 1 && out.metric) *out.metric = settle_metric_half(E, S) + settle_metric_half(E, S);    // env without settle substeps
    }
  } else {
    for (int s = A.s0; s < A.s1; s++) {
      substep<SHAPE>(E, S, statR, rs, nullptr, 0);
      work += S.last_iterations * (12 + 2 * rs.n_pts);
    }
    if (A.last) {
      if (out.metric) *out.metric = m0 + settle_metric_half(E, S);
      write_pose(S, out.pose + (N - 1) * 7);
    }
  }
  if (!(A.kind == PH_SETTLE && A.last) && !(A.kind != PH_SETTLE && A.j == N - 1 && !A.settle_after)) {
    st.S = S; st.carry = carry; st.m0 = m0;
  }
  st.work = work;
}

// ---- descending counting sort of the rollouts by the solver work of the phase just run (bucket = work >> shift)
enum { SORT_BUCKETS = 2048 };

static __global__ void work_hist_kernel(const RolloutState* __restrict__ st, int K, int shift, int* __restrict__ hist) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int b = st[k].work >> shift;
  if (b >= SORT_BUCKETS) b = SORT_BUCKETS - 1;
  atomicAdd(hist + b, 1);
}

// offsets[b] = number of rollouts in heavier buckets (single block, SORT_BUCKETS threads... done by 1024 threads x 2)
static __global__ void work_scan_kernel(int* __restrict__ hist) {
  __shared__ int s[SORT_BUCKETS];
  for (int i = threadIdx.x; i < SORT_BUCKETS; i += blockDim.x) s[i] = hist[SORT_BUCKETS - 1 - i];      // reversed: heaviest bucket first
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < SORT_BUCKETS; i++) { const int c = s[i]; s[i] = run; run += c; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SORT_BUCKETS; i += blockDim.x) hist[SORT_BUCKETS - 1 - i] = s[i];
}

static __global__ void work_scatter_kernel(const RolloutState* __restrict__ st, int K, int shift, int* __restrict__ offsets, int* __restrict__ order) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int b = st[k].work >> shift;
  if (b >= SORT_BUCKETS) b = SORT_BUCKETS - 1;
  order[atomicAdd(offsets + b, 1)] = k;
}

static int sort_by_work(cudaStream_t st, const RolloutState* states, int K, int* hist, int* order) {
  // work of a 50-substep phase <= 50 substeps x 3 islands x 50 sweeps x (12 + 48) rows < 2^19: 2048 buckets of 256
  const int shift = 8;
  PG_CUDA(cudaMemsetAsync(hist, 0, SORT_BUCKETS * sizeof(int), st));
  work_hist_kernel<<<(K + 255) / 256, 256, 0, st>>>(states, K, shift, hist); count_launch();
  work_scan_kernel<<<1, 1024, 0, st>>>(hist); count_launch();
  work_scatter_kernel<<<(K + 255) / 256, 256, 0, st>>>(states, K, shift, hist, order); count_launch();
  return PG_OK;
}

// the whole batch, phase by phase; `train` rollouts settle after the last env step
template <int SHAPE>
static int launch_rollout_phased(cudaStream_t st, const EnvDev* env, int settle_substeps, void* states_v, int* hist, int* order,
                                 const double* u, const float* noise, int K, int N, const int* path, double max_force,
                                 double* pose, double* fins, double* metric, const int* prob) {
  RolloutState* states = (RolloutState*)states_v;
  const size_t smem = SM_DOUBLES * PG_BLOCK_THREADS * sizeof(double);
  PG_CUDA(cudaFuncSetAttribute(phase_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (K + PG_BLOCK_THREADS - 1) / PG_BLOCK_THREADS;      // a CTA's warps take sorted warps b, b + blocks, ...: covers [0, 4 * 32 * blocks)
  const int chunk = 50;
  for (int j = 0; j < N; j++) {
    PhaseArgs A;
    A.kind = j == 0 ? PH_FIRST : PH_CONTROL; A.j = j; A.s0 = A.s1 = 0; A.last = 0; A.K = K; A.N = N;
    A.settle_after = (j == N - 1 && settle_substeps > 0) ? 1 : 0;
    if (j > 0) { int rc = sort_by_work(st, states, K, hist, order); if (rc) return rc; }
    phase_kernel<SHAPE><<<blocks, PG_BLOCK_THREADS, smem, st>>>(env, states, j > 0 ? order : nullptr, A, u, noise, path, prob, max_force, pose, fins, metric);
    count_launch();
  }
  for (int s0 = 0; s0 < settle_substeps; s0 += chunk) {
    PhaseArgs A;
    A.kind = PH_SETTLE; A.j = N - 1; A.s0 = s0; A.s1 = s0 + chunk < settle_substeps ? s0 + chunk : settle_substeps;
    A.last = A.s1 == settle_substeps ? 1 : 0; A.settle_after = 0; A.K = K; A.N = N;
    int rc = sort_by_work(st, states, K, hist, order); if (rc) return rc;
    phase_kernel<SHAPE><<<blocks, PG_BLOCK_THREADS, smem, st>>>(env, states, order, A, u, noise, path, prob, max_force, pose, fins, metric);
    count_launch();
  }
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

static inline size_t rollout_state_bytes() { return sizeof(RolloutState); }

}  // namespace pg
