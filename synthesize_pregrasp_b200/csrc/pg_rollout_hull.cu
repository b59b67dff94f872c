// Hull-object instantiation of the rollout kernel (plate, waterbottle).  Compiled with -fmad=false: the reference engine
// rounds every product, and the GJK / expanding-polytope decisions (support ties, near-equal depths, simplex region
// tests) turn a 1-ulp difference into a different contact point; without contraction the kernel agrees with the
// CPU restatement to 1e-12 on 42/48 random plate rollouts (23/48 with contraction).  ~15 % slower, by choice.
// CTA shape of the hull kernel: ONE large CTA per SM whose warps meet on CTA barriers inside every substep (at its start,
// before each narrowphase query, before the solver set-up).  The kernel stalls on instruction fetch (profiles/README.md):
// each substep walks ~15 k rarely-executed instructions of collision code, and warps that reach them at different times
// each pay the misses.  All rollouts run the same number of substeps, so the barriers are free of deadlock and line the
// warps up: 2108 -> 1440 ms on the plate workload with the first barrier, 1110 ms with all three (64-thread CTAs: no effect).
#ifndef PG_BLOCK_THREADS
#define PG_BLOCK_THREADS 448
#endif
#ifndef RK_MINBLOCKS
#define RK_MINBLOCKS 1
#endif
#include "pg_rollout_kernel.cuh"

namespace pg {

int launch_rollout_hull(cudaStream_t st, const EnvDev* env, const double* u, const float* noise, int K, int N, const int* path, double max_force,
                        double* pose, double* fins, double* metric, double* trace, int* iters, const int* prob, void* regroup) {
  return launch_rollout_sized<SHAPE_HULL>(st, env, u, noise, K, N, path, max_force, pose, fins, metric, trace, iters, prob, regroup);
}

int launch_episode_hull(cudaStream_t st, const EnvDev* env, void* state, const double* a, int cstate, double max_force, int settle,
                        double* pose7, double* fins6, double* metric, double* trace) {
  return launch_episode_sized<SHAPE_HULL>(st, env, state, a, cstate, max_force, settle, pose7, fins6, metric, trace);
}

}  // namespace pg
