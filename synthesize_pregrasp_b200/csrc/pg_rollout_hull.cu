// Hull-object instantiation of the rollout kernel (plate, waterbottle).  Compiled with -fmad=false: the reference engine
// rounds every product, and the GJK / expanding-polytope decisions (support ties, near-equal depths, simplex region
// tests) turn a 1-ulp difference into a different contact point; without contraction the kernel agrees with the
// CPU restatement to 1e-12 on 42/48 random plate rollouts (23/48 with contraction).  ~15 % slower, by choice.
// CTA shape of the hull kernel: ONE large CTA per SM whose warps meet on CTA barriers inside every substep (at its start,
// before each narrowphase query, before the solver set-up).  The kernel stalls on instruction fetch (profiles/README.md):
// each substep walks ~15 k rarely-executed instructions of collision code, and warps that reach them at different times
// each pay the misses.  All rollouts run the same number of substeps, so the barriers are free of deadlock and line the
// warps up: 2108 -> 1440 ms on the plate workload with the first barrier, 1110 ms with all three (64-thread CTAs: no effect).
#define PG_SUBSTEP_BARRIER 1
#ifndef PG_BLOCK_THREADS
#define PG_BLOCK_THREADS 448
#endif
#ifndef RK_MINBLOCKS
#define RK_MINBLOCKS 1
#endif
#include "pg_rollout_kernel.cuh"

namespace pg {

void launch_rollout_hull(int blocks_hint, int threads_hint, size_t smem_hint, cudaStream_t st, const EnvDev* env, const double* u, const float* noise,
                         int K, int N, const int* path, double max_force, double* pose, double* fins, double* metric, double* trace, int* iters,
                         const int* prob) {
  (void)threads_hint; (void)smem_hint;
  // the shared-memory slice is strided by the compile-time CTA size, so any launch size up to it is valid; small batches get
  // smaller CTAs so that every SM has work (K = 360: 12 CTAs of 32; 30 x 360: 113 CTAs of 96; K = 65 536: 147 CTAs of 448)
  const size_t smem = SM_DOUBLES * RK_THREADS * sizeof(double);
  static int sm_of[64] = {0};                                  // per device: SM count, and the one-time shared-memory opt-in
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sm_of[dev]) {
    cudaDeviceGetAttribute(&sm_of[dev], cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(rollout_kernel<SHAPE_HULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const int n_sm = sm_of[dev];
  int threads = ((K + n_sm - 1) / n_sm + 31) / 32 * 32;
  if (threads < 32) threads = 32;
  if (threads > RK_THREADS) threads = RK_THREADS;
  int blocks = (K + threads - 1) / threads;
  if (blocks > RK_MINBLOCKS * n_sm) blocks = RK_MINBLOCKS * n_sm;
  if (blocks_hint < (K + 63) / 64) {                            // caller-side residency cap (PG_ROLLOUT_CTAS tuning knob)
    int capped = (blocks_hint * 64 + threads - 1) / threads;
    if (capped < blocks) blocks = capped < 1 ? 1 : capped;
  }
  rollout_kernel<SHAPE_HULL><<<blocks, threads, smem, st>>>(env, u, noise, K, N, path, max_force, pose, fins, metric, trace, iters, prob);
}

}  // namespace pg
