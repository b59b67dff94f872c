// Hull-object instantiation of the rollout kernel (plate, waterbottle).  Compiled with -fmad=false: the reference engine
// rounds every product, and the GJK / expanding-polytope decisions (support ties, near-equal depths, simplex region
// tests) turn a 1-ulp difference into a different contact point; without contraction the kernel agrees with the
// CPU restatement to 1e-12 on 42/48 random plate rollouts (23/48 with contraction).  ~15 % slower, by choice.
#include "pg_rollout_kernel.cuh"

namespace pg {

void launch_rollout_hull(int blocks, int threads, size_t smem, cudaStream_t st, const EnvDev* env, const double* u, const float* noise,
                         int K, int N, const int* path, double max_force, double* pose, double* fins, double* metric, double* trace, int* iters,
                         const int* prob) {
  rollout_kernel<SHAPE_HULL><<<blocks, threads, smem, st>>>(env, u, noise, K, N, path, max_force, pose, fins, metric, trace, iters, prob);
}

}  // namespace pg
