// Phased rollouts, hull objects (plate, waterbottle); compiled with -fmad=false like pg_rollout_hull.cu (GJK / polytope decisions
// are rounding-sensitive).  Own translation unit: PG_BLOCK_THREADS = 128.
#define PG_BLOCK_THREADS 128
#ifndef PH_MINBLOCKS
#define PH_MINBLOCKS 3
#endif
#include "pg_rollout_phase.cuh"

namespace pg {

int launch_rollout_phased_hull(cudaStream_t st, const EnvDev* env, int settle_substeps, void* states, int* hist, int* order, const double* u,
                               const float* noise, int K, int N, const int* path, double max_force, double* pose, double* fins,
                               double* metric, const int* prob) {
  return launch_rollout_phased<SHAPE_HULL>(st, env, settle_substeps, states, hist, order, u, noise, K, N, path, max_force, pose, fins, metric, prob);
}

}  // namespace pg
