// TEST-ONLY: compiles the product's rollout logic (csrc/pg_rollout.cuh, the code the CUDA kernel runs per
// thread) for the HOST so the container's CPU tests can compare it with the independent oracle before any
// GPU time is spent.  Not part of the product, not a fallback: the shipped library has no CPU path.
#include "../../synthesize_pregrasp_b200/csrc/pg_envbuild.h"
#include "../../synthesize_pregrasp_b200/csrc/pg_rollout.cuh"
#include "../../synthesize_pregrasp_b200/csrc/pg_qp.cuh"

#ifdef PG_HOST_STATS
namespace pg { long g_stat_pts[32], g_stat_nc[16]; double g_stat_flops[4]; }
extern "C" void hostsim_flops(double* out, int reset) { for (int i = 0; i < 4; i++) { out[i] = pg::g_stat_flops[i]; if (reset) pg::g_stat_flops[i] = 0; } }
extern "C" void hostsim_stats(long* pts, long* nc) { for (int i = 0; i < 32; i++) pts[i] = pg::g_stat_pts[i]; for (int i = 0; i < 16; i++) nc[i] = pg::g_stat_nc[i]; }
#endif
static int* g_settle = nullptr;
extern "C" void hostsim_set_settle_buffer(int* p) { g_settle = p; }
extern "C" int hostsim_rollout(const PgEnvSpec* sp, const double* u, const float* noise, int K, int N, const int* path,
                               double max_force, int train, double* pose, double* fins, double* metric, double* trace,
                               int* iters) {
  pg::EnvDev E;
  std::string err;
  int rc = pg::build_envdev(sp, E, err);
  if (rc) { fprintf(stderr, "hostsim: %s\n", err.c_str()); return rc; }
  for (int k = 0; k < K; k++) {
    pg::RolloutOut out;
    out.pose = pose + (size_t)k * N * 7;
    out.fins = fins + (size_t)k * N * 6;
    out.metric = metric ? metric + k : nullptr;
    out.trace = trace ? trace + (size_t)k * N * 50 * 7 : nullptr;
    out.iters = iters ? iters + (size_t)k * N * 50 : nullptr;
    out.settle_iters = g_settle ? g_settle + (size_t)k * 512 : nullptr;
    out.sync_cta = 0; out.regroup = nullptr; out.k = k; out.k0 = 0;
    double sm[pg::SM_DOUBLES];
    if (E.obj_shape != pg::SHAPE_BOX) pg::rollout_one<pg::SHAPE_HULL>(E, u, noise ? noise + (size_t)k * N * 48 : nullptr, N, path, max_force, train, out, sm, 1);
    else pg::rollout_one<pg::SHAPE_BOX>(E, u, noise ? noise + (size_t)k * N * 48 : nullptr, N, path, max_force, train, out, sm, 1);
  }
  return 0;
}

// wrench-closure QP objective of one contact triple (csrc/pg_qp.cuh), host build
extern "C" double hostsim_qp(const double* p9, const double* n9, double mu, double f_min, int* iters) {
  pg::V3 p[3], n[3];
  for (int i = 0; i < 3; i++) { p[i] = pg::V3(p9[3 * i], p9[3 * i + 1], p9[3 * i + 2]); n[i] = pg::V3(n9[3 * i], n9[3 * i + 1], n9[3 * i + 2]); }
  return pg::qp::wrench_closure_objective(p, n, mu, f_min, iters);
}

extern "C" int hostsim_nnls(const double* A, const double* y, double* x) {
  double M[pg::qp::ROWS][pg::qp::COLS];
  for (int r = 0; r < pg::qp::ROWS; r++) for (int c = 0; c < pg::qp::COLS; c++) M[r][c] = A[r * pg::qp::COLS + c];
  return pg::qp::nnls(M, y, x);
}
extern "C" void hostsim_ls(const double* A, const double* y, const int* idx, int np, double* s) {
  double M[pg::qp::ROWS][pg::qp::COLS];
  for (int r = 0; r < pg::qp::ROWS; r++) for (int c = 0; c < pg::qp::COLS; c++) M[r][c] = A[r * pg::qp::COLS + c];
  pg::qp::ls_passive(M, y, idx, np, s);
}
