// C ABI (include/pregrasp_b200.h): handle management, constant upload, host-buffer wrappers.
#include <math.h>
#include <string.h>
#include <stddef.h>
#include <stdlib.h>
#include <atomic>
#include <vector>
#include <algorithm>
#include "pg_internal.h"
#include "pg_envbuild.h"

namespace pg {
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return PG_ERR_CUDA;
}
void count_launch() { g_launches++; }
}  // namespace pg

using namespace pg;

extern "C" const char* pg_last_error(void) { return g_err.c_str(); }
extern "C" const char* pg_version(void) { return "pregrasp_b200 0.1 (sm_100a)"; }
extern "C" uint64_t pg_launch_count(void) { return g_launches.load(); }
extern "C" int pg_set_rollout_mode(PgEnv* e, int mode) {
  if (!e || mode < 0 || mode > 2) { set_error("pg_set_rollout_mode: bad arguments"); return PG_ERR_INVALID; }
  e->rollout_mode = mode;
  return PG_OK;
}


static void pk_quat_to_matrix(const double* q_xyzw, double* R) {
  double x = q_xyzw[0], y = q_xyzw[1], z = q_xyzw[2], w = q_xyzw[3];
  double two_s = 2.0 / (w * w + x * x + y * y + z * z);
  R[0] = 1 - two_s * (y * y + z * z); R[1] = two_s * (x * y - z * w); R[2] = two_s * (x * z + y * w);
  R[3] = two_s * (x * y + z * w); R[4] = 1 - two_s * (x * x + z * z); R[5] = two_s * (y * z - x * w);
  R[6] = two_s * (x * z - y * w); R[7] = two_s * (y * z + x * w); R[8] = 1 - two_s * (x * x + y * y);
}

struct Blob {
  std::vector<unsigned char> h;
  size_t add(const void* p, size_t bytes) {
    size_t off = (h.size() + 255) & ~size_t(255);
    h.resize(off + bytes);
    memcpy(h.data() + off, p, bytes);
    return off;
  }
};

extern "C" int pg_env_create(const PgEnvSpec* sp, int device, PgEnv** out) {
  if (!sp || !out) { set_error("pg_env_create: null argument"); return PG_ERR_INVALID; }
  if (sp->n_static < 1 || sp->n_static > MAX_STATICS || sp->n_regions > MAX_REGIONS || sp->n_states > MAX_STATES ||
      sp->n_cloud != P_CLOUD || sp->n_crop < 1 || sp->n_crop > 2 || sp->n_dist_box < 0 || sp->n_dist_box > 4 || sp->n_dist_tri < 0 || sp->n_dist_prism < 0 || sp->n_dist_prism > 2 ||
      (sp->n_dist_prism > 0 && (sp->dist_prism[0][5] < 3 || sp->dist_prism[0][5] > 32)) || (sp->n_dist_prism > 1 && (sp->dist_prism[1][5] < 3 || sp->dist_prism[1][5] > 32)) ||
      (sp->n_dist_tri > 0 && !sp->dist_tri) || !sp->cloud || !sp->csg) {
    set_error("pg_env_create: spec out of supported bounds"); return PG_ERR_INVALID;
  }
  int ndev = 0;
  PG_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("pg_env_create: no such CUDA device"); return PG_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(device));
  PgEnv* e = new PgEnv();
  memset(e, 0, sizeof(PgEnv));
  e->device = device;
  EnvDev& E = e->host_env;
  { std::string err; int rc = build_envdev(sp, E, err); if (rc) { set_error("pg_env_create: " + err); delete e; return rc; } }
  // ---- cost constants
  const PgScoreWeights& W = sp->weights;
  if (!W.conv1_w || !W.out_b) { set_error("pg_env_create: score weights missing"); delete e; return PG_ERR_INVALID; }
  std::vector<float> a1((size_t)P_CLOUD * C1), w1d(C1), w2t(C1 * C2), w3at(C2 * C3), w3bt(C2 * C3);
  for (int p = 0; p < P_CLOUD; p++) {
    float x = (float)sp->cloud[3 * p], y = (float)sp->cloud[3 * p + 1], z = (float)sp->cloud[3 * p + 2];
    for (int j = 0; j < C1; j++) {
      float v = W.conv1_b[j];
      v = fmaf(W.conv1_w[4 * j], x, v); v = fmaf(W.conv1_w[4 * j + 1], y, v); v = fmaf(W.conv1_w[4 * j + 2], z, v);
      a1[(size_t)p * C1 + j] = v;
    }
  }
  for (int j = 0; j < C1; j++) w1d[j] = W.conv1_w[4 * j + 3];
  for (int c = 0; c < C2; c++) for (int j = 0; j < C1; j++) w2t[j * C2 + c] = W.conv2_w[c * C1 + j];
  for (int o = 0; o < C3; o++) for (int k = 0; k < C2; k++) { w3at[k * C3 + o] = W.conv3_w[o * 2 * C2 + k]; w3bt[k * C3 + o] = W.conv3_w[o * 2 * C2 + C2 + k]; }
  // tensor-core operands (pg_cost_tc.cu): tf32 hi + lo splits in the canonical K-major layout [(k/4)][n][k%4].
  // conv2 has no activation (network.py:402-404), so conv3's point half is folded:  W3a (W2 h1 + b2) = (W3a W2) h1 + W3a b2
  auto tf32_rna = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
  const int N4 = 32;
  std::vector<float> w32hi(C1 * C3), w32lo(C1 * C3), w2hi(C1 * C2), w2lo(C1 * C2), w4hi(C3 * N4, 0.f), w4lo(C3 * N4, 0.f), cbase(C3);
  for (int o = 0; o < C3; o++) {
    double bb = 0;
    for (int c = 0; c < C2; c++) bb += (double)W.conv3_w[o * 2 * C2 + c] * (double)W.conv2_b[c];
    cbase[o] = (float)((double)W.conv3_b[o] + bb);
    for (int j = 0; j < C1; j++) {
      double acc = 0;
      for (int c = 0; c < C2; c++) acc += (double)W.conv3_w[o * 2 * C2 + c] * (double)W.conv2_w[c * C1 + j];
      const float wv = (float)acc, hi = tf32_rna(wv);
      const size_t idx = ((size_t)(j / 4) * C3 + o) * 4 + (j % 4);
      w32hi[idx] = hi; w32lo[idx] = wv - hi;
    }
  }
  for (int c = 0; c < C2; c++) for (int j = 0; j < C1; j++) {
    const float wv = W.conv2_w[c * C1 + j], hi = tf32_rna(wv);
    const size_t idx = ((size_t)(j / 4) * C2 + c) * 4 + (j % 4);
    w2hi[idx] = hi; w2lo[idx] = wv - hi;
  }
  for (int q = 0; q < C4; q++) for (int o = 0; o < C3; o++) {
    const float wv = W.conv4_w[q * C3 + o], hi = tf32_rna(wv);
    const size_t idx = ((size_t)(o / 4) * N4 + q) * 4 + (o % 4);
    w4hi[idx] = hi; w4lo[idx] = wv - hi;
  }
  e->use_tensor_cores = 1;
  Blob b;
  size_t o_env = b.add(&E, sizeof(EnvDev));
  size_t o_cloud = b.add(sp->cloud, sizeof(double) * 3 * P_CLOUD);
  size_t o_a1 = b.add(a1.data(), a1.size() * 4), o_w1d = b.add(w1d.data(), C1 * 4), o_w2t = b.add(w2t.data(), w2t.size() * 4);
  size_t o_b2 = b.add(W.conv2_b, C2 * 4), o_w3at = b.add(w3at.data(), w3at.size() * 4), o_w3bt = b.add(w3bt.data(), w3bt.size() * 4);
  size_t o_b3 = b.add(W.conv3_b, C3 * 4), o_w4 = b.add(W.conv4_w, C4 * C3 * 4), o_b4 = b.add(W.conv4_b, C4 * 4);
  size_t o_f1w = b.add(W.fc1_w, F1 * LAT * 4), o_f1b = b.add(W.fc1_b, F1 * 4), o_f2w = b.add(W.fc2_w, F2 * F1 * 4), o_f2b = b.add(W.fc2_b, F2 * 4);
  size_t o_f3w = b.add(W.fc3_w, F3 * F2 * 4), o_f3b = b.add(W.fc3_b, F3 * 4), o_ow = b.add(W.out_w, F3 * 4), o_ob = b.add(W.out_b, 4);
  size_t o_w1 = b.add(W.conv1_w, C1 * 4 * 4), o_b1 = b.add(W.conv1_b, C1 * 4);
  size_t o_w32hi = b.add(w32hi.data(), w32hi.size() * 4), o_w32lo = b.add(w32lo.data(), w32lo.size() * 4);
  size_t o_w4hi = b.add(w4hi.data(), w4hi.size() * 4), o_w4lo = b.add(w4lo.data(), w4lo.size() * 4);
  size_t o_w2hi = b.add(w2hi.data(), w2hi.size() * 4), o_w2lo = b.add(w2lo.data(), w2lo.size() * 4), o_cbase = b.add(cbase.data(), cbase.size() * 4);
  size_t o_tri = sp->n_dist_tri > 0 ? b.add(sp->dist_tri, sizeof(double) * 9 * sp->n_dist_tri) : 0;
  std::vector<double> tri_sph((size_t)4 * (sp->n_dist_tri > 0 ? sp->n_dist_tri : 0));
  for (int t = 0; t < sp->n_dist_tri; t++) {
    const double* v = sp->dist_tri + 9 * t;
    double c[3], r2 = 0;
    for (int k = 0; k < 3; k++) c[k] = (v[k] + v[3 + k] + v[6 + k]) / 3.0;
    for (int q = 0; q < 3; q++) { double dd = 0; for (int k = 0; k < 3; k++) dd += (v[3 * q + k] - c[k]) * (v[3 * q + k] - c[k]); if (dd > r2) r2 = dd; }
    tri_sph[4 * t] = c[0]; tri_sph[4 * t + 1] = c[1]; tri_sph[4 * t + 2] = c[2]; tri_sph[4 * t + 3] = sqrt(r2) * (1.0 + 1e-12);
  }
  size_t o_sph = sp->n_dist_tri > 0 ? b.add(tri_sph.data(), sizeof(double) * tri_sph.size()) : 0;
  const int TG = 16, n_grp = (sp->n_dist_tri + TG - 1) / TG;
  std::vector<double> grp_sph((size_t)4 * n_grp);
  for (int g = 0; g < n_grp; g++) {
    int t0 = g * TG, t1 = std::min(sp->n_dist_tri, t0 + TG);
    double c[3] = {0, 0, 0}, r = 0;
    for (int t = t0; t < t1; t++) for (int k = 0; k < 3; k++) c[k] += tri_sph[4 * t + k] / (t1 - t0);
    for (int t = t0; t < t1; t++) for (int q = 0; q < 3; q++) {
      double dd = 0;
      for (int k = 0; k < 3; k++) dd += (sp->dist_tri[9 * t + 3 * q + k] - c[k]) * (sp->dist_tri[9 * t + 3 * q + k] - c[k]);
      if (sqrt(dd) > r) r = sqrt(dd);
    }
    grp_sph[4 * g] = c[0]; grp_sph[4 * g + 1] = c[1]; grp_sph[4 * g + 2] = c[2]; grp_sph[4 * g + 3] = r * (1.0 + 1e-12);
  }
  size_t o_gsph = n_grp > 0 ? b.add(grp_sph.data(), sizeof(double) * grp_sph.size()) : 0;
  unsigned char* d = nullptr;
  cudaError_t ce = cudaMalloc(&d, b.h.size());
  if (ce != cudaSuccess) { delete e; return cuda_fail(ce, "cudaMalloc(env constants)"); }
  ce = cudaMemcpy(d, b.h.data(), b.h.size(), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { cudaFree(d); delete e; return cuda_fail(ce, "cudaMemcpy(env constants)"); }
  e->blob = d;
  e->dev_env = reinterpret_cast<EnvDev*>(d + o_env);
  CostDev& c = e->cost;
  c.cloud = reinterpret_cast<const double*>(d + o_cloud);
#define FP(off) reinterpret_cast<const float*>(d + (off))
  c.a1 = FP(o_a1); c.w1d = FP(o_w1d); c.w2t = FP(o_w2t); c.b2 = FP(o_b2); c.w3at = FP(o_w3at); c.w3bt = FP(o_w3bt);
  c.b3 = FP(o_b3); c.w4 = FP(o_w4); c.b4 = FP(o_b4); c.fc1_w = FP(o_f1w); c.fc1_b = FP(o_f1b); c.fc2_w = FP(o_f2w);
  c.fc2_b = FP(o_f2b); c.fc3_w = FP(o_f3w); c.fc3_b = FP(o_f3b); c.out_w = FP(o_ow); c.out_b = FP(o_ob);
  c.w1 = FP(o_w1); c.b1 = FP(o_b1);
  e->tcw.w32_hi = FP(o_w32hi); e->tcw.w32_lo = FP(o_w32lo); e->tcw.w4_hi = FP(o_w4hi); e->tcw.w4_lo = FP(o_w4lo);
  e->tcw.w2_hi = FP(o_w2hi); e->tcw.w2_lo = FP(o_w2lo); e->tcw.cbase = FP(o_cbase);
#undef FP
  c.n_cloud = P_CLOUD; c.n_crop = sp->n_crop; c.n_dist_box = sp->n_dist_box; c.n_dist_tri = sp->n_dist_tri;
  memcpy(c.crop, sp->crop, sizeof(c.crop));
  memcpy(c.dist_box, sp->dist_box, sizeof(c.dist_box));
  c.dist_tri = sp->n_dist_tri > 0 ? reinterpret_cast<const double*>(d + o_tri) : nullptr;
  c.tri_sphere = sp->n_dist_tri > 0 ? reinterpret_cast<const double*>(d + o_sph) : nullptr;
  c.tri_group_sphere = n_grp > 0 ? reinterpret_cast<const double*>(d + o_gsph) : nullptr;
  c.n_dist_prism = sp->n_dist_prism;
  for (int i = 0; i < sp->n_dist_prism; i++) {
    const double* pr = sp->dist_prism[i];
    const int n = (int)pr[5];
    c.prism_sides[i] = n; c.prism_hh[i] = pr[4];
    for (int k = 0; k < 3; k++) c.prism_c[i][k] = pr[k];
    const double step = 2.0 * 3.14159265358979323846 / n;
    for (int j = 0; j <= n; j++) { c.prism_v[i][j][0] = cos(step * (j % n)) * pr[3]; c.prism_v[i][j][1] = sin(step * (j % n)) * pr[3]; }
  }
  c.clearance_h = sp->clearance_h; c.project_uses_clearance = sp->project_uses_clearance;
  c.clear_n = sp->n_clear_walls;
  for (int w = 0; w < sp->n_clear_walls && w < 2; w++) {
    const PgBox& bx = sp->statics[sp->clear_walls[w]];
    for (int i = 0; i < 3; i++) c.clear_pos[w][i] = bx.pos[i];
    pk_quat_to_matrix(bx.quat, c.clear_R[w]);
  }
  for (int i = 0; i < 3; i++) c.clear_size[i] = sp->clearance_wall_size[i];
  ce = cudaStreamCreateWithFlags(&e->host_stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { cudaFree(d); delete e; return cuda_fail(ce, "cudaStreamCreate"); }
  *out = e;
  return PG_OK;
}

extern "C" int pg_env_destroy(PgEnv* e) {
  if (!e) return PG_OK;
  cudaSetDevice(e->device);
  cudaFree(e->blob); cudaFree(e->pose); cudaFree(e->fins); cudaFree(e->logit); cudaFree(e->latent); cudaFree(e->path_dev);
  cudaFree(e->u_stage); cudaFree(e->noise_stage); cudaFree(e->J_stage); cudaFree(e->metric_stage); cudaFree(e->trace_stage);
  cudaFree(e->regroup); cudaFree(e->ph_states); cudaFree(e->ph_order); cudaFree(e->ph_hist);
  if (e->host_stream) cudaStreamDestroy(e->host_stream);
  delete e;
  return PG_OK;
}

static int ensure_scratch(PgEnv* e, size_t evals) {
  if (e->cap_evals >= evals && e->path_dev) return PG_OK;
  cudaFree(e->pose); cudaFree(e->fins); cudaFree(e->logit); cudaFree(e->latent);
  e->pose = e->fins = nullptr; e->logit = e->latent = nullptr;
  e->cap_evals = 0;                                  // a failed allocation below must not leave a stale capacity behind
  PG_CUDA(cudaMalloc(&e->pose, evals * 7 * sizeof(double)));
  PG_CUDA(cudaMalloc(&e->fins, evals * 6 * sizeof(double)));
  PG_CUDA(cudaMalloc(&e->logit, evals * sizeof(float)));
  PG_CUDA(cudaMalloc(&e->latent, evals * LAT * sizeof(float)));
  if (!e->path_dev) { PG_CUDA(cudaMalloc(&e->path_dev, 64 * sizeof(int))); e->cap_path = 64; }
  e->cap_evals = evals;
  return PG_OK;
}

static int check_rollout_args(PgEnv* e, const double* u, int K, int N, const int32_t* path) {
  if (!e || !u || !path || K <= 0 || N <= 0 || N > 32) { set_error("pg_rollout: bad arguments"); return PG_ERR_INVALID; }
  for (int j = 0; j < N; j++) if (path[j] < 0 || path[j] >= e->host_env.n_states) { set_error("pg_rollout: path state out of range"); return PG_ERR_INVALID; }
  return PG_OK;
}

extern "C" int pg_physics(PgEnv* e, const double* u_dev, const float* noise_dev, int K, int N, const int32_t* path_host,
                          double max_force, double* pose_dev, double* fins_dev, double* metric_dev, double* trace_dev,
                          int32_t* iters_dev, void* stream) {
  int rc = check_rollout_args(e, u_dev, K, N, path_host);
  if (rc) return rc;
  if (!pose_dev || !fins_dev) { set_error("pg_physics: null output"); return PG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PG_CUDA(cudaSetDevice(e->device));
  rc = ensure_scratch(e, (size_t)K * N);
  if (rc) return rc;
  PG_CUDA(cudaMemcpyAsync(e->path_dev, path_host, N * sizeof(int), cudaMemcpyHostToDevice, st));
  return launch_physics(e, u_dev, noise_dev, K, N, e->path_dev, max_force, pose_dev, fins_dev, metric_dev, trace_dev, iters_dev, st);
}

extern "C" int pg_cost(PgEnv* e, const double* pose_dev, const double* fins_dev, int B, float* logit_dev, void* stream) {
  if (!e || !pose_dev || !fins_dev || !logit_dev || B <= 0) { set_error("pg_cost: bad arguments"); return PG_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(e->device));
  int rc = ensure_scratch(e, (size_t)B);
  if (rc) return rc;
  return launch_cost(e, pose_dev, fins_dev, B, e->latent, logit_dev, (cudaStream_t)stream);
}

extern "C" int pg_score(PgEnv* e, const float* pts_dev, const float* df_dev, const uint8_t* mask_dev, const float* grasp_dev,
                        int B, int P, float* logit_dev, void* stream) {
  if (!e || !pts_dev || !df_dev || !grasp_dev || !logit_dev || B <= 0 || P <= 0) { set_error("pg_score: bad arguments"); return PG_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(e->device));
  int rc = ensure_scratch(e, (size_t)B);
  if (rc) return rc;
  return launch_score(e, pts_dev, df_dev, mask_dev, grasp_dev, B, P, e->latent, logit_dev, (cudaStream_t)stream);
}

extern "C" int pg_rollout(PgEnv* e, const double* u_dev, const float* noise_dev, int K, int N, const int32_t* path_host,
                          double max_force, double* J_dev, double* metric_dev, double* trace_dev, void* stream) {
  int rc = check_rollout_args(e, u_dev, K, N, path_host);
  if (rc) return rc;
  if (!J_dev) { set_error("pg_rollout: null J"); return PG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PG_CUDA(cudaSetDevice(e->device));
  rc = ensure_scratch(e, (size_t)K * N);
  if (rc) return rc;
  PG_CUDA(cudaMemcpyAsync(e->path_dev, path_host, N * sizeof(int), cudaMemcpyHostToDevice, st));
  rc = launch_physics(e, u_dev, noise_dev, K, N, e->path_dev, max_force, e->pose, e->fins, metric_dev, trace_dev, nullptr, st);
  if (rc) return rc;
  rc = launch_cost(e, e->pose, e->fins, K * N, e->latent, e->logit, st);
  if (rc) return rc;
  return launch_accumulate_cost(e->logit, K, N, J_dev, st);
}

// model_optimize.py:74-117 runs one MPPI problem per (contact-state path, repetition) one after the other; they are
// independent, so Q of them can share every launch: rollout k belongs to problem prob[k]
extern "C" int pg_rollout_multi(PgEnv* e, const double* u_dev, const float* noise_dev, int K, int N, int Q, const int32_t* prob_dev,
                                const int32_t* paths_host, double max_force, double* J_dev, double* metric_dev, void* stream) {
  if (!e || !u_dev || !paths_host || !prob_dev || !J_dev || K <= 0 || N <= 0 || N > 32 || Q <= 0) { set_error("pg_rollout_multi: bad arguments"); return PG_ERR_INVALID; }
  for (int i = 0; i < Q * N; i++) if (paths_host[i] < 0 || paths_host[i] >= e->host_env.n_states) { set_error("pg_rollout_multi: path state out of range"); return PG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PG_CUDA(cudaSetDevice(e->device));
  int rc = ensure_scratch(e, (size_t)K * N);
  if (rc) return rc;
  if (e->cap_path < (size_t)Q * N) {
    PG_CUDA(cudaStreamSynchronize(st));
    cudaFree(e->path_dev); e->path_dev = nullptr; e->cap_path = 0;
    PG_CUDA(cudaMalloc(&e->path_dev, (size_t)Q * N * sizeof(int)));
    e->cap_path = (size_t)Q * N;
  }
  PG_CUDA(cudaMemcpyAsync(e->path_dev, paths_host, (size_t)Q * N * sizeof(int), cudaMemcpyHostToDevice, st));
  rc = launch_physics(e, u_dev, noise_dev, K, N, e->path_dev, max_force, e->pose, e->fins, metric_dev, nullptr, nullptr, st, prob_dev);
  if (rc) return rc;
  rc = launch_cost(e, e->pose, e->fins, K * N, e->latent, e->logit, st);
  if (rc) return rc;
  return launch_accumulate_cost(e->logit, K, N, J_dev, st);
}

extern "C" int pg_rollout_host(PgEnv* e, const double* u_host, const float* noise_host, int K, int N,
                               const int32_t* path_host, double max_force, double* J_host, double* metric_host,
                               double* trace_host) {
  int rc = check_rollout_args(e, u_host, K, N, path_host);
  if (rc) return rc;
  if (!J_host) { set_error("pg_rollout_host: null J"); return PG_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(e->device));
  if (e->cap_stage_K < (size_t)K) {
    cudaFree(e->noise_stage); cudaFree(e->J_stage); cudaFree(e->metric_stage); cudaFree(e->u_stage);
    e->noise_stage = nullptr; e->J_stage = e->metric_stage = e->u_stage = nullptr;
    PG_CUDA(cudaMalloc(&e->noise_stage, (size_t)K * 32 * ADIM * sizeof(float)));
    PG_CUDA(cudaMalloc(&e->J_stage, (size_t)K * sizeof(double)));
    PG_CUDA(cudaMalloc(&e->metric_stage, (size_t)K * sizeof(double)));
    PG_CUDA(cudaMalloc(&e->u_stage, 32 * ADIM * sizeof(double)));
    e->cap_stage_K = K;
  }
  size_t tr = trace_host ? (size_t)K * N * SKIP * 7 : 0;
  if (tr > e->cap_stage_trace) {
    cudaFree(e->trace_stage); e->trace_stage = nullptr;
    PG_CUDA(cudaMalloc(&e->trace_stage, tr * sizeof(double)));
    e->cap_stage_trace = tr;
  }
  cudaStream_t st = e->host_stream;                 // the handle's own stream: host-buffer calls on different handles overlap
  PG_CUDA(cudaMemcpyAsync(e->u_stage, u_host, (size_t)N * ADIM * sizeof(double), cudaMemcpyHostToDevice, st));
  if (noise_host) PG_CUDA(cudaMemcpyAsync(e->noise_stage, noise_host, (size_t)K * N * ADIM * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = pg_rollout(e, e->u_stage, noise_host ? e->noise_stage : nullptr, K, N, path_host, max_force, e->J_stage,
                  metric_host ? e->metric_stage : nullptr, trace_host ? e->trace_stage : nullptr, st);
  if (rc) return rc;
  PG_CUDA(cudaMemcpyAsync(J_host, e->J_stage, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (metric_host) PG_CUDA(cudaMemcpyAsync(metric_host, e->metric_stage, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (trace_host) PG_CUDA(cudaMemcpyAsync(trace_host, e->trace_stage, tr * sizeof(double), cudaMemcpyDeviceToHost, st));
  PG_CUDA(cudaStreamSynchronize(st));
  return PG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-sample gym-style stepping (SURVEY 8(b) pg_env_step_single + PgCarry): one episode advanced one env step per call.
struct PgCarry {
  PgEnv* env;
  void* state;          // pg::EpisodeState on the device
  double* io;           // device staging: a[48] | pose[7] | fins[6] | metric[1] | trace[50*7]
  float* logit;         // [1]
  float* latent;        // [26]
  int step;
};
enum { IO_A = 0, IO_POSE = ADIM, IO_FINS = ADIM + 7, IO_METRIC = ADIM + 13, IO_TRACE = ADIM + 14, IO_DOUBLES = ADIM + 14 + SKIP * 7 };

extern "C" int pg_carry_reset(PgCarry* c, double* pose7_host) {
  if (!c) { set_error("pg_carry_reset: null carry"); return PG_ERR_INVALID; }
  PgEnv* e = c->env;
  PG_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = e->host_stream;
  int rc = launch_episode(e, c->state, nullptr, 0, 0.0, 0, c->io + IO_POSE, nullptr, nullptr, nullptr, st);
  if (rc) return rc;
  if (pose7_host) PG_CUDA(cudaMemcpyAsync(pose7_host, c->io + IO_POSE, 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
  PG_CUDA(cudaStreamSynchronize(st));
  c->step = 0;
  return PG_OK;
}

extern "C" int pg_carry_create(PgEnv* e, PgCarry** out) {
  if (!e || !out) { set_error("pg_carry_create: null argument"); return PG_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(e->device));
  PgCarry* c = new PgCarry();
  memset(c, 0, sizeof(PgCarry));
  c->env = e;
  cudaError_t ce = cudaMalloc(&c->state, episode_state_bytes());
  if (ce == cudaSuccess) ce = cudaMalloc(&c->io, IO_DOUBLES * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc(&c->logit, sizeof(float));
  if (ce == cudaSuccess) ce = cudaMalloc(&c->latent, LAT * sizeof(float));
  if (ce != cudaSuccess) { cudaFree(c->state); cudaFree(c->io); cudaFree(c->logit); cudaFree(c->latent); delete c; return cuda_fail(ce, "cudaMalloc(carry)"); }
  int rc = pg_carry_reset(c, nullptr);
  if (rc) { cudaFree(c->state); cudaFree(c->io); cudaFree(c->logit); cudaFree(c->latent); delete c; return rc; }
  *out = c;
  return PG_OK;
}

extern "C" int pg_carry_destroy(PgCarry* c) {
  if (!c) return PG_OK;
  cudaSetDevice(c->env->device);
  cudaFree(c->state); cudaFree(c->io); cudaFree(c->logit); cudaFree(c->latent);
  delete c;
  return PG_OK;
}

extern "C" int pg_env_step_single(PgEnv* e, PgCarry* c, const double* a_host, int32_t state, double max_force, int32_t last_step,
                                  double* reward_host, double* metric_host, double* pose7_host, double* trace_host) {
  if (!e || !c || c->env != e || !a_host) { set_error("pg_env_step_single: bad arguments"); return PG_ERR_INVALID; }
  if (state < 0 || state >= e->host_env.n_states) { set_error("pg_env_step_single: contact state out of range"); return PG_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = e->host_stream;
  PG_CUDA(cudaMemcpyAsync(c->io + IO_A, a_host, ADIM * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = launch_episode(e, c->state, c->io + IO_A, state, max_force, last_step ? 1 : 0, c->io + IO_POSE, c->io + IO_FINS,
                          c->io + IO_METRIC, trace_host ? c->io + IO_TRACE : nullptr, st);
  if (rc) return rc;
  // reward = 100 x score logit on the pose after this step (after the settle substeps on the last one), plate_env:608-612
  rc = launch_cost(e, c->io + IO_POSE, c->io + IO_FINS, 1, c->latent, c->logit, st);
  if (rc) return rc;
  float logit = 0.f;
  PG_CUDA(cudaMemcpyAsync(&logit, c->logit, sizeof(float), cudaMemcpyDeviceToHost, st));
  if (pose7_host) PG_CUDA(cudaMemcpyAsync(pose7_host, c->io + IO_POSE, 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (metric_host && last_step) PG_CUDA(cudaMemcpyAsync(metric_host, c->io + IO_METRIC, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (trace_host) PG_CUDA(cudaMemcpyAsync(trace_host, c->io + IO_TRACE, SKIP * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
  PG_CUDA(cudaStreamSynchronize(st));
  if (metric_host && !last_step) *metric_host = 0.0;
  if (reward_host) *reward_host = 100.0 * (double)logit;
  c->step += 1;
  return PG_OK;
}
