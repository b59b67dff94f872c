// Host-side definition of the env handle shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/pregrasp_b200.h"
#include "pg_env.cuh"

namespace pg {

enum { P_CLOUD = 2048, C1 = 32, C2 = 64, C3 = 256, C4 = 20, LAT = 26, F1 = 512, F2 = 512, F3 = 256 };

struct CostDev {                 // device pointers of the cost stage constants
  const double* cloud;           // [P,3] f64
  const float* a1;               // [P,32]  W1[:, :3] xyz + b1  (rollout-invariant part of conv1)
  const float* w1d;              // [32]    W1[:, 3]
  const float* w2t;              // [32,64] conv2 weight transposed
  const float* b2;               // [64]
  const float* w3at;             // [64,256] conv3 weight, point half, transposed
  const float* w3bt;             // [64,256] conv3 weight, global-feature half, transposed
  const float* b3;               // [256]
  const float* w4;               // [20,256]
  const float* b4;               // [20]
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *fc3_w, *fc3_b, *out_w, *out_b;
  const float* w1;               // [32,4] full conv1 (pg_score path with explicit points)
  const float* b1;
  int n_cloud, n_crop, n_dist_box, n_dist_tri;
  double crop[2][2][3];
  double dist_box[4][2][3];
  const double* dist_tri;
  const double* tri_sphere;      // [n_dist_tri][4]: centroid and radius of each triangle's bounding sphere
  const double* tri_group_sphere; // [ceil(n_dist_tri/16)][4]: bounding sphere of each group of 16 consecutive triangles
  int n_dist_prism;               // right prisms (regular polygon cross-section, axis z): tessellated cylinders in closed form
  int prism_sides[2];
  double prism_c[2][3], prism_hh[2];
  double prism_v[2][33][2];       // polygon vertices (closed: v[sides] = v[0]), counter-clockwise
  double clearance_h;
  int project_uses_clearance;     // 1: checkClearance(local point) incl. the walls (bookshelf_graph_env.py:683-689)
  int clear_n;
  double clear_pos[2][3], clear_R[2][9], clear_size[3];
};

// tensor-core operands of the cost kernel: tf32 hi / lo splits in the UMMA canonical K-major layout [(k/4)][n][k%4]
struct TcWeights {
  const float *w32_hi, *w32_lo;  // (W3a W2): conv3's point half with conv2 folded in, [256 x 32]
  const float *w4_hi, *w4_lo;    // conv4 padded to 32 outputs, [32 x 256]
  const float *w2_hi, *w2_lo;    // conv2, [64 x 32]
  const float* cbase;            // [256]  b3 + W3a b2
};

}  // namespace pg

struct PgEnv {
  int device;
  pg::EnvDev host_env;
  pg::EnvDev* dev_env;
  pg::CostDev cost;
  void* blob;                    // one allocation holding every constant table
  pg::TcWeights tcw;             // tensor-core operands of pcn_tc_kernel
  int use_tensor_cores;
  // scratch (grown on demand)
  double *pose, *fins;           // [K,N,7], [K,N,2,3]
  float* logit;                  // [K*N]
  float* latent;                 // [K*N,26]
  int* path_dev; size_t cap_path;
  size_t cap_evals;
  double *u_stage; float* noise_stage; double *J_stage, *metric_stage, *trace_stage;   // pg_rollout_host staging
  size_t cap_stage_K, cap_stage_trace;
  cudaStream_t host_stream;            // pg_rollout_host: copies + kernels of one handle run on its own stream
  void* regroup; size_t cap_regroup;   // settle-phase exchange buffer of the rollout kernel ([K] RegroupSlot)
  void* ph_states; size_t cap_ph;      // phased rollouts: [K] RolloutState between phases, sort scratch
  int *ph_order, *ph_hist;
  int rollout_mode;                    // 0 auto (phased for large batches), 1 persistent kernel, 2 phased (pg_set_rollout_mode)
};

namespace pg {
void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();
#define PG_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return pg::cuda_fail(_e, #x); } while (0)
int launch_physics(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                   double* pose, double* fins, double* metric, double* trace, int* iters, cudaStream_t st,
                   const int* prob_dev = nullptr);
int launch_physics_phased(PgEnv* env, const double* u, const float* noise, int K, int N, const int* path_dev, double max_force,
                          double* pose, double* fins, double* metric, cudaStream_t st, const int* prob_dev);
size_t episode_state_bytes();
int launch_episode(PgEnv* env, void* state, const double* a, int cstate, double max_force, int settle, double* pose7, double* fins6,
                   double* metric, double* trace, cudaStream_t st);
int launch_cost(PgEnv* env, const double* pose, const double* fins, int B, float* latent, float* logit, cudaStream_t st);
int launch_score(PgEnv* env, const float* pts, const float* df, const uint8_t* mask, const float* grasp, int B, int P,
                 float* latent, float* logit, cudaStream_t st);
int launch_cost_tc(PgEnv* env, const double* pose, const double* fins, int B, float* latent, cudaStream_t st);
int launch_mlp(PgEnv* env, const float* latent, int B, float* logit, cudaStream_t st);
int launch_accumulate_cost(const float* logit, int K, int N, double* J, cudaStream_t st);
}
