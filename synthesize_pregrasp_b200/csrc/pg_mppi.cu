// Kernel 3: exp-weighted (softmax) MPPI update, stoch_traj_opt.py:192-204.
//   J -= min(J); S = exp(-kappa J); u[j,d] += alpha * sum_k S_k E[k,j,d] / sum_k S_k
// HBM-bound: streams the fp32 noise once (4*N*D bytes per rollout).  Three launches, deterministic:
// (1) block minima of J, (2) per-block weighted partial sums with a fixed intra-block reduction order,
// (3) one block folds the partials in index order, writes (min, sum w, sum J, sum w*E) for the multi-GPU
// merge and optionally applies the update.
#include "pg_internal.h"

namespace pg {

#define MP_THREADS 384          // 4 row-groups x 96 columns when N*D = 96

__global__ void mppi_min_kernel(const double* __restrict__ J, int K, double* __restrict__ block_min) {
  double m = INFINITY;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) m = fmin(m, J[k]);
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x + 31) / 32; w++) m = fmin(m, sm[w]);
    block_min[blockIdx.x] = m;
  }
}

// partial layout per block: [0] sum w, [1] sum J, [2..2+ND) sum w*E
__global__ void __launch_bounds__(MP_THREADS)
mppi_accum_kernel(const double* __restrict__ J, const float* __restrict__ noise, int K, int ND, double kappa,
                  const double* __restrict__ block_min, int n_min, int rows_per_block, double* __restrict__ partial) {
  __shared__ double s_min;
  __shared__ double s_acc[MP_THREADS];
  __shared__ double s_w[4], s_j[4];
  if (threadIdx.x == 0) {
    double m = block_min[0];
    for (int i = 1; i < n_min; i++) m = fmin(m, block_min[i]);
    s_min = m;
  }
  __syncthreads();
  const double jmin = s_min;
  const int groups = MP_THREADS / ND;            // row groups handled concurrently
  const int g = threadIdx.x / ND, d = threadIdx.x % ND;
  int r0 = blockIdx.x * rows_per_block, r1 = min(K, r0 + rows_per_block);
  double acc = 0, sw = 0, sj = 0;
  if (g < groups) {
    for (int r = r0 + g; r < r1; r += groups) {
      double jr = J[r];
      double w = exp(-kappa * (jr - jmin));      // +inf cost -> weight 0
      acc += w * (double)__ldg(noise + (size_t)r * ND + d);
      if (d == 0) { sw += w; sj += jr; }
    }
  }
  s_acc[threadIdx.x] = acc;
  if (g < groups && d == 0) { s_w[g] = sw; s_j[g] = sj; }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * (2 + ND);
  if (threadIdx.x < ND) {
    double v = 0;
    for (int q = 0; q < groups; q++) v += s_acc[q * ND + threadIdx.x];
    out[2 + threadIdx.x] = v;
  }
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int q = 0; q < groups; q++) { a += s_w[q]; b += s_j[q]; }
    out[0] = a; out[1] = b;
  }
}

__global__ void mppi_final_kernel(const double* __restrict__ partial, int nblocks, int ND, const double* __restrict__ block_min,
                                  int n_min, double alpha, double* __restrict__ u, double* __restrict__ out, int apply) {
  __shared__ double s_sw;
  int t = threadIdx.x;
  if (t == 0) {
    double m = block_min[0];
    for (int i = 1; i < n_min; i++) m = fmin(m, block_min[i]);
    double a = 0, b = 0;
    for (int i = 0; i < nblocks; i++) { a += partial[(size_t)i * (2 + ND)]; b += partial[(size_t)i * (2 + ND) + 1]; }
    s_sw = a;
    if (out) { out[0] = m; out[1] = a; out[2] = b; }
  }
  __syncthreads();
  for (int d = t; d < ND; d += blockDim.x) {
    double v = 0;
    for (int i = 0; i < nblocks; i++) v += partial[(size_t)i * (2 + ND) + 2 + d];
    if (out) out[3 + d] = v;
    if (apply) u[d] = u[d] + alpha * (v / s_sw);
  }
}

}  // namespace pg

using namespace pg;

static double* g_scratch[16] = {nullptr};
static size_t g_scratch_cap[16] = {0};

extern "C" int pg_mppi_update(const double* J_dev, const float* noise_dev, int K, int N, int D, double kappa, double alpha,
                              double* u_inout_dev, double* partial_dev, int apply, void* stream) {
  if (!J_dev || !noise_dev || K <= 0 || N <= 0 || D <= 0) { set_error("pg_mppi_update: bad arguments"); return PG_ERR_INVALID; }
  int ND = N * D;
  if (ND > MP_THREADS) { set_error("pg_mppi_update: N*D > 384 not supported"); return PG_ERR_INVALID; }
  if (apply && !u_inout_dev) { set_error("pg_mppi_update: apply without u"); return PG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int nblocks = sms * 4;
  int rows_per_block = (K + nblocks - 1) / nblocks;
  if (rows_per_block < 32) { rows_per_block = 32; }
  nblocks = (K + rows_per_block - 1) / rows_per_block;
  int n_min = (K + 1023) / 1024;
  if (n_min > sms) n_min = sms;
  size_t need = (size_t)nblocks * (2 + ND) + n_min;
  if (dev < 16 && g_scratch_cap[dev] < need) {
    if (g_scratch[dev]) cudaFree(g_scratch[dev]);
    PG_CUDA(cudaMalloc(&g_scratch[dev], need * sizeof(double)));
    g_scratch_cap[dev] = need;
  }
  double* partial = g_scratch[dev];
  double* bmin = partial + (size_t)nblocks * (2 + ND);
  mppi_min_kernel<<<n_min, 256, 0, st>>>(J_dev, K, bmin); count_launch();
  mppi_accum_kernel<<<nblocks, MP_THREADS, 0, st>>>(J_dev, noise_dev, K, ND, kappa, bmin, n_min, rows_per_block, partial); count_launch();
  mppi_final_kernel<<<1, 128, 0, st>>>(partial, nblocks, ND, bmin, n_min, alpha, u_inout_dev, partial_dev, apply); count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

extern "C" int pg_mppi_merge_apply(const double* P, int G, int N, int D, double kappa, double alpha, double* u) {
  if (!P || !u || G <= 0) { set_error("pg_mppi_merge_apply: bad arguments"); return PG_ERR_INVALID; }
  int ND = N * D, stride = 3 + ND;
  double m = P[0];
  for (int g = 1; g < G; g++) m = fmin(m, P[g * stride]);
  double S = 0;
  for (int g = 0; g < G; g++) S += P[g * stride + 1] * exp(-kappa * (P[g * stride] - m));
  for (int d = 0; d < ND; d++) {
    double v = 0;
    for (int g = 0; g < G; g++) v += P[g * stride + 3 + d] * exp(-kappa * (P[g * stride] - m));
    u[d] = u[d] + alpha * (v / S);
  }
  return PG_OK;
}
