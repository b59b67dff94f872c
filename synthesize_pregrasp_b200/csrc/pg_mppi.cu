// Kernel 3: exp-weighted (softmax) MPPI update, stoch_traj_opt.py:192-204.
//   J -= min(J); S = exp(-kappa J); u[j,d] += alpha * sum_k S_k E[k,j,d] / sum_k S_k
// HBM-bound: streams the fp32 noise once (4*N*D bytes per rollout).  Three launches, deterministic:
// (1) block minima of J, (2) per-block weighted partial sums with a fixed intra-block reduction order,
// (3) one block folds the partials in index order, writes (min, sum w, sum J, sum w*E) for the multi-GPU
// merge and optionally applies the update.
#include "pg_internal.h"

namespace pg {

#define MP_THREADS 256
#define MP_MAX_ND 1536          // N <= 32 env steps x D = 48 (the rollout ABI's limit)
#define FIN_PARTS 8             // mppi_final_kernel: threads per column

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

// Streams a contiguous chunk of rows (rollouts).  Per tile of MP_THREADS rows every thread computes ONE weight (fp64 exp is
// ~100 instructions; the columns of a row share it through shared memory), then thread (g, c) accumulates column slot c
// (VEC consecutive floats, one 16-byte load when VEC = 4) over the rows g, g + groups, ... of the tile.  Fixed order
// everywhere: the result does not depend on scheduling.   partial layout per block: [0] sum w, [1] sum J, [2..2+ND) sum w*E
template <int VEC>
__global__ void __launch_bounds__(MP_THREADS)
mppi_accum_kernel(const double* __restrict__ J, const float* __restrict__ noise, int K, int ND, double kappa,
                  const double* __restrict__ block_min, int n_min, int rows_per_block, double* __restrict__ partial) {
  __shared__ double s_min;
  __shared__ double s_wt[MP_THREADS];            // weights of the current tile of rows
  __shared__ double s_acc[MP_MAX_ND];            // [groups][ND] partial sums (groups * ND <= max(MP_THREADS * VEC, ND))
  __shared__ double s_red[2][MP_THREADS / 32];
  const int t = threadIdx.x;
  if (t == 0) {
    double m = block_min[0];
    for (int i = 1; i < n_min; i++) m = fmin(m, block_min[i]);
    s_min = m;
  }
  __syncthreads();
  const double jmin = s_min;
  const int slots = ND / VEC;                              // column slots
  const int ct = slots < MP_THREADS ? slots : MP_THREADS;  // threads per row group
  const int groups = MP_THREADS / ct;
  const int g = t / ct, c_in = t % ct;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(K, r0 + rows_per_block);
  double sw = 0, sj = 0;
  double* out = partial + (size_t)blockIdx.x * (2 + ND);
  for (int c0 = 0; c0 < slots; c0 += ct) {                 // one pass unless N*D/VEC > MP_THREADS
    const int c = c0 + c_in;
    const bool col_ok = (g < groups) && (c < slots);
    double acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = 0;
    for (int tile = r0; tile < r1; tile += MP_THREADS) {
      const int nrow = min(MP_THREADS, r1 - tile);
      __syncthreads();                                     // previous tile's weights are consumed
      double w = 0;
      if (t < nrow) {
        const double jr = J[tile + t];
        w = exp(-kappa * (jr - jmin));                     // +inf cost -> weight 0
        if (c0 == 0) { sw += w; sj += jr; }
      }
      s_wt[t] = w;
      __syncthreads();
      if (col_ok) {
        const float* base = noise + (size_t)tile * ND + (size_t)c * VEC;
        for (int r = g; r < nrow; r += groups) {
          const double wr = s_wt[r];
          if constexpr (VEC == 4) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * ND));
            acc[0] += wr * (double)e.x; acc[1] += wr * (double)e.y; acc[2] += wr * (double)e.z; acc[3] += wr * (double)e.w;
          } else {
            acc[0] += wr * (double)__ldg(base + (size_t)r * ND);
          }
        }
      }
    }
    __syncthreads();
    if (col_ok) {
#pragma unroll
      for (int v = 0; v < VEC; v++) s_acc[g * (ct * VEC) + c_in * VEC + v] = acc[v];
    }
    __syncthreads();
    for (int i = t; i < ct * VEC && c0 * VEC + i < ND; i += MP_THREADS) {
      double v = 0;
      for (int q = 0; q < groups; q++) v += s_acc[q * (ct * VEC) + i];
      out[2 + c0 * VEC + i] = v;
    }
    __syncthreads();                                       // s_acc is rewritten by the next column pass
  }
  // sum w, sum J: warp shuffle tree, then the warps in index order
  for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(0xffffffffu, sw, o); sj += __shfl_xor_sync(0xffffffffu, sj, o); }
  if ((t & 31) == 0) { s_red[0][t >> 5] = sw; s_red[1][t >> 5] = sj; }
  __syncthreads();
  if (t == 0) {
    double a = 0, b = 0;
    for (int q = 0; q < MP_THREADS / 32; q++) { a += s_red[0][q]; b += s_red[1][q]; }
    out[0] = a; out[1] = b;
  }
}

// folds the per-block partials: thread d sums column d over the blocks in index order; warp 0 sums (sum w, sum J) with a
// fixed lane-strided order + shuffle tree
__global__ void mppi_final_kernel(const double* __restrict__ partial, int nblocks, int ND, const double* __restrict__ block_min,
                                  int n_min, double alpha, double* __restrict__ u, double* __restrict__ out, int apply) {
  __shared__ double s_sw;
  const int t = threadIdx.x;
  if (t < 32) {
    double a = 0, b = 0;
    for (int i = t; i < nblocks; i += 32) { a += partial[(size_t)i * (2 + ND)]; b += partial[(size_t)i * (2 + ND) + 1]; }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    double m = INFINITY;
    for (int i = t; i < n_min; i += 32) m = fmin(m, block_min[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (t == 0) { s_sw = a; if (out) { out[0] = m; out[1] = a; out[2] = b; } }
  }
  __syncthreads();
  // column d: FIN_PARTS threads sum contiguous ranges of the blocks (in index order), lane 0 of the group adds the FIN_PARTS
  // partial sums in order -- a fixed tree, so the result does not depend on scheduling; eight times shorter dependent chain of
  // L2 loads than one thread per column (this kernel was 52 us of a 175 us update at K = 2^20)
  const int part = t % FIN_PARTS, per = (nblocks + FIN_PARTS - 1) / FIN_PARTS;
  const int i0 = part * per, i1 = min(nblocks, i0 + per);
  for (int d0 = 0; d0 < ND; d0 += blockDim.x / FIN_PARTS) {
    const int d = d0 + t / FIN_PARTS;
    double v = 0;
    if (d < ND) for (int i = i0; i < i1; i++) v += partial[(size_t)i * (2 + ND) + 2 + d];
    double tot = v;                                  // FIN_PARTS consecutive lanes of one warp hold one column
    for (int q = 1; q < FIN_PARTS; q++) { const double o = __shfl_sync(0xffffffffu, v, (t & 31) - part + q); if (part == 0) tot += o; }
    if (d < ND && part == 0) {
      if (out) out[3 + d] = tot;
      if (apply) u[d] = u[d] + alpha * (tot / s_sw);
    }
  }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_mppi_update(const double* J_dev, const float* noise_dev, int K, int N, int D, double kappa, double alpha,
                              double* u_inout_dev, double* partial_dev, int apply, void* stream) {
  if (!J_dev || !noise_dev || K <= 0 || N <= 0 || D <= 0) { set_error("pg_mppi_update: bad arguments"); return PG_ERR_INVALID; }
  const int ND = N * D;
  if (ND > MP_MAX_ND) { set_error("pg_mppi_update: N*D > 1536 not supported"); return PG_ERR_INVALID; }
  if (apply && !u_inout_dev) { set_error("pg_mppi_update: apply without u"); return PG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  PG_CUDA(cudaGetDevice(&dev));
  int sms = 148;
  PG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int nblocks = sms * 4;
  int rows_per_block = (K + nblocks - 1) / nblocks;
  if (rows_per_block < 32) rows_per_block = 32;
  nblocks = (K + rows_per_block - 1) / rows_per_block;
  int n_min = (K + 1023) / 1024;
  if (n_min > sms) n_min = sms;
  // scratch from the stream-ordered pool: no process-global state, safe with several devices / threads / streams
  const size_t need = (size_t)nblocks * (2 + ND) + n_min;
  double* partial = nullptr;
  PG_CUDA(cudaMallocAsync(&partial, need * sizeof(double), st));
  double* bmin = partial + (size_t)nblocks * (2 + ND);
  mppi_min_kernel<<<n_min, 256, 0, st>>>(J_dev, K, bmin); count_launch();
  const bool vec4 = (ND % 4 == 0) && ((reinterpret_cast<uintptr_t>(noise_dev) & 15) == 0);
  if (vec4) mppi_accum_kernel<4><<<nblocks, MP_THREADS, 0, st>>>(J_dev, noise_dev, K, ND, kappa, bmin, n_min, rows_per_block, partial);
  else mppi_accum_kernel<1><<<nblocks, MP_THREADS, 0, st>>>(J_dev, noise_dev, K, ND, kappa, bmin, n_min, rows_per_block, partial);
  count_launch();
  mppi_final_kernel<<<1, 1024, 0, st>>>(partial, nblocks, ND, bmin, n_min, alpha, u_inout_dev, partial_dev, apply); count_launch();
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(partial, st);
  if (le != cudaSuccess) return cuda_fail(le, "pg_mppi_update launch");
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
