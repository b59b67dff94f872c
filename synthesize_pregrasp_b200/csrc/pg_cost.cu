// Kernel 2: fused per-step cost.  One CTA per (rollout, env-step) evaluation:
//   object pose -> world cloud -> de-occlusion crop mask -> exact distance to the env boxes (the reference's
//   "distance field") -> nearest cloud point of each finger -> PCN encoder -> [20] latent;  a second small
//   kernel runs the 26->512->512->256->1 MLP over groups of evaluations.
// Replaces envs/plate_contact_graph_env.py:393-437 (calc_reward_value_stage_one), :661-684, the multi-box
// variants envs/bookshelf_graph_env.py:645-690, neurals/distancefield_utils.py:15-22 and
// neurals/network.py:398-412 / :443-455, which the reference runs at batch 1 through Open3D + torch.
// Exact algebra used to cut work (no approximation): (i) conv1 over xyz is rollout-invariant, precomputed per
// point; (ii) conv3(concat(x, g)) = W3a x + W3b g with the g half evaluated once per cloud; (iii) cropped
// points are masked to -inf before both max-pools; [P,256] activations never leave the SM.
#include "pg_internal.h"
#include "pg_cost_geom.cuh"

namespace pg {

#define NT 256
#define TILE 64

struct Smem {
  float w2t[C1 * C2];
  float w3at[C2 * C3];
  float w4t[C3 * C4];      // [256][20]
  float h2T[C2 * TILE];    // [k][pt]
  float red[NT * 8];       // scratch
  float df[P_CLOUD];
  float msk[P_CLOUD];      // 0 or -inf
  float c[C3];
  float g[C2];
  float b2[C2], w1d[C1], b4[C4 + 4];
  float w1[C1 * 4], b1[C1];
  double pose[7], R[9];
  double fin[6];
  float lat[LAT + 2];
  int count;
};

__device__ __forceinline__ float box_dist(double x, double y, double z, const double* lo, const double* hi) {
  double dx = fmax(fmax(lo[0] - x, x - hi[0]), 0.0), dy = fmax(fmax(lo[1] - y, y - hi[1]), 0.0),
         dz = fmax(fmax(lo[2] - z, z - hi[2]), 0.0);
  double outside = sqrt(dx * dx + dy * dy + dz * dz);
  double inside = fmin(fmin(fmin(x - lo[0], hi[0] - x), fmin(y - lo[1], hi[1] - y)), fmin(z - lo[2], hi[2] - z));
  return (float)(outside > 0 ? outside : fmax(inside, 0.0));
}


// h1 = relu(conv1) for one point; FROM_POSE uses the precomputed xyz part
template <bool FROM_POSE>
__device__ __forceinline__ void point_h1(const Smem& s, const CostDev& cd, const float* pts, int p, float df, float* h1) {
  if (FROM_POSE) {
    const float4* a = reinterpret_cast<const float4*>(cd.a1 + (size_t)p * C1);
#pragma unroll
    for (int j = 0; j < C1 / 4; j++) {
      float4 v = __ldg(a + j);
      h1[4 * j + 0] = fmaxf(fmaf(s.w1d[4 * j + 0], df, v.x), 0.f);
      h1[4 * j + 1] = fmaxf(fmaf(s.w1d[4 * j + 1], df, v.y), 0.f);
      h1[4 * j + 2] = fmaxf(fmaf(s.w1d[4 * j + 2], df, v.z), 0.f);
      h1[4 * j + 3] = fmaxf(fmaf(s.w1d[4 * j + 3], df, v.w), 0.f);
    }
  } else {
    float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
#pragma unroll
    for (int j = 0; j < C1; j++) {
      float v = s.b1[j];
      v = fmaf(s.w1[4 * j], x, v); v = fmaf(s.w1[4 * j + 1], y, v); v = fmaf(s.w1[4 * j + 2], z, v); v = fmaf(s.w1[4 * j + 3], df, v);
      h1[j] = fmaxf(v, 0.f);
    }
  }
}

template <bool FROM_POSE>
__global__ void __launch_bounds__(NT, 1)
pcn_kernel(CostDev cd, const double* __restrict__ pose, const double* __restrict__ fins, const float* __restrict__ pts_in,
           const float* __restrict__ df_in, const uint8_t* __restrict__ mask_in, const float* __restrict__ grasp_in,
           int B, int P, float* __restrict__ latent) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  // ---- stage the weights once per CTA (persistent over evaluations)
  for (int i = tid; i < C1 * C2; i += NT) s.w2t[i] = cd.w2t[i];
  for (int i = tid; i < C2 * C3; i += NT) s.w3at[i] = cd.w3at[i];
  for (int i = tid; i < C3 * C4; i += NT) { int ch = i / C4, o = i % C4; s.w4t[i] = cd.w4[o * C3 + ch]; }
  if (tid < C2) s.b2[tid] = cd.b2[tid];
  if (tid < C1) { s.w1d[tid] = cd.w1d[tid]; s.b1[tid] = cd.b1[tid]; }
  if (tid < C1 * 4) s.w1[tid] = cd.w1[tid];
  if (tid < C4) s.b4[tid] = cd.b4[tid];
  __syncthreads();

  for (int e = blockIdx.x; e < B; e += gridDim.x) {
    const float* pts = FROM_POSE ? nullptr : pts_in + (size_t)e * P * 3;
    // ---------------- stage 0: geometry (crop mask, distance field, finger projection)
    if (FROM_POSE) {
      if (tid < 7) s.pose[tid] = pose[(size_t)e * 7 + tid];
      if (tid < 6) s.fin[tid] = fins[(size_t)e * 6 + tid];
      __syncthreads();
      if (tid == 0) {
        // pytorch_kinematics.quaternion_to_matrix on (w,x,y,z)  (utils/helper.py:304-307)
        double x = s.pose[3], y = s.pose[4], z = s.pose[5], w = s.pose[6];
        double two_s = 2.0 / (w * w + x * x + y * y + z * z);
        s.R[0] = 1 - two_s * (y * y + z * z); s.R[1] = two_s * (x * y - z * w); s.R[2] = two_s * (x * z + y * w);
        s.R[3] = two_s * (x * y + z * w); s.R[4] = 1 - two_s * (x * x + z * z); s.R[5] = two_s * (y * z - x * w);
        s.R[6] = two_s * (x * z - y * w); s.R[7] = two_s * (y * z + x * w); s.R[8] = 1 - two_s * (x * x + y * y);
      }
      __syncthreads();
      for (int p = tid; p < P; p += NT) {
        double cx = cd.cloud[3 * p], cy = cd.cloud[3 * p + 1], cz = cd.cloud[3 * p + 2];
        double wx = s.R[0] * cx + s.R[1] * cy + s.R[2] * cz + s.pose[0];
        double wy = s.R[3] * cx + s.R[4] * cy + s.R[5] * cz + s.pose[1];
        double wz = s.R[6] * cx + s.R[7] * cy + s.R[8] * cz + s.pose[2];
        bool keep = false;
        for (int b = 0; b < cd.n_crop; b++) {
          const double* lo = cd.crop[b][0];
          const double* hi = cd.crop[b][1];
          keep |= (wx >= lo[0] && wx <= hi[0] && wy >= lo[1] && wy <= hi[1] && wz >= lo[2] && wz <= hi[2]);
        }
        double qx = (double)(float)wx, qy = (double)(float)wy, qz = (double)(float)wz;   // compute_distance(points.astype(float32))
        float d = INFINITY;
        for (int b = 0; b < cd.n_dist_box; b++) d = fminf(d, box_dist(qx, qy, qz, cd.dist_box[b][0], cd.dist_box[b][1]));
        if (cd.n_dist_tri > 0) {
          d = tri_field(cd, qx, qy, qz, d);
        }
        if (cd.n_dist_prism > 0) d = prism_field(cd, qx, qy, qz, d);
        s.df[p] = d;
        s.msk[p] = keep ? 0.f : -INFINITY;
      }
      // nearest cloud point of each finger's local position (project_to_pcd, plate_env:668-674)
      for (int f = 0; f < 2; f++) {
        double fx = s.fin[3 * f], fy = s.fin[3 * f + 1], fz = s.fin[3 * f + 2];
        double best = INFINITY;
        int bi = 0x7fffffff;
        for (int p = tid; p < P; p += NT) {
          double dx = cd.cloud[3 * p] - fx, dy = cd.cloud[3 * p + 1] - fy, dz = cd.cloud[3 * p + 2] - fz;
          double dd = dx * dx + dy * dy + dz * dz;
          if (dd < best) { best = dd; bi = p; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          double ob = __shfl_xor_sync(0xffffffffu, best, o);
          int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        double* rb = reinterpret_cast<double*>(s.red);
        int* ri = reinterpret_cast<int*>(s.red + 32);
        __syncthreads();
        if ((tid & 31) == 0) { rb[tid >> 5] = best; ri[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
          for (int w = 1; w < NT / 32; w++) if (rb[w] < best || (rb[w] == best && ri[w] < bi)) { best = rb[w]; bi = ri[w]; }
          bool ok = fz > cd.clearance_h;                  // plate rule, plate_env:669
          if (cd.project_uses_clearance) {                // checkClearance(point) on the LOCAL point, as written
            ok = !(fz < cd.clearance_h);
            for (int w = 0; w < cd.clear_n && ok; w++) {
              const double* Rw = cd.clear_R[w];
              double rx = fx - cd.clear_pos[w][0], ry = fy - cd.clear_pos[w][1], rz = fz - cd.clear_pos[w][2];
              double lx = Rw[0] * rx + Rw[3] * ry + Rw[6] * rz, ly = Rw[1] * rx + Rw[4] * ry + Rw[7] * rz,
                     lz = Rw[2] * rx + Rw[5] * ry + Rw[8] * rz;
              double dx = fmax(fabs(lx) - cd.clear_size[0], 0.0), dy = fmax(fabs(ly) - cd.clear_size[1], 0.0),
                     dz = fmax(fabs(lz) - cd.clear_size[2], 0.0);
              if (sqrt(dx * dx + dy * dy + dz * dz) < cd.clearance_h) ok = false;
            }
          }
          for (int c = 0; c < 3; c++) s.lat[C4 + 3 * f + c] = ok ? (float)cd.cloud[3 * bi + c] : 100.f;
        }
        __syncthreads();
      }
    } else {
      if (tid < 6) s.lat[C4 + tid] = grasp_in[(size_t)e * 6 + tid];
      __syncthreads();
    }
    const float* dfp = FROM_POSE ? s.df : df_in + (size_t)e * P;
    auto mask_of = [&](int p) -> float {
      if (p >= P) return -INFINITY;
      if (FROM_POSE) return s.msk[p];
      return mask_in ? (mask_in[(size_t)e * P + p] ? 0.f : -INFINITY) : 0.f;
    };

    // ---------------- stage 1: global feature g = max_p conv2(relu(conv1(x_p)))
    float gmax[C2];
#pragma unroll
    for (int c = 0; c < C2; c++) gmax[c] = -INFINITY;
    int cnt = 0;
    for (int p = tid; p < P; p += NT) {
      float m = mask_of(p);
      if (m != 0.f) continue;
      cnt++;
      float h1[C1];
      point_h1<FROM_POSE>(s, cd, pts, p, dfp[p], h1);
      float acc[C2];
#pragma unroll
      for (int c = 0; c < C2; c++) acc[c] = s.b2[c];
#pragma unroll 4
      for (int j = 0; j < C1; j++) {
        float h = h1[j];
        const float4* w = reinterpret_cast<const float4*>(s.w2t + j * C2);
#pragma unroll
        for (int c4 = 0; c4 < C2 / 4; c4++) {
          float4 wv = w[c4];
          acc[4 * c4] = fmaf(wv.x, h, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(wv.y, h, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(wv.z, h, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(wv.w, h, acc[4 * c4 + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < C2; c++) gmax[c] = fmaxf(gmax[c], acc[c]);
    }
    // block max-reduce of gmax[64] (+ kept-point count)
#pragma unroll
    for (int c = 0; c < C2; c++) {
      float v = gmax[c];
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      gmax[c] = v;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if ((tid & 31) == 0) {
      for (int c = 0; c < C2; c++) s.red[(tid >> 5) * C2 + c] = gmax[c];
      reinterpret_cast<int*>(s.red + 8 * C2)[tid >> 5] = cnt;
    }
    __syncthreads();
    if (tid < C2) {
      float v = s.red[tid];
      for (int w = 1; w < NT / 32; w++) v = fmaxf(v, s.red[w * C2 + tid]);
      s.g[tid] = v;
    }
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < NT / 32; w++) t += reinterpret_cast<int*>(s.red + 8 * C2)[w];
      s.count = t;
    }
    __syncthreads();
    if (s.count == 0) {       // every point cropped away: the reference would raise; report +inf cost
      if (tid < LAT) latent[(size_t)e * LAT + tid] = NAN;
      __syncthreads();
      continue;
    }
    // c = W3b g + b3
    {
      float v = cd.b3[tid];
      for (int k = 0; k < C2; k++) v = fmaf(__ldg(cd.w3bt + k * C3 + tid), s.g[k], v);
      s.c[tid] = v;
    }
    __syncthreads();

    // ---------------- stage 2: per tile  h2 -> conv3 (+c, relu) -> conv4 -> masked max
    const int tx = tid & 15, ty = tid >> 4;
    float rmax[C4];
#pragma unroll
    for (int o = 0; o < C4; o++) rmax[o] = -INFINITY;
    for (int t0 = 0; t0 < P; t0 += TILE) {
      {   // h2 tile, thread -> (point, 16 channels)
        int pt = tid & (TILE - 1), q = tid >> 6, p = t0 + pt;
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] = s.b2[q * 16 + c];
        if (p < P) {
          float h1[C1];
          point_h1<FROM_POSE>(s, cd, pts, p, dfp[p], h1);
#pragma unroll 4
          for (int j = 0; j < C1; j++) {
            float h = h1[j];
            const float4* w = reinterpret_cast<const float4*>(s.w2t + j * C2 + q * 16);
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
              float4 wv = w[c4];
              acc[4 * c4] = fmaf(wv.x, h, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(wv.y, h, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(wv.z, h, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(wv.w, h, acc[4 * c4 + 3]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 16; c++) s.h2T[(q * 16 + c) * TILE + pt] = acc[c];
      }
      __syncthreads();
      // conv3 point half: thread (tx,ty) -> points ty*4..+3, channels (i*16+tx)*4+cc
      float acc[4][16];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 16; c++) acc[r][c] = 0.f;
#pragma unroll 2
      for (int k = 0; k < C2; k++) {
        float4 a = *reinterpret_cast<const float4*>(s.h2T + k * TILE + ty * 4);
        float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float4 b = *reinterpret_cast<const float4*>(s.w3at + k * C3 + (i * 16 + tx) * 4);
#pragma unroll
          for (int r = 0; r < 4; r++) {
            acc[r][4 * i] = fmaf(av[r], b.x, acc[r][4 * i]); acc[r][4 * i + 1] = fmaf(av[r], b.y, acc[r][4 * i + 1]);
            acc[r][4 * i + 2] = fmaf(av[r], b.z, acc[r][4 * i + 2]); acc[r][4 * i + 3] = fmaf(av[r], b.w, acc[r][4 * i + 3]);
          }
        }
      }
      // + c, relu, then this thread's share of conv4
      float out[4][C4];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int o = 0; o < C4; o++) out[r][o] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          int ch = (i * 16 + tx) * 4 + cc;
          float cv = s.c[ch];
          float h[4];
#pragma unroll
          for (int r = 0; r < 4; r++) h[r] = fmaxf(acc[r][4 * i + cc] + cv, 0.f);
          const float4* w = reinterpret_cast<const float4*>(s.w4t + ch * C4);
#pragma unroll
          for (int o4 = 0; o4 < C4 / 4; o4++) {
            float4 wv = w[o4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
              out[r][4 * o4] = fmaf(wv.x, h[r], out[r][4 * o4]); out[r][4 * o4 + 1] = fmaf(wv.y, h[r], out[r][4 * o4 + 1]);
              out[r][4 * o4 + 2] = fmaf(wv.z, h[r], out[r][4 * o4 + 2]); out[r][4 * o4 + 3] = fmaf(wv.w, h[r], out[r][4 * o4 + 3]);
            }
          }
        }
      // reduce the 16 channel-slices (lanes differing in tx) and fold into the running masked max
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float m = mask_of(t0 + ty * 4 + r);
#pragma unroll
        for (int o = 0; o < C4; o++) {
          float v = out[r][o];
          v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8);
          rmax[o] = fmaxf(rmax[o], v + s.b4[o] + m);
        }
      }
      __syncthreads();
    }
    // reduce rmax over the 16 ty groups (tx==0 lanes hold the values; all lanes of a group agree)
    if (tx == 0) for (int o = 0; o < C4; o++) s.red[ty * C4 + o] = rmax[o];
    __syncthreads();
    if (tid < C4) {
      float v = s.red[tid];
      for (int w = 1; w < 16; w++) v = fmaxf(v, s.red[w * C4 + tid]);
      s.lat[tid] = v;
    }
    __syncthreads();
    if (tid < LAT) latent[(size_t)e * LAT + tid] = s.lat[tid];
    __syncthreads();
  }
}

// ---- MLP 26 -> 512 -> 512 -> 256 -> 1 over groups of G evaluations per CTA (weights streamed from L2 once per group)
#define MLP_G 8
__global__ void __launch_bounds__(256) mlp_kernel(CostDev cd, const float* __restrict__ latent, int B, float* __restrict__ logit) {
  __shared__ float xin[MLP_G][LAT + 2];
  __shared__ float h1[MLP_G][F1];
  __shared__ float h2[MLP_G][F2];
  __shared__ float h3[MLP_G][F3];
  const int tid = threadIdx.x;
  for (int g0 = blockIdx.x * MLP_G; g0 < B; g0 += gridDim.x * MLP_G) {
    int ng = min(MLP_G, B - g0);
    for (int i = tid; i < MLP_G * LAT; i += 256) { int g = i / LAT, c = i % LAT; xin[g][c] = g < ng ? latent[(size_t)(g0 + g) * LAT + c] : 0.f; }
    __syncthreads();
    for (int o = tid; o < F1; o += 256) {
      float acc[MLP_G];
      float b = cd.fc1_b[o];
#pragma unroll
      for (int g = 0; g < MLP_G; g++) acc[g] = b;
      for (int k = 0; k < LAT; k++) {
        float w = __ldg(cd.fc1_w + o * LAT + k);
#pragma unroll
        for (int g = 0; g < MLP_G; g++) acc[g] = fmaf(w, xin[g][k], acc[g]);
      }
#pragma unroll
      for (int g = 0; g < MLP_G; g++) h1[g][o] = fmaxf(acc[g], 0.f);
    }
    __syncthreads();
    for (int o = tid; o < F2; o += 256) {
      float acc[MLP_G];
      float b = cd.fc2_b[o];
#pragma unroll
      for (int g = 0; g < MLP_G; g++) acc[g] = b;
      const float4* wr = reinterpret_cast<const float4*>(cd.fc2_w + (size_t)o * F1);
      for (int k4 = 0; k4 < F1 / 4; k4++) {
        float4 w = __ldg(wr + k4);
#pragma unroll
        for (int g = 0; g < MLP_G; g++) {
          acc[g] = fmaf(w.x, h1[g][4 * k4], acc[g]); acc[g] = fmaf(w.y, h1[g][4 * k4 + 1], acc[g]);
          acc[g] = fmaf(w.z, h1[g][4 * k4 + 2], acc[g]); acc[g] = fmaf(w.w, h1[g][4 * k4 + 3], acc[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < MLP_G; g++) h2[g][o] = fmaxf(acc[g], 0.f);
    }
    __syncthreads();
    for (int o = tid; o < F3; o += 256) {
      float acc[MLP_G];
      float b = cd.fc3_b[o];
#pragma unroll
      for (int g = 0; g < MLP_G; g++) acc[g] = b;
      const float4* wr = reinterpret_cast<const float4*>(cd.fc3_w + (size_t)o * F2);
      for (int k4 = 0; k4 < F2 / 4; k4++) {
        float4 w = __ldg(wr + k4);
#pragma unroll
        for (int g = 0; g < MLP_G; g++) {
          acc[g] = fmaf(w.x, h2[g][4 * k4], acc[g]); acc[g] = fmaf(w.y, h2[g][4 * k4 + 1], acc[g]);
          acc[g] = fmaf(w.z, h2[g][4 * k4 + 2], acc[g]); acc[g] = fmaf(w.w, h2[g][4 * k4 + 3], acc[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < MLP_G; g++) h3[g][o] = fmaxf(acc[g], 0.f);
    }
    __syncthreads();
    {   // output neuron: one warp per evaluation
      int w = tid >> 5, lane = tid & 31;
      for (int g = w; g < ng; g += 8) {
        float v = 0.f;
        for (int k = lane; k < F3; k += 32) v = fmaf(__ldg(cd.out_w + k), h3[g][k], v);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) logit[g0 + g] = v + cd.out_b[0];
      }
    }
    __syncthreads();
  }
}

__global__ void accumulate_cost_kernel(const float* __restrict__ logit, int K, int N, double* __restrict__ J) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double cost = 0;
  for (int j = 0; j < N; j++) {
    float l = logit[(size_t)k * N + j];
    // reward = 100 * score (plate_env:612); cost += -reward (stoch_traj_opt.py:133-134); NaN/empty cloud -> +inf
    cost += isnan(l) ? INFINITY : -(double)(100.0 * (double)l);
  }
  J[k] = cost;
}

static int sm_count(int device) {
  static int cached[16] = {0};
  if (device < 16 && cached[device]) return cached[device];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (device < 16) cached[device] = n;
  return n;
}

int launch_mlp(PgEnv* env, const float* latent, int B, float* logit, cudaStream_t st) {
  int g2 = (B + MLP_G - 1) / MLP_G;
  if (g2 > 4 * sm_count(env->device)) g2 = 4 * sm_count(env->device);
  mlp_kernel<<<g2, 256, 0, st>>>(env->cost, latent, B, logit);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

int launch_cost(PgEnv* env, const double* pose, const double* fins, int B, float* latent, float* logit, cudaStream_t st) {
  if (env->use_tensor_cores) {
    int rc = launch_cost_tc(env, pose, fins, B, latent, st);
    if (rc) return rc;
    return launch_mlp(env, latent, B, logit, st);
  }
  size_t smem = sizeof(Smem);
  // per-device opt-in, set on every launch (an env may live on any device of the process)
  PG_CUDA(cudaFuncSetAttribute(pcn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PG_CUDA(cudaFuncSetAttribute(pcn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count(env->device);
  if (grid > B) grid = B;
  pcn_kernel<true><<<grid, NT, smem, st>>>(env->cost, pose, fins, nullptr, nullptr, nullptr, nullptr, B, env->cost.n_cloud, latent);
  count_launch();
  PG_CUDA(cudaGetLastError());
  int g2 = (B + MLP_G - 1) / MLP_G;
  if (g2 > 4 * sm_count(env->device)) g2 = 4 * sm_count(env->device);
  mlp_kernel<<<g2, 256, 0, st>>>(env->cost, latent, B, logit);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

int launch_score(PgEnv* env, const float* pts, const float* df, const uint8_t* mask, const float* grasp, int B, int P,
                 float* latent, float* logit, cudaStream_t st) {
  size_t smem = sizeof(Smem);
  PG_CUDA(cudaFuncSetAttribute(pcn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count(env->device);
  if (grid > B) grid = B;
  pcn_kernel<false><<<grid, NT, smem, st>>>(env->cost, nullptr, nullptr, pts, df, mask, grasp, B, P, latent);
  count_launch();
  PG_CUDA(cudaGetLastError());
  int g2 = (B + MLP_G - 1) / MLP_G;
  if (g2 > 4 * sm_count(env->device)) g2 = 4 * sm_count(env->device);
  mlp_kernel<<<g2, 256, 0, st>>>(env->cost, latent, B, logit);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

int launch_accumulate_cost(const float* logit, int K, int N, double* J, cudaStream_t st) {
  accumulate_cost_kernel<<<(K + 255) / 256, 256, 0, st>>>(logit, K, N, J);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
