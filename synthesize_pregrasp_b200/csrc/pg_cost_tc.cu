// Kernel 2 (tensor-core version): fused per-step cost with the dominant contraction on tcgen05.
//
// Same algebra and interface as pcn_kernel<true> in pg_cost.cu (see its header for the reference citations).
// Split of the work per 128-point tile of one evaluation:
//   CUDA cores : conv1 (precomputed xyz part + df), conv2 32->64 (weights as constant-memory operands),
//                +c / ReLU / conv4 256->20 on the accumulator rows read back from TMEM, masked max-pools
//   tcgen05    : conv3 point half  D[128x256] = h2[128x64] x W3a^T[64x256], kind::tf32, fp32 accumulate in TMEM.
// fp32 fidelity on tf32 tensor cores: both operands are split a = a_hi + a_lo (a_hi = cvt.rna.tf32, a_lo = a - a_hi
// exactly) and three MMAs a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulate into the same TMEM tile (dropped term
// a_lo*b_lo <= 2^-22 relative), because MPPI weights exp(-5*100*dlogit) do not tolerate plain-tf32 logits.
// Operands sit in shared memory in the no-swizzle K-major canonical layout ([k-chunk of 16 B][row][16 B]:
// SBO = 128 B between 8-row groups, LBO = rows*16 B between the two 16-byte halves of one K=8 step).
#include "pg_internal.h"
#include "pg_cost_geom.cuh"

namespace pg {

#ifndef TC_THREADS
#define TC_THREADS 256
#endif
#define TC_M 128
#define TC_NS (TC_THREADS / TC_M)          // column groups: each thread owns one point row and 1/TC_NS of the channels
#define TC_CH (C2 / TC_NS)                // conv2 channels per thread
#define TC_COLS (C3 / TC_NS)              // accumulator columns per thread


struct TcSmem {
  float b_hi[C2 * C3];                 // W3a, tf32-rounded, canonical K-major: [(k/4)][n][k%4]
  float b_lo[C2 * C3];
  float a_buf[C2 * TC_M];              // h2 tile (hi part, then lo part): [(k/4)][row][k%4]; also the conv4 partial exchange
  float w2t[C1 * C2];                  // conv2 [j][c]
  float w4t[C3 * C4];                  // conv4 [ch][o]
  float b2[C2], w1d[C1], b4[C4 + 4];
  float df[P_CLOUD];
  float c[C3];
  float g[C2];
  float red[8 * C2];
  float lat[LAT + 2];
  double pose[7], R[9], fin[6];
  unsigned char msk[P_CLOUD];
  unsigned long long mbar;
  unsigned int tmem_base;
  int count;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// two independent fp32 FMAs in one instruction (FFMA2): d = a * b + d, bit-identical to two fmaf
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a), bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  return d;                            // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// kind::tf32, fp32 accumulate, K-major A and B, M=128, N=256 (cute::UMMA::InstrDescriptor bit layout)
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C3 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}

__device__ __forceinline__ float tc_box_dist(double x, double y, double z, const double* lo, const double* hi) {
  double dx = fmax(fmax(lo[0] - x, x - hi[0]), 0.0), dy = fmax(fmax(lo[1] - y, y - hi[1]), 0.0),
         dz = fmax(fmax(lo[2] - z, z - hi[2]), 0.0);
  double outside = sqrt(dx * dx + dy * dy + dz * dz);
  double inside = fmin(fmin(fmin(x - lo[0], hi[0] - x), fmin(y - lo[1], hi[1] - y)), fmin(z - lo[2], hi[2] - z));
  return (float)(outside > 0 ? outside : fmax(inside, 0.0));
}

// conv1 + TC_CH of the 64 conv2 channels (group G selects which) for point p; weights broadcast from shared memory
template <int G>
__device__ __forceinline__ void point_h2_grp_t(const TcSmem& s, const CostDev& cd, int p, float df, float* out) {
  float h1[C1];
  const float4* a = reinterpret_cast<const float4*>(cd.a1 + (size_t)p * C1);
#pragma unroll
  for (int j = 0; j < C1 / 4; j++) {
    float4 v = __ldg(a + j);
    h1[4 * j + 0] = fmaxf(fmaf(s.w1d[4 * j + 0], df, v.x), 0.f);
    h1[4 * j + 1] = fmaxf(fmaf(s.w1d[4 * j + 1], df, v.y), 0.f);
    h1[4 * j + 2] = fmaxf(fmaf(s.w1d[4 * j + 2], df, v.z), 0.f);
    h1[4 * j + 3] = fmaxf(fmaf(s.w1d[4 * j + 3], df, v.w), 0.f);
  }
#pragma unroll
  for (int c = 0; c < TC_CH; c++) out[c] = s.b2[G * TC_CH + c];
#pragma unroll
  for (int j = 0; j < C1; j++) {
    const float4* w = reinterpret_cast<const float4*>(s.w2t + j * C2 + G * TC_CH);
#pragma unroll
    const float2 hh = make_float2(h1[j], h1[j]);
    for (int c4 = 0; c4 < TC_CH / 4; c4++) {
      float4 wv = w[c4];
      float2* o2 = reinterpret_cast<float2*>(out + 4 * c4);
      ffma2(o2[0], make_float2(wv.x, wv.y), hh);
      ffma2(o2[1], make_float2(wv.z, wv.w), hh);
    }
  }
}
__device__ __forceinline__ void point_h2_grp(const TcSmem& s, const CostDev& cd, int p, float df, int grp, float* out) {   // grp is warp-uniform
  switch (grp) {
    case 0: point_h2_grp_t<0>(s, cd, p, df, out); break;
    case 1: point_h2_grp_t<1>(s, cd, p, df, out); break;
#if TC_NS > 2
    case 2: point_h2_grp_t<2>(s, cd, p, df, out); break;
    default: point_h2_grp_t<3>(s, cd, p, df, out); break;
#else
    default: break;
#endif
  }
}
// +c, ReLU and the conv4 partial sums of 32 accumulator columns [cbase, cbase+32)
__device__ __forceinline__ void conv4_block(const TcSmem& s, int cbase, const uint32_t* v, float* out) {
#pragma unroll
  for (int j = 0; j < 32; j++) {
    float h = fmaxf(__uint_as_float(v[j]) + s.c[cbase + j], 0.f);
    const float4* w = reinterpret_cast<const float4*>(s.w4t + (cbase + j) * C4);
#pragma unroll
    const float2 hh = make_float2(h, h);
    for (int o4 = 0; o4 < C4 / 4; o4++) {
      float4 wv = w[o4];
      float2* o2 = reinterpret_cast<float2*>(out + 4 * o4);
      ffma2(o2[0], make_float2(wv.x, wv.y), hh);
      ffma2(o2[1], make_float2(wv.z, wv.w), hh);
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
pcn_tc_kernel(CostDev cd, const float* __restrict__ w3a_hi, const float* __restrict__ w3a_lo,
              const double* __restrict__ pose, const double* __restrict__ fins, int B, float* __restrict__ latent) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (TC_M - 1), grp = tid >> 7;        // TMEM lane / point row of this thread, column group
  const int P = cd.n_cloud;

  // ---- one-time setup: weights into shared memory (W3a in the canonical layout), TMEM allocation, mbarrier
  for (int i = tid; i < C2 * C3; i += TC_THREADS) { s.b_hi[i] = w3a_hi[i]; s.b_lo[i] = w3a_lo[i]; }
  for (int i = tid; i < C1 * C2; i += TC_THREADS) s.w2t[i] = cd.w2t[i];
  for (int i = tid; i < C3 * C4; i += TC_THREADS) { int ch = i / C4, o = i % C4; s.w4t[i] = cd.w4[o * C3 + ch]; }
  if (tid < C2) s.b2[tid] = cd.b2[tid];
  if (tid < C1) s.w1d[tid] = cd.w1d[tid];
  if (tid < C4) s.b4[tid] = cd.b4[tid];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = s.tmem_base;
  const uint32_t idesc = make_idesc();
  const uint32_t a_addr = smem_u32(s.a_buf), b_hi_addr = smem_u32(s.b_hi), b_lo_addr = smem_u32(s.b_lo);
  const uint32_t mbar = smem_u32(&s.mbar);
  uint32_t phase = 0;

  for (int e = blockIdx.x; e < B; e += gridDim.x) {
    // ---------------- stage 0: geometry (identical to pcn_kernel<true>)
    if (tid < 7) s.pose[tid] = pose[(size_t)e * 7 + tid];
    if (tid < 6) s.fin[tid] = fins[(size_t)e * 6 + tid];
    __syncthreads();
    if (tid == 0) {
      double x = s.pose[3], y = s.pose[4], z = s.pose[5], w = s.pose[6];
      double two_s = 2.0 / (w * w + x * x + y * y + z * z);
      s.R[0] = 1 - two_s * (y * y + z * z); s.R[1] = two_s * (x * y - z * w); s.R[2] = two_s * (x * z + y * w);
      s.R[3] = two_s * (x * y + z * w); s.R[4] = 1 - two_s * (x * x + z * z); s.R[5] = two_s * (y * z - x * w);
      s.R[6] = two_s * (x * z - y * w); s.R[7] = two_s * (y * z + x * w); s.R[8] = 1 - two_s * (x * x + y * y);
    }
    __syncthreads();
    int cnt = 0;
    for (int p = tid; p < P; p += TC_THREADS) {
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
      double qx = (double)(float)wx, qy = (double)(float)wy, qz = (double)(float)wz;
      float d = INFINITY;
      for (int b = 0; b < cd.n_dist_box; b++) d = fminf(d, tc_box_dist(qx, qy, qz, cd.dist_box[b][0], cd.dist_box[b][1]));
      if (cd.n_dist_tri > 0) d = tri_field(cd, qx, qy, qz, d);
      s.df[p] = d;
      s.msk[p] = keep ? 1 : 0;
      cnt += keep;
    }
    for (int f = 0; f < 2; f++) {
      double fx = s.fin[3 * f], fy = s.fin[3 * f + 1], fz = s.fin[3 * f + 2];
      double best = INFINITY;
      int bi = 0x7fffffff;
      for (int p = tid; p < P; p += TC_THREADS) {
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
      if (lane == 0) { rb[warp] = best; ri[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < TC_THREADS / 32; w++) if (rb[w] < best || (rb[w] == best && ri[w] < bi)) { best = rb[w]; bi = ri[w]; }
        bool ok = fz > cd.clearance_h;
        if (cd.project_uses_clearance) {
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
    // ---------------- stage 1: g = masked max over points of conv2 (this thread: 32 channels of its rows)
    float gmax[TC_CH];
#pragma unroll
    for (int c = 0; c < TC_CH; c++) gmax[c] = -INFINITY;
    for (int t0 = 0; t0 < P; t0 += TC_M) {
      int p = t0 + row;
      if (p < P && s.msk[p]) {
        float h2[TC_CH];
        point_h2_grp(s, cd, p, s.df[p], grp, h2);
#pragma unroll
        for (int c = 0; c < TC_CH; c++) gmax[c] = fmaxf(gmax[c], h2[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < TC_CH; c++) {
      float v = gmax[c];
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      gmax[c] = v;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if (lane == 0) {
      for (int c = 0; c < TC_CH; c++) s.red[warp * TC_CH + c] = gmax[c];     // the four warps of group g hold channels g*TC_CH ..
      reinterpret_cast<int*>(s.red + 8 * 32)[warp] = cnt;
    }
    __syncthreads();
    if (tid < C2) {
      int h = tid / TC_CH, c = tid % TC_CH;
      float v = s.red[(4 * h) * TC_CH + c];
      for (int w = 1; w < 4; w++) v = fmaxf(v, s.red[(4 * h + w) * TC_CH + c]);
      s.g[tid] = v;
    }
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < TC_THREADS / 32; w++) t += reinterpret_cast<int*>(s.red + 8 * 32)[w];
      s.count = t;
    }
    __syncthreads();
    if (s.count == 0) {
      if (tid < LAT) latent[(size_t)e * LAT + tid] = NAN;
      __syncthreads();
      continue;
    }
    if (tid < C3) {
      float v = cd.b3[tid];
      for (int k = 0; k < C2; k++) v = fmaf(__ldg(cd.w3bt + k * C3 + tid), s.g[k], v);
      s.c[tid] = v;
    }
    __syncthreads();

    // ---------------- stage 2: per 128-point tile: conv2 -> tcgen05 conv3 -> (+c, relu, conv4) -> masked max
    float rmax[C4];
#pragma unroll
    for (int o = 0; o < C4; o++) rmax[o] = -INFINITY;
    for (int t0 = 0; t0 < P; t0 += TC_M) {
      const int p = t0 + row;
      {
        float h2[TC_CH];
        if (p < P) point_h2_grp(s, cd, p, s.df[p], grp, h2);
        else {
#pragma unroll
          for (int c = 0; c < TC_CH; c++) h2[c] = 0.f;
        }
        // split a = hi + lo; hi goes to the operand buffer now, lo is kept in registers for the second phase
        float lo[TC_CH];
#pragma unroll
        for (int q = 0; q < TC_CH / 4; q++) {
          float4 hi;
          hi.x = tf32_rna(h2[4 * q]); hi.y = tf32_rna(h2[4 * q + 1]); hi.z = tf32_rna(h2[4 * q + 2]); hi.w = tf32_rna(h2[4 * q + 3]);
          lo[4 * q] = h2[4 * q] - hi.x; lo[4 * q + 1] = h2[4 * q + 1] - hi.y; lo[4 * q + 2] = h2[4 * q + 2] - hi.z; lo[4 * q + 3] = h2[4 * q + 3] - hi.w;
          *reinterpret_cast<float4*>(s.a_buf + ((grp * (TC_CH / 4) + q) * TC_M + row) * 4) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        const uint32_t lbo_a = TC_M * 16, lbo_b = C3 * 16, sbo = 128;
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
          for (int ks = 0; ks < C2 / 8; ks++) {                     // K = 8 tf32 per instruction = 2 chunks of 16 B
            uint64_t ah = make_desc(a_addr + ks * 2 * lbo_a, lbo_a, sbo);
            mma_tf32(tmem_d, ah, make_desc(b_hi_addr + ks * 2 * lbo_b, lbo_b, sbo), idesc, ks > 0 ? 1u : 0u);   // a_hi * b_hi
            mma_tf32(tmem_d, ah, make_desc(b_lo_addr + ks * 2 * lbo_b, lbo_b, sbo), idesc, 1u);                 // a_hi * b_lo
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        // second phase: the same buffer now carries a_lo
#pragma unroll
        for (int q = 0; q < TC_CH / 4; q++)
          *reinterpret_cast<float4*>(s.a_buf + ((grp * (TC_CH / 4) + q) * TC_M + row) * 4) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
          for (int ks = 0; ks < C2 / 8; ks++)
            mma_tf32(tmem_d, make_desc(a_addr + ks * 2 * lbo_a, lbo_a, sbo), make_desc(b_hi_addr + ks * 2 * lbo_b, lbo_b, sbo), idesc, 1u);  // a_lo * b_hi
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
      }
      // epilogue: this thread owns accumulator row `row`, columns grp*TC_COLS .. +TC_COLS
      float out[C4];
#pragma unroll
      for (int o = 0; o < C4; o++) out[o] = 0.f;
#pragma unroll
      for (int cb = 0; cb < TC_COLS / 32; cb++) {
        uint32_t v[32];
        uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(grp * TC_COLS + cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        conv4_block(s, grp * TC_COLS + cb * 32, v, out);
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();                                             // all TMEM reads done; A buffers free
      float* xch = s.a_buf;                                        // exchange the partial sums of groups 1..TC_NS-1
      if (grp > 0) {
#pragma unroll
        for (int o = 0; o < C4; o++) xch[((grp - 1) * C4 + o) * TC_M + row] = out[o];
      }
      __syncthreads();
      if (grp == 0) {
        float m = (p < P && s.msk[p]) ? 0.f : -INFINITY;
#pragma unroll
        for (int o = 0; o < C4; o++) {
          float acc = out[o];
#pragma unroll
          for (int gg = 0; gg < TC_NS - 1; gg++) acc += xch[(gg * C4 + o) * TC_M + row];
          rmax[o] = fmaxf(rmax[o], acc + s.b4[o] + m);
        }
      }
      __syncthreads();
    }
    // reduce rmax over the 128 rows (threads of half 0)
#pragma unroll
    for (int o = 0; o < C4; o++) {
      float v = rmax[o];
      for (int k = 16; k > 0; k >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, k));
      rmax[o] = v;
    }
    if (grp == 0 && lane == 0) for (int o = 0; o < C4; o++) s.red[warp * C4 + o] = rmax[o];
    __syncthreads();
    if (tid < C4) {
      float v = s.red[tid];
      for (int w = 1; w < 4; w++) v = fmaxf(v, s.red[w * C4 + tid]);
      s.lat[tid] = v;
    }
    __syncthreads();
    if (tid < LAT) latent[(size_t)e * LAT + tid] = s.lat[tid];
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(256));
}

// ---------------------------------------------------------------------------------------------------------
int launch_cost_tc(PgEnv* env, const double* pose, const double* fins, int B, float* latent, cudaStream_t st) {
  size_t smem = sizeof(TcSmem);
  // the opt-in is per device and an env may live on any device of the process: set it on every launch (host-side, cheap)
  PG_CUDA(cudaFuncSetAttribute(pcn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, env->device);
  int grid = sms < B ? sms : B;
  pcn_tc_kernel<<<grid, TC_THREADS, smem, st>>>(env->cost, env->w3a_hi, env->w3a_lo, pose, fins, B, latent);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
