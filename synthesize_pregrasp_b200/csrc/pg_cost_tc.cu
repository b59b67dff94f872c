// Kernel 2 (tensor-core version): fused per-step cost, every convolution of the point network on tcgen05.
//
// Same interface and geometry stage as pcn_kernel<true> in pg_cost.cu (see its header for the reference citations:
// envs/plate_contact_graph_env.py:393-437 crop / distance field / projection, neurals/network.py:398-412 PCN).
// Algebra per 128-point tile of one evaluation (exact identities, conv2 has no activation -- network.py:402-404):
//   h1  = relu(conv1)                                   CUDA cores (4 -> 32: xyz part precomputed per point, + w1d * df)
//   D2  = h1 . W2^T                 [128x32].[32x64]    tcgen05   -> global max-pool g (pass 1 over the cloud)
//   D3  = h1 . (W3a W2)^T           [128x32].[32x256]   tcgen05   conv3's point half with conv2 folded in:  W3a (W2 h1 + b2)
//   z   = relu(D3 + c),  c = b3 + W3a b2 + W3b g        epilogue on the TMEM rows; z is written straight back to shared
//                                                       memory as the A operand of the next GEMM (back-to-back fusion)
//   D4  = z . W4^T                  [128x256].[256x32]  tcgen05   (20 outputs padded to N = 32), accumulated over 8 K-chunks
//   lat = masked max over points of D4 + b4
// so all of the network's point-wise FLOPs (the ones SURVEY 8(d) counts: 164.1 MFLOP per evaluation) run on the tensor cores;
// the [K, P, 256] activations never leave the SM.
// fp32 fidelity on tf32 tensor cores: both operands are split a = a_hi + a_lo (a_hi = cvt.rna.tf32, a_lo = a - a_hi
// exactly) and three MMAs a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulate into the same TMEM tile (dropped term
// a_lo*b_lo <= 2^-22 relative), because MPPI weights exp(-5*100*dlogit) do not tolerate plain-tf32 logits.
// Operands sit in shared memory in the no-swizzle K-major canonical layout ([k-chunk of 16 B][row][16 B]:
// SBO = 128 B between 8-row groups, LBO = rows*16 B between the two 16-byte halves of one K=8 step).  The weight operands
// (hi / lo, already in that layout) are brought in once per CTA by TMA bulk copies (cp.async.bulk -> mbarrier).
#include "pg_internal.h"
#include "pg_cost_geom.cuh"

namespace pg {

#define TC_THREADS 256
#define TC_M 128
#define TC_KC 32                         // conv4 K-chunk: channels of z staged per back-to-back step
#define TC_N4 32                         // conv4 outputs padded to a legal MMA N

struct TcSmem {
  float w32_hi[C1 * C3], w32_lo[C1 * C3];   // (W3a W2)  [256 x 32]  canonical K-major: [(k/4)][n][k%4]
  float w4_hi[C3 * TC_N4], w4_lo[C3 * TC_N4];   // W4 (padded) [32 x 256]
  float w2_hi[C1 * C2], w2_lo[C1 * C2];     // W2        [64 x 32]
  float a1_hi[C1 * TC_M], a1_lo[C1 * TC_M]; // h1 tile   [128 x 32]  [(k/4)][row][k%4]
  float a4_hi[TC_KC * TC_M], a4_lo[TC_KC * TC_M];   // z chunk [128 x 32]
  float df[P_CLOUD];
  float c[C3];
  float g[C2];
  float b2[C2], w1d[C1], b4[TC_N4];
  float red[8 * C2];
  float lat[LAT + 2];
  double pose[7], R[9], fin[6];
  unsigned char msk[P_CLOUD];
  unsigned long long mbar, mbar_w, mbar_b[2];
  unsigned int tmem_base;
  int count;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// kind::tf32, fp32 accumulate, K-major A and B, M = 128, N = n (cute::UMMA::InstrDescriptor bit layout)
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// D[128 x N] (+)= A[128 x K] . B[N x K]^T with the 3x-split: per K = 8 step  a_hi b_hi + a_hi b_lo + a_lo b_hi.
// A: rows = TC_M, B: rows = N, both canonical K-major; issued by ONE thread, which is the critical path of every step that
// waits for the tensor core -- so the four operand descriptors are built ONCE (desc_* below) and a K step only advances their
// 14-bit start-address field (address >> 4; the operands stay far below the field's 256 KB range, so the add never carries).
struct OpDesc { uint64_t hi, lo; };                                   // descriptors of the hi / lo part of one operand at K step 0
template <int ROWS>
__device__ __forceinline__ OpDesc make_opdesc(uint32_t hi_addr, uint32_t lo_addr) {
  OpDesc d;
  d.hi = make_desc(hi_addr, ROWS * 16, 128);
  d.lo = make_desc(lo_addr, ROWS * 16, 128);
  return d;
}
template <int N, int K>
__device__ __forceinline__ void gemm_split3(uint32_t tmem_d, OpDesc a, OpDesc b, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(N);
  constexpr uint64_t step_a = (2u * TC_M * 16u) >> 4, step_b = (2u * N * 16u) >> 4;   // one K = 8 step = two 16-byte chunks
#pragma unroll
  for (int ks = 0; ks < K / 8; ks++) {
    mma_tf32(tmem_d, a.hi + ks * step_a, b.hi + ks * step_b, idesc, (ks > 0 || accumulate) ? 1u : 0u);
    mma_tf32(tmem_d, a.hi + ks * step_a, b.lo + ks * step_b, idesc, 1u);
    mma_tf32(tmem_d, a.lo + ks * step_a, b.hi + ks * step_b, idesc, 1u);
  }
}

__device__ __forceinline__ void mma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}

// TMA bulk copy global -> shared, completion on an mbarrier (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

#define TC_LD16(v, taddr)                                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),     \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                        \
               : "r"(taddr))
#define TC_LD_WAIT() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ float tc_box_dist(double x, double y, double z, const double* lo, const double* hi) {
  double dx = fmax(fmax(lo[0] - x, x - hi[0]), 0.0), dy = fmax(fmax(lo[1] - y, y - hi[1]), 0.0),
         dz = fmax(fmax(lo[2] - z, z - hi[2]), 0.0);
  double outside = sqrt(dx * dx + dy * dy + dz * dz);
  double inside = fmin(fmin(fmin(x - lo[0], hi[0] - x), fmin(y - lo[1], hi[1] - y)), fmin(z - lo[2], hi[2] - z));
  return (float)(outside > 0 ? outside : fmax(inside, 0.0));
}

// conv1 + ReLU for 16 of the 32 channels of point p (thread (row, grp): channels 16 grp ..), split into tf32 hi / lo and
// stored as the A operand of the next GEMM; rows beyond the cloud are zero
__device__ __forceinline__ void stage_h1(TcSmem& s, const CostDev& cd, int p, int P, int row, int grp, float* dst_hi, float* dst_lo) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int j = 16 * grp + 4 * q;
    float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
    if (p < P) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(cd.a1 + (size_t)p * C1 + j));
      const float df = s.df[p];
      const float h0 = fmaxf(fmaf(s.w1d[j], df, v.x), 0.f), h1 = fmaxf(fmaf(s.w1d[j + 1], df, v.y), 0.f);
      const float h2 = fmaxf(fmaf(s.w1d[j + 2], df, v.z), 0.f), h3 = fmaxf(fmaf(s.w1d[j + 3], df, v.w), 0.f);
      hi = make_float4(tf32_rna(h0), tf32_rna(h1), tf32_rna(h2), tf32_rna(h3));
      lo = make_float4(h0 - hi.x, h1 - hi.y, h2 - hi.z, h3 - hi.w);
    }
    const int o = ((j >> 2) * TC_M + row) * 4;
    *reinterpret_cast<float4*>(dst_hi + o) = hi;
    *reinterpret_cast<float4*>(dst_lo + o) = lo;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
pcn_tc_kernel(CostDev cd, TcWeights tw, const double* __restrict__ pose, const double* __restrict__ fins, int B, float* __restrict__ latent) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (TC_M - 1), grp = tid >> 7;        // TMEM lane / point row of this thread, column half
  const int P = cd.n_cloud;
  const uint32_t mbar = smem_u32(&s.mbar), mbar_w = smem_u32(&s.mbar_w);

  // ---- one-time setup: TMEM allocation, mbarriers, weight operands by TMA bulk copies
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_w));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s.mbar_b[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s.mbar_b[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    const uint32_t bytes = 4u * (2 * C1 * C3 + 2 * C3 * TC_N4 + 2 * C1 * C2);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_w), "r"(bytes) : "memory");
    bulk_g2s(s.w32_hi, tw.w32_hi, 4u * C1 * C3, mbar_w); bulk_g2s(s.w32_lo, tw.w32_lo, 4u * C1 * C3, mbar_w);
    bulk_g2s(s.w4_hi, tw.w4_hi, 4u * C3 * TC_N4, mbar_w); bulk_g2s(s.w4_lo, tw.w4_lo, 4u * C3 * TC_N4, mbar_w);
    bulk_g2s(s.w2_hi, tw.w2_hi, 4u * C1 * C2, mbar_w); bulk_g2s(s.w2_lo, tw.w2_lo, 4u * C1 * C2, mbar_w);
  }
  if (tid < C2) s.b2[tid] = cd.b2[tid];
  if (tid < C1) s.w1d[tid] = cd.w1d[tid];
  if (tid < TC_N4) s.b4[tid] = tid < C4 ? cd.b4[tid] : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  mbar_wait(mbar_w, 0);                                    // weights landed (async proxy writes, visible to the tensor core)
  const uint32_t tmem = s.tmem_base;
  const uint32_t tm_d3 = tmem, tm_d4 = tmem + C3, tm_d2 = tmem;          // D3: columns 0..255, D4: 256..287, D2 (pass 1) aliases D3
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const OpDesc d_a1 = make_opdesc<TC_M>(smem_u32(s.a1_hi), smem_u32(s.a1_lo)), d_a4 = make_opdesc<TC_M>(smem_u32(s.a4_hi), smem_u32(s.a4_lo));
  const OpDesc d_w32 = make_opdesc<C3>(smem_u32(s.w32_hi), smem_u32(s.w32_lo)), d_w2 = make_opdesc<C2>(smem_u32(s.w2_hi), smem_u32(s.w2_lo));
  const OpDesc d_w4 = make_opdesc<TC_N4>(smem_u32(s.w4_hi), smem_u32(s.w4_lo));
  const uint32_t mbar_b0 = smem_u32(&s.mbar_b[0]), mbar_b1 = smem_u32(&s.mbar_b[1]);
  uint32_t phase = 0, ph0 = 0, ph1 = 0;     // parities of mbar (D3 GEMM) and of the two operand-buffer barriers

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
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) reinterpret_cast<int*>(s.red)[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < TC_THREADS / 32; w++) t += reinterpret_cast<int*>(s.red)[w];
      s.count = t;
    }
    __syncthreads();
    if (s.count == 0) {
      if (tid < LAT) latent[(size_t)e * LAT + tid] = NAN;
      __syncthreads();
      continue;
    }

    // ---------------- pass 1: g = masked max over points of conv2 (tcgen05, N = 64); thread (row, grp): channels 32 grp ..
    float gmax[32];
#pragma unroll
    for (int c = 0; c < 32; c++) gmax[c] = -INFINITY;
    // software pipeline over the tiles: two operand buffers (a1, a4) and two accumulator regions (D2[0], D2[1]); the tensor
    // core works on tile t + 1 while the CUDA cores reduce tile t
    {
      const int T = (P + TC_M - 1) / TC_M;
      stage_h1(s, cd, row, P, row, grp, s.a1_hi, s.a1_lo);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        gemm_split3<C2, C1>(tm_d2, d_a1, d_w2, 0u);
        mma_commit(mbar_b0);
      }
      for (int t = 0; t < T; t++) {
        const int b = t & 1;
        if (t + 1 < T) {                                             // stage and launch tile t + 1 into the other buffer / region
          stage_h1(s, cd, (t + 1) * TC_M + row, P, row, grp, b ? s.a1_hi : s.a4_hi, b ? s.a1_lo : s.a4_lo);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;");
          __syncthreads();                                           // (also: everyone has read D2[1 - b] of tile t - 1)
          if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            gemm_split3<C2, C1>(tm_d2 + (b ? 0u : (uint32_t)C2), b ? d_a1 : d_a4, d_w2, 0u);
            mma_commit(b ? mbar_b0 : mbar_b1);
          }
        }
        if (b) { mbar_wait(mbar_b1, ph1); ph1 ^= 1; } else { mbar_wait(mbar_b0, ph0); ph0 ^= 1; }
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int p = t * TC_M + row;
        const bool keep = p < P && s.msk[p];
#pragma unroll
        for (int half = 0; half < 2; half++) {
          uint32_t v[16];
          TC_LD16(v, tm_d2 + (uint32_t)(b * C2) + lane_base + (uint32_t)(32 * grp + 16 * half));
          TC_LD_WAIT();
          if (keep) {
#pragma unroll
            for (int i = 0; i < 16; i++) gmax[16 * half + i] = fmaxf(gmax[16 * half + i], __uint_as_float(v[i]) + s.b2[32 * grp + 16 * half + i]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 32; c++) {
      float v = gmax[c];
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      gmax[c] = v;
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 32; c++) s.red[warp * 32 + c] = gmax[c];   // the four warps of half `grp` hold channels 32 grp ..
    }
    __syncthreads();
    if (tid < C2) {
      const int h = tid / 32, c = tid % 32;
      float v = s.red[(4 * h) * 32 + c];
      for (int w = 1; w < 4; w++) v = fmaxf(v, s.red[(4 * h + w) * 32 + c]);
      s.g[tid] = v;
    }
    __syncthreads();
    if (tid < C3) {                                                  // c = (b3 + W3a b2) + W3b g
      float wk[C2];                                                  // all 64 loads in flight at once (they were a dependent L2-latency chain)
#pragma unroll
      for (int k = 0; k < C2; k++) wk[k] = __ldg(cd.w3bt + k * C3 + tid);
      float v = tw.cbase[tid];
#pragma unroll
      for (int k = 0; k < C2; k++) v = fmaf(wk[k], s.g[k], v);
      s.c[tid] = v;
    }
    __syncthreads();

    // ---------------- pass 2: per 128-point tile  h1 -> [tcgen05 D3] -> z chunks -> [tcgen05 D4] -> masked max
    float rmax[16];                                                  // thread (row, grp): conv4 outputs 16 grp .. (outputs >= 20 are padding)
#pragma unroll
    for (int o = 0; o < 16; o++) rmax[o] = -INFINITY;
    for (int t0 = 0; t0 < P; t0 += TC_M) {
      const int p = t0 + row;
      stage_h1(s, cd, p, P, row, grp, s.a1_hi, s.a1_lo);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();                                               // (also: D3 / D4 of the previous tile have been read by everyone)
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        gemm_split3<C3, C1>(tm_d3, d_a1, d_w32, 0u);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase); phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;");
      // back-to-back: z chunk ch (channels 32 ch ..) = relu(D3 + c) -> shared memory (tf32 hi / lo) -> D4 += z_chunk . W4_chunk^T.
      // Two operand buffers (a4, and a1 -- free once D3 is complete): the tensor core works on chunk ch while the CUDA cores
      // prepare chunk ch + 1; a buffer is rewritten only after the MMAs of two chunks earlier have completed.
#pragma unroll 1
      for (int ch = 0; ch < C3 / TC_KC; ch++) {
        const int b = ch & 1;
        uint32_t v[16];
        TC_LD16(v, tm_d3 + lane_base + (uint32_t)(TC_KC * ch + 16 * grp));
        TC_LD_WAIT();
        float4 hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 cc = *reinterpret_cast<const float4*>(s.c + TC_KC * ch + 16 * grp + 4 * q);
          const float z0 = fmaxf(__uint_as_float(v[4 * q]) + cc.x, 0.f), z1 = fmaxf(__uint_as_float(v[4 * q + 1]) + cc.y, 0.f);
          const float z2 = fmaxf(__uint_as_float(v[4 * q + 2]) + cc.z, 0.f), z3 = fmaxf(__uint_as_float(v[4 * q + 3]) + cc.w, 0.f);
          hi[q] = make_float4(tf32_rna(z0), tf32_rna(z1), tf32_rna(z2), tf32_rna(z3));
          lo[q] = make_float4(z0 - hi[q].x, z1 - hi[q].y, z2 - hi[q].z, z3 - hi[q].w);
        }
        if (ch >= 2) { if (b) { mbar_wait(mbar_b1, ph1); ph1 ^= 1; } else { mbar_wait(mbar_b0, ph0); ph0 ^= 1; } }
        float* dh = b ? s.a1_hi : s.a4_hi;
        float* dl = b ? s.a1_lo : s.a4_lo;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int o = ((4 * grp + q) * TC_M + row) * 4;
          *reinterpret_cast<float4*>(dh + o) = hi[q];
          *reinterpret_cast<float4*>(dl + o) = lo[q];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;");
          // W4 chunk: K rows 32 ch .. of the [32 x 256] operand = byte offset (32 ch / 4) * (N * 16), in descriptor units (>> 4)
          const uint64_t boff = (uint64_t)((TC_KC * ch / 4) * (TC_N4 * 16) >> 4);
          OpDesc w4c; w4c.hi = d_w4.hi + boff; w4c.lo = d_w4.lo + boff;
          gemm_split3<TC_N4, TC_KC>(tm_d4, b ? d_a1 : d_a4, w4c, ch > 0 ? 1u : 0u);
          mma_commit(b ? mbar_b1 : mbar_b0);
        }
      }
      mbar_wait(mbar_b0, ph0); ph0 ^= 1;                             // the last two chunks' MMAs
      mbar_wait(mbar_b1, ph1); ph1 ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;");
      {
        uint32_t v[16];
        TC_LD16(v, tm_d4 + lane_base + (uint32_t)(16 * grp));
        TC_LD_WAIT();
        const float m = (p < P && s.msk[p]) ? 0.f : -INFINITY;
#pragma unroll
        for (int o = 0; o < 16; o++) rmax[o] = fmaxf(rmax[o], __uint_as_float(v[o]) + s.b4[16 * grp + o] + m);
      }
    }
    // reduce rmax over the 128 rows
#pragma unroll
    for (int o = 0; o < 16; o++) {
      float v = rmax[o];
      for (int k = 16; k > 0; k >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, k));
      rmax[o] = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 16; o++) s.red[warp * 16 + o] = rmax[o];
    }
    __syncthreads();
    if (tid < C4) {
      const int h = tid / 16, o = tid % 16;
      float v = s.red[(4 * h) * 16 + o];
      for (int w = 1; w < 4; w++) v = fmaxf(v, s.red[(4 * h + w) * 16 + o]);
      s.lat[tid] = v;
    }
    __syncthreads();
    if (tid < LAT) latent[(size_t)e * LAT + tid] = s.lat[tid];
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------
int launch_cost_tc(PgEnv* env, const double* pose, const double* fins, int B, float* latent, cudaStream_t st) {
  size_t smem = sizeof(TcSmem);
  // the opt-in is per device and an env may live on any device of the process: set it on every launch (host-side, cheap)
  PG_CUDA(cudaFuncSetAttribute(pcn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, env->device);
  int grid = sms < B ? sms : B;
  pcn_tc_kernel<<<grid, TC_THREADS, smem, st>>>(env->cost, env->tcw, pose, fins, B, latent);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
