// Kernel 2 (tensor-core version): fused per-step cost, every convolution of the point network on tcgen05.
//
// Same interface and geometry stage as pcn_kernel<true> in pg_cost.cu (see its header for the reference citations:
// envs/plate_contact_graph_env.py:393-437 crop / distance field / projection, neurals/network.py:398-412 PCN).
// Algebra per 128-point tile of one evaluation (exact identities, conv2 has no activation -- network.py:402-404):
//   h1  = relu(conv1)                                   CUDA cores (4 -> 32: xyz part precomputed per point, + w1d * df)
//   D2  = h1 . W2^T                 [128x32].[32x64]    tcgen05   -> global max-pool g (pass 1 over the cloud)
//   D3  = h1 . (W3a W2)^T           [128x32].[32x256]   tcgen05   conv3's point half with conv2 folded in:  W3a (W2 h1 + b2)
//   z   = relu(D3 + c),  c = b3 + W3a b2 + W3b g        epilogue on the TMEM rows; z is written straight back to TENSOR
//                                                       MEMORY as the A operand of the next GEMM (back-to-back fusion)
//   D4  = z . W4^T                  [128x256].[256x32]  tcgen05   (20 outputs padded to N = 32), accumulated over 8 K-chunks
//   lat = masked max over points of D4 + b4
// so all of the network's point-wise FLOPs (the ones SURVEY 8(d) counts: 164.1 MFLOP per evaluation) run on the tensor cores;
// the [K, P, 256] activations never leave the SM.
// fp32 fidelity on tf32 tensor cores: both operands are split a = a_hi + a_lo (a_hi = cvt.rna.tf32, a_lo = a - a_hi
// exactly) and three MMAs a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulate into the same TMEM tile (dropped term
// a_lo*b_lo <= 2^-22 relative), because MPPI weights exp(-5*100*dlogit) do not tolerate plain-tf32 logits.
// A operands (h1 tiles, z chunks) live in tensor memory (tcgen05.st by the producer warps, tcgen05.mma with [a_tmem]); the weight
// operands sit in shared memory in the no-swizzle K-major canonical layout ([k-chunk of 16 B][row][16 B]: SBO = 128 B between
// 8-row groups, LBO = rows*16 B between the two 16-byte halves of one K=8 step), hi / lo already in that layout, brought in once
// per CTA by TMA bulk copies (cp.async.bulk -> mbarrier).
#include "pg_internal.h"
#include "pg_cost_geom.cuh"

namespace pg {

#define TC_M 128
#define TC_KC 32                         // conv4 K-chunk: channels of z staged per back-to-back step
#define TC_N4 32                         // conv4 outputs padded to a legal MMA N
#ifndef TC_NG
#define TC_NG 4                          // column groups: threads per point row of a tile (the epilogues are latency bound: more warps)
#endif
#define TC_THREADS (TC_M * TC_NG)        // producer / epilogue threads
#define TC_W1 (C1 / TC_NG)               // conv1 channels staged per thread
#define TC_W2 (C2 / TC_NG)               // conv2 channels reduced per thread (pass 1)
#define TC_WC (TC_KC / TC_NG)            // z channels per thread and K-chunk
#define TC_W4 (TC_N4 / TC_NG)            // conv4 outputs per thread
#define TC_NB 3                          // A-operand buffers in tensor memory (64 columns each: 32 hi + 32 lo)
#define TC_NT (TC_THREADS + 32)           // + one warp whose lane 0 issues every tcgen05.mma

struct TcSmem {
  float w32_hi[C1 * C3], w32_lo[C1 * C3];   // (W3a W2)  [256 x 32]  canonical K-major: [(k/4)][n][k%4]
  float w4_hi[C3 * TC_N4], w4_lo[C3 * TC_N4];   // W4 (padded) [32 x 256]
  float w2_hi[C1 * C2], w2_lo[C1 * C2];     // W2        [64 x 32]
  float df[P_CLOUD];
  float c[C3];
  float g[C2];
  float b2[C2], w1d[C1], b4[TC_N4];
  float red[8 * C2];
  float lat[LAT + 2];
  double pose[7], R[9], fin[6];
  unsigned char msk[P_CLOUD];
  unsigned long long mbar_w, full[TC_NB], done[TC_NB];   // weights landed | A buffer b written by all producer warps | MMAs reading buffer b complete
  unsigned int tmem_base;
  int count;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// round-to-nearest (ties away) to tf32 = what cvt.rna.tf32.f32 returns for finite inputs, in two integer operations
// (the compiler expands the cvt with an inf / nan path that the epilogues issue 16 times per thread and K-chunk)
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

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

// tcgen05.mma with the A operand in TENSOR MEMORY (lane = row, one tf32 per 32-bit column) and B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// D[128 x N] (+)= A[128 x K] . B[N x K]^T with the 3x-split: per K = 8 step  a_hi b_hi + a_hi b_lo + a_lo b_hi.
// A: tensor memory, hi part in columns [a, a + K), lo part in [a + 32, a + 32 + K) (written by the producers with tcgen05.st: the
// operand never touches shared memory, no proxy fence, and the MMA's A read leaves the shared-memory bandwidth to B).
// B: rows = N, canonical K-major in shared memory.  Issued by ONE thread, which is on the critical path of every hand-off: the B
// descriptors are built once (OpDesc) and a K step only advances their 14-bit start-address field (address >> 4).
struct OpDesc { uint64_t hi, lo; };                                   // descriptors of the hi / lo part of one operand at K step 0
template <int ROWS>
__device__ __forceinline__ OpDesc make_opdesc(uint32_t hi_addr, uint32_t lo_addr) {
  OpDesc d;
  d.hi = make_desc(hi_addr, ROWS * 16, 128);
  d.lo = make_desc(lo_addr, ROWS * 16, 128);
  return d;
}
template <int N, int K>
__device__ __forceinline__ void gemm_split3(uint32_t tmem_d, uint32_t tmem_a, OpDesc b, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(N);
  constexpr uint64_t step_b = (2u * N * 16u) >> 4;                    // one K = 8 step = two 16-byte chunks of B
#pragma unroll
  for (int ks = 0; ks < K / 8; ks++) {
    mma_tf32_ts(tmem_d, tmem_a + 8 * ks, b.hi + ks * step_b, idesc, (ks > 0 || accumulate) ? 1u : 0u);
    mma_tf32_ts(tmem_d, tmem_a + 8 * ks, b.lo + ks * step_b, idesc, 1u);
    mma_tf32_ts(tmem_d, tmem_a + 32 + 8 * ks, b.hi + ks * step_b, idesc, 1u);
  }
}

__device__ __forceinline__ void mma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
// a producer warp publishes its part of an A buffer: its tcgen05.st have completed, its TMEM reads are ordered, then one arrival
// per warp (the barriers count the producer warps)
__device__ __forceinline__ void publish(uint32_t mbar, int lane) {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncwarp();
  if (lane == 0) mbar_arrive(mbar);
}
#define TC_PRODUCER_SYNC() asm volatile("bar.sync 1, %0;" :: "n"(TC_THREADS) : "memory")     // the producer warps only

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
#define TC_LD8(v, taddr)                                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                  \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr))
template <int W> __device__ __forceinline__ void tc_ld(uint32_t* v, uint32_t taddr) {
  if (W == 16) { TC_LD16(v, taddr); } else { TC_LD8(v, taddr); }
}
#define TC_ST8(taddr, v)                                                                                                        \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"                                   \
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory")
#define TC_LD_WAIT() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ float tc_box_dist(double x, double y, double z, const double* lo, const double* hi) {
  double dx = fmax(fmax(lo[0] - x, x - hi[0]), 0.0), dy = fmax(fmax(lo[1] - y, y - hi[1]), 0.0),
         dz = fmax(fmax(lo[2] - z, z - hi[2]), 0.0);
  double outside = sqrt(dx * dx + dy * dy + dz * dz);
  double inside = fmin(fmin(fmin(x - lo[0], hi[0] - x), fmin(y - lo[1], hi[1] - y)), fmin(z - lo[2], hi[2] - z));
  return (float)(outside > 0 ? outside : fmax(inside, 0.0));
}

// conv1 + ReLU for TC_W1 = 8 of the 32 channels of point p (thread (row, grp): channels 8 grp ..), split into tf32 hi / lo and
// stored into A buffer `abuf` (tensor memory, this thread's lane: hi in columns 8 grp .., lo 32 columns further); rows beyond the
// cloud are zero
__device__ __forceinline__ void stage_h1(const TcSmem& s, const CostDev& cd, int p, int P, int grp, uint32_t abuf_lane) {
  static_assert(TC_W1 == 8, "one x8 tensor-memory store per half");
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int j = 8 * grp + 4 * q;
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < P) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(cd.a1 + (size_t)p * C1 + j));
      const float df = s.df[p];
      h[0] = fmaxf(fmaf(s.w1d[j], df, v.x), 0.f); h[1] = fmaxf(fmaf(s.w1d[j + 1], df, v.y), 0.f);
      h[2] = fmaxf(fmaf(s.w1d[j + 2], df, v.z), 0.f); h[3] = fmaxf(fmaf(s.w1d[j + 3], df, v.w), 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float hh = tf32_rna(h[i]);
      hi[4 * q + i] = __float_as_uint(hh); lo[4 * q + i] = __float_as_uint(h[i] - hh);
    }
  }
  TC_ST8(abuf_lane + (uint32_t)(8 * grp), hi);
  TC_ST8(abuf_lane + (uint32_t)(32 + 8 * grp), lo);
}

__global__ void __launch_bounds__(TC_NT, 1)
pcn_tc_kernel(CostDev cd, TcWeights tw, const double* __restrict__ pose, const double* __restrict__ fins, int B, float* __restrict__ latent) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (TC_M - 1), grp = (tid >> 7) & (TC_NG - 1);  // producers: TMEM lane / point row of this thread, column group
  const bool issuer = warp == TC_THREADS / 32;             // the extra warp: lane 0 issues the MMAs, never touches operands or TMEM data
  const int P = cd.n_cloud;
  const uint32_t mbar_w = smem_u32(&s.mbar_w);
  const uint32_t full_b = smem_u32(&s.full[0]), done_b = smem_u32(&s.done[0]);     // barrier b at + 8 b

  // ---- one-time setup: TMEM allocation, mbarriers, weight operands by TMA bulk copies
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_w));
    for (int b = 0; b < TC_NB; b++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(full_b + 8 * b), "n"(TC_THREADS / 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(done_b + 8 * b));
    }
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
  // tensor-memory map (512 columns): D3 0..255 | D4 256..287 | A buffers 288 + 64 b (32 hi + 32 lo columns each); pass 1: D2[0|1] alias D3
  const uint32_t tm_d3 = tmem, tm_d4 = tmem + C3, tm_d2 = tmem, tm_a = tmem + C3 + TC_N4;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const OpDesc d_w32 = make_opdesc<C3>(smem_u32(s.w32_hi), smem_u32(s.w32_lo)), d_w2 = make_opdesc<C2>(smem_u32(s.w2_hi), smem_u32(s.w2_lo));
  const OpDesc d_w4 = make_opdesc<TC_N4>(smem_u32(s.w4_hi), smem_u32(s.w4_lo));
  // Producer / consumer protocol.  full[b]: every producer warp has written A buffer b (one arrival per warp); done[b]: the MMAs
  // that read buffer b have completed (tcgen05.commit by the issuer).  Each side keeps the parities of the barriers it waits on
  // (bit b of pf / pd); both walk the same sequence of tiles / K-chunks.
  uint32_t pf = 0, pd = 0;
#define TC_WAIT_FULL(b) { mbar_wait(full_b + 8 * (b), (pf >> (b)) & 1u); pf ^= 1u << (b); }
#define TC_WAIT_DONE(b) { mbar_wait(done_b + 8 * (b), (pd >> (b)) & 1u); pd ^= 1u << (b); }
  const int T = (P + TC_M - 1) / TC_M;

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
    for (int p = tid; p < P; p += TC_NT) {
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
      if (cd.n_dist_prism > 0) d = prism_field(cd, qx, qy, qz, d);
      s.df[p] = d;
      s.msk[p] = keep ? 1 : 0;
      cnt += keep;
    }
    for (int f = 0; f < 2; f++) {
      double fx = s.fin[3 * f], fy = s.fin[3 * f + 1], fz = s.fin[3 * f + 2];
      double best = INFINITY;
      int bi = 0x7fffffff;
      for (int p = tid; p < P; p += TC_NT) {
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
      int* ri = reinterpret_cast<int*>(s.red + 64);             // behind the (up to 17) doubles of rb
      __syncthreads();
      if (lane == 0) { rb[warp] = best; ri[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < TC_NT / 32; w++) if (rb[w] < best || (rb[w] == best && ri[w] < bi)) { best = rb[w]; bi = ri[w]; }
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
      for (int w = 0; w < TC_NT / 32; w++) t += reinterpret_cast<int*>(s.red)[w];
      s.count = t;
    }
    __syncthreads();
    if (s.count == 0) {
      if (tid < LAT) latent[(size_t)e * LAT + tid] = NAN;
      __syncthreads();
      continue;
    }

    if (issuer) {
      // ================= MMA issuer (one thread): waits for A buffers, feeds the tensor core, signals completion
      if (lane == 0) {
        for (int t = 0; t < T; t++) {                                  // pass 1: D2[t & 1] = h1 tile . W2^T, A buffer t & 1
          const int b = t & 1;
          TC_WAIT_FULL(b);
          asm volatile("tcgen05.fence::after_thread_sync;");
          gemm_split3<C2, C1>(tm_d2 + (uint32_t)(b * C2), tm_a + 64u * b, d_w2, 0u);
          mma_commit(done_b + 8 * b);
        }
        for (int t = 0; t < T; t++) {                                  // pass 2
          TC_WAIT_FULL(2);
          asm volatile("tcgen05.fence::after_thread_sync;");
          gemm_split3<C3, C1>(tm_d3, tm_a + 128u, d_w32, 0u);          // D3 = h1 tile (A buffer 2) . (W3a W2)^T
          mma_commit(done_b + 16);
#pragma unroll 1
          for (int ch = 0; ch < C3 / TC_KC; ch++) {                    // D4 += z chunk (A buffer ch % 3) . W4 chunk^T
            const int b = ch % TC_NB;
            TC_WAIT_FULL(b);
            asm volatile("tcgen05.fence::after_thread_sync;");
            // W4 chunk: K rows 32 ch .. of the [32 x 256] operand = byte offset (32 ch / 4) * (N * 16), in descriptor units (>> 4)
            const uint64_t boff = (uint64_t)((TC_KC * ch / 4) * (TC_N4 * 16) >> 4);
            OpDesc w4c; w4c.hi = d_w4.hi + boff; w4c.lo = d_w4.lo + boff;
            gemm_split3<TC_N4, TC_KC>(tm_d4, tm_a + 64u * b, w4c, ch > 0 ? 1u : 0u);
            mma_commit(done_b + 8 * b);
          }
        }
      }
      __syncwarp();
    } else {
      // ================= producers / epilogue (TC_NG warps per TMEM lane quadrant; thread (row, grp))
      // ---------------- pass 1: g = masked max over points of conv2 (N = 64); this thread: channels TC_W2 grp .. of its row.
      // Software pipeline over the tiles: tile t + 1 is staged (other A buffer, other accumulator region) before tile t is reduced.
      float gmax[TC_W2];
#pragma unroll
      for (int c = 0; c < TC_W2; c++) gmax[c] = -INFINITY;
      stage_h1(s, cd, row, P, grp, tm_a + lane_base);
      publish(full_b, lane);
      for (int t = 0; t < T; t++) {
        const int b = t & 1;
        if (t + 1 < T) {                 // A buffer 1 - b / D2[1 - b] were released when this thread finished tile t - 1
          stage_h1(s, cd, (t + 1) * TC_M + row, P, grp, tm_a + 64u * (1 - b) + lane_base);
          publish(full_b + 8 * (1 - b), lane);
        }
        TC_WAIT_DONE(b);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int p = t * TC_M + row;
        const bool keep = p < P && s.msk[p];
#pragma unroll
        for (int part = 0; part < TC_W2 / 16; part++) {
          uint32_t v[16];
          TC_LD16(v, tm_d2 + (uint32_t)(b * C2) + lane_base + (uint32_t)(TC_W2 * grp + 16 * part));
          TC_LD_WAIT();
          if (keep) {
#pragma unroll
            for (int i = 0; i < 16; i++) gmax[16 * part + i] = fmaxf(gmax[16 * part + i], __uint_as_float(v[i]) + s.b2[TC_W2 * grp + 16 * part + i]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < TC_W2; c++) {
        float v = gmax[c];
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        gmax[c] = v;
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < TC_W2; c++) s.red[warp * TC_W2 + c] = gmax[c];   // the four warps of group `grp` hold channels TC_W2 grp ..
      }
      TC_PRODUCER_SYNC();
      if (tid < C2) {
        const int h = tid / TC_W2, c = tid % TC_W2;
        float v = s.red[(4 * h) * TC_W2 + c];
        for (int w = 1; w < 4; w++) v = fmaxf(v, s.red[(4 * h + w) * TC_W2 + c]);
        s.g[tid] = v;
      }
      TC_PRODUCER_SYNC();
      if (tid < C3) {                                                  // c = (b3 + W3a b2) + W3b g
        float v = tw.cbase[tid];
#pragma unroll
        for (int k0 = 0; k0 < C2; k0 += 32) {                          // 32 loads in flight at a time (they were a dependent L2-latency chain)
          float wk[32];
#pragma unroll
          for (int k = 0; k < 32; k++) wk[k] = __ldg(cd.w3bt + (k0 + k) * C3 + tid);
#pragma unroll
          for (int k = 0; k < 32; k++) v = fmaf(wk[k], s.g[k0 + k], v);
        }
        s.c[tid] = v;
      }
      TC_PRODUCER_SYNC();

      // ---------------- pass 2: per 128-point tile  h1 -> [D3] -> z chunks -> [D4] -> masked max
      float rmax[TC_W4];                                               // conv4 outputs TC_W4 grp .. of this row (outputs >= 20 are padding)
#pragma unroll
      for (int o = 0; o < TC_W4; o++) rmax[o] = -INFINITY;
      for (int t = 0; t < T; t++) {
        const int p = t * TC_M + row;
        stage_h1(s, cd, p, P, grp, tm_a + 128u + lane_base);          // A buffer 2 is free: its last reader (chunk 5 of the previous tile) was waited for
        publish(full_b + 16, lane);
        TC_WAIT_DONE(2);                                               // D3 complete
        asm volatile("tcgen05.fence::after_thread_sync;");
        // back-to-back: z chunk ch (channels 32 ch ..) = relu(D3 + c) -> tensor memory (tf32 hi / lo) -> D4 += z_chunk . W4_chunk^T.
        // Three A buffers in flight: a buffer is rewritten only after the MMAs of three chunks earlier have completed.
#pragma unroll 1
        for (int ch = 0; ch < C3 / TC_KC; ch++) {
          const int b = ch % TC_NB;
          uint32_t v[TC_WC];
          tc_ld<TC_WC>(v, tm_d3 + lane_base + (uint32_t)(TC_KC * ch + TC_WC * grp));
          TC_LD_WAIT();
          uint32_t hi[TC_WC], lo[TC_WC];
#pragma unroll
          for (int q = 0; q < TC_WC / 4; q++) {
            const float4 cc = *reinterpret_cast<const float4*>(s.c + TC_KC * ch + TC_WC * grp + 4 * q);
            const float z[4] = {fmaxf(__uint_as_float(v[4 * q]) + cc.x, 0.f), fmaxf(__uint_as_float(v[4 * q + 1]) + cc.y, 0.f),
                                fmaxf(__uint_as_float(v[4 * q + 2]) + cc.z, 0.f), fmaxf(__uint_as_float(v[4 * q + 3]) + cc.w, 0.f)};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float zh = tf32_rna(z[i]);
              hi[4 * q + i] = __float_as_uint(zh); lo[4 * q + i] = __float_as_uint(z[i] - zh);
            }
          }
          if (ch >= TC_NB) TC_WAIT_DONE(b);
          static_assert(TC_WC == 8, "one x8 tensor-memory store per half");
          TC_ST8(tm_a + 64u * b + lane_base + (uint32_t)(TC_WC * grp), hi);
          TC_ST8(tm_a + 64u * b + lane_base + (uint32_t)(32 + TC_WC * grp), lo);
          publish(full_b + 8 * b, lane);
        }
        TC_WAIT_DONE(2);                                               // the last three chunks' MMAs (chunks 5, 6, 7: buffers 2, 0, 1)
        TC_WAIT_DONE(0);
        TC_WAIT_DONE(1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        {
          uint32_t v[TC_W4];
          tc_ld<TC_W4>(v, tm_d4 + lane_base + (uint32_t)(TC_W4 * grp));
          TC_LD_WAIT();
          const float m = (p < P && s.msk[p]) ? 0.f : -INFINITY;
#pragma unroll
          for (int o = 0; o < TC_W4; o++) rmax[o] = fmaxf(rmax[o], __uint_as_float(v[o]) + s.b4[TC_W4 * grp + o] + m);
        }
      }
      // reduce rmax over the 128 rows
#pragma unroll
      for (int o = 0; o < TC_W4; o++) {
        float v = rmax[o];
        for (int k = 16; k > 0; k >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, k));
        rmax[o] = v;
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      if (lane == 0) {
#pragma unroll
        for (int o = 0; o < TC_W4; o++) s.red[warp * TC_W4 + o] = rmax[o];
      }
      TC_PRODUCER_SYNC();
      if (tid < C4) {
        const int h = tid / TC_W4, o = tid % TC_W4;
        float v = s.red[(4 * h) * TC_W4 + o];
        for (int w = 1; w < 4; w++) v = fmaxf(v, s.red[(4 * h + w) * TC_W4 + o]);
        s.lat[tid] = v;
      }
      TC_PRODUCER_SYNC();
      if (tid < LAT) latent[(size_t)e * LAT + tid] = s.lat[tid];
    }
    __syncthreads();                                                   // producers and issuer meet; shared scratch is free for the next evaluation
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
  pcn_tc_kernel<<<grid, TC_NT, smem, st>>>(env->cost, env->tcw, pose, fins, B, latent);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
