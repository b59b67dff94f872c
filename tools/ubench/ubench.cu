// Micro-benchmarks that size the rollout kernel's design space on the B200 this runs on ("measure an FFMA / DFMA
// micro-benchmark on the box and record it beside the others").  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
//   * dfma / ffma peak: W warps x ILP independent FMA chains per thread on every SM  -> TFLOP/s (the measured FP64 / FP32 peaks)
//   * dfma latency: one warp, one dependent chain, cycles per DFMA (clock64)
//   * dfma issue: one warp per SM sub-partition with ILP 1..8: what a SINGLE warp can extract from the FP64 pipe
//   * lds latency: dependent shared-memory pointer chase
//   * tmem scratch: tcgen05.st / tcgen05.ld round trip of per-thread data (correctness), dependent ld latency, ld/st throughput
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 123.456f) out[0] = s;
}

__global__ void dfma_latency_kernel(double* out, long long* cycles, int iters, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) x = fma(x, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  if (x == 123.456) out[0] = x;
}

// mixed chain as in a Gauss-Seidel row: dependent add -> compare -> select -> fma
__global__ void dchain_kernel(double* out, long long* cycles, int iters, double a, double b) {
  double x = threadIdx.x * 0.001, imp = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      double sum = imp + x;
      if (sum < 0.0) { x = 0.0 - imp; imp = 0.0; } else if (sum > 1e10) { x = 1e10 - imp; imp = 1e10; } else imp = sum;
      x = fma(x, a, b);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  if (x == 123.456) out[0] = x + imp;
}

__global__ void lds_latency_kernel(int* out, long long* cycles, int iters) {
  __shared__ int chain[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) chain[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) p = chain[p];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  if (p == -1) out[0] = p;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// TMEM as a per-thread scratchpad: lane (32*(warp%4) + lane) x 512 columns x 4 B = 2 KB per thread at 128 threads per CTA.
// mode 0: correctness (st all columns, ld back), mode 1: dependent st->ld latency, mode 2: ld.x32 throughput, mode 3: st.x32 throughput
__global__ void __launch_bounds__(128, 1) tmem_kernel(int mode, int iters, unsigned* errs, long long* cycles, double* sink) {
  __shared__ unsigned tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned base = tmem_base_s + ((unsigned)(warp * 32) << 16);     // lane field in bits 31:16
  unsigned bad = 0;
  long long t0 = 0, t1 = 0;
  double acc = 0;
  if (mode == 0) {
    for (int c = 0; c < 512; c += 2) {
      double v = (double)(threadIdx.x * 1000 + c) * 1.000000000001;
      unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(base + c), "r"(lo), "r"(hi) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int c = 0; c < 512; c += 2) {
      unsigned lo, hi;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(base + c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      double v = __hiloint2double((int)hi, (int)lo);
      if (v != (double)(threadIdx.x * 1000 + c) * 1.000000000001) bad++;
    }
    atomicAdd(errs, bad);
  } else if (mode == 1) {
    unsigned lo = threadIdx.x, hi = 7;
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(base + ((it * 2) & 510)), "r"(lo), "r"(hi) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(base + ((it * 2) & 510)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      lo += 1;
    }
    t1 = clock64();
    acc = lo + hi;
  } else if (mode == 4) {      // dependent ld only (address depends on the loaded value)
    for (int c = 0; c < 512; c += 2) {
      unsigned nxt = (c + 34) & 510, z = 0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(base + c), "r"(nxt), "r"(z) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    unsigned p = 0, hi;
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(p), "=r"(hi) : "r"(base + p) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    t1 = clock64();
    acc = p;
  } else if (mode == 2) {
    unsigned r[32];
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const unsigned col = (it * 32) & 480;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(base + col) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; i++) acc += r[i];
    }
    t1 = clock64();
  } else if (mode == 3) {
    unsigned r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = threadIdx.x + i;
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const unsigned col = (it * 32) & 480;
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
          :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
            "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
            "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(base + col) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    t1 = clock64();
  }
  if (threadIdx.x == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(512));
}

template <typename F>
static float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  double* dout; float* fout; long long* cyc; unsigned* errs; int* iout;
  CK(cudaMalloc(&dout, 64)); CK(cudaMalloc(&fout, 64)); CK(cudaMalloc(&cyc, 8 * 1024)); CK(cudaMalloc(&errs, 4)); CK(cudaMalloc(&iout, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %.0f", prop.name, sms, clock_khz / 1000.0);

  // ---- peaks: 8 warps per SM sub-partition, ILP 8
  {
    const int iters = 20000, threads = 1024, blocks = sms * 2;
    float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(dout, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 8 * 8 * (double)iters * threads * blocks;
    printf(", \"dfma_peak_tflops\": %.3f", fl / ms * 1e-9);
    float msf = time_ms([&] { ffma_kernel<8><<<blocks, threads>>>(fout, iters * 4, 1.0000001f, 1e-9f); });
    double flf = 2.0 * 8 * 8 * (double)iters * 4 * threads * blocks;
    printf(", \"ffma_peak_tflops\": %.3f", flf / msf * 1e-9);
  }
  // ---- what few warps can extract from the FP64 pipe: W warps per SM (1 CTA per SM), ILP chains per thread
  printf(", \"dfma_per_sm_clk\": {");
  {
    bool first = true;
    const int iters = 20000;
    for (int warps : {1, 2, 3, 4, 8, 14}) {
      auto run = [&](int ilp) {
        float ms = 0;
        if (ilp == 1) ms = time_ms([&] { dfma_kernel<1><<<sms, warps * 32>>>(dout, iters, 1.0000001, 1e-9); });
        if (ilp == 2) ms = time_ms([&] { dfma_kernel<2><<<sms, warps * 32>>>(dout, iters, 1.0000001, 1e-9); });
        if (ilp == 4) ms = time_ms([&] { dfma_kernel<4><<<sms, warps * 32>>>(dout, iters, 1.0000001, 1e-9); });
        if (ilp == 8) ms = time_ms([&] { dfma_kernel<8><<<sms, warps * 32>>>(dout, iters, 1.0000001, 1e-9); });
        double fmas = (double)ilp * 8 * iters * warps * 32;          // per SM
        return fmas / (ms * 1e-3 * 1.965e9);                        // DFMA lanes per clock per SM at 1965 MHz
      };
      for (int ilp : {1, 2, 4, 8}) { printf("%s\"w%d_ilp%d\": %.2f", first ? "" : ", ", warps, ilp, run(ilp)); first = false; }
    }
  }
  printf("}");
  // ---- latencies (cycles)
  {
    long long h;
    dfma_latency_kernel<<<1, 32>>>(dout, cyc, 4096, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dfma_latency_cycles\": %.2f", (double)h / (4096.0 * 16));
    dchain_kernel<<<1, 32>>>(dout, cyc, 4096, 0.999999, 1e-9); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"gs_clamp_chain_cycles_per_step\": %.2f", (double)h / (4096.0 * 8));
    lds_latency_kernel<<<1, 32>>>(iout, cyc, 4096); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"lds_latency_cycles\": %.2f", (double)h / (4096.0 * 16));
  }
  // ---- TMEM as scratch
  {
    unsigned zero = 0, e = 0; long long h;
    CK(cudaMemcpy(errs, &zero, 4, cudaMemcpyHostToDevice));
    tmem_kernel<<<sms, 128>>>(0, 0, errs, nullptr, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&e, errs, 4, cudaMemcpyDeviceToHost));
    printf(", \"tmem_roundtrip_mismatches\": %u", e);
    tmem_kernel<<<1, 128>>>(1, 4096, errs, cyc, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"tmem_st_wait_ld_wait_cycles\": %.1f", (double)h / 4096.0);
    tmem_kernel<<<1, 128>>>(4, 4096, errs, cyc, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"tmem_dependent_ld_cycles\": %.1f", (double)h / 4096.0);
    tmem_kernel<<<1, 128>>>(2, 4096, errs, cyc, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"tmem_ld_x32_4warps_bytes_per_clk_sm\": %.1f", 4096.0 * 128 * 128 / (double)h);
    tmem_kernel<<<1, 128>>>(3, 4096, errs, cyc, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"tmem_st_x32_4warps_bytes_per_clk_sm\": %.1f", 4096.0 * 128 * 128 / (double)h);
  }
  printf("}\n");
  return 0;
}
