// Kernel 4: geometric force/torque-closure test, utils/dyn_feasibility_check.py:124-235, one thread per item
// (a grasp with M<=8 contacts: first passing triple of C(M,3), :226-235) and the exhaustive C(P,3) sweep of
// find_dynamically_feasible_contacts :239-303 with the :260-265 pairwise-distance prefilter, enumerating the
// triple index space on the fly (no materialised triples in HBM) with on-device compaction.
#include "pg_internal.h"
#include "pg_qp.cuh"

namespace pg {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double a, double b, double c) { D3 r; r.x = a; r.y = b; r.z = c; return r; }
// explicit round-to-nearest mul/add (no FMA contraction): the sign tests below compare products that are exactly
// zero for axis-aligned contacts, so the arithmetic must be plain IEEE like the reference's numpy ufuncs
__device__ __forceinline__ D3 sub(D3 a, D3 b) { return d3(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)); }
__device__ __forceinline__ D3 scl(D3 a, double s) { return d3(__dmul_rn(a.x, s), __dmul_rn(a.y, s), __dmul_rn(a.z, s)); }
__device__ __forceinline__ double dt3(D3 a, D3 b) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z));
}
__device__ __forceinline__ D3 crs(D3 a, D3 b) {
  return d3(__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
            __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ double nrm3(D3 a) { return sqrt(dt3(a, a)); }

__device__ __forceinline__ D3 project2norm(D3 v, D3 n) {        // :124-125
  double l = nrm3(n);
  return scl(n, dt3(v, n) / (l * l));
}

// pk.quaternion_apply(pk.axis_angle_to_quaternion(aa), p)  (:191-196)
__device__ D3 rotate_axis_angle(D3 aa, D3 p) {
  double ang = nrm3(aa), half = 0.5 * ang, s;
  if (fabs(ang) < 1e-6) s = 0.5 - (ang * ang) / 48; else s = sin(half) / ang;
  double qw = cos(half), qx = __dmul_rn(aa.x, s), qy = __dmul_rn(aa.y, s), qz = __dmul_rn(aa.z, s);
  // q * (0,p)
#define M_(a, b) __dmul_rn(a, b)
#define A_(a, b) __dadd_rn(a, b)
#define S_(a, b) __dsub_rn(a, b)
  // pytorch3d quaternion_raw_multiply(q, (0,p)) then with conj(q); term order as in the published definition
  double aw = qw, ax = qx, ay = qy, az = qz, bw = 0.0, bx = p.x, by = p.y, bz = p.z;
  double tw = S_(S_(S_(M_(aw, bw), M_(ax, bx)), M_(ay, by)), M_(az, bz));
  double tx = S_(A_(A_(M_(aw, bx), M_(ax, bw)), M_(ay, bz)), M_(az, by));
  double ty = A_(A_(S_(M_(aw, by), M_(ax, bz)), M_(ay, bw)), M_(az, bx));
  double tz = A_(S_(A_(M_(aw, bz), M_(ax, by)), M_(ay, bx)), M_(az, bw));
  double cw = qw, cx = -qx, cy = -qy, cz = -qz;
  double ox = S_(A_(A_(M_(tw, cx), M_(tx, cw)), M_(ty, cz)), M_(tz, cy));
  double oy = A_(A_(S_(M_(tw, cy), M_(tx, cz)), M_(ty, cw)), M_(tz, cx));
  double oz = A_(S_(A_(M_(tw, cz), M_(tx, cy)), M_(ty, cx)), M_(tz, cw));
#undef M_
#undef A_
#undef S_
  return d3(ox, oy, oz);
}

__device__ __forceinline__ bool in_triangle(D3 a, D3 b, D3 c) {  // :127-136, point = origin, keeps the `c1<0 and c1<0` typo
  D3 o = d3(0, 0, 0);
  D3 r0 = crs(sub(b, a), sub(o, a)), r1 = crs(sub(c, b), sub(o, b)), r2 = crs(sub(a, c), sub(o, c));
  double c1 = dt3(r0, r1), c2 = dt3(r0, r2);
  return (c1 > 0 && c2 > 0) || (c1 < 0 && c1 < 0);
}

__device__ bool simple_3p(const D3* P, const D3* N) {            // :153-224
  const double MU = 5.0;
  D3 n = crs(sub(P[1], P[0]), sub(P[2], P[0]));
  double ln = nrm3(n);
  if (ln < 1e-4) return false;
  n = d3(n.x / ln, n.y / ln, n.z / ln);
  D3 pjn[3], pjp[3], cone[6];
  for (int i = 0; i < 3; i++) { pjn[i] = project2norm(N[i], n); pjp[i] = sub(N[i], pjn[i]); }
  for (int i = 0; i < 3; i++) if (nrm3(pjn[i]) / nrm3(pjp[i]) > MU) return false;
  for (int i = 0; i < 3; i++) {
    double ni = nrm3(N[i]);
    double temp = ni / nrm3(pjp[i]) * nrm3(pjn[i]);
    double angle = atan(sqrt(MU * MU - temp * temp) / sqrt(temp * temp + ni * ni));
    double lp = nrm3(pjp[i]);
    D3 p = d3(pjp[i].x / lp, pjp[i].y / lp, pjp[i].z / lp);
    cone[2 * i] = rotate_axis_angle(scl(n, angle), p);
    cone[2 * i + 1] = rotate_axis_angle(scl(n, -angle), p);
  }
  bool ok = false;
  for (int m = 0; m < 8 && !ok; m++) {
    // order of :201-216: (00,10,20) (00,11,20) (00,10,21) (00,11,21) (01,10,20) (01,11,20) (01,10,21) (01,11,21)
    int a = m >> 2, b = m & 1, c = (m >> 1) & 1;
    ok = in_triangle(cone[a], cone[2 + b], cone[4 + c]);
  }
  if (!ok) return false;
  // check_sign_change :138-151
  D3 center = d3((P[0].x + P[1].x + P[2].x) / 3, (P[0].y + P[1].y + P[2].y) / 3, (P[0].z + P[1].z + P[2].z) / 3);
  D3 prev = d3(0, 0, 0);
  bool have_prev = false;
  for (int i = 0; i < 3; i++) {
    D3 r = sub(P[i], center);
    D3 c1 = crs(r, cone[2 * i]), c2 = crs(r, cone[2 * i + 1]);
    if (!have_prev) { prev = c1; have_prev = true; }
    if (dt3(prev, c1) < 0 || dt3(c1, c2) < 0) return true;
    prev = c2;
  }
  return false;
}

__global__ void dyn_simple_kernel(const double* __restrict__ pts, const double* __restrict__ nrm, const int* __restrict__ count,
                                  int B, int M, uint8_t* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int m = count ? count[b] : M;
  const double* p = pts + (size_t)b * M * 3;
  const double* n = nrm + (size_t)b * M * 3;
  bool res = false;
  for (int i = 0; i < m && !res; i++)
    for (int j = i + 1; j < m && !res; j++)
      for (int k = j + 1; k < m && !res; k++) {
        D3 P[3] = {d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), d3(p[3 * j], p[3 * j + 1], p[3 * j + 2]), d3(p[3 * k], p[3 * k + 1], p[3 * k + 2])};
        D3 Nn[3] = {d3(n[3 * i], n[3 * i + 1], n[3 * i + 2]), d3(n[3 * j], n[3 * j + 1], n[3 * j + 2]), d3(n[3 * k], n[3 * k + 1], n[3 * k + 2])};
        res = simple_3p(P, Nn);
      }
  out[b] = res ? 1 : 0;
}

__device__ __forceinline__ unsigned long long c3(unsigned long long n) { return n < 3 ? 0ull : n * (n - 1) * (n - 2) / 6; }
__device__ __forceinline__ unsigned long long c2(unsigned long long n) { return n < 2 ? 0ull : n * (n - 1) / 2; }

// lexicographic rank -> (i<j<k), the order of itertools.combinations(range(P), 3)  (:293)
__device__ void unrank3(unsigned long long r, int P, int& i, int& j, int& k) {
  unsigned long long total = c3(P);
  int lo = 0, hi = P - 3;                      // largest i with (#triples whose first < i) <= r
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (total - c3(P - mid) <= r) lo = mid; else hi = mid - 1;
  }
  i = lo;
  r -= total - c3(P - i);
  int n2 = P - i - 1;                          // remaining pool for (j,k)
  unsigned long long t2 = c2(n2);
  lo = 0; hi = n2 - 2;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (t2 - c2(n2 - mid) <= r) lo = mid; else hi = mid - 1;
  }
  j = i + 1 + lo;
  r -= t2 - c2(n2 - lo);
  k = j + 1 + (int)r;
}

__global__ void dyn_sweep_kernel(const double* __restrict__ pts, const double* __restrict__ nrm, int P, double min_dist,
                                 unsigned long long first, unsigned long long last, int* __restrict__ triples,
                                 unsigned long long cap, unsigned long long* __restrict__ count) {
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long r = first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; r < last; r += stride) {
    int i, j, k;
    unrank3(r, P, i, j, k);
    D3 Pp[3] = {d3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), d3(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]), d3(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2])};
    if (nrm3(sub(Pp[0], Pp[1])) < min_dist || nrm3(sub(Pp[0], Pp[2])) < min_dist || nrm3(sub(Pp[1], Pp[2])) < min_dist) continue;
    D3 Nn[3] = {d3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), d3(nrm[3 * j], nrm[3 * j + 1], nrm[3 * j + 2]), d3(nrm[3 * k], nrm[3 * k + 1], nrm[3 * k + 2])};
    if (simple_3p(Pp, Nn)) {
      unsigned long long slot = atomicAdd(count, 1ull);
      if (triples && slot < cap) { triples[3 * slot] = i; triples[3 * slot + 1] = j; triples[3 * slot + 2] = k; }
    }
  }
}

// ---- wrench-closure QP criterion (check_dyn_feasible :77-112 and the sweep :239-303 as written, with the QP) -------------
__device__ __forceinline__ V3 ld3(const double* p, int i) { return V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

// one thread per grasp of M contacts: first triple (itertools.combinations order) whose objective is <= tol decides;
// normals are normalised per contact as :94 does.  obj_out (nullable) [B]: the smallest objective met before the decision.
__global__ void dyn_qp_kernel(const double* __restrict__ pts, const double* __restrict__ nrm, const int* __restrict__ count,
                              int B, int M, double mu, double f_min, double tol, uint8_t* __restrict__ out, double* __restrict__ obj_out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int m = count ? count[b] : M;
  const double* p = pts + (size_t)b * M * 3;
  const double* n = nrm + (size_t)b * M * 3;
  bool res = false;
  double best = INFINITY;
  for (int i = 0; i < m && !res; i++)
    for (int j = i + 1; j < m && !res; j++)
      for (int k = j + 1; k < m && !res; k++) {
        V3 P[3] = {ld3(p, i), ld3(p, j), ld3(p, k)}, Nn[3] = {ld3(n, i), ld3(n, j), ld3(n, k)};
        for (int q = 0; q < 3; q++) Nn[q] = Nn[q] / len(Nn[q]);
        double o = qp::wrench_closure_objective(P, Nn, mu, f_min, nullptr);
        if (o < best) best = o;
        res = o <= tol;
      }
  out[b] = res ? 1 : 0;
  if (obj_out) obj_out[b] = best;
}

__global__ void dyn_sweep_qp_kernel(const double* __restrict__ pts, const double* __restrict__ nrm, int P, double min_dist,
                                    double mu, double f_min, double tol, unsigned long long first, unsigned long long last,
                                    int* __restrict__ triples, unsigned long long cap, unsigned long long* __restrict__ count) {
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long r = first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; r < last; r += stride) {
    int i, j, k;
    unrank3(r, P, i, j, k);
    V3 Pp[3] = {ld3(pts, i), ld3(pts, j), ld3(pts, k)};
    if (len(Pp[0] - Pp[1]) < min_dist || len(Pp[0] - Pp[2]) < min_dist || len(Pp[1] - Pp[2]) < min_dist) continue;   // :259-265
    V3 Nn[3] = {ld3(nrm, i), ld3(nrm, j), ld3(nrm, k)};                                   // cloud normals used as given (:271)
    if (qp::wrench_closure_objective(Pp, Nn, mu, f_min, nullptr) <= tol) {
      unsigned long long slot = atomicAdd(count, 1ull);
      if (triples && slot < cap) { triples[3 * slot] = i; triples[3 * slot + 1] = j; triples[3 * slot + 2] = k; }
    }
  }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_dyn_feasible_simple(const double* pts_dev, const double* nrm_dev, const int32_t* count_dev, int B, int M,
                                      uint8_t* out_dev, void* stream) {
  if (!pts_dev || !nrm_dev || !out_dev || B < 0 || M < 3 || M > 8) { set_error("pg_dyn_feasible_simple: bad arguments"); return PG_ERR_INVALID; }
  if (B == 0) return PG_OK;
  dyn_simple_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pts_dev, nrm_dev, count_dev, B, M, out_dev);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

extern "C" int pg_dyn_sweep_simple(const double* cloud_pts_dev, const double* cloud_nrm_dev, int P, double min_dist,
                                   uint64_t first, uint64_t last, int32_t* triples_out_dev, uint64_t cap,
                                   unsigned long long* count_dev, void* stream) {
  if (!cloud_pts_dev || !cloud_nrm_dev || !count_dev || P < 3) { set_error("pg_dyn_sweep_simple: bad arguments"); return PG_ERR_INVALID; }
  uint64_t total = (uint64_t)P * (P - 1) * (P - 2) / 6;
  if (last > total) last = total;
  if (first >= last) return PG_OK;
  int dev = 0, sms = 148;
  PG_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint64_t n = last - first;
  int blocks = (int)((n + 127) / 128 < (uint64_t)sms * 16 ? (n + 127) / 128 : (uint64_t)sms * 16);
  dyn_sweep_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(cloud_pts_dev, cloud_nrm_dev, P, min_dist, first, last,
                                                            triples_out_dev, cap, count_dev);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

extern "C" int pg_dyn_feasible_qp(const double* pts_dev, const double* nrm_dev, const int32_t* count_dev, int B, int M, double mu,
                                  double f_min, double tol, uint8_t* out_dev, double* obj_out_dev, void* stream) {
  if (!pts_dev || !nrm_dev || !out_dev || B < 0 || M < 3 || M > 8 || !(mu > 0) || !(f_min > 0)) { set_error("pg_dyn_feasible_qp: bad arguments"); return PG_ERR_INVALID; }
  if (B == 0) return PG_OK;
  dyn_qp_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(pts_dev, nrm_dev, count_dev, B, M, mu, f_min, tol, out_dev, obj_out_dev);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

extern "C" int pg_dyn_sweep_qp(const double* cloud_pts_dev, const double* cloud_nrm_dev, int P, double min_dist, double mu, double f_min,
                               double tol, uint64_t first, uint64_t last, int32_t* triples_out_dev, uint64_t cap,
                               unsigned long long* count_dev, void* stream) {
  if (!cloud_pts_dev || !cloud_nrm_dev || !count_dev || P < 3 || !(mu > 0) || !(f_min > 0)) { set_error("pg_dyn_sweep_qp: bad arguments"); return PG_ERR_INVALID; }
  uint64_t total = (uint64_t)P * (P - 1) * (P - 2) / 6;
  if (last > total) last = total;
  if (first >= last) return PG_OK;
  int dev = 0, sms = 148;
  PG_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint64_t n = last - first;
  int blocks = (int)((n + 63) / 64 < (uint64_t)sms * 16 ? (n + 63) / 64 : (uint64_t)sms * 16);
  dyn_sweep_qp_kernel<<<blocks, 64, 0, (cudaStream_t)stream>>>(cloud_pts_dev, cloud_nrm_dev, P, min_dist, mu, f_min, tol, first, last,
                                                              triples_out_dev, cap, count_dev);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}
