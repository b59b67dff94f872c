// Small fixed-size fp64 linear algebra used by the rollout kernels (device + host-test builds).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#define PG_HDN static __host__ __device__ __noinline__
#else
#define PG_HD inline
#define PG_HDN static
#endif

namespace pg {

struct V3 {
  double x, y, z;
  PG_HD V3() : x(0), y(0), z(0) {}
  PG_HD V3(double a, double b, double c) : x(a), y(b), z(c) {}
  PG_HD double& operator[](int i) { return (&x)[i]; }
  PG_HD double operator[](int i) const { return (&x)[i]; }
};
PG_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
PG_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
PG_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
PG_HD V3 operator*(V3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
PG_HD V3 operator*(double s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
PG_HD V3 operator/(V3 a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
PG_HD V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
PG_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PG_HD V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PG_HD double len(V3 a) { return sqrt(dot(a, a)); }
PG_HD V3 mulc(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
PG_HD V3 divc(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }

struct Q4 { double x, y, z, w; };
PG_HD Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}

// pybullet.multiplyTransforms(posA, ornA, posB, identity)[0] exactly as PyBullet computes it: b3Transform with b3Scalar = FLOAT
// (b3MultiplyTransforms; the extension has double-precision Bullet but single-precision Bullet3Common).  The env feeds these
// float32 results into the physics -- fingertip teleport position / constraint pivot (plate_env:462-474) and the servo targets
// (:526-530) -- and the recorded fingertip positions are reproduced bit for bit only this way.  Every operation is a single
// IEEE float operation in b3's order (no contraction: intrinsics on the device, -ffp-contract=off in the host test build).
#if defined(__CUDA_ARCH__)
#define PG_FMUL(a, b) __fmul_rn((a), (b))
#define PG_FADD(a, b) __fadd_rn((a), (b))
#define PG_FSUB(a, b) __fsub_rn((a), (b))
#define PG_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define PG_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define PG_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define PG_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define PG_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#endif
PG_HD V3 b3_multiply_transforms(V3 posA, Q4 ornA, V3 posB) {
  const float x = (float)ornA.x, y = (float)ornA.y, z = (float)ornA.z, w = (float)ornA.w;
  const float d = PG_FADD(PG_FADD(PG_FADD(PG_FMUL(x, x), PG_FMUL(y, y)), PG_FMUL(z, z)), PG_FMUL(w, w));   // b3Quaternion::length2
  const float s = PG_FDIV(2.0f, d);                                                                         // b3Matrix3x3::setRotation
  const float xs = PG_FMUL(x, s), ys = PG_FMUL(y, s), zs = PG_FMUL(z, s);
  const float wx = PG_FMUL(w, xs), wy = PG_FMUL(w, ys), wz = PG_FMUL(w, zs);
  const float xx = PG_FMUL(x, xs), xy = PG_FMUL(x, ys), xz = PG_FMUL(x, zs);
  const float yy = PG_FMUL(y, ys), yz = PG_FMUL(y, zs), zz = PG_FMUL(z, zs);
  const float r00 = PG_FSUB(1.0f, PG_FADD(yy, zz)), r01 = PG_FSUB(xy, wz), r02 = PG_FADD(xz, wy);
  const float r10 = PG_FADD(xy, wz), r11 = PG_FSUB(1.0f, PG_FADD(xx, zz)), r12 = PG_FSUB(yz, wx);
  const float r20 = PG_FSUB(xz, wy), r21 = PG_FADD(yz, wx), r22 = PG_FSUB(1.0f, PG_FADD(xx, yy));
  const float v0 = (float)posB.x, v1 = (float)posB.y, v2 = (float)posB.z;
  // b3Transform::operator(): dot3 (products summed left to right) + origin
  const float o0 = PG_FADD(PG_FADD(PG_FADD(PG_FMUL(v0, r00), PG_FMUL(v1, r01)), PG_FMUL(v2, r02)), (float)posA.x);
  const float o1 = PG_FADD(PG_FADD(PG_FADD(PG_FMUL(v0, r10), PG_FMUL(v1, r11)), PG_FMUL(v2, r12)), (float)posA.y);
  const float o2 = PG_FADD(PG_FADD(PG_FADD(PG_FMUL(v0, r20), PG_FMUL(v1, r21)), PG_FMUL(v2, r22)), (float)posA.z);
  return V3((double)o0, (double)o1, (double)o2);
}

PG_HD Q4 qnormalize(Q4 q) {
  double l = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  Q4 r; r.x = q.x / l; r.y = q.y / l; r.z = q.z / l; r.w = q.w / l;
  return r;
}

// 3x3, row major; columns are the body axes
struct M3 {
  double m[9];
  PG_HD double operator()(int i, int j) const { return m[3 * i + j]; }
  PG_HD double& operator()(int i, int j) { return m[3 * i + j]; }
  PG_HD V3 row(int i) const { return V3(m[3 * i], m[3 * i + 1], m[3 * i + 2]); }
  PG_HD V3 col(int j) const { return V3(m[j], m[3 + j], m[6 + j]); }
};
PG_HD M3 qmat(Q4 q) {   // same expression tree as btMatrix3x3::setRotation
  double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  double s = 2.0 / d;
  double xs = q.x * s, ys = q.y * s, zs = q.z * s;
  double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
  double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
  double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
  M3 r;
  r.m[0] = 1.0 - (yy + zz); r.m[1] = xy - wz; r.m[2] = xz + wy;
  r.m[3] = xy + wz; r.m[4] = 1.0 - (xx + zz); r.m[5] = yz - wx;
  r.m[6] = xz - wy; r.m[7] = yz + wx; r.m[8] = 1.0 - (xx + yy);
  return r;
}
PG_HD V3 mul(const M3& a, V3 v) { return V3(dot(a.row(0), v), dot(a.row(1), v), dot(a.row(2), v)); }
PG_HD V3 mulT(const M3& a, V3 v) { return V3(dot(a.col(0), v), dot(a.col(1), v), dot(a.col(2), v)); }

PG_HD void plane_space1(V3 n, V3& p, V3& q) {
  const double SQRT12 = 0.7071067811865475244008443621048490;
  if (fabs(n.z) > SQRT12) {
    double a = n.y * n.y + n.z * n.z;
    double k = 1.0 / sqrt(a);
    p = V3(0, -n.z * k, n.y * k);
    q = V3(a * k, -n.x * p.z, n.x * p.y);
  } else {
    double a = n.x * n.x + n.y * n.y;
    double k = 1.0 / sqrt(a);
    p = V3(-n.y * k, n.x * k, 0);
    q = V3(-n.z * p.y, n.z * p.x, a * k);
  }
}

}  // namespace pg
