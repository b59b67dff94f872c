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
