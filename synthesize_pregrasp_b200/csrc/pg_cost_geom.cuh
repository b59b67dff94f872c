// Exact point-to-triangle distance shared by the two cost kernels (the waterbottle's tessellated holder cylinders,
// neurals/distancefield_utils.py:62-68).
#pragma once

namespace pg {

enum { TRI_GROUP = 16 };     // triangles per culling group (consecutive triangles of a tessellation are neighbours)

__device__ __noinline__ static double tri_dist2(double px, double py, double pz, const double* t) {
  // squared distance from p to triangle (a,b,c), Ericson "Real-Time Collision Detection" 5.1.5
  double ax = t[0], ay = t[1], az = t[2], bx = t[3], by = t[4], bz = t[5], cx = t[6], cy = t[7], cz = t[8];
  double abx = bx - ax, aby = by - ay, abz = bz - az, acx = cx - ax, acy = cy - ay, acz = cz - az;
  double apx = px - ax, apy = py - ay, apz = pz - az;
  double d1 = abx * apx + aby * apy + abz * apz, d2 = acx * apx + acy * apy + acz * apz;
  double qx, qy, qz;
  if (d1 <= 0 && d2 <= 0) { qx = ax; qy = ay; qz = az; }
  else {
    double bpx = px - bx, bpy = py - by, bpz = pz - bz;
    double d3 = abx * bpx + aby * bpy + abz * bpz, d4 = acx * bpx + acy * bpy + acz * bpz;
    if (d3 >= 0 && d4 <= d3) { qx = bx; qy = by; qz = bz; }
    else {
      double vc = d1 * d4 - d3 * d2;
      if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); qx = ax + v * abx; qy = ay + v * aby; qz = az + v * abz; }
      else {
        double cpx = px - cx, cpy = py - cy, cpz = pz - cz;
        double d5 = abx * cpx + aby * cpy + abz * cpz, d6 = acx * cpx + acy * cpy + acz * cpz;
        if (d6 >= 0 && d5 <= d6) { qx = cx; qy = cy; qz = cz; }
        else {
          double vb = d5 * d2 - d1 * d6;
          if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); qx = ax + w * acx; qy = ay + w * acy; qz = az + w * acz; }
          else {
            double va = d3 * d6 - d5 * d4;
            if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
              double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
              qx = bx + w * (cx - bx); qy = by + w * (cy - by); qz = bz + w * (cz - bz);
            } else {
              double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
              qx = ax + abx * v + acx * w; qy = ay + aby * v + acy * w; qz = az + abz * v + acz * w;
            }
          }
        }
      }
    }
  }
  double ex = px - qx, ey = py - qy, ez = pz - qz;
  return ex * ex + ey * ey + ez * ez;
}

// exact minimum distance (as float, like RaycastingScene.compute_distance) over the triangle soup, starting from the
// box distance d: triangles whose bounding sphere is farther than the best distance so far are skipped (conservative)
__device__ __forceinline__ float tri_field(const CostDev& cd, double qx, double qy, double qz, float d) {
  double best = INFINITY, ub = (double)d;
  const int G = (cd.n_dist_tri + TRI_GROUP - 1) / TRI_GROUP;
  for (int g = 0; g < G; g++) {                                   // two-level culling: groups of consecutive triangles, then triangles
    const double* gs = cd.tri_group_sphere + 4 * g;
    double gx = qx - gs[0], gy = qy - gs[1], gz = qz - gs[2], gthr = ub + gs[3];
    if (gx * gx + gy * gy + gz * gz > gthr * gthr * (1.0 + 1e-12)) continue;
    const int t1 = min(cd.n_dist_tri, (g + 1) * TRI_GROUP);
    for (int t = g * TRI_GROUP; t < t1; t++) {
      const double* sp = cd.tri_sphere + 4 * t;
      double dx = qx - sp[0], dy = qy - sp[1], dz = qz - sp[2], thr = ub + sp[3];
      if (dx * dx + dy * dy + dz * dz > thr * thr * (1.0 + 1e-12)) continue;
      double d2 = tri_dist2(qx, qy, qz, cd.dist_tri + 9 * t);
      if (d2 < best) { best = d2; ub = fmin(ub, sqrt(d2)); }
    }
  }
  return fminf(d, (float)sqrt(best));
}

// exact unsigned distance to the SURFACE of the right prisms (the tessellated cylinders of distancefield_utils.py:62-68 in
// closed form: caps and side strips of the tessellation are planar, so the triangle list and the prism are one surface).
// 2D distance to the polygon boundary (min over its edges), inside test by the edge cross products, then the axial part.
__device__ __forceinline__ float prism_field(const CostDev& cd, double qx, double qy, double qz, float d) {
  for (int i = 0; i < cd.n_dist_prism; i++) {
    const double x = qx - cd.prism_c[i][0], y = qy - cd.prism_c[i][1], z = qz - cd.prism_c[i][2];
    double best = INFINITY;
    bool inside = true;
    const int n = cd.prism_sides[i];
    for (int j = 0; j < n; j++) {
      const double ax = cd.prism_v[i][j][0], ay = cd.prism_v[i][j][1];
      const double ex = cd.prism_v[i][j + 1][0] - ax, ey = cd.prism_v[i][j + 1][1] - ay, wx = x - ax, wy = y - ay;
      const double t = fmin(1.0, fmax(0.0, (wx * ex + wy * ey) / (ex * ex + ey * ey)));
      const double dx = wx - t * ex, dy = wy - t * ey;
      best = fmin(best, dx * dx + dy * dy);
      inside = inside && (ex * wy - ey * wx >= 0);
    }
    const double dz = fabs(z) - cd.prism_hh[i];
    double dist;
    if (inside) dist = dz <= 0 ? fmin(sqrt(best), -dz) : dz;
    else dist = dz <= 0 ? sqrt(best) : sqrt(best + dz * dz);
    d = fminf(d, (float)dist);
  }
  return d;
}

}  // namespace pg
