// Camera model: world -> image projection with rational radial (k1..k6) and tangential (p1, p2)
// distortion, and its inverse.  Follows reference camera.py:591-663, 1138-1264, 1305-1337,
// 1435-1519 operation by operation (unfused) so that the only rounding difference to the reference
// is the 3x3 rotation product, which NumPy hands to BLAS.
#pragma once
#include "common.cuh"

namespace gb {

// Distortion of camera coordinates (camera.py:1180-1196), zero coefficients skipped exactly as
// camera.py:1147-1161 does.
__device__ __forceinline__ void distort(const gb_camera& c, double x, double y, double& xd, double& yd) {
  const bool any_k = (c.k[0] != 0.0) | (c.k[1] != 0.0) | (c.k[2] != 0.0) | (c.k[3] != 0.0) | (c.k[4] != 0.0) |
                     (c.k[5] != 0.0);
  const bool any_p = (c.p[0] != 0.0) | (c.p[1] != 0.0);
  xd = x;
  yd = y;
  if (!any_k && !any_p) return;
  const double r2 = add(mul(x, x), mul(y, y));
  if (any_k) {
    double dr = 1.0;
    if (c.k[0] != 0.0) dr = add(dr, mul(c.k[0], r2));
    if (c.k[1] != 0.0) dr = add(dr, mul(mul(c.k[1], r2), r2));
    if (c.k[2] != 0.0) dr = add(dr, mul(mul(mul(c.k[2], r2), r2), r2));
    if ((c.k[3] != 0.0) | (c.k[4] != 0.0) | (c.k[5] != 0.0)) {
      double den = 1.0;
      if (c.k[3] != 0.0) den = add(den, mul(c.k[3], r2));
      if (c.k[4] != 0.0) den = add(den, mul(mul(c.k[4], r2), r2));
      if (c.k[5] != 0.0) den = add(den, mul(mul(mul(c.k[5], r2), r2), r2));
      dr = quo(dr, den);
    }
    xd = mul(xd, dr);
    yd = mul(yd, dr);
  }
  if (any_p) {
    const double xty = mul(x, y);
    const double two_xty = mul(2.0, xty);
    const double tx = add(mul(two_xty, c.p[0]), mul(c.p[1], add(r2, mul(2.0, mul(x, x)))));
    const double ty = add(mul(c.p[0], add(r2, mul(2.0, mul(y, y)))), mul(two_xty, c.p[1]));
    xd = add(xd, tx);
    yd = add(yd, ty);
  }
}

__device__ __forceinline__ void project_direction(const gb_camera& c, double dx, double dy, double dz, double& u, double& v);

// Camera.xyz_to_uv (camera.py:591-628): NaN behind the camera (camera.py:1466-1467).
__device__ __forceinline__ void project(const gb_camera& c, double px, double py, double pz, double& u, double& v) {
  if (c.affine) {  // raster observer (raster.py:423-445): (xy - (xlim[0], ylim[0])) / d, z unused
    u = quo(sub(px, c.xyz[0]), c.f[0]);
    v = quo(sub(py, c.xyz[1]), c.f[1]);
    return;
  }
  const double dx = sub(px, c.xyz[0]);
  const double dy = sub(py, c.xyz[1]);
  double dz = sub(pz, c.xyz[2]);
  if (c.has_corr) {
    // helpers.elevation_corrections (helpers.py:1790): (refraction - 1) * d2 / (2 * radius)
    const double d2 = add(mul(dx, dx), mul(dy, dy));
    dz = add(dz, quo(mul(c.corr_c1, d2), c.corr_c2));
  }
  project_direction(c, dx, dy, dz, u, v);
}

// Camera.xyz_to_uv(directions=True) (camera.py:1448-1449): a ray from the camera, no translation and no correction.
__device__ __forceinline__ void project_direction(const gb_camera& c, double dx, double dy, double dz, double& u, double& v) {
  const double xc = fma(c.R[2], dz, fma(c.R[1], dy, c.R[0] * dx));
  const double yc = fma(c.R[5], dz, fma(c.R[4], dy, c.R[3] * dx));
  const double zc = fma(c.R[8], dz, fma(c.R[7], dy, c.R[6] * dx));
  double x = quo(xc, zc), y = quo(yc, zc);
  if (zc <= 0.0) {
    x = CUDART_NAN;
    y = CUDART_NAN;
  }
  double xd, yd;
  distort(c, x, y, xd, yd);
  u = add(mul(xd, c.f[0]), c.cc[0]);
  v = add(mul(yd, c.f[1]), c.cc[1]);
}

// Camera + distortion flags as the step kernel receives it (kernel-parameter / constant bank).
struct CamK {
  gb_camera c;
  int32_t any_k, any_p, has_den, pad_;
};

__host__ inline void camk_from(const gb_camera& c, CamK& out) {
  out.c = c;
  out.any_k = (c.k[0] != 0.0) || (c.k[1] != 0.0) || (c.k[2] != 0.0) || (c.k[3] != 0.0) || (c.k[4] != 0.0) || (c.k[5] != 0.0);
  out.has_den = (c.k[3] != 0.0) || (c.k[4] != 0.0) || (c.k[5] != 0.0);
  out.any_p = (c.p[0] != 0.0) || (c.p[1] != 0.0);
  out.pad_ = 0;
}

// Same mapping as project(), arranged for throughput in the fused step kernel: one division
// instead of three and fused multiply-adds.  With s = xc^2 + yc^2 and z2 = zc^2 the rational radial
// factor is N / D, N = z2^3 + k1 s z2^2 + k2 s^2 z2 + k3 s^3 (D likewise with k4..k6), so
//   x' = xc N / (zc D) + tx / z2,   q = 1 / (z2 D)  ->  1 / (zc D) = q zc,  1 / z2 = q D.
// Differs from project() by a few ulp (~1e-12 px at map scale; tests/test_gpu_track.py bounds it).
__device__ __forceinline__ void project_fast(const CamK& k, double px, double py, double pz, double& u, double& v) {
  const gb_camera& c = k.c;
  const double dx = px - c.xyz[0], dy = py - c.xyz[1];
  if (c.affine) {  // raster observer: (xy - origin) / cell size, as the reference divides (raster.py:445)
    u = dx / c.f[0];
    v = dy / c.f[1];
    return;
  }
  double dz = pz - c.xyz[2];
  if (c.has_corr) dz = add(dz, quo(mul(c.corr_c1, add(mul(dx, dx), mul(dy, dy))), c.corr_c2));
  const double xc = fma(c.R[2], dz, fma(c.R[1], dy, c.R[0] * dx));
  const double yc = fma(c.R[5], dz, fma(c.R[4], dy, c.R[3] * dx));
  const double zc = fma(c.R[8], dz, fma(c.R[7], dy, c.R[6] * dx));
  double xd, yd;
  if (!k.any_k && !k.any_p) {
    const double iz = 1.0 / zc;
    xd = xc * iz;
    yd = yc * iz;
  } else {
    const double z2 = zc * zc, s = fma(xc, xc, yc * yc);
    const double s2 = s * s, z4 = z2 * z2;
    const double s3 = s2 * s, z6 = z4 * z2;
    const double sz4 = s * z4, s2z2 = s2 * z2;
    const double Nn = fma(c.k[2], s3, fma(c.k[1], s2z2, fma(c.k[0], sz4, z6)));
    const double Dd = k.has_den ? fma(c.k[5], s3, fma(c.k[4], s2z2, fma(c.k[3], sz4, z6))) : z6;
    const double q = 1.0 / (z2 * Dd);
    const double izd = q * zc * Nn;  // radial factor / zc
    xd = xc * izd;
    yd = yc * izd;
    if (k.any_p) {
      const double iz2 = q * Dd;
      const double xy2 = 2.0 * xc * yc;
      xd = fma(fma(c.p[1], fma(2.0 * xc, xc, s), xy2 * c.p[0]), iz2, xd);
      yd = fma(fma(c.p[0], fma(2.0 * yc, yc, s), xy2 * c.p[1]), iz2, yd);
    }
  }
  u = fma(xd, c.f[0], c.cc[0]);
  v = fma(yd, c.f[1], c.cc[1]);
  if (!(zc > 0.0)) {
    u = CUDART_NAN;
    v = CUDART_NAN;
  }
}

// Inverse distortion (camera.py:1198-1264, 1305-1337).
__device__ inline void undistort(const gb_camera& c, double x, double y, double& xu, double& yu) {
  const bool any_k = (c.k[0] != 0.0) | (c.k[1] != 0.0) | (c.k[2] != 0.0) | (c.k[3] != 0.0) | (c.k[4] != 0.0) |
                     (c.k[5] != 0.0);
  const bool any_p = (c.p[0] != 0.0) | (c.p[1] != 0.0);
  xu = x;
  yu = y;
  if (!any_k && !any_p) return;
  const bool only_k1 = (c.k[0] != 0.0) && !((c.k[1] != 0.0) | (c.k[2] != 0.0) | (c.k[3] != 0.0) | (c.k[4] != 0.0) |
                                          (c.k[5] != 0.0)) && !any_p;
  if (only_k1) {
    // closed-form cubic r^3 + r / k1 - r' / k1 = 0 (camera.py:1232-1264)
    const double k1 = c.k[0];
    const double phi = atan2(y, x);
    const double cphi = cos(phi), sphi = sin(phi);
    const double Q = -1.0 / (3.0 * k1);
    const double R = -x / (2.0 * k1 * cphi);
    const double Q3 = Q * Q * Q;  // numpy: Q ** 3
    double r;
    if (R * R < Q3) {
      const double th = acos(R * pow(Q, -1.5));
      r = -2.0 * sqrt(Q) * cos((th - 2.0 * CUDART_PI) / 3.0);
    } else {
      const double sgn = (R > 0.0) ? 1.0 : ((R < 0.0) ? -1.0 : 0.0);
      const double A = -sgn * pow(fabs(R) + sqrt(R * R - Q3), 1.0 / 3.0);
      const double B = (A != 0.0) ? Q / A : 0.0;
      r = A + B;
    }
    xu = cphi * r;
    yu = sphi * r;
    return;
  }
  // 20 fixed-point iterations, no early exit (camera.py:1305-1337; tolerance = 0)
  double ex = x, ey = y;
  for (int it = 0; it < 20; ++it) {
    const double r2 = add(mul(ex, ex), mul(ey, ey));
    const double xty = mul(ex, ey);
    const double two_xty = mul(2.0, xty);
    const double tx = add(mul(two_xty, c.p[0]), mul(c.p[1], add(r2, mul(2.0, mul(ex, ex)))));
    const double ty = add(mul(c.p[0], add(r2, mul(2.0, mul(ey, ey)))), mul(two_xty, c.p[1]));
    if (any_p && !any_k) {
      ex = sub(x, tx);
      ey = sub(y, ty);
    } else {
      double dr = 1.0;
      if (c.k[0] != 0.0) dr = add(dr, mul(c.k[0], r2));
      if (c.k[1] != 0.0) dr = add(dr, mul(mul(c.k[1], r2), r2));
      if (c.k[2] != 0.0) dr = add(dr, mul(mul(mul(c.k[2], r2), r2), r2));
      if ((c.k[3] != 0.0) | (c.k[4] != 0.0) | (c.k[5] != 0.0)) {
        double den = 1.0;
        if (c.k[3] != 0.0) den = add(den, mul(c.k[3], r2));
        if (c.k[4] != 0.0) den = add(den, mul(mul(c.k[4], r2), r2));
        if (c.k[5] != 0.0) den = add(den, mul(mul(mul(c.k[5], r2), r2), r2));
        dr = quo(dr, den);
      }
      const double inv = quo(1.0, dr);
      ex = mul(sub(x, tx), inv);
      ey = mul(sub(y, ty), inv);
    }
  }
  xu = ex;
  yu = ey;
}

// Camera.uv_to_xyz (camera.py:630-663): ray direction with unit depth along the optical axis.
__device__ inline void unproject(const gb_camera& c, double u, double v, double& dx, double& dy, double& dz) {
  if (c.affine) {  // Grid.uv_to_xyz (raster.py:447-459): uv * d + origin, z is NaN; the caller adds xyz = origin when asked for points
    dx = mul(u, c.f[0]);
    dy = mul(v, c.f[1]);
    dz = CUDART_NAN;
    return;
  }
  // (uv - (imgsz * 0.5 + c)) * (1 / f)   (camera.py:1517)
  const double x = mul(sub(u, c.cc[0]), quo(1.0, c.f[0]));
  const double y = mul(sub(v, c.cc[1]), quo(1.0, c.f[1]));
  double xu, yu;
  undistort(c, x, y, xu, yu);
  // xy . R[0:2, :] + R[2, :]   (camera.py:1487-1489)
  dx = add(fma(yu, c.R[3], xu * c.R[0]), c.R[6]);
  dy = add(fma(yu, c.R[4], xu * c.R[1]), c.R[7]);
  dz = add(fma(yu, c.R[5], xu * c.R[2]), c.R[8]);
}

}  // namespace gb
