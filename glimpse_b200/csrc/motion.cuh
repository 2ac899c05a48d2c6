// Motion models, surfaces and random draws.
//   CartesianMotion / CylindricalMotion: reference track/motion.py:92-311
//   TangentCartesianMotion / TangentCylindricalMotion: reference track/motion.py:314-522
//   Raster.sample point mode (DEM, DEM sigma, viewshed): reference raster.py:891-1027
#pragma once
#include "common.cuh"

namespace gb {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11) + Box-Muller.
// ---------------------------------------------------------------------------------------------
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
  }
  __device__ __forceinline__ void generate(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c[0];
    out[1] = c[1];
    out[2] = c[2];
    out[3] = c[3];
  }
};

// Three standard normals for (point, time, particle); `stream` separates init (0..1) from steps (2).
// `need` masks the pairs that are actually used (bit 0: z0/z1, bit 1: z2): a normal that only multiplies a
// zero sigma is not generated (the stream of the others is unchanged).
__device__ __forceinline__ void philox_normals3(uint64_t seed, uint64_t point, uint32_t time, uint32_t particle,
                                                uint32_t stream, double& z0, double& z1, double& z2, int need = 3) {
  Philox g{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  g.generate(particle, time, (uint32_t)point, ((uint32_t)(point >> 32) << 8) | stream, r);
  const float two_neg32 = 2.3283064365386963e-10f;
  z0 = z1 = z2 = 0.0;
  if (need & 1) {
    const float u0 = ((float)r[0] + 0.5f) * two_neg32;
    const float a0 = fminf(fmaxf(u0, 1.0e-10f), 1.0f);  // u in (0, 1]
    // sqrt(x) as x * rsqrt(x) (MUFU.RSQ); x = -2 ln(u) is floored at 1e-30 so that u == 1 gives 0, not NaN
    const float e0 = fmaxf(-2.0f * __logf(a0), 1.0e-30f);
    const float rad0 = e0 * rsqrtf(e0);
    float s0, c0;
    __sincosf(6.283185307179586f * ((float)(r[1] >> 8) * 5.9604644775390625e-08f), &s0, &c0);
    z0 = (double)(rad0 * c0);
    z1 = (double)(rad0 * s0);
  }
  if (need & 2) {
    const float u1 = ((float)r[2] + 0.5f) * two_neg32;
    const float a1 = fminf(fmaxf(u1, 1.0e-10f), 1.0f);
    const float e1 = fmaxf(-2.0f * __logf(a1), 1.0e-30f);
    const float rad1 = e1 * rsqrtf(e1);
    z2 = (double)(rad1 * __cosf(6.283185307179586f * ((float)(r[3] >> 8) * 5.9604644775390625e-08f)));
  }
}

__device__ __forceinline__ bool motion_is_tangent(const gb_motion& m) { return m.kind >= GB_MOTION_TANGENT_CARTESIAN; }
__device__ __forceinline__ bool motion_is_cylindrical(const gb_motion& m) {
  return m.kind == GB_MOTION_CYLINDRICAL || m.kind == GB_MOTION_TANGENT_CYLINDRICAL;
}

// Which normals evolve_particle really consumes for this motion model.
__device__ __forceinline__ int evolve_needs(const gb_motion& m) {
  const double third = motion_is_tangent(m) ? m.slope_sigma : m.a_sigma[2];
  return ((m.a_sigma[0] != 0.0 || m.a_sigma[1] != 0.0) ? 1 : 0) | (third != 0.0 ? 2 : 0);
}

__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t point, uint32_t time) {
  Philox g{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  g.generate(0xFFFFFFFFu, time, (uint32_t)point, ((uint32_t)(point >> 32) << 8) | 3u, r);
  // 53-bit uniform in [0, 1) like np.random.random()
  const uint64_t bits = (((uint64_t)r[0] >> 5) << 26) | ((uint64_t)r[1] >> 6);
  return (double)bits * (1.0 / 9007199254740992.0);
}

// One uniform per (point, time, particle): the stratified resampler's per-stratum draw (stream 4).
__device__ __forceinline__ double philox_uniform_particle(uint64_t seed, uint64_t point, uint32_t time, uint32_t particle) {
  Philox g{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  g.generate(particle, time, (uint32_t)point, ((uint32_t)(point >> 32) << 8) | 4u, r);
  const uint64_t bits = (((uint64_t)r[0] >> 5) << 26) | ((uint64_t)r[1] >> 6);
  return (double)bits * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------
// Surfaces
// ---------------------------------------------------------------------------------------------
// Point-mode Raster.sample with bounds_error=True (raster.py:913-1027): `oob` is set when the
// reference would raise "Some of the sampling coordinates are out of bounds" (NaN coordinates
// fail the comparison, as in NumPy).  order 1: RegularGridInterpolator linear with linear
// extrapolation in the half-cell rim; order 0: nearest (ties to the lower cell).
__device__ inline double surface_sample(const gb_surface& s, double x, double y, int order, bool& oob) {
  oob = !((x >= s.xmin) & (x <= s.xmax) & (y >= s.ymin) & (y <= s.ymax));
  if (s.z == nullptr) return s.value;
  if (oob) return CUDART_NAN;
  // index of the grid interval: searchsorted(grid, v) - 1 clipped to [0, n - 2]
  double fx = (x - s.x0) / s.dx, fy = (y - s.y0) / s.dy;
  int ix = (int)ceil(fx) - 1, iy = (int)ceil(fy) - 1;
  ix = max(0, min(ix, s.nx - 2));
  iy = max(0, min(iy, s.ny - 2));
  const double tx = fx - (double)ix, ty = fy - (double)iy;
  if (order == 0) {
    const int jx = (tx <= 0.5) ? ix : ix + 1, jy = (ty <= 0.5) ? iy : iy + 1;
    return s.z[(int64_t)jx * s.ny + jy];
  }
  const double z00 = s.z[(int64_t)ix * s.ny + iy], z01 = s.z[(int64_t)ix * s.ny + iy + 1];
  const double z10 = s.z[(int64_t)(ix + 1) * s.ny + iy], z11 = s.z[(int64_t)(ix + 1) * s.ny + iy + 1];
  return z00 * (1.0 - tx) * (1.0 - ty) + z01 * (1.0 - tx) * ty + z10 * tx * (1.0 - ty) + z11 * tx * ty;
}

// ---------------------------------------------------------------------------------------------
// Particle initialisation and evolution
// ---------------------------------------------------------------------------------------------
// initialize_particles (motion.py:149-163, 260-283, 378-390, 485-505).  zn = the six normals of this particle
// in the reference's draw order: randn(n,2) -> xy, randn(n) -> z, randn(n,3) -> velocity (tangent kinds:
// randn(n,2), zn[5] unused and vz = 0).
__device__ inline void init_particle(const gb_motion& m, const gb_surface* surfaces, const double (&zn)[6],
                                     double (&s)[6], uint32_t& flags) {
  s[0] = add(m.xy[0], mul(m.xy_sigma[0], zn[0]));
  s[1] = add(m.xy[1], mul(m.xy_sigma[1], zn[1]));
  bool oob0, oob1;
  const double z = surface_sample(surfaces[m.dem], s[0], s[1], 1, oob0);
  const double zs = surface_sample(surfaces[m.dem_sigma], s[0], s[1], 1, oob1);
  if (oob0 | oob1) flags |= GB_F_DEM_OOB;
  s[2] = add(z, mul(zs, zn[2]));
  const double v0 = add(m.v[0], mul(m.v_sigma[0], zn[3]));
  const double v1 = add(m.v[1], mul(m.v_sigma[1], zn[4]));
  if (motion_is_cylindrical(m)) {
    s[3] = mul(v0, cos(v1));
    s[4] = mul(v0, sin(v1));
  } else {
    s[3] = v0;
    s[4] = v1;
  }
  s[5] = motion_is_tangent(m) ? 0.0 : add(m.v[2], mul(m.v_sigma[2], zn[5]));
}

// Height of a tangent-model particle after a horizontal step (dx, dy) from (x, y, z) (motion.py:402-409):
// its offset above the DEM, widened by slope_sigma * z2 * |dxy|, on top of the DEM at the new position.
__device__ __forceinline__ double tangent_height(const gb_surface& dem, double slope_sigma, double z2, double x, double y,
                                              double z, double dx, double dy, uint32_t& flags) {
  bool oob0, oob1;
  double zoff = sub(z, surface_sample(dem, x, y, 1, oob0));
  zoff = add(zoff, mul(mul(slope_sigma, z2), sqrt(add(mul(dx, dx), mul(dy, dy)))));
  const double znew = add(surface_sample(dem, add(x, dx), add(y, dy), 1, oob1), zoff);
  if (oob0 | oob1) flags |= GB_F_EVOLVE_OOB;
  return znew;
}

// evolve_particles (motion.py:165-179, 285-311): position uses the old velocity.  Tangent kinds
// (motion.py:392-420, 507-522): the step is horizontal, the particle keeps its offset above the DEM, widened by
// slope_sigma * z2 * |dxy|; z0, z1 are the randn(n,2) draw and z2 the randn(n) draw.
// TAN = false compiles the tangent branch out (kernels specialised for tracks without tangent models).
template <bool TAN = true>
__device__ __forceinline__ void evolve_particle(const gb_motion& m, const gb_surface* surfaces, double tau, double tau2,
                                                double z0, double z1, double z2, double (&s)[6], uint32_t& flags) {
  double a0 = add(m.a[0], mul(m.a_sigma[0], z0));
  double a1 = add(m.a[1], mul(m.a_sigma[1], z1));
  if (motion_is_cylindrical(m)) {
    const double vx = s[3], vy = s[4];
    const double vr = sqrt(add(mul(vx, vx), mul(vy, vy)));
    const double ax = sub(mul(a0, quo(vx, vr)), mul(vy, a1));
    const double ay = add(mul(a0, quo(vy, vr)), mul(vx, a1));
    a0 = ax;
    a1 = ay;
  }
  const double dx = add(mul(tau, s[3]), mul(mul(0.5, a0), tau2));
  const double dy = add(mul(tau, s[4]), mul(mul(0.5, a1), tau2));
  if (TAN && motion_is_tangent(m)) {
    s[2] = tangent_height(surfaces[m.dem], m.slope_sigma, z2, s[0], s[1], s[2], dx, dy, flags);
    s[0] = add(s[0], dx);
    s[1] = add(s[1], dy);
    s[3] = add(s[3], mul(tau, a0));
    s[4] = add(s[4], mul(tau, a1));
    return;
  }
  const double a2 = add(m.a[2], mul(m.a_sigma[2], z2));
  s[0] = add(s[0], dx);
  s[1] = add(s[1], dy);
  s[2] = add(s[2], add(mul(tau, s[5]), mul(mul(0.5, a2), tau2)));
  s[3] = add(s[3], mul(tau, a0));
  s[4] = add(s[4], mul(tau, a1));
  s[5] = add(s[5], mul(tau, a2));
}

// compute_log_likelihoods (motion.py:181-204): (dem(xy) - z)^2 / (2 sigma^2) where sigma != 0.
__device__ __forceinline__ double surface_log_likelihood(const gb_motion& m, const gb_surface* surfaces, double x,
                                                         double y, double z, uint32_t& flags) {
  bool oob0, oob1;
  const double zd = surface_sample(surfaces[m.dem], x, y, 1, oob0);
  const double zs = surface_sample(surfaces[m.dem_sigma], x, y, 1, oob1);
  if (oob0 | oob1) flags |= GB_F_DEM_OOB;
  if (zs == 0.0) return 0.0;
  const double dz = sub(zd, z);
  return mul(quo(1.0, mul(2.0, mul(zs, zs))), mul(dz, dz));
}

}  // namespace gb
