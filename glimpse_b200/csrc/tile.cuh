// Per-(point, observer) tile pipeline, executed by one CTA out of shared memory:
//   raw window -> histogram -> CDF-matching look-up table -> 5x5 median high-pass -> SSD surface
//   -> bicubic not-a-knot spline in Hermite form.
// Reference: Tracker.extract_tile (track/tracker.py:494-534), helpers.normalize / compute_cdf /
// match_cdf (helpers.py:324-344, 433-493), scipy.ndimage.median_filter(size=5x5, mode='reflect'),
// cv2.matchTemplate(TM_SQDIFF) / (w h) (tracker.py:609-614), RectBivariateSpline(kx=ky=3, s=0)
// (track/observer.py:178-214).
//
// Two identities keep the work integer-valued until the last moment (frames are uint8 band sums):
//  * CDF matching only depends on pixel ranks, so the matched tile is LUT[raw] with one LUT entry
//    per grey level: LUT[g] = interp(count(raw <= g) / size, template_quantiles, template_values).
//  * LUT is non-decreasing and the median of 25 values is one of them, so
//    median5x5(LUT[raw]) == LUT[median5x5(raw)]; the median runs on integers.
#pragma once
#include "common.cuh"
#include "median25.cuh"

namespace gb {

// Thomas-algorithm factors of the not-a-knot slope system (rows [1 2], [1 4 1]..., [2 1]); they
// depend only on the row index, not on the system size.  Filled once per device by the host.
__constant__ double c_spline_cp[256];
__constant__ double c_spline_inv[256];

struct TileWork {
  float4* herm;    // [Mv * Mu] (F, dF/du, dF/dv, d2F/dudv)
  double* lut;     // [nbins]
  double* tq;      // [nvals] template quantiles
  double* tv;      // [nvals] template values
  float* hp;       // [Sv * Su] high-passed search tile
  float* tmpl;     // [th * tw] high-passed template
  uint32_t* hist;  // [nbins]
  uint16_t* raw;   // [Sv * Su]
  int Su, Sv, Mu, Mv, nbins, nvals, tw, th;
};

__host__ __device__ inline int64_t tile_bytes_needed(int Su, int Sv, int tw, int th, int nbins, int nvals) {
  const int64_t Mu = Su - tw + 1, Mv = Sv - th + 1;
  int64_t b = 0;
  b += ((Mu * Mv * 16 + 15) / 16) * 16;
  b += (int64_t)nbins * 8 + (int64_t)nvals * 16;
  b += ((int64_t)Su * Sv * 4 + 15) / 16 * 16;
  b += ((int64_t)tw * th * 4 + 15) / 16 * 16;
  b += (int64_t)nbins * 4;
  b += ((int64_t)Su * Sv * 2 + 15) / 16 * 16;
  return b;
}

__device__ inline void tile_carve(char* base, TileWork& w) {
  char* p = base;
  w.herm = reinterpret_cast<float4*>(p);
  p += (((int64_t)w.Mu * w.Mv * 16 + 15) / 16) * 16;
  w.lut = reinterpret_cast<double*>(p);
  p += (int64_t)w.nbins * 8;
  w.tq = reinterpret_cast<double*>(p);
  p += (int64_t)w.nvals * 8;
  w.tv = reinterpret_cast<double*>(p);
  p += (int64_t)w.nvals * 8;
  w.hp = reinterpret_cast<float*>(p);
  p += ((int64_t)w.Su * w.Sv * 4 + 15) / 16 * 16;
  w.tmpl = reinterpret_cast<float*>(p);
  p += ((int64_t)w.tw * w.th * 4 + 15) / 16 * 16;
  w.hist = reinterpret_cast<uint32_t*>(p);
  p += (int64_t)w.nbins * 4;
  w.raw = reinterpret_cast<uint16_t*>(p);
}

// np.interp(q, xp, fp) for one abscissa (numpy compiled_base.c arr_interp): clamped at the ends,
// exact hit returns fp[j], otherwise slope * (q - xp[j]) + fp[j] unfused.
__device__ inline double interp_clamped(double q, const double* xp, const double* fp, int n) {
  if (q > xp[n - 1]) return fp[n - 1];
  if (q < xp[0]) return fp[0];
  int lo = 0, hi = n;  // largest j with xp[j] <= q
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= q) lo = mid; else hi = mid;
  }
  const int j = lo;
  if (j == n - 1 || xp[j] == q) return fp[j];
  const double slope = quo(sub(fp[j + 1], fp[j]), sub(xp[j + 1], xp[j]));
  return add(mul(slope, sub(q, xp[j])), fp[j]);
}

__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy.ndimage 'reflect': d c b a | a b c d | d c b a
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - i - 1;
  return i;
}

// Median of the 5x5 neighbourhood of (r, c) in an integer tile with reflected borders.
__device__ __forceinline__ int median5x5(const uint16_t* raw, int Su, int Sv, int r, int c) {
  int v[25];
  int cols[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) cols[j] = reflect_index(c + j - 2, Su);
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const uint16_t* row = raw + reflect_index(r + i - 2, Sv) * Su;
#pragma unroll
    for (int j = 0; j < 5; ++j) v[i * 5 + j] = row[cols[j]];
  }
  return gb_median25(v);
}

// Solve the not-a-knot slope system along one line of the Hermite array (stride in floats).
// `in` holds the samples, `out` receives the slopes; intermediate values pass through `out`.
__device__ inline void spline_slopes_line(const float* in, float* out, int stride, int m) {
  const double f0 = in[0], f1 = in[stride], f2 = in[2 * stride];
  double d = 0.5 * (5.0 * (f1 - f0) + (f2 - f1));
  out[0] = (float)d;
  double prev = f0, cur = f1;
  for (int i = 1; i < m - 1; ++i) {
    const double next = in[(i + 1) * stride];
    const double rhs = 3.0 * (next - prev);
    d = (rhs - d) * c_spline_inv[i];
    out[i * stride] = (float)d;
    prev = cur;
    cur = next;
  }
  {
    const double fa = in[(m - 3) * stride], fb = in[(m - 2) * stride], fc = in[(m - 1) * stride];
    const double rhs = 0.5 * (5.0 * (fc - fb) + (fb - fa));
    d = (rhs - 2.0 * d) / (1.0 - 2.0 * c_spline_cp[m - 2]);
  }
  double s = d;
  out[(m - 1) * stride] = (float)s;
  for (int i = m - 2; i >= 0; --i) {
    s = (double)out[i * stride] - c_spline_cp[i] * s;
    out[i * stride] = (float)s;
  }
}

// Bicubic Hermite evaluation at (x, y) measured from the first cell centre, in cell units.
__device__ __forceinline__ double hermite_eval(const float4* herm, int Mu, int Mv, double x, double y) {
  int j = min((int)floor(x), Mu - 2), i = min((int)floor(y), Mv - 2);
  j = max(j, 0);
  i = max(i, 0);
  const double tx = x - (double)j, ty = y - (double)i;
  const double tx2 = tx * tx, tx3 = tx2 * tx, ty2 = ty * ty, ty3 = ty2 * ty;
  const double a0 = 2.0 * tx3 - 3.0 * tx2 + 1.0, a1 = tx3 - 2.0 * tx2 + tx, a2 = -2.0 * tx3 + 3.0 * tx2, a3 = tx3 - tx2;
  const double b0 = 2.0 * ty3 - 3.0 * ty2 + 1.0, b1 = ty3 - 2.0 * ty2 + ty, b2 = -2.0 * ty3 + 3.0 * ty2, b3 = ty3 - ty2;
  const float4 h00 = herm[i * Mu + j], h01 = herm[i * Mu + j + 1];
  const float4 h10 = herm[(i + 1) * Mu + j], h11 = herm[(i + 1) * Mu + j + 1];
  // rows of the patch: value and u-slope interpolated along u, for f and for df/dv
  const double top_f = a0 * h00.x + a2 * h01.x + a1 * h00.y + a3 * h01.y;
  const double bot_f = a0 * h10.x + a2 * h11.x + a1 * h10.y + a3 * h11.y;
  const double top_v = a0 * h00.z + a2 * h01.z + a1 * h00.w + a3 * h01.w;
  const double bot_v = a0 * h10.z + a2 * h11.z + a1 * h10.w + a3 * h11.w;
  return b0 * top_f + b2 * bot_f + b1 * top_v + b3 * bot_v;
}

// Build the Hermite surface for one search window.  All threads of the CTA participate.
// `box` = (left, top, right, bottom); template data in global memory.  The caller has verified
// the capacity and carved `w`.
__device__ inline void tile_build_surface(const gb_image* img, const int* box, const double* g_tmpl,
                                          const double* g_tq, const double* g_tv, TileWork& w, float* dump_search,
                                          float* dump_sse, int64_t dump_cap, long long* clk) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int Su = w.Su, Sv = w.Sv, Mu = w.Mu, Mv = w.Mv;
  const int area = Su * Sv;
  // 1. raw window, template, template CDF; clear histogram
  {
    const uint16_t* gray = img->gray;
    const int pitch = img->pitch, left = box[0], top = box[1];
    for (int i = tid; i < area; i += nthr) {
      const int r = i / Su, c = i - r * Su;
      w.raw[i] = gray[(int64_t)(top + r) * pitch + left + c];
    }
    for (int i = tid; i < w.nbins; i += nthr) w.hist[i] = 0u;
    for (int i = tid; i < w.tw * w.th; i += nthr) w.tmpl[i] = (float)g_tmpl[i];
    for (int i = tid; i < w.nvals; i += nthr) {
      w.tq[i] = g_tq[i];
      w.tv[i] = g_tv[i];
    }
  }
  __syncthreads();
  // 2. histogram of grey levels
  for (int i = tid; i < area; i += nthr) atomicAdd(&w.hist[w.raw[i]], 1u);
  __syncthreads();
  // 3. inclusive cumulative counts (one warp; nbins <= 1024)
  if (tid < 32) {
    const int per = (w.nbins + 31) / 32;
    const int b0 = tid * per, b1 = min(b0 + per, w.nbins);
    uint32_t sum = 0;
    for (int b = b0; b < b1; ++b) sum += w.hist[b];
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (tid >= d) incl += o;
    }
    uint32_t run = incl - sum;
    for (int b = b0; b < b1; ++b) {
      const uint32_t h = w.hist[b];
      run += h;
      // keep the count of the level in the top bit-free range; mark empty levels with 0
      w.hist[b] = h ? run : 0u;
    }
  }
  __syncthreads();
  // 4. look-up table: matched value of every occupied grey level (helpers.py:488-493)
  for (int b = tid; b < w.nbins; b += nthr) {
    const uint32_t cle = w.hist[b];
    if (cle) w.lut[b] = interp_clamped(quo((double)cle, (double)area), w.tq, w.tv, w.nvals);
  }
  __syncthreads();
  if (clk && threadIdx.x == 0) clk[0] = clock64();
  // 5. high-pass: matched value minus the matched 5x5 median (tracker.py:530-531), cast to
  //    float32 as the reference does for matchTemplate (tracker.py:610)
  for (int i = tid; i < area; i += nthr) {
    const int r = i / Su, c = i - r * Su;
    const int med = median5x5(w.raw, Su, Sv, r, c);
    const float v = (float)sub(w.lut[w.raw[i]], w.lut[med]);
    w.hp[i] = v;
    if (dump_search && i < dump_cap) dump_search[i] = v;
  }
  __syncthreads();
  if (clk && threadIdx.x == 0) clk[1] = clock64();
  // 6. area-normalised sum of squared differences (tracker.py:609-614)
  {
    const double inv_area = 1.0 / (double)(w.tw * w.th);
    const int tw = w.tw, th = w.th;
    for (int o = tid; o < Mu * Mv; o += nthr) {
      const int r = o / Mu, c = o - r * Mu;
      float acc = 0.0f;
      for (int i = 0; i < th; ++i) {
        const float* srow = w.hp + (r + i) * Su + c;
        const float* trow = w.tmpl + i * tw;
        float racc = 0.0f;
        for (int j = 0; j < tw; ++j) {
          const float d = srow[j] - trow[j];
          racc = fmaf(d, d, racc);
        }
        acc += racc;
      }
      const float sse = (float)((double)acc * inv_area);
      w.herm[o] = make_float4(sse, 0.0f, 0.0f, 0.0f);
      if (dump_sse && o < dump_cap) dump_sse[o] = sse;
    }
  }
  __syncthreads();
  if (clk && threadIdx.x == 0) clk[2] = clock64();
  // 7. Hermite data: dF/du along rows and dF/dv along columns, then the cross derivative
  {
    float* base = reinterpret_cast<float*>(w.herm);
    const int Mvp = 32 * ((Mv + 31) / 32);  // rows and columns on separate warps
    for (int line = tid; line < Mvp + Mu; line += nthr) {
      if (line < Mv) {
        spline_slopes_line(base + (int64_t)line * Mu * 4, base + (int64_t)line * Mu * 4 + 1, 4, Mu);
      } else if (line >= Mvp) {
        const int c = line - Mvp;
        spline_slopes_line(base + (int64_t)c * 4, base + (int64_t)c * 4 + 2, Mu * 4, Mv);
      }
    }
    __syncthreads();
    for (int c = tid; c < Mu; c += nthr) spline_slopes_line(base + (int64_t)c * 4 + 1, base + (int64_t)c * 4 + 3, Mu * 4, Mv);
  }
  __syncthreads();
}

}  // namespace gb
