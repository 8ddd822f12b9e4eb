// Ray sampling kernels (bandwidth-bound, bit-exact with the reference's float32 arithmetic).
//   reference: misc.py:234-247 sample_pt_coarse, misc.py:249-261 zero_invalid_pts,
//              mg_Img_Eval.py:57-60 / Eval_Tools_2.py:255-258 (shadow-march ray tops).
#include "common.cuh"
#include "api.h"

namespace snb {

// One warp per ray; lanes stride over samples so that the [S,3] slab of a ray is written as
// contiguous 12-byte records (coalesced: a warp writes 384 contiguous bytes per pass).
// All arithmetic uses explicit round-to-nearest intrinsics so that nvcc cannot contract
// a*b+c*d into an FMA (the reference does two rounded products and a rounded sum).
__global__ void __launch_bounds__(256) sample_rays_kernel(const float* __restrict__ top, const float* __restrict__ bot,
                                                          const float* __restrict__ ts, int N, int S, int zero_oob,
                                                          float* __restrict__ pts, float* __restrict__ deltas) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float fS = (float)S;
  for (int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5); n < N; n += gridDim.x * warps_per_block) {
    const float tx = __ldg(top + 3 * n), ty = __ldg(top + 3 * n + 1), tz = __ldg(top + 3 * n + 2);
    const float bx = __ldg(bot + 3 * n), by = __ldg(bot + 3 * n + 1), bz = __ldg(bot + 3 * n + 2);
    // deltas = sqrt(sum((top-bot)**2, 1)) / n_course      (misc.py:243)
    const float dx = __fsub_rn(tx, bx), dy = __fsub_rn(ty, by), dz = __fsub_rn(tz, bz);
    const float ss = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float delta = __fdiv_rn(__fsqrt_rn(ss), fS);
    for (int s = lane; s < S; s += 32) {
      const float t1 = __ldg(ts + s);
      const float t0 = __fsub_rn(1.0f, t1);  // (1 - ts)
      // pts = top*(1-ts) + bot*ts              (misc.py:244)
      const float px = __fadd_rn(__fmul_rn(tx, t0), __fmul_rn(bx, t1));
      const float py = __fadd_rn(__fmul_rn(ty, t0), __fmul_rn(by, t1));
      const float pz = __fadd_rn(__fmul_rn(tz, t0), __fmul_rn(bz, t1));
      const long long o = (long long)n * S + s;
      if (pts) {
        pts[3 * o] = px;
        pts[3 * o + 1] = py;
        pts[3 * o + 2] = pz;
      }
      float d = delta;
      if (zero_oob) {  // misc.py:257: good = all coords in [-1, 1] (bounds inclusive)
        const bool good = (px <= 1.f) && (py <= 1.f) && (pz <= 1.f) && (px >= -1.f) && (py >= -1.f) && (pz >= -1.f);
        if (!good) d = 0.f;
      }
      deltas[o] = d;
    }
  }
}

// new_top = p + ((1 - p_z)/sun_z) * sun, evaluated in float64 then rounded (CLI path) or in float32 (engine path).
__global__ void __launch_bounds__(256) solar_tops_kernel(const float* __restrict__ pts, long long M, double sx, double sy,
                                                         double sz, int f64, float* __restrict__ tops) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const float px = pts[3 * m], py = pts[3 * m + 1], pz = pts[3 * m + 2];
    if (f64) {
      // S = (1. - new_bots[:,2]) / sun[2] : float32 tensor op float64 numpy scalar -> stays float32 in torch
      // (python/numpy scalars do not promote tensors); then S.reshape * sun.reshape (float64 ndarray) promotes
      // to float64, the sum is float64 and .float() rounds once.      (mg_Img_Eval.py:58-60)
      const float k = __fdiv_rn(__fsub_rn(1.0f, pz), (float)sz);
      const double kd = (double)k;
      tops[3 * m] = (float)((double)px + kd * sx);
      tops[3 * m + 1] = (float)((double)py + kd * sy);
      tops[3 * m + 2] = (float)((double)pz + kd * sz);
    } else {
      // Eval_Tools_2.py:257-258: all float32
      const float fx = (float)sx, fy = (float)sy, fz = (float)sz;
      const float k = __fdiv_rn(__fsub_rn(1.0f, pz), fz);
      tops[3 * m] = __fadd_rn(px, __fmul_rn(k, fx));
      tops[3 * m + 1] = __fadd_rn(py, __fmul_rn(k, fy));
      tops[3 * m + 2] = __fadd_rn(pz, __fmul_rn(k, fz));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Rays of an affine-approximated RPC camera on the device: the closed-form 2x2 solve of P_img_Pinhole.invert_P
// (pre_NeRF/P_Img.py:133-147) for every pixel at z = z_top and z = z_bot, plus the bounds filter of
// mg_Pt_holder.py:180-187 / mg_Img_Eval.py:83-84.  float64 with explicit round-to-nearest intrinsics in numpy's
// evaluation order (no FMA contraction), so x, y are bit-identical to the reference's float64 arrays; the float32
// cast is the reference's `t.tensor(...).float()`.
struct CamP {
  double p[12];        // row-major 3x4 projection (already normalised / scaled by the caller)
  double bounds[4];    // x_min, x_max, y_min, y_max (inclusive)
};

__device__ __forceinline__ void invert_P_d(const CamP& c, double row, double col, double h, double& x, double& y) {
  const double* P = c.p;
  // P[1,2]*h + P[1,3] - P[2,2]*h*col - P[2,3]*col
  const double A = __dsub_rn(__dsub_rn(__dadd_rn(__dmul_rn(P[6], h), P[7]), __dmul_rn(__dmul_rn(P[10], h), col)),
                             __dmul_rn(P[11], col));
  // P[0,2]*h + P[0,3] - P[2,2]*h*row - P[2,3]*row
  const double B = __dsub_rn(__dsub_rn(__dadd_rn(__dmul_rn(P[2], h), P[3]), __dmul_rn(__dmul_rn(P[10], h), row)),
                             __dmul_rn(P[11], row));
  const double P11mP31x = __dsub_rn(P[0], __dmul_rn(P[8], row));
  const double P22mP32y = __dsub_rn(P[5], __dmul_rn(P[9], col));
  const double P12mP32x = __dsub_rn(P[1], __dmul_rn(P[9], row));
  const double P21mP31y = __dsub_rn(P[4], __dmul_rn(P[8], col));
  const double den = __dsub_rn(__dmul_rn(P11mP31x, P22mP32y), __dmul_rn(P12mP32x, P21mP31y));
  x = __ddiv_rn(__dsub_rn(__dmul_rn(P12mP32x, A), __dmul_rn(P22mP32y, B)), den);
  y = __ddiv_rn(__dadd_rn(__dmul_rn(-P11mP31x, A), __dmul_rn(P21mP31y, B)), den);
}

// pixel i: explicit (rows[i], cols[i]) or the raster grid (i / W * ds, i % W * ds)
__global__ void __launch_bounds__(256) camera_rays_kernel(CamP cam, const int* __restrict__ rows, const int* __restrict__ cols,
                                                          long long n, int W, int ds, double z_top, double z_bot,
                                                          float* __restrict__ tops, float* __restrict__ bots,
                                                          double* __restrict__ xy64, unsigned char* __restrict__ good) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double row = rows ? (double)rows[i] : (double)((i / W) * ds);
    const double col = cols ? (double)cols[i] : (double)((i % W) * ds);
    double tx, ty, bx, by;
    invert_P_d(cam, row, col, z_top, tx, ty);
    invert_P_d(cam, row, col, z_bot, bx, by);
    tops[3 * i] = (float)tx, tops[3 * i + 1] = (float)ty, tops[3 * i + 2] = (float)z_top;
    bots[3 * i] = (float)bx, bots[3 * i + 1] = (float)by, bots[3 * i + 2] = (float)z_bot;
    if (xy64) xy64[4 * i] = tx, xy64[4 * i + 1] = ty, xy64[4 * i + 2] = bx, xy64[4 * i + 3] = by;
    if (good) {
      const double x0 = cam.bounds[0], x1 = cam.bounds[1], y0 = cam.bounds[2], y1 = cam.bounds[3];
      good[i] = (tx <= x1) && (x0 <= tx) && (ty <= y1) && (y0 <= ty) && (bx <= x1) && (x0 <= bx) && (by <= y1) && (y0 <= by);
    }
  }
}

// create_solor_rays_uniform (Eval_Tools_2.py:72-108) for n rays at once: the reference's per-ray Python loop over
// world_angle_2_local_vec (all_NeRF/mg_unit_converter.py:5-9,29-34,59-68) in float64, like numpy, then
//   delta = 2 v / v_z (float64),  start = (2 u_x - 1, 2 u_y - 1, 1) (float32),  end = float(start - delta),
//   time  = (cos f0, sin f0, cos f1, sin f1),  f = (u * 2) * fl32(pi)   (float32, like `t.rand([n,2]) * 2 * np.pi`).
struct SolarGeo {
  double wc[3];
  double h[12];   // first three rows of World2Local_H
};

__global__ void __launch_bounds__(128) solar_rays_kernel(SolarGeo g, const double* __restrict__ az_el, const float* __restrict__ u_xy,
                                                         const float* __restrict__ u_time, int n, float* __restrict__ starts,
                                                         float* __restrict__ ends, float* __restrict__ vec,
                                                         float* __restrict__ times) {
  const double kDeg = 0.017453292519943295;      // pi / 180
  const double kRadDeg = 57.29577951308232;      // 180 / pi
  const double kR = 1000. * 6378.137;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double az = az_el[2 * i], el = az_el[2 * i + 1];
    // LLA_get_vec(center, theta = az, rho = el)
    const double Y = cos(az * kDeg), X = sin(az * kDeg);
    const double Z = tan(el * kDeg) * sqrt(X * X + Y * Y);
    const double norm = sqrt(X * X + Y * Y + Z * Z) / 1000.;
    const double xn = X / norm, yn = Y / norm, zn = Z / norm;
    const double lat = g.wc[0] + (yn / kR) * kRadDeg;
    const double lon = g.wc[1] + (xn / (kR * cos(g.wc[0] * kDeg))) * kRadDeg;
    const double alt = g.wc[2] + zn;
    // homogeneous transform, first three rows; then unit length
    const double t0 = lat * g.h[0] + lon * g.h[1] + alt * g.h[2] + g.h[3];
    const double t1 = lat * g.h[4] + lon * g.h[5] + alt * g.h[6] + g.h[7];
    const double t2 = lat * g.h[8] + lon * g.h[9] + alt * g.h[10] + g.h[11];
    const double len = sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    const double v0 = t0 / len, v1 = t1 / len, v2 = t2 / len;
    const double d0 = 2. * (v0 / v2), d1 = 2. * (v1 / v2), d2 = 2. * (v2 / v2);
    const float sx = __fadd_rn(__fmul_rn(2.0f, u_xy[2 * i]), -1.0f), sy = __fadd_rn(__fmul_rn(2.0f, u_xy[2 * i + 1]), -1.0f);
    starts[3 * i] = sx, starts[3 * i + 1] = sy, starts[3 * i + 2] = 1.0f;
    ends[3 * i] = (float)((double)sx - d0), ends[3 * i + 1] = (float)((double)sy - d1), ends[3 * i + 2] = (float)(1.0 - d2);
    vec[3 * i] = (float)v0, vec[3 * i + 1] = (float)v1, vec[3 * i + 2] = (float)v2;
    if (times) {
      const float f0 = __fmul_rn(__fmul_rn(u_time[2 * i], 2.0f), 3.14159274101257324f);
      const float f1 = __fmul_rn(__fmul_rn(u_time[2 * i + 1], 2.0f), 3.14159274101257324f);
      times[4 * i] = cosf(f0), times[4 * i + 1] = sinf(f0), times[4 * i + 2] = cosf(f1), times[4 * i + 3] = sinf(f1);
    }
  }
}

// T_NeRF.Supervised_Sample (T_NeRF_net_v2.py:175-181): density implied by the prior height map at every sample point.
//   idx = long(((xy + 1) / 2) * (shape - 1))   float32 arithmetic, truncation;   P = min(0.99, hm[idx] >= z)  (hm float64)
//   rho = -log(1 - P) / delta  = k_hit / delta  or  -0 / delta     (k_hit = -log(1 - fl32(0.99)) evaluated by the caller)
// Negative indices wrap like torch's; anything still outside the map is clamped (torch would raise).
__global__ void __launch_bounds__(256) supervised_sample_kernel(const float* __restrict__ pts, const float* __restrict__ delta,
                                                                const double* __restrict__ hm, int H, int W, float k_hit,
                                                                long long M, float* __restrict__ out) {
  const float ch = (float)(H - 1), cw = (float)(W - 1);
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const float x = pts[3 * m], y = pts[3 * m + 1], z = pts[3 * m + 2];
    long long ix = (long long)__fmul_rn(__fdiv_rn(__fadd_rn(x, 1.0f), 2.0f), ch);
    long long iy = (long long)__fmul_rn(__fdiv_rn(__fadd_rn(y, 1.0f), 2.0f), cw);
    if (ix < 0) ix += H;
    if (iy < 0) iy += W;
    ix = ix < 0 ? 0 : (ix >= H ? H - 1 : ix);
    iy = iy < 0 ? 0 : (iy >= W ? W - 1 : iy);
    const bool hit = hm[ix * W + iy] >= (double)z;
    out[m] = __fdiv_rn(hit ? k_hit : -0.0f, delta[m]);
  }
}

}  // namespace snb

extern "C" int snb_sample_rays(const float* top, const float* bot, const float* ts, int N, int S, int zero_oob,
                               float* pts, float* deltas, void* stream) {
  SNB_CHECK_ARG(top && bot && ts && deltas && N >= 0 && S > 0);
  if (N == 0) return SNB_OK;
  const int wpb = 8;
  int grid = snb::grid_for(N, wpb, 16);
  snb::sample_rays_kernel<<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(top, bot, ts, N, S, zero_oob, pts, deltas);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_solar_tops(const float* pts, long long M, const double* sun3_host, int f64, float* tops,
                              void* stream) {
  SNB_CHECK_ARG(pts && sun3_host && tops && M >= 0);
  if (M == 0) return SNB_OK;
  int grid = snb::grid_for(M, 256, 16);
  snb::solar_tops_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pts, M, sun3_host[0], sun3_host[1], sun3_host[2],
                                                                  f64, tops);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_camera_rays(const double* P, const int* rows, const int* cols, long long n, int W, int ds, double z_top,
                               double z_bot, const double* bounds, float* tops, float* bots, double* xy64,
                               unsigned char* good, void* stream) {
  SNB_CHECK_ARG(P && tops && bots && n >= 0 && ((rows == nullptr) == (cols == nullptr)) && (rows || (W > 0 && ds > 0)));
  SNB_CHECK_ARG(!good || bounds);
  if (n == 0) return SNB_OK;
  snb::CamP cam;
  for (int i = 0; i < 12; ++i) cam.p[i] = P[i];           // host pointers: 16 doubles travel as a kernel argument
  for (int i = 0; i < 4; ++i) cam.bounds[i] = bounds ? bounds[i] : 0.0;
  snb::camera_rays_kernel<<<snb::grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(cam, rows, cols, n, W, ds, z_top, z_bot, tops,
                                                                                      bots, xy64, good);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_solar_rays(const double* world_center, const double* W2L_H, const double* az_el, const float* u_xy,
                              const float* u_time, int n, float* starts, float* ends, float* vec, float* times,
                              void* stream) {
  SNB_CHECK_ARG(world_center && W2L_H && az_el && u_xy && starts && ends && vec && n >= 0 && ((times == nullptr) == (u_time == nullptr)));
  if (n == 0) return SNB_OK;
  snb::SolarGeo g;
  for (int i = 0; i < 3; ++i) g.wc[i] = world_center[i];    // host pointers: travel as a kernel argument
  for (int i = 0; i < 12; ++i) g.h[i] = W2L_H[i];
  snb::solar_rays_kernel<<<snb::grid_for(n, 128, 8), 128, 0, (cudaStream_t)stream>>>(g, az_el, u_xy, u_time, n, starts, ends, vec, times);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_supervised_sample(const float* pts, const float* delta, const double* hm, int H, int W, float k_hit,
                                     long long M, float* out, void* stream) {
  SNB_CHECK_ARG(pts && delta && hm && out && H >= 1 && W >= 1 && M >= 0);
  if (M == 0) return SNB_OK;
  snb::supervised_sample_kernel<<<snb::grid_for(M, 256, 16), 256, 0, (cudaStream_t)stream>>>(pts, delta, hm, H, W, k_hit, M, out);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
