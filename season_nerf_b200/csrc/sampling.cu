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
