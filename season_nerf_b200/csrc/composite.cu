// Alpha-compositing with a transmittance scan: forward, fused backward, shadow-march transmittance,
// float64 CLI composite and the fused year sweep.  One warp per ray; lanes stride over the samples
// so that every load/store of a warp covers one contiguous span of the ray's record; the
// exclusive transmittance prefix is a warp-shuffle scan carried across 32-sample chunks.
//   reference: Eval_Tools_2.py:13-16 (get_PV), :187-215 (eval), mg_Img_Eval.py:68-70, :123-228.
#include <cstdlib>
#include "common.cuh"
#include "api.h"

namespace snb {

constexpr int kMaxChunks = 8;  // S <= 256 for the register-resident backward

struct RayAcc {
  float a0, a1, a2;  // sum PS*col
  float r0, r1, r2;  // classic: sum PS*col*(vis+(1-vis)*sky)
  float vs;          // sum vis*PS
  float k0, k1, k2;  // sum_s sky (per-sample sky)
};

// Unrolled forward (S <= 32*kChunks): every load of the ray (6 per sample chunk) is issued before the first scan, so
// one warp keeps 6*kChunks 128-byte requests in flight instead of waiting on each chunk's shuffle chain in turn.
template <bool kClassic, int kChunks>
__global__ void __launch_bounds__(256)
composite_fwd_unrolled_kernel(const float* __restrict__ rho, const float* __restrict__ deltas, const float* __restrict__ col,
                              const float* __restrict__ vis, const float* __restrict__ sky, int N, int S,
                              float* __restrict__ PV, float* __restrict__ PE, float* __restrict__ PS,
                              float* __restrict__ albedo, float* __restrict__ rendered, float* __restrict__ vis_sum) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    float y[kChunks], c0[kChunks], c1[kChunks], c2[kChunks], v[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const long long o = base + s;
      const float r = ok ? __ldg(rho + o) : 0.f, d = ok ? __ldg(deltas + o) : 0.f;
      c0[c] = ok ? __ldg(col + 3 * o) : 0.f, c1[c] = ok ? __ldg(col + 3 * o + 1) : 0.f, c2[c] = ok ? __ldg(col + 3 * o + 2) : 0.f;
      v[c] = ok ? __ldg(vis + o) : 0.f;
      y[c] = r * d;
    }
    const float ks0 = __ldg(sky + 3 * n), ks1 = __ldg(sky + 3 * n + 1), ks2 = __ldg(sky + 3 * n + 2);
    float carry = 0.f, a0 = 0, a1 = 0, a2 = 0, vs = 0, r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const long long o = base + s;
      const float incl = warp_scan_incl(y[c], lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      const float pv = expf(-(carry + prev));
      const float pe = 1.f - expf(-y[c]);
      const float ps = pv * pe;
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (s < S) {
        if (PV) PV[o] = pv;
        if (PE) PE[o] = pe;
        if (PS) PS[o] = ps;
        a0 += ps * c0[c], a1 += ps * c1[c], a2 += ps * c2[c];
        vs += v[c] * ps;
        if (kClassic) {
          r0 += ps * c0[c] * (v[c] + (1.f - v[c]) * ks0);
          r1 += ps * c1[c] * (v[c] + (1.f - v[c]) * ks1);
          r2 += ps * c2[c] * (v[c] + (1.f - v[c]) * ks2);
        }
      }
    }
    a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2), vs = warp_sum(vs);
    if (kClassic) r0 = warp_sum(r0), r1 = warp_sum(r1), r2 = warp_sum(r2);
    if (lane == 0) {
      albedo[3 * n] = a0, albedo[3 * n + 1] = a1, albedo[3 * n + 2] = a2;
      if (vis_sum) vis_sum[n] = vs;
      if (kClassic) {
        rendered[3 * n] = r0, rendered[3 * n + 1] = r1, rendered[3 * n + 2] = r2;
      } else {
        const float sv3 = sigmoidf_((vs - .2f) * 30.f);  // Eval_Tools_2.py:214
        rendered[3 * n] = a0 * (sv3 + (1.f - sv3) * ks0);
        rendered[3 * n + 1] = a1 * (sv3 + (1.f - sv3) * ks1);
        rendered[3 * n + 2] = a2 * (sv3 + (1.f - sv3) * ks2);
      }
    }
  }
}

template <bool kClassic, bool kSkyPerSample>
__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float* __restrict__ rho, const float* __restrict__ deltas, const float* __restrict__ col,
                     const float* __restrict__ vis, const float* __restrict__ sky, int N, int S,
                     float* __restrict__ PV, float* __restrict__ PE, float* __restrict__ PS,
                     float* __restrict__ albedo, float* __restrict__ rendered, float* __restrict__ vis_sum) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    float carry = 0.f;
    RayAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float ks0 = 0, ks1 = 0, ks2 = 0;
    if (!kSkyPerSample) {
      ks0 = __ldg(sky + 3 * n), ks1 = __ldg(sky + 3 * n + 1), ks2 = __ldg(sky + 3 * n + 2);
    }
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < S;
      const long long o = base + s;
      const float y = ok ? __ldg(rho + o) * __ldg(deltas + o) : 0.f;
      const float incl = warp_scan_incl(y, lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      const float pv = expf(-(carry + prev));
      const float pe = 1.f - expf(-y);
      const float ps = pv * pe;
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (ok) {
        if (PV) PV[o] = pv;
        if (PE) PE[o] = pe;
        if (PS) PS[o] = ps;
        const float c0 = __ldg(col + 3 * o), c1 = __ldg(col + 3 * o + 1), c2 = __ldg(col + 3 * o + 2);
        const float v = __ldg(vis + o);
        acc.a0 += ps * c0, acc.a1 += ps * c1, acc.a2 += ps * c2;
        acc.vs += v * ps;
        float q0 = ks0, q1 = ks1, q2 = ks2;
        if (kSkyPerSample) {
          q0 = __ldg(sky + 3 * o), q1 = __ldg(sky + 3 * o + 1), q2 = __ldg(sky + 3 * o + 2);
          acc.k0 += q0, acc.k1 += q1, acc.k2 += q2;
        }
        if (kClassic) {
          acc.r0 += ps * c0 * (v + (1.f - v) * q0);
          acc.r1 += ps * c1 * (v + (1.f - v) * q1);
          acc.r2 += ps * c2 * (v + (1.f - v) * q2);
        }
      }
    }
    acc.a0 = warp_sum(acc.a0), acc.a1 = warp_sum(acc.a1), acc.a2 = warp_sum(acc.a2);
    acc.vs = warp_sum(acc.vs);
    if (kClassic) acc.r0 = warp_sum(acc.r0), acc.r1 = warp_sum(acc.r1), acc.r2 = warp_sum(acc.r2);
    if (kSkyPerSample) {
      ks0 = warp_sum(acc.k0) / (float)S, ks1 = warp_sum(acc.k1) / (float)S, ks2 = warp_sum(acc.k2) / (float)S;
    }
    if (lane == 0) {
      albedo[3 * n] = acc.a0, albedo[3 * n + 1] = acc.a1, albedo[3 * n + 2] = acc.a2;
      if (vis_sum) vis_sum[n] = acc.vs;
      if (kClassic) {
        rendered[3 * n] = acc.r0, rendered[3 * n + 1] = acc.r1, rendered[3 * n + 2] = acc.r2;
      } else {
        const float sv3 = sigmoidf_((acc.vs - .2f) * 30.f);  // Eval_Tools_2.py:214
        rendered[3 * n] = acc.a0 * (sv3 + (1.f - sv3) * ks0);
        rendered[3 * n + 1] = acc.a1 * (sv3 + (1.f - sv3) * ks1);
        rendered[3 * n + 2] = acc.a2 * (sv3 + (1.f - sv3) * ks2);
      }
    }
  }
}

// Fused backward.  Sweep 1 (forward order) rebuilds PV/PE per sample into registers and the per-ray sums;
// sweep 2 (reverse order) runs the suffix scan  sum_{t>s} dPV_t*PV_t  with shuffles and emits all gradients.
template <bool kClassic, bool kSkyPerSample, int kChunks>
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float* __restrict__ rho, const float* __restrict__ deltas, const float* __restrict__ col,
                     const float* __restrict__ vis, const float* __restrict__ sky, int N, int S,
                     const float* __restrict__ d_rendered, const float* __restrict__ d_albedo,
                     const float* __restrict__ dPE, const float* __restrict__ dPV, const float* __restrict__ dPS,
                     float* __restrict__ d_rho, float* __restrict__ d_col, float* __restrict__ d_sky,
                     float* __restrict__ d_vis) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    float pv[kChunks], pe[kChunks];
    float cc0[kChunks], cc1[kChunks], cc2[kChunks], vv[kChunks], dd[kChunks];   // colour / vis / delta of sweep 1, reused by sweep 2
    float carry = 0.f, a0 = 0, a1 = 0, a2 = 0, vs = 0, k0 = 0, k1 = 0, k2 = 0;
    if (!kSkyPerSample) k0 = __ldg(sky + 3 * n), k1 = __ldg(sky + 3 * n + 1), k2 = __ldg(sky + 3 * n + 2);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const long long o = base + s;
      dd[c] = ok ? __ldg(deltas + o) : 0.f;
      cc0[c] = ok ? __ldg(col + 3 * o) : 0.f, cc1[c] = ok ? __ldg(col + 3 * o + 1) : 0.f, cc2[c] = ok ? __ldg(col + 3 * o + 2) : 0.f;
      vv[c] = ok ? __ldg(vis + o) : 0.f;
      const float y = ok ? __ldg(rho + o) * dd[c] : 0.f;
      const float incl = warp_scan_incl(y, lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      pv[c] = expf(-(carry + prev));
      pe[c] = 1.f - expf(-y);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (ok && !kClassic) {
        const float ps = pv[c] * pe[c];
        a0 += ps * cc0[c], a1 += ps * cc1[c], a2 += ps * cc2[c];
        vs += vv[c] * ps;
        if (kSkyPerSample) k0 += __ldg(sky + 3 * o), k1 += __ldg(sky + 3 * o + 1), k2 += __ldg(sky + 3 * o + 2);
      }
    }
    float g0 = d_rendered ? __ldg(d_rendered + 3 * n) : 0.f, g1 = d_rendered ? __ldg(d_rendered + 3 * n + 1) : 0.f,
          g2 = d_rendered ? __ldg(d_rendered + 3 * n + 2) : 0.f;
    const float e0 = d_albedo ? __ldg(d_albedo + 3 * n) : 0.f, e1 = d_albedo ? __ldg(d_albedo + 3 * n + 1) : 0.f,
                e2 = d_albedo ? __ldg(d_albedo + 3 * n + 2) : 0.f;
    float dA0 = e0, dA1 = e1, dA2 = e2;  // gradient w.r.t. the albedo sum
    float dk0 = 0, dk1 = 0, dk2 = 0;     // gradient w.r.t. the mean sky colour (non-classic)
    float dvs = 0.f;                     // gradient w.r.t. sum_s vis*PS: vis is detached but PS is not (Eval_Tools_2.py:214)
    if (!kClassic) {
      a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2), vs = warp_sum(vs);
      if (kSkyPerSample) k0 = warp_sum(k0) / (float)S, k1 = warp_sum(k1) / (float)S, k2 = warp_sum(k2) / (float)S;
      const float sv3 = sigmoidf_((vs - .2f) * 30.f);
      dA0 += g0 * (sv3 + (1.f - sv3) * k0), dA1 += g1 * (sv3 + (1.f - sv3) * k1), dA2 += g2 * (sv3 + (1.f - sv3) * k2);
      dk0 = g0 * a0 * (1.f - sv3), dk1 = g1 * a1 * (1.f - sv3), dk2 = g2 * a2 * (1.f - sv3);
      dvs = (g0 * a0 * (1.f - k0) + g1 * a1 * (1.f - k1) + g2 * a2 * (1.f - k2)) * sv3 * (1.f - sv3) * 30.f;
      if (!kSkyPerSample && lane == 0 && d_sky) d_sky[3 * n] = dk0, d_sky[3 * n + 1] = dk1, d_sky[3 * n + 2] = dk2;
      if (kSkyPerSample) dk0 /= (float)S, dk1 /= (float)S, dk2 /= (float)S;
    }
    float suffix = 0.f;                    // sum over later chunks of dPV_t*PV_t
    float dsk0 = 0, dsk1 = 0, dsk2 = 0;    // classic + per-ray sky accumulation
#pragma unroll
    for (int c = kChunks - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const long long o = base + s;
      float q = 0.f, dpe_tot = 0.f;
      if (ok) {
        const float c0 = cc0[c], c1 = cc1[c], c2 = cc2[c];
        const float ps = pv[c] * pe[c];
        float dps = dPS ? __ldg(dPS + o) : 0.f;
        if (kClassic) {
          const float v = vv[c];
          float q0 = k0, q1 = k1, q2 = k2;
          if (kSkyPerSample) q0 = __ldg(sky + 3 * o), q1 = __ldg(sky + 3 * o + 1), q2 = __ldg(sky + 3 * o + 2);
          const float sh0 = v + (1.f - v) * q0, sh1 = v + (1.f - v) * q1, sh2 = v + (1.f - v) * q2;
          dps += g0 * c0 * sh0 + g1 * c1 * sh1 + g2 * c2 * sh2 + e0 * c0 + e1 * c1 + e2 * c2;
          d_col[3 * o] = ps * (g0 * sh0 + e0), d_col[3 * o + 1] = ps * (g1 * sh1 + e1), d_col[3 * o + 2] = ps * (g2 * sh2 + e2);
          if (d_vis) d_vis[o] = ps * (g0 * c0 * (1.f - q0) + g1 * c1 * (1.f - q1) + g2 * c2 * (1.f - q2));
          const float t0 = g0 * ps * c0 * (1.f - v), t1 = g1 * ps * c1 * (1.f - v), t2 = g2 * ps * c2 * (1.f - v);
          if (kSkyPerSample) {
            if (d_sky) d_sky[3 * o] = t0, d_sky[3 * o + 1] = t1, d_sky[3 * o + 2] = t2;
          } else {
            dsk0 += t0, dsk1 += t1, dsk2 += t2;
          }
        } else {
          dps += dA0 * c0 + dA1 * c1 + dA2 * c2 + dvs * vv[c];
          d_col[3 * o] = ps * dA0, d_col[3 * o + 1] = ps * dA1, d_col[3 * o + 2] = ps * dA2;
          if (kSkyPerSample && d_sky) d_sky[3 * o] = dk0, d_sky[3 * o + 1] = dk1, d_sky[3 * o + 2] = dk2;
        }
        dpe_tot = dps * pv[c] + (dPE ? __ldg(dPE + o) : 0.f);
        const float dpv_tot = dps * pe[c] + (dPV ? __ldg(dPV + o) : 0.f);
        q = dpv_tot * pv[c];
      }
      // exclusive suffix sum of q over lanes (reverse scan) + later chunks
      float incl = q;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, incl, off);
        if (lane + off < 32) incl += t;
      }
      const float excl = incl - q + suffix;
      suffix += __shfl_sync(0xffffffffu, incl, 0);
      if (ok) {
        const float dy = dpe_tot * (1.f - pe[c]) - excl;
        d_rho[o] = dy * dd[c];
      }
    }
    if (kClassic && !kSkyPerSample && d_sky) {
      dsk0 = warp_sum(dsk0), dsk1 = warp_sum(dsk1), dsk2 = warp_sum(dsk2);
      if (lane == 0) d_sky[3 * n] = dsk0, d_sky[3 * n + 1] = dsk1, d_sky[3 * n + 2] = dsk2;
    }
  }
}

// out[m] = exp(-sum_{k<S-1} rho[m,k]*deltas[m,k])   (mg_Img_Eval.py:68-70)
__global__ void __launch_bounds__(256) march_transmittance_kernel(const float* __restrict__ rho,
                                                                  const float* __restrict__ deltas, long long M, int S,
                                                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long m = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); m < M; m += (long long)gridDim.x * wpb) {
    float acc = 0.f;
    for (int s = lane; s < S - 1; s += 32) acc += __ldg(rho + m * S + s) * __ldg(deltas + m * S + s);
    acc = warp_sum(acc);
    if (lane == 0) out[m] = expf(-acc);
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double sigd(double x) { return 1.0 / (1.0 + exp(-x)); }

// float64 exclusive-prefix transmittance * emission for one chunk of 32 samples
__device__ __forceinline__ double ps_chunk_d(double y, int lane, double& carry) {
  double incl = y;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  double prev = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) prev = 0.0;
  const double ps = exp(-(carry + prev)) * (1.0 - exp(-y));
  carry += __shfl_sync(0xffffffffu, incl, 31);
  return ps;
}

// mg_Img_Eval.py:123-190 in float64: Base_Img, Season_Adj_Img, per-class extreme images, raw shadow sums.
template <typename T>
__global__ void __launch_bounds__(128)
cli_composite_kernel(const T* __restrict__ rho, const T* __restrict__ deltas, const T* __restrict__ base,
                     const T* __restrict__ vis, const T* __restrict__ adj, const double* __restrict__ cls,
                     const T* __restrict__ exact_vis, int N, int S, int C, double* __restrict__ base_img,
                     double* __restrict__ season_img, double* __restrict__ extreme, double* __restrict__ raw_shadow,
                     double* __restrict__ raw_shadow_exact) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    double carry = 0.0, b[3] = {0, 0, 0}, se[3] = {0, 0, 0}, sh = 0, she = 0;
    double ex[8][3];
#pragma unroll
    for (int c = 0; c < 8; ++c) ex[c][0] = ex[c][1] = ex[c][2] = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < S;
      const long long o = (long long)n * S + s;
      const double ps = ps_chunk_d(ok ? (double)rho[o] * (double)deltas[o] : 0.0, lane, carry);
      if (ok) {
        sh += ps * (double)vis[o];
        if (exact_vis) she += ps * (double)exact_vis[o];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double bc = (double)base[3 * o + d];
          b[d] += ps * sigd(bc);
          double mix = 0;
          for (int c = 0; c < C; ++c) {
            const double a = (double)adj[(o * C + c) * 3 + d];
            mix += cls[c] * a;
            if (c < 8) ex[c][d] += ps * sigd(bc + a);
          }
          se[d] += ps * sigd(bc + mix);
        }
      }
    }
    sh = warp_sum_d(sh);
    if (exact_vis) she = warp_sum_d(she);
#pragma unroll
    for (int d = 0; d < 3; ++d) b[d] = warp_sum_d(b[d]), se[d] = warp_sum_d(se[d]);
    for (int c = 0; c < C && c < 8; ++c)
#pragma unroll
      for (int d = 0; d < 3; ++d) ex[c][d] = warp_sum_d(ex[c][d]);
    if (lane == 0) {
      raw_shadow[n] = sh;
      if (exact_vis && raw_shadow_exact) raw_shadow_exact[n] = she;
      for (int d = 0; d < 3; ++d) base_img[3 * n + d] = b[d], season_img[3 * n + d] = se[d];
      for (int c = 0; c < C && c < 8; ++c)
        for (int d = 0; d < 3; ++d) extreme[((long long)c * N + n) * 3 + d] = ex[c][d];
    }
  }
}

// The output image of the render CLI (main_run_Season_NeRF.py:90-92 = Season_Adj_Img * Shadow_Adjust) straight from the
// RAW network heads of a block of rays: pos4 [M,4] = (sigma, base colour logits), vis_raw [M], adj [M,C,3], deltas [M]
// (already zero outside the cube).  softplus / sigmoid are applied in float32 like the network's own activations
// (T_NeRF_net_v2.py:139-151), the sums run in float64 like get_imgs_from_Img_Dict (mg_Img_Eval.py:123-160); the base /
// extreme images of the component API are not formed (3 float64 sigmoids per sample instead of 18), and the eight
// element-wise torch launches + activated copies of the component path (render.py _internal_render) disappear.
__global__ void __launch_bounds__(128)
render_composite_raw_kernel(const float* __restrict__ pos4, const float* __restrict__ vis_raw, const float* __restrict__ adj,
                            const float* __restrict__ deltas, const double* __restrict__ cls,
                            const float* __restrict__ exact_vis, int N, int S, int C, double* __restrict__ season_img,
                            double* __restrict__ raw_shadow, double* __restrict__ raw_shadow_exact) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  double cw[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) cw[c] = c < C ? cls[c] : 0.0;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    double carry = 0.0, se[3] = {0, 0, 0}, sh = 0, she = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < S;
      const long long o = (long long)n * S + s;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) p = __ldg(reinterpret_cast<const float4*>(pos4) + o);
      const double ps = ps_chunk_d(ok ? (double)softplusf_(p.x) * (double)deltas[o] : 0.0, lane, carry);
      if (ok) {
        sh += ps * (double)sigmoidf_(__ldg(vis_raw + o));
        if (exact_vis) she += ps * (double)exact_vis[o];
        const float bc[3] = {p.y, p.z, p.w};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double mix = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < C) mix += cw[c] * (double)__ldg(adj + (o * C + c) * 3 + d);
          se[d] += ps * sigd((double)bc[d] + mix);
        }
      }
    }
    sh = warp_sum_d(sh);
    if (exact_vis) she = warp_sum_d(she);
#pragma unroll
    for (int d = 0; d < 3; ++d) se[d] = warp_sum_d(se[d]);
    if (lane == 0) {
      raw_shadow[n] = sh;
      if (exact_vis && raw_shadow_exact) raw_shadow_exact[n] = she;
      for (int d = 0; d < 3; ++d) season_img[3 * n + d] = se[d];
    }
  }
}

// mg_Img_Eval.py:166-181 (use_classic_shadows): classic[n,:] = sum_s PS * sigmoid(base + class . adjust) * (vis + (1 - vis) * sky)
// with the per-sample sky colour of the component dict, float64 sums like the reference's numpy.
template <typename T>
__global__ void __launch_bounds__(128)
cli_classic_shadow_kernel(const T* __restrict__ rho, const T* __restrict__ deltas, const T* __restrict__ base,
                          const T* __restrict__ vis, const T* __restrict__ adj, const T* __restrict__ sky,
                          const double* __restrict__ cls, int N, int S, int C, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    double carry = 0.0, acc[3] = {0, 0, 0};
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool ok = s < S;
      const long long o = (long long)n * S + s;
      const double ps = ps_chunk_d(ok ? (double)rho[o] * (double)deltas[o] : 0.0, lane, carry);
      if (ok) {
        const double v = (double)vis[o];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double mix = 0;
          for (int c = 0; c < C; ++c) mix += cls[c] * (double)adj[(o * C + c) * 3 + d];
          const double col = sigd((double)base[3 * o + d] + mix) * (v + (1.0 - v) * (double)sky[3 * o + d]);
          acc[d] += ps * col;
        }
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[d] = warp_sum_d(acc[d]);
    if (lane == 0)
      for (int d = 0; d < 3; ++d) out[3 * (long long)n + d] = acc[d];
  }
}

// mg_Img_Eval.py:192-228 fused over the T class vectors: PS, base and adjust are read ONCE per ray and kept
// in registers (3 samples per lane at S=96); the T recombinations run out of registers.
//   out[t, n, :] = shade[n, :] * sum_s PS[n,s] * sigmoid(base[n,s,:] + sum_c cls[t,c] * adj[n,s,c,:])
// Arithmetic type of the T-loop = element type of the components: float32 network outputs are recombined in float32
// (class mix, sigmoid and the three per-lane terms; |error| < 3e-7 on a [0,1] colour) and reduced across the warp in
// float64; float64 components (arrays a caller modified on the host) keep the reference's float64 numpy arithmetic.
// The transmittance scan is float64 in both cases.
template <typename CT> __device__ __forceinline__ CT sig_ct(CT x);
template <> __device__ __forceinline__ double sig_ct<double>(double x) { return 1.0 / (1.0 + exp(-x)); }
template <> __device__ __forceinline__ float sig_ct<float>(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

template <typename TI, int kChunks>
__global__ void __launch_bounds__(128)
year_sweep_kernel(const TI* __restrict__ rho, const TI* __restrict__ deltas, const TI* __restrict__ base,
                  const TI* __restrict__ adj, const double* __restrict__ cls, const double* __restrict__ shade,
                  const TI* __restrict__ ps_weight, int N, int S, int C, int T, double* __restrict__ out) {
  typedef TI CT;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  extern __shared__ double cls_s[];          // [T*C] class vectors, converted once per block
  CT* wsm = reinterpret_cast<CT*>(cls_s);
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) wsm[i] = (CT)cls[i];
  __syncthreads();
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    double carry = 0.0;
    CT ps[kChunks], bc[kChunks][3], ad[kChunks][4][3];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const long long o = (long long)n * S + s;
      double psd = ps_chunk_d(ok ? (double)rho[o] * (double)deltas[o] : 0.0, lane, carry);
      if (ps_weight && ok) psd *= (double)ps_weight[o];     // classic-shadow alignment: PS * vis (mg_Img_Eval.py:448-449)
      ps[c] = ok ? (CT)psd : (CT)0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        bc[c][d] = ok ? (CT)base[3 * o + d] : (CT)0;
#pragma unroll
        for (int k = 0; k < 4; ++k) ad[c][k][d] = (ok && k < C) ? (CT)adj[(o * C + k) * 3 + d] : (CT)0;
      }
    }
    double sh[3] = {1.0, 1.0, 1.0};
    if (shade) sh[0] = shade[3 * (long long)n], sh[1] = shade[3 * (long long)n + 1], sh[2] = shade[3 * (long long)n + 2];
    for (int t = 0; t < T; ++t) {
      CT w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = k < C ? wsm[t * C + k] : (CT)0;
      CT acc[3] = {0, 0, 0};
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const CT mix = w[0] * ad[c][0][d] + w[1] * ad[c][1][d] + w[2] * ad[c][2][d] + w[3] * ad[c][3][d];
          acc[d] += ps[c] * sig_ct<CT>(bc[c][d] + mix);
        }
      double r[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) r[d] = warp_sum_d((double)acc[d]);
      if (lane < 3) out[((long long)t * N + n) * 3 + lane] = (lane == 0 ? r[0] : lane == 1 ? r[1] : r[2]) * sh[lane];
    }
  }
}


// ---- year sweep, float32 components: one lane per TIME STEP ------------------------------------------------------------
// The kernel above gives each lane 3 samples of the ray and pays a cross-lane float64 reduction (3 x 5 double shuffles) per
// (ray, time step); ncu (profiles/r02_ncu_year_sweep_v1.txt): 97 ms for 365 x 1024^2, issue slots 76 % busy, XU 57 % -
// instruction-issue bound.  Here a warp still owns one ray, but its lanes own 32 different time steps: the ray's
// 96 x (PS, base[3], adj[4][3]) land in shared memory once (6 KB per warp), every lane walks the samples in order and
// accumulates its own time steps, so there is NO cross-lane reduction, the ray data arrive as 4 broadcast LDS.128 per sample
// for kTJ time steps per lane, and the three sigmoids of a sample share one reciprocal:
//     1/(1+a), 1/(1+b), 1/(1+c) = r*(1+b)(1+c), r*(1+a)(1+c), r*(1+a)(1+b),  r = 1/((1+a)(1+b)(1+c))
// (4 MUFU per 3 sigmoids instead of 6; the logit is clamped at -28 so that the product stays finite: sigmoid(-28) = 7e-13).
// Base colours and class vectors are pre-scaled by -log2(e): the class mix lands directly in the exponent of ex2.approx.
// Per-lane sums run in float32 over 16 samples and are flushed into float64 accumulators (|error| of a [0,1] colour < 5e-7).
// kSweepTJ = time steps per lane and pass (32 * kSweepTJ per pass: 4 -> three passes for a year)
__device__ __forceinline__ float ex2_approx(float x) {      // one MUFU.EX2 (2 ulp), no range fix-up: |x| <= 28*log2(e) here
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kSweepTJ>
__global__ void __launch_bounds__(128)
year_sweep_lanes_kernel(const float* __restrict__ rho, const float* __restrict__ deltas, const float* __restrict__ base,
                        const float* __restrict__ adj, const double* __restrict__ cls, const double* __restrict__ shade,
                        const float* __restrict__ ps_weight, int raw_pos4, int N, int S, int C, int T, int T_pad,
                        double* __restrict__ out) {
  // raw_pos4: `rho` points at the network's raw pos4 [M,4] = (sigma, base colour logits) and `base` is unused: softplus is
  // applied here (float32, like the network's own activation) and no activated copy of the heads is ever made
  extern __shared__ float4 sweep_sm[];
  float4* wsm = sweep_sm;                                        // [T_pad] class vectors * -log2(e), zero beyond T and C
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float4* ray = sweep_sm + T_pad + (size_t)warp * S * 4;         // [S][4] float4: (PS, b0', b1', b2'), adj[0..11]
  float* rayf = reinterpret_cast<float*>(ray);
  const float kNegLog2e = -1.4426950408889634f;
  for (int i = threadIdx.x; i < T_pad; i += blockDim.x) {
    float w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = (i < T && k < C) ? kNegLog2e * (float)cls[(long long)i * C + k] : 0.f;
    wsm[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const int chunks = (S + 31) >> 5;
  for (int n = blockIdx.x * wpb + warp; n < N; n += gridDim.x * wpb) {
    // ---- stage the ray: float64 transmittance scan (as above), then the 16 floats of every sample ----
    double carry = 0.0;
    for (int c = 0; c < chunks; ++c) {
      const int s_ = c * 32 + lane;
      const bool ok = s_ < S;
      const long long o = (long long)n * S + s_;
      float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) p4 = raw_pos4 ? __ldg(reinterpret_cast<const float4*>(rho) + o)
                            : make_float4(rho[o], base[3 * o + 0], base[3 * o + 1], base[3 * o + 2]);
      const float rho_s = raw_pos4 ? softplusf_(p4.x) : p4.x;
      double psd = ps_chunk_d(ok ? (double)rho_s * (double)deltas[o] : 0.0, lane, carry);
      if (ps_weight && ok) psd *= (double)ps_weight[o];
      if (ok) {
        rayf[s_ * 16 + 0] = (float)psd;
        rayf[s_ * 16 + 1] = kNegLog2e * p4.y;
        rayf[s_ * 16 + 2] = kNegLog2e * p4.z;
        rayf[s_ * 16 + 3] = kNegLog2e * p4.w;
      }
    }
    {
      // adj [S][C][3] is contiguous per ray: coalesced reads, scattered into the 12 floats after (PS, base) of each sample
      const float* a = adj + (long long)n * S * C * 3;
      const int per = C * 3;
      for (int i = lane; i < S * per; i += 32) {
        const int s_ = i / per, r = i - s_ * per;
        rayf[s_ * 16 + 4 + r] = a[i];
      }
      if (C < 4)
        for (int i = lane; i < S * (12 - per); i += 32) {
          const int s_ = i / (12 - per), r = i - s_ * (12 - per);
          rayf[s_ * 16 + 4 + per + r] = 0.f;
        }
    }
    double sh[3] = {1.0, 1.0, 1.0};
    if (shade) sh[0] = shade[3 * (long long)n], sh[1] = shade[3 * (long long)n + 1], sh[2] = shade[3 * (long long)n + 2];
    __syncwarp();
    for (int t0 = 0; t0 < T; t0 += 32 * kSweepTJ) {
      float4 w[kSweepTJ];
      float acc[kSweepTJ][3];
      double accd[kSweepTJ][3];
#pragma unroll
      for (int j = 0; j < kSweepTJ; ++j) {
        w[j] = wsm[t0 + j * 32 + lane];
#pragma unroll
        for (int d = 0; d < 3; ++d) acc[j][d] = 0.f, accd[j][d] = 0.0;
      }
      for (int s_ = 0; s_ < S; ++s_) {
        const float4 p = ray[s_ * 4], a0 = ray[s_ * 4 + 1], a1 = ray[s_ * 4 + 2], a2 = ray[s_ * 4 + 3];
#pragma unroll
        for (int j = 0; j < kSweepTJ; ++j) {
          // x_d = -log2(e) * (base_d + sum_k cls_k * adj[k][d]);  adj floats: k0:(a0.x a0.y a0.z) k1:(a0.w a1.x a1.y)
          // k2:(a1.z a1.w a2.x) k3:(a2.y a2.z a2.w)
          float x0 = fmaf(w[j].x, a0.x, fmaf(w[j].y, a0.w, fmaf(w[j].z, a1.z, fmaf(w[j].w, a2.y, p.y))));
          float x1 = fmaf(w[j].x, a0.y, fmaf(w[j].y, a1.x, fmaf(w[j].z, a1.w, fmaf(w[j].w, a2.z, p.z))));
          float x2 = fmaf(w[j].x, a0.z, fmaf(w[j].y, a1.y, fmaf(w[j].z, a2.x, fmaf(w[j].w, a2.w, p.w))));
          const float kClamp = 28.f * 1.4426950408889634f;          // -log2(e) * (-28)
          x0 = fminf(x0, kClamp), x1 = fminf(x1, kClamp), x2 = fminf(x2, kClamp);
          const float A = 1.f + ex2_approx(x0), B = 1.f + ex2_approx(x1), Cc = 1.f + ex2_approx(x2);
          const float AB = A * B;
          const float r = rcp_approx(AB * Cc);
          acc[j][0] = fmaf(p.x, r * (B * Cc), acc[j][0]);
          acc[j][1] = fmaf(p.x, r * (A * Cc), acc[j][1]);
          acc[j][2] = fmaf(p.x, r * AB, acc[j][2]);
        }
        if ((s_ & 15) == 15 || s_ == S - 1) {
#pragma unroll
          for (int j = 0; j < kSweepTJ; ++j)
#pragma unroll
            for (int d = 0; d < 3; ++d) accd[j][d] += (double)acc[j][d], acc[j][d] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < kSweepTJ; ++j) {
        const int t = t0 + j * 32 + lane;
        if (t < T) {
          double* o = out + ((long long)t * N + n) * 3;
          o[0] = accd[j][0] * sh[0], o[1] = accd[j][1] * sh[1], o[2] = accd[j][2] * sh[2];
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace snb

using namespace snb;

extern "C" int snb_composite_fwd(const float* rho, const float* deltas, const float* col, const float* vis,
                                 const float* sky, int sky_per_sample, int N, int S, int classic, float* PV, float* PE,
                                 float* PS, float* albedo, float* rendered, float* vis_sum, void* stream) {
  SNB_CHECK_ARG(rho && deltas && col && vis && sky && albedo && rendered && N >= 0 && S > 0);
  if (N == 0) return SNB_OK;
  const int grid = grid_for(N, 8, 16);
  cudaStream_t st = (cudaStream_t)stream;
#define SNB_FWD(C_, K_) \
  composite_fwd_kernel<C_, K_><<<grid, 256, 0, st>>>(rho, deltas, col, vis, sky, N, S, PV, PE, PS, albedo, rendered, vis_sum)
#define SNB_FWDU(C_, CH) \
  composite_fwd_unrolled_kernel<C_, CH><<<grid, 256, 0, st>>>(rho, deltas, col, vis, sky, N, S, PV, PE, PS, albedo, rendered, vis_sum)
  const int chunks = (S + 31) / 32;
  if (!sky_per_sample && chunks <= 4) {
    if (classic) { if (chunks == 1) SNB_FWDU(true, 1); else if (chunks == 2) SNB_FWDU(true, 2); else if (chunks == 3) SNB_FWDU(true, 3); else SNB_FWDU(true, 4); }
    else { if (chunks == 1) SNB_FWDU(false, 1); else if (chunks == 2) SNB_FWDU(false, 2); else if (chunks == 3) SNB_FWDU(false, 3); else SNB_FWDU(false, 4); }
  } else if (classic) { if (sky_per_sample) SNB_FWD(true, true); else SNB_FWD(true, false); }
  else { if (sky_per_sample) SNB_FWD(false, true); else SNB_FWD(false, false); }
#undef SNB_FWD
#undef SNB_FWDU
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

template <bool C_, bool K_>
static int launch_bwd(int chunks, int grid, cudaStream_t st, const float* rho, const float* deltas, const float* col,
                      const float* vis, const float* sky, int N, int S, const float* d_rendered, const float* d_albedo,
                      const float* dPE, const float* dPV, const float* dPS, float* d_rho, float* d_col, float* d_sky,
                      float* d_vis) {
#define SNB_BWD(CH) \
  composite_bwd_kernel<C_, K_, CH><<<grid, 256, 0, st>>>(rho, deltas, col, vis, sky, N, S, d_rendered, d_albedo, dPE, dPV, dPS, d_rho, d_col, d_sky, d_vis)
  if (chunks <= 1) SNB_BWD(1);
  else if (chunks == 2) SNB_BWD(2);
  else if (chunks == 3) SNB_BWD(3);
  else if (chunks == 4) SNB_BWD(4);
  else SNB_BWD(8);
#undef SNB_BWD
  return 0;
}

extern "C" int snb_composite_bwd(const float* rho, const float* deltas, const float* col, const float* vis,
                                 const float* sky, int sky_per_sample, int N, int S, int classic,
                                 const float* d_rendered, const float* d_albedo, const float* dPE, const float* dPV,
                                 const float* dPS, float* d_rho, float* d_col, float* d_sky, float* d_vis,
                                 void* stream) {
  SNB_CHECK_ARG(rho && deltas && col && vis && sky && d_rho && d_col && N >= 0 && S > 0);
  if (S > 32 * kMaxChunks) return SNB_ERR_UNSUPPORTED;
  if (N == 0) return SNB_OK;
  const int chunks = (S + 31) / 32;
  const int grid = grid_for(N, 8, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (classic) {
    if (sky_per_sample) launch_bwd<true, true>(chunks, grid, st, rho, deltas, col, vis, sky, N, S, d_rendered, d_albedo, dPE, dPV, dPS, d_rho, d_col, d_sky, d_vis);
    else launch_bwd<true, false>(chunks, grid, st, rho, deltas, col, vis, sky, N, S, d_rendered, d_albedo, dPE, dPV, dPS, d_rho, d_col, d_sky, d_vis);
  } else {
    if (sky_per_sample) launch_bwd<false, true>(chunks, grid, st, rho, deltas, col, vis, sky, N, S, d_rendered, d_albedo, dPE, dPV, dPS, d_rho, d_col, d_sky, d_vis);
    else launch_bwd<false, false>(chunks, grid, st, rho, deltas, col, vis, sky, N, S, d_rendered, d_albedo, dPE, dPV, dPS, d_rho, d_col, d_sky, d_vis);
  }
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_march_transmittance(const float* rho, const float* deltas, long long M, int S, float* out,
                                       void* stream) {
  SNB_CHECK_ARG(rho && deltas && out && M >= 0 && S > 0);
  if (M == 0) return SNB_OK;
  march_transmittance_kernel<<<grid_for(M, 8, 16), 256, 0, (cudaStream_t)stream>>>(rho, deltas, M, S, out);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_cli_composite(const void* rho, const void* deltas, const void* base, const void* vis,
                                 const void* adj, const double* cls, const void* exact_vis, int in_dtype, int N, int S,
                                 int C, double* base_img, double* season_img, double* extreme, double* raw_shadow,
                                 double* raw_shadow_exact, void* stream) {
  SNB_CHECK_ARG(rho && deltas && base && vis && adj && cls && base_img && season_img && extreme && raw_shadow);
  SNB_CHECK_ARG(N >= 0 && S > 0 && C >= 1 && C <= 8 && (in_dtype == SNB_F32 || in_dtype == SNB_F64));
  if (N == 0) return SNB_OK;
  const int grid = grid_for(N, 4, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == SNB_F64)
    cli_composite_kernel<double><<<grid, 128, 0, st>>>((const double*)rho, (const double*)deltas, (const double*)base,
        (const double*)vis, (const double*)adj, cls, (const double*)exact_vis, N, S, C, base_img, season_img, extreme,
        raw_shadow, raw_shadow_exact);
  else
    cli_composite_kernel<float><<<grid, 128, 0, st>>>((const float*)rho, (const float*)deltas, (const float*)base,
        (const float*)vis, (const float*)adj, cls, (const float*)exact_vis, N, S, C, base_img, season_img, extreme,
        raw_shadow, raw_shadow_exact);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_render_composite_raw(const float* pos4, const float* vis_raw, const float* adj, const float* deltas,
                                        const double* cls, const float* exact_vis, int N, int S, int C, double* season_img,
                                        double* raw_shadow, double* raw_shadow_exact, void* stream) {
  SNB_CHECK_ARG(pos4 && vis_raw && adj && deltas && cls && season_img && raw_shadow && N >= 0 && S > 0);
  SNB_CHECK_ARG((((uintptr_t)pos4) & 15) == 0);
  if (C < 1 || C > 4) return SNB_ERR_UNSUPPORTED;
  if (N == 0) return SNB_OK;
  render_composite_raw_kernel<<<grid_for(N, 4, 16), 128, 0, (cudaStream_t)stream>>>(pos4, vis_raw, adj, deltas, cls, exact_vis, N, S, C,
                                                                                   season_img, raw_shadow, raw_shadow_exact);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_cli_classic_shadow(const void* rho, const void* deltas, const void* base, const void* vis, const void* adj,
                                      const void* sky, const double* cls, int in_dtype, int N, int S, int C, double* out,
                                      void* stream) {
  SNB_CHECK_ARG(rho && deltas && base && vis && adj && sky && cls && out);
  SNB_CHECK_ARG(N >= 0 && S > 0 && C >= 1 && C <= 8 && (in_dtype == SNB_F32 || in_dtype == SNB_F64));
  if (N == 0) return SNB_OK;
  const int grid = grid_for(N, 4, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == SNB_F64)
    cli_classic_shadow_kernel<double><<<grid, 128, 0, st>>>((const double*)rho, (const double*)deltas, (const double*)base,
        (const double*)vis, (const double*)adj, (const double*)sky, cls, N, S, C, out);
  else
    cli_classic_shadow_kernel<float><<<grid, 128, 0, st>>>((const float*)rho, (const float*)deltas, (const float*)base,
        (const float*)vis, (const float*)adj, (const float*)sky, cls, N, S, C, out);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

template <typename TI>
static void launch_sweep(int chunks, int grid, cudaStream_t st, const void* rho, const void* deltas, const void* base,
                         const void* adj, const double* cls, const double* shade, const void* ps_weight, int N, int S, int C,
                         int T, double* out) {
  const TI *r = (const TI*)rho, *d = (const TI*)deltas, *b = (const TI*)base, *a = (const TI*)adj, *w = (const TI*)ps_weight;
  const size_t sm = (size_t)T * C * sizeof(double);
  if (chunks <= 1) year_sweep_kernel<TI, 1><<<grid, 128, sm, st>>>(r, d, b, a, cls, shade, w, N, S, C, T, out);
  else if (chunks == 2) year_sweep_kernel<TI, 2><<<grid, 128, sm, st>>>(r, d, b, a, cls, shade, w, N, S, C, T, out);
  else if (chunks == 3) year_sweep_kernel<TI, 3><<<grid, 128, sm, st>>>(r, d, b, a, cls, shade, w, N, S, C, T, out);
  else year_sweep_kernel<TI, 4><<<grid, 128, sm, st>>>(r, d, b, a, cls, shade, w, N, S, C, T, out);
}

static int year_sweep_impl(const void* rho, const void* deltas, const void* base, const void* adj, const double* cls,
                           const double* shade, const void* ps_weight, int in_dtype, int raw_pos4, int N, int S, int C, int T,
                           double* out, void* stream);

extern "C" int snb_year_sweep(const void* rho, const void* deltas, const void* base, const void* adj,
                              const double* cls, const double* shade, const void* ps_weight, int in_dtype, int N, int S,
                              int C, int T, double* out, void* stream) {
  SNB_CHECK_ARG(base);
  return year_sweep_impl(rho, deltas, base, adj, cls, shade, ps_weight, in_dtype, 0, N, S, C, T, out, stream);
}

extern "C" int snb_year_sweep_raw(const float* pos4, const float* deltas, const float* adj, const double* cls,
                                  const double* shade, int N, int S, int C, int T, double* out, void* stream) {
  SNB_CHECK_ARG(pos4 && (((uintptr_t)pos4) & 15) == 0);
  return year_sweep_impl(pos4, deltas, pos4, adj, cls, shade, nullptr, SNB_F32, 1, N, S, C, T, out, stream);
}

static int year_sweep_impl(const void* rho, const void* deltas, const void* base, const void* adj, const double* cls,
                           const double* shade, const void* ps_weight, int in_dtype, int raw_pos4, int N, int S, int C, int T,
                           double* out, void* stream) {
  SNB_CHECK_ARG(rho && deltas && base && adj && cls && out && N >= 0 && S > 0 && T >= 0);
  SNB_CHECK_ARG(in_dtype == SNB_F32 || in_dtype == SNB_F64);
  if (C < 1 || C > 4 || S > 128 || (long long)T * C * 8 > 40 * 1024) return SNB_ERR_UNSUPPORTED;
  if (N == 0 || T == 0) return SNB_OK;
  if (in_dtype == SNB_F32) {
    // float32 components (the resident network outputs): lane-per-time-step kernel
    static int tj = 0;
    if (tj == 0) {
      // A/B switch for measurements: 4 (default), 6 or 8 time steps per lane.  Measured for 365 x 1024^2 on B200
      // (profiles/r02_year_sweep_tj.txt): 54.6 / 60.2 / 77.8 ms - occupancy (96 / 128 / 168 registers) beats the amortised
      // ray loads
      const char* e = getenv("SNB_SWEEP_TJ");
      tj = e ? atoi(e) : 4;
      if (tj != 6 && tj != 8) tj = 4;
    }
    const int per_pass = 32 * tj;
    const int T_pad = (T + per_pass - 1) / per_pass * per_pass;
    const size_t sm = ((size_t)T_pad + (size_t)4 * S * 4) * sizeof(float4);          // class vectors + 4 warps x [S][4] float4
    if (sm > 96 * 1024) return SNB_ERR_UNSUPPORTED;
    const int grid = grid_for(N, 4, 16);
#define SNB_SWEEP_LANES(TJ)                                                                                                  \
  do {                                                                                                                       \
    static bool attr_set = false;                                                                                            \
    if (!attr_set) {                                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(year_sweep_lanes_kernel<TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
      if (e != cudaSuccess) return (int)e;                                                                                   \
      attr_set = true;                                                                                                       \
    }                                                                                                                        \
    year_sweep_lanes_kernel<TJ><<<grid, 128, sm, (cudaStream_t)stream>>>(                                                    \
        (const float*)rho, (const float*)deltas, (const float*)base, (const float*)adj, cls, shade, (const float*)ps_weight,     \
        raw_pos4, N, S, C, T, T_pad, out);                                                                                                   \
  } while (0)
    if (tj == 4) SNB_SWEEP_LANES(4);
    else if (tj == 8) SNB_SWEEP_LANES(8);
    else SNB_SWEEP_LANES(6);
#undef SNB_SWEEP_LANES
    count_launch();
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  const int grid = grid_for(N, 4, 16);
  const int chunks = (S + 31) / 32;
  if (in_dtype == SNB_F64) launch_sweep<double>(chunks, grid, (cudaStream_t)stream, rho, deltas, base, adj, cls, shade, ps_weight, N, S, C, T, out);
  else launch_sweep<float>(chunks, grid, (cudaStream_t)stream, rho, deltas, base, adj, cls, shade, ps_weight, N, S, C, T, out);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
