// Head activations fused with the compositing scan (training fast path and sharded render).
//
// The network leaves RAW heads in HBM: pos4 [M,4] = (sigma, colour logits), vis [M], adj [M,C,3] per sample point and
// sky [N,3], class logits [N,C] per ray.  The reference activates them with a dozen element-wise torch ops
// (T_NeRF_net_v2.py:87-98: softplus, softmax, the class mix sum_c adj[c]*class[c], three sigmoids) and composites the
// activated copies (Eval_Tools_2.py:187-215).  Here one warp per ray reads the raw heads ONCE (68 B per sample), applies
// the activations in registers and runs the transmittance scan; the backward kernel recomputes them and emits the
// gradients of the raw heads directly (softplus', sigmoid', softmax Jacobian included) - no activated copy, no
// per-op autograd node, ~65 small launches less per training step.
//
//   heads_composite   image pass: albedo, rendered colour (both solar conventions), activated sky, sum vis*PS
//   solar_loss        solar pass (Eval_Tools_2.py:297-337 + :353-368): per-ray  sum_s (vis - PV)^2  and  1 - sum_s PE*PV*vis
#include "common.cuh"
#include "api.h"

namespace snb {

constexpr int kHeadsMaxChunks = 4;     // S <= 128 (register-resident backward)

__device__ __forceinline__ void softmax4(const float* __restrict__ logits, int C, float (&cl)[4]) {
  float m = -3.0e38f;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < C) m = fmaxf(m, __ldg(logits + c));
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cl[c] = c < C ? expf(__ldg(logits + c) - m) : 0.f;
    sum += cl[c];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int c = 0; c < 4; ++c) cl[c] *= inv;
}

// raw heads of one sample -> activated (rho, colour[3], vis); adj12 = this sample's [C,3] block (zero padded to 4 classes)
struct SampleRaw {
  float4 p;          // sigma_raw, c0, c1, c2
  float a[12];       // adj[c*3 + d]
  float v, d;        // vis_raw, delta
};

struct HeadPitch {
  int pos, vis, adj;      // row pitch (floats) of the raw head matrices: contiguous [M,4] / [M] / [M,3C] = 4 / 1 / 3C; the
};                        // layer-wise training path hands over 16-float-wide padded rows

__device__ __forceinline__ void load_sample(SampleRaw& r, bool ok, long long o, const float* __restrict__ pos,
                                            const float* __restrict__ vis, const float* __restrict__ adj,
                                            const float* __restrict__ deltas, int C, const HeadPitch ld) {
  if (!ok) {
    r.p = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 12; ++i) r.a[i] = 0.f;
    r.v = 0.f, r.d = 0.f;
    return;
  }
  r.p = __ldg(reinterpret_cast<const float4*>(pos + o * ld.pos));
  r.v = __ldg(vis + o * ld.vis);
  r.d = __ldg(deltas + o);
  if (C == 4) {
    const float4* a4 = reinterpret_cast<const float4*>(adj + o * ld.adj);
    const float4 x = __ldg(a4), y = __ldg(a4 + 1), z = __ldg(a4 + 2);
    r.a[0] = x.x, r.a[1] = x.y, r.a[2] = x.z, r.a[3] = x.w, r.a[4] = y.x, r.a[5] = y.y, r.a[6] = y.z, r.a[7] = y.w;
    r.a[8] = z.x, r.a[9] = z.y, r.a[10] = z.z, r.a[11] = z.w;
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) r.a[i] = i < 3 * C ? __ldg(adj + o * ld.adj + i) : 0.f;
  }
}

__device__ __forceinline__ void activate(const SampleRaw& r, const float (&cl)[4], float& rho, float& c0, float& c1,
                                         float& c2, float& v) {
  rho = softplusf_(r.p.x);
  const float m0 = cl[0] * r.a[0] + cl[1] * r.a[3] + cl[2] * r.a[6] + cl[3] * r.a[9];
  const float m1 = cl[0] * r.a[1] + cl[1] * r.a[4] + cl[2] * r.a[7] + cl[3] * r.a[10];
  const float m2 = cl[0] * r.a[2] + cl[1] * r.a[5] + cl[2] * r.a[8] + cl[3] * r.a[11];
  c0 = sigmoidf_(r.p.y + m0), c1 = sigmoidf_(r.p.z + m1), c2 = sigmoidf_(r.p.w + m2);
  v = sigmoidf_(r.v);
}

template <bool kClassic, int kChunks>
__global__ void __launch_bounds__(256)
heads_composite_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ vis_raw, const float* __restrict__ adj,
                           const float* __restrict__ sky_raw, const float* __restrict__ cls_logits,
                           const float* __restrict__ deltas, const HeadPitch ld, int N, int S, int C, float* __restrict__ albedo,
                           float* __restrict__ rendered, float* __restrict__ sky_act, float* __restrict__ vis_sum,
                           float* __restrict__ PV, float* __restrict__ PE, float* __restrict__ PS) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    SampleRaw raw[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) load_sample(raw[c], c * 32 + lane < S, base + c * 32 + lane, pos, vis_raw, adj, deltas, C, ld);
    float cl[4];
    softmax4(cls_logits + (long long)n * C, C, cl);
    const float ks0 = sigmoidf_(__ldg(sky_raw + 3 * n)), ks1 = sigmoidf_(__ldg(sky_raw + 3 * n + 1)),
                ks2 = sigmoidf_(__ldg(sky_raw + 3 * n + 2));
    float carry = 0.f, a0 = 0, a1 = 0, a2 = 0, vs = 0, r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const long long o = base + s;
      float rho, c0, c1, c2, v;
      activate(raw[c], cl, rho, c0, c1, c2, v);
      const float y = s < S ? rho * raw[c].d : 0.f;
      const float incl = warp_scan_incl(y, lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      const float pv = expf(-(carry + prev));
      const float pe = 1.f - expf(-y);
      const float ps = pv * pe;
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (s < S) {
        if (PV) PV[o] = pv;
        if (PE) PE[o] = pe;
        if (PS) PS[o] = ps;
        a0 += ps * c0, a1 += ps * c1, a2 += ps * c2;
        vs += v * ps;
        if (kClassic) {
          r0 += ps * c0 * (v + (1.f - v) * ks0);
          r1 += ps * c1 * (v + (1.f - v) * ks1);
          r2 += ps * c2 * (v + (1.f - v) * ks2);
        }
      }
    }
    a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2), vs = warp_sum(vs);
    if (kClassic) r0 = warp_sum(r0), r1 = warp_sum(r1), r2 = warp_sum(r2);
    if (lane == 0) {
      albedo[3 * n] = a0, albedo[3 * n + 1] = a1, albedo[3 * n + 2] = a2;
      if (sky_act) sky_act[3 * n] = ks0, sky_act[3 * n + 1] = ks1, sky_act[3 * n + 2] = ks2;
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

// Backward of the above w.r.t. the RAW heads.  Sweep 1 = the forward (activations, PV / PE and the per-ray sums in
// registers); sweep 2 (reverse) = suffix scan of the transmittance gradient as in composite_bwd_kernel, then the chain
// through softplus / sigmoid / class mix; the softmax Jacobian and the sky sigmoid are applied once per ray.
// vis is detached in the non-classic colour formula (Eval_Tools_2.py:214) while PS is not: d_vis_raw only if classic.
template <bool kClassic, int kChunks>
__global__ void __launch_bounds__(256)
heads_composite_bwd_kernel(const float* __restrict__ pos, const float* __restrict__ vis_raw, const float* __restrict__ adj,
                           const float* __restrict__ sky_raw, const float* __restrict__ cls_logits,
                           const float* __restrict__ deltas, const HeadPitch ld, int N, int S, int C, const float* __restrict__ d_albedo,
                           const float* __restrict__ d_rendered, const float* __restrict__ d_sky_act,
                           float* __restrict__ d_pos, float* __restrict__ d_vis_raw, float* __restrict__ d_adj,
                           float* __restrict__ d_sky_raw, float* __restrict__ d_cls_logits) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    SampleRaw raw[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) load_sample(raw[c], c * 32 + lane < S, base + c * 32 + lane, pos, vis_raw, adj, deltas, C, ld);
    float cl[4];
    softmax4(cls_logits + (long long)n * C, C, cl);
    const float k0 = sigmoidf_(__ldg(sky_raw + 3 * n)), k1 = sigmoidf_(__ldg(sky_raw + 3 * n + 1)),
                k2 = sigmoidf_(__ldg(sky_raw + 3 * n + 2));
    float pv[kChunks], pe[kChunks], cc0[kChunks], cc1[kChunks], cc2[kChunks], vv[kChunks];
    float carry = 0.f, a0 = 0, a1 = 0, a2 = 0, vs = 0;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      float rho;
      activate(raw[c], cl, rho, cc0[c], cc1[c], cc2[c], vv[c]);
      const float y = s < S ? rho * raw[c].d : 0.f;
      const float incl = warp_scan_incl(y, lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      pv[c] = expf(-(carry + prev));
      pe[c] = 1.f - expf(-y);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (s < S && !kClassic) {
        const float ps = pv[c] * pe[c];
        a0 += ps * cc0[c], a1 += ps * cc1[c], a2 += ps * cc2[c];
        vs += vv[c] * ps;
      }
    }
    const float g0 = d_rendered ? __ldg(d_rendered + 3 * n) : 0.f, g1 = d_rendered ? __ldg(d_rendered + 3 * n + 1) : 0.f,
                g2 = d_rendered ? __ldg(d_rendered + 3 * n + 2) : 0.f;
    const float e0 = d_albedo ? __ldg(d_albedo + 3 * n) : 0.f, e1 = d_albedo ? __ldg(d_albedo + 3 * n + 1) : 0.f,
                e2 = d_albedo ? __ldg(d_albedo + 3 * n + 2) : 0.f;
    float dA0 = e0, dA1 = e1, dA2 = e2;
    float dk0 = 0, dk1 = 0, dk2 = 0, dvs = 0.f;
    if (!kClassic) {
      a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2), vs = warp_sum(vs);
      const float sv3 = sigmoidf_((vs - .2f) * 30.f);
      dA0 += g0 * (sv3 + (1.f - sv3) * k0), dA1 += g1 * (sv3 + (1.f - sv3) * k1), dA2 += g2 * (sv3 + (1.f - sv3) * k2);
      dk0 = g0 * a0 * (1.f - sv3), dk1 = g1 * a1 * (1.f - sv3), dk2 = g2 * a2 * (1.f - sv3);
      dvs = (g0 * a0 * (1.f - k0) + g1 * a1 * (1.f - k1) + g2 * a2 * (1.f - k2)) * sv3 * (1.f - sv3) * 30.f;
    }
    float suffix = 0.f;
    float dsk0 = 0, dsk1 = 0, dsk2 = 0;          // classic: gradient w.r.t. the activated per-ray sky
    float dcl[4] = {0.f, 0.f, 0.f, 0.f};         // gradient w.r.t. the class probabilities (this lane's samples)
#pragma unroll
    for (int c = kChunks - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const long long o = base + s;
      float q = 0.f, dpe_tot = 0.f;
      float dc0 = 0.f, dc1 = 0.f, dc2 = 0.f, dv = 0.f;       // gradients w.r.t. the ACTIVATED colour / vis
      if (ok) {
        const float c0 = cc0[c], c1 = cc1[c], c2 = cc2[c];
        const float ps = pv[c] * pe[c];
        float dps;
        if (kClassic) {
          const float v = vv[c];
          const float sh0 = v + (1.f - v) * k0, sh1 = v + (1.f - v) * k1, sh2 = v + (1.f - v) * k2;
          dps = g0 * c0 * sh0 + g1 * c1 * sh1 + g2 * c2 * sh2 + e0 * c0 + e1 * c1 + e2 * c2;
          dc0 = ps * (g0 * sh0 + e0), dc1 = ps * (g1 * sh1 + e1), dc2 = ps * (g2 * sh2 + e2);
          dv = ps * (g0 * c0 * (1.f - k0) + g1 * c1 * (1.f - k1) + g2 * c2 * (1.f - k2));
          dsk0 += g0 * ps * c0 * (1.f - v), dsk1 += g1 * ps * c1 * (1.f - v), dsk2 += g2 * ps * c2 * (1.f - v);
        } else {
          dps = dA0 * c0 + dA1 * c1 + dA2 * c2 + dvs * vv[c];
          dc0 = ps * dA0, dc1 = ps * dA1, dc2 = ps * dA2;
        }
        dpe_tot = dps * pv[c];
        q = dps * pe[c] * pv[c];
      }
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
        const float drho = dy * raw[c].d;
        // softplus'(x) = sigmoid(x) (torch: 1 above the threshold 20)
        const float dsig = raw[c].p.x > 20.f ? drho : drho * sigmoidf_(raw[c].p.x);
        const float z0 = dc0 * cc0[c] * (1.f - cc0[c]), z1 = dc1 * cc1[c] * (1.f - cc1[c]), z2 = dc2 * cc2[c] * (1.f - cc2[c]);
        reinterpret_cast<float4*>(d_pos)[o] = make_float4(dsig, z0, z1, z2);
        if (kClassic && d_vis_raw) d_vis_raw[o] = dv * vv[c] * (1.f - vv[c]);
#pragma unroll
        for (int k = 0; k < 4; ++k) dcl[k] += z0 * raw[c].a[3 * k] + z1 * raw[c].a[3 * k + 1] + z2 * raw[c].a[3 * k + 2];
        if (C == 4) {
          float4* d4 = reinterpret_cast<float4*>(d_adj) + 3 * o;
          d4[0] = make_float4(z0 * cl[0], z1 * cl[0], z2 * cl[0], z0 * cl[1]);
          d4[1] = make_float4(z1 * cl[1], z2 * cl[1], z0 * cl[2], z1 * cl[2]);
          d4[2] = make_float4(z2 * cl[2], z0 * cl[3], z1 * cl[3], z2 * cl[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < C) {
              d_adj[(o * C + k) * 3] = z0 * cl[k], d_adj[(o * C + k) * 3 + 1] = z1 * cl[k], d_adj[(o * C + k) * 3 + 2] = z2 * cl[k];
            }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dcl[k] = warp_sum(dcl[k]);
    if (kClassic) dsk0 = warp_sum(dsk0), dsk1 = warp_sum(dsk1), dsk2 = warp_sum(dsk2);
    if (lane == 0) {
      float t0 = kClassic ? dsk0 : dk0, t1 = kClassic ? dsk1 : dk1, t2 = kClassic ? dsk2 : dk2;
      if (d_sky_act) t0 += __ldg(d_sky_act + 3 * n), t1 += __ldg(d_sky_act + 3 * n + 1), t2 += __ldg(d_sky_act + 3 * n + 2);
      d_sky_raw[3 * n] = t0 * k0 * (1.f - k0), d_sky_raw[3 * n + 1] = t1 * k1 * (1.f - k1), d_sky_raw[3 * n + 2] = t2 * k2 * (1.f - k2);
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) dot += cl[k] * dcl[k];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < C) d_cls_logits[(long long)n * C + k] = cl[k] * (dcl[k] - dot);
    }
  }
}

// ---- solar pass ---------------------------------------------------------------------------------------------------------
// err[n] = sum_s (sigmoid(vis_raw) - PV)^2,  absorb[n] = 1 - sum_s PE*PV*sigmoid(vis_raw)  with rho = softplus(rho_raw);
// PV, PE are detached in both terms (Eval_Tools_2.py:353-368): the only gradient is w.r.t. vis_raw.
template <int kChunks>
__global__ void __launch_bounds__(256)
solar_loss_fwd_kernel(const float* __restrict__ rho_raw, const float* __restrict__ vis_raw, const float* __restrict__ deltas,
                      int ld_rho, int ld_vis, int N, int S, float* __restrict__ err, float* __restrict__ absorb) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    float y[kChunks], v[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      y[c] = ok ? softplusf_(__ldg(rho_raw + (base + s) * ld_rho)) * __ldg(deltas + base + s) : 0.f;
      v[c] = ok ? sigmoidf_(__ldg(vis_raw + (base + s) * ld_vis)) : 0.f;
    }
    float carry = 0.f, e = 0.f, ab = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const float incl = warp_scan_incl(y[c], lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      const float pv = expf(-(carry + prev));
      const float pe = 1.f - expf(-y[c]);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (c * 32 + lane < S) {
        e += (v[c] - pv) * (v[c] - pv);
        ab += pe * pv * v[c];
      }
    }
    e = warp_sum(e), ab = warp_sum(ab);
    if (lane == 0) err[n] = e, absorb[n] = 1.f - ab;
  }
}

template <int kChunks>
__global__ void __launch_bounds__(256)
solar_loss_bwd_kernel(const float* __restrict__ rho_raw, const float* __restrict__ vis_raw, const float* __restrict__ deltas,
                      int ld_rho, int ld_vis, int N, int S, const float* __restrict__ g_err, const float* __restrict__ g_abs,
                      float* __restrict__ d_vis_raw) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const long long base = (long long)n * S;
    const float ge = g_err ? __ldg(g_err + n) : 0.f, ga = g_abs ? __ldg(g_abs + n) : 0.f;
    float carry = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      const float yv = ok ? softplusf_(__ldg(rho_raw + (base + s) * ld_rho)) * __ldg(deltas + base + s) : 0.f;
      const float v = ok ? sigmoidf_(__ldg(vis_raw + (base + s) * ld_vis)) : 0.f;
      const float incl = warp_scan_incl(yv, lane);
      float prev = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) prev = 0.f;
      const float pv = expf(-(carry + prev));
      const float pe = 1.f - expf(-yv);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (ok) d_vis_raw[base + s] = (ge * 2.f * (v - pv) - ga * pe * pv) * v * (1.f - v);
    }
  }
}

}  // namespace snb

using namespace snb;

#define SNB_HEADS_DISPATCH(KERNEL, ...)                                          \
  do {                                                                           \
    if (classic) {                                                               \
      if (chunks <= 1) KERNEL<true, 1><<<grid, 256, 0, st>>>(__VA_ARGS__);       \
      else if (chunks == 2) KERNEL<true, 2><<<grid, 256, 0, st>>>(__VA_ARGS__);  \
      else if (chunks == 3) KERNEL<true, 3><<<grid, 256, 0, st>>>(__VA_ARGS__);  \
      else KERNEL<true, 4><<<grid, 256, 0, st>>>(__VA_ARGS__);                   \
    } else {                                                                     \
      if (chunks <= 1) KERNEL<false, 1><<<grid, 256, 0, st>>>(__VA_ARGS__);      \
      else if (chunks == 2) KERNEL<false, 2><<<grid, 256, 0, st>>>(__VA_ARGS__); \
      else if (chunks == 3) KERNEL<false, 3><<<grid, 256, 0, st>>>(__VA_ARGS__); \
      else KERNEL<false, 4><<<grid, 256, 0, st>>>(__VA_ARGS__);                  \
    }                                                                            \
  } while (0)

extern "C" int snb_heads_composite_fwd(const float* pos4, int ld_pos, const float* vis_raw, int ld_vis, const float* adj,
                                       int ld_adj, const float* sky_raw, const float* cls_logits, const float* deltas, int N,
                                       int S, int C, int classic,
                                       float* albedo, float* rendered, float* sky_act, float* vis_sum, float* PV, float* PE,
                                       float* PS, void* stream) {
  SNB_CHECK_ARG(pos4 && vis_raw && adj && sky_raw && cls_logits && deltas && albedo && rendered && N >= 0 && S > 0);
  SNB_CHECK_ARG((((uintptr_t)pos4) & 15) == 0 && (((uintptr_t)adj) & 15) == 0 && ld_pos >= 4 && (ld_pos & 3) == 0 && ld_vis >= 1);
  if (C < 1 || C > 4 || S > 32 * kHeadsMaxChunks) return SNB_ERR_UNSUPPORTED;
  SNB_CHECK_ARG(ld_adj >= 3 * C && (C != 4 || (ld_adj & 3) == 0));
  if (N == 0) return SNB_OK;
  const int chunks = (S + 31) / 32;
  const int grid = grid_for(N, 8, 16);
  cudaStream_t st = (cudaStream_t)stream;
  const HeadPitch ld = {ld_pos, ld_vis, ld_adj};
  SNB_HEADS_DISPATCH(heads_composite_fwd_kernel, pos4, vis_raw, adj, sky_raw, cls_logits, deltas, ld, N, S, C, albedo, rendered,
                     sky_act, vis_sum, PV, PE, PS);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_heads_composite_bwd(const float* pos4, int ld_pos, const float* vis_raw, int ld_vis, const float* adj,
                                       int ld_adj, const float* sky_raw, const float* cls_logits, const float* deltas, int N,
                                       int S, int C, int classic,
                                       const float* d_albedo, const float* d_rendered, const float* d_sky_act, float* d_pos4,
                                       float* d_vis_raw, float* d_adj, float* d_sky_raw, float* d_cls_logits, void* stream) {
  SNB_CHECK_ARG(pos4 && vis_raw && adj && sky_raw && cls_logits && deltas && d_pos4 && d_adj && d_sky_raw && d_cls_logits);
  SNB_CHECK_ARG(N >= 0 && S > 0 && (((uintptr_t)pos4) & 15) == 0 && (((uintptr_t)adj) & 15) == 0 &&
                (((uintptr_t)d_pos4) & 15) == 0 && (((uintptr_t)d_adj) & 15) == 0 && ld_pos >= 4 && (ld_pos & 3) == 0 && ld_vis >= 1);
  if (C < 1 || C > 4 || S > 32 * kHeadsMaxChunks) return SNB_ERR_UNSUPPORTED;
  SNB_CHECK_ARG(ld_adj >= 3 * C && (C != 4 || (ld_adj & 3) == 0));
  if (N == 0) return SNB_OK;
  const int chunks = (S + 31) / 32;
  const int grid = grid_for(N, 8, 16);
  cudaStream_t st = (cudaStream_t)stream;
  const HeadPitch ld = {ld_pos, ld_vis, ld_adj};
  SNB_HEADS_DISPATCH(heads_composite_bwd_kernel, pos4, vis_raw, adj, sky_raw, cls_logits, deltas, ld, N, S, C, d_albedo, d_rendered,
                     d_sky_act, d_pos4, d_vis_raw, d_adj, d_sky_raw, d_cls_logits);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_solar_loss_fwd(const float* rho_raw, int ld_rho, const float* vis_raw, int ld_vis, const float* deltas, int N,
                                  int S, float* err, float* absorb, void* stream) {
  SNB_CHECK_ARG(rho_raw && vis_raw && deltas && err && absorb && N >= 0 && S > 0 && ld_rho >= 1 && ld_vis >= 1);
  if (S > 32 * kHeadsMaxChunks) return SNB_ERR_UNSUPPORTED;
  if (N == 0) return SNB_OK;
  const int chunks = (S + 31) / 32;
  const int grid = grid_for(N, 8, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (chunks <= 1) solar_loss_fwd_kernel<1><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, err, absorb);
  else if (chunks == 2) solar_loss_fwd_kernel<2><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, err, absorb);
  else if (chunks == 3) solar_loss_fwd_kernel<3><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, err, absorb);
  else solar_loss_fwd_kernel<4><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, err, absorb);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_solar_loss_bwd(const float* rho_raw, int ld_rho, const float* vis_raw, int ld_vis, const float* deltas, int N,
                                  int S, const float* g_err, const float* g_abs, float* d_vis_raw, void* stream) {
  SNB_CHECK_ARG(rho_raw && vis_raw && deltas && d_vis_raw && N >= 0 && S > 0 && ld_rho >= 1 && ld_vis >= 1);
  if (S > 32 * kHeadsMaxChunks) return SNB_ERR_UNSUPPORTED;
  if (N == 0) return SNB_OK;
  const int chunks = (S + 31) / 32;
  const int grid = grid_for(N, 8, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (chunks <= 1) solar_loss_bwd_kernel<1><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, g_err, g_abs, d_vis_raw);
  else if (chunks == 2) solar_loss_bwd_kernel<2><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, g_err, g_abs, d_vis_raw);
  else if (chunks == 3) solar_loss_bwd_kernel<3><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, g_err, g_abs, d_vis_raw);
  else solar_loss_bwd_kernel<4><<<grid, 256, 0, st>>>(rho_raw, vis_raw, deltas, ld_rho, ld_vis, N, S, g_err, g_abs, d_vis_raw);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
