// The O(N) loss terms of one training step in ONE forward and ONE backward kernel (Eval_Tools_2.py:353-443 for the
// default configuration: Barron adaptive colour loss, solar correction terms, no prior DSM).
//
// The reference evaluates them with ~100 element-wise / reduction torch ops on [N,3] tensors (N = 4096 rays) and autograd
// replays as many: ~270 launches of 2-3 us each on the critical path between the forward and the backward sweep of the
// network.  Here:
//
//   loss_tail_fwd   one CTA.  Phase 1: log Z(alpha_c) and d log Z / d alpha by the same 768-node Gauss-Legendre rule under
//                   x = tan(theta) as adaptive_loss.py (float64).  Phase 2: one sweep over the rays - Barron's rho(x, alpha, s)
//                   and its derivatives w.r.t. alpha and s, the squared error, the sky and albedo regularisers, the two
//                   solar sums.  Phase 3: the scalar terms, their weights and the weighted total.
//   loss_tail_bwd   element-wise: gradients w.r.t. the rendered colour, the albedo (argmin rows only), the activated sky,
//                   the per-ray solar sums, alpha and scale, for given upstream multipliers of the five differentiable terms.
//
// The formulas are those autograd derives from adaptive_loss.lossfun (clamp / where semantics included: alpha == 2 and
// alpha == 0 select their closed forms and cut the gradient w.r.t. alpha; |alpha - 2| < eps freezes b).  alpha and scale
// come in as tensors computed by torch from the latent parameters, so that the branch taken is the one torch takes.
#include "common.cuh"
#include "api.h"

namespace snb {

constexpr int kLossThreads = 512;        // 128 registers per thread: the float64 pow / exp / log of the quadrature need them
constexpr int kQuadNodes = 768;
constexpr float kEpsF = 1.1920928955078125e-07f;      // numpy.finfo(float32).eps, the clamp floor of adaptive_loss.py

// rho(x; alpha, scale = 1) and d rho / d alpha in the precision of T, with lossfun's clamp / where semantics
template <typename T>
__device__ __forceinline__ void barron_alpha(T sq, T alpha, T& rho, T& drho_dalpha, T& drho_dsq) {
  const T eps = (T)kEpsF;
  if (alpha == (T)2) {
    rho = (T)0.5 * sq, drho_dalpha = (T)0, drho_dsq = (T)0.5;
    return;
  }
  if (alpha == (T)0) {
    rho = log1p((T)0.5 * sq), drho_dalpha = (T)0, drho_dsq = (T)0.5 / ((T)1 + (T)0.5 * sq);
    return;
  }
  const T d2 = fabs(alpha - (T)2), aa = fabs(alpha);
  const T b = d2 > eps ? d2 : eps;
  const T sgn = alpha >= (T)0 ? (T)1 : (T)-1;
  const T a = sgn * (aa > eps ? aa : eps);
  const T db = d2 >= eps ? (alpha > (T)2 ? (T)1 : (alpha < (T)2 ? (T)-1 : (T)0)) : (T)0;      // clamp passes the gradient where |.| >= eps
  const T da = aa >= eps ? (T)1 : (T)0;                                                         // sgn * sgn(alpha) = 1
  const T u = sq / b + (T)1;
  const T e = (T)0.5 * alpha;
  const T pw = pow(u, e);
  const T pwm1 = pow(u, e - (T)1);
  const T boa = b / a;
  rho = boa * (pw - (T)1);
  drho_dsq = boa * (e * pwm1) / b;
  drho_dalpha = boa * pw * log(u) * (T)0.5                    // through the exponent
                + boa * (e * pwm1) * (-sq / (b * b)) * db     // through b inside u
                + (pw - (T)1) * (db / a - b / (a * a) * da);  // through the factor b / a
}

struct LossTailParams {
  int N;
  const float* rendered;   // [N,3]
  const float* gt;         // [N,3]
  const float* albedo;     // [N,3]
  const float* sky;        // [N,3] activated sky colour of the ray
  const float* err;        // [N]  sum_s (vis - PV)^2
  const float* absorb;     // [N]  1 - sum_s PE*PV*vis
  const float* alpha;      // [3]
  const float* scale;      // [3]
  const double* theta;     // [768] quadrature nodes * pi/2
  const double* qw;        // [768] quadrature weights * pi/2
  float sc_lambda;
  int solar_type2;
  float* vals;             // [16] out, see kV*
  float* aux;              // [32] out, for the backward kernel, see kA*
};
// vals: 0 Color_ada, 1 Color_alpha, 2 Color_width, 3 Color (mse), 4 Solar_Correction, 5 Solar_Correction_2, 6 Sky_Color_Var,
//       7 Albedo_Color, 8 scale^2, 9 solar weight = sc_lambda / scale^2, 10 weighted total
// aux:  0-2 sum_n d rho/d alpha, 3-5 sum_n d rho/d scale, 6-8 d logZ/d alpha, 9-11 albedo min value, 12-14 albedo argmin (as float bits)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum of `v` over the block -> every thread (kLossThreads threads, scratch [32])
__device__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = lane < kLossThreads / 32 ? scratch[lane] : 0.0;
  r = warp_sum(r);
  return r;
}

__global__ void __launch_bounds__(kLossThreads, 1) loss_tail_fwd_kernel(const LossTailParams p) {
  __shared__ double scratch[32];
  __shared__ float s_alpha[3], s_scale[3];
  __shared__ double s_logz[3], s_dlogz[3];
  __shared__ float s_minv[32 * 3];
  __shared__ int s_mini[32 * 3];
  const int tid = threadIdx.x;
  if (tid < 3) s_alpha[tid] = __ldg(p.alpha + tid), s_scale[tid] = __ldg(p.scale + tid);
  __syncthreads();

  // ---- phase 1: log partition function and its derivative, float64 (adaptive_loss.AdaptiveLossFunction.log_partition) ----
  for (int c = 0; c < 3; ++c) {
    const double alpha = (double)s_alpha[c];
    double z = 0.0, dz = 0.0;
    for (int i = tid; i < kQuadNodes; i += kLossThreads) {
      const double th = p.theta[i], x = tan(th), cs = cos(th);
      double rho, dra, drs;
      barron_alpha<double>(x * x, alpha, rho, dra, drs);
      const double term = exp(-rho) / (cs * cs) * p.qw[i];
      z += term, dz -= term * dra;
    }
    z = block_sum(z, scratch);
    dz = block_sum(dz, scratch);
    if (tid == 0) s_logz[c] = log(z), s_dlogz[c] = dz / z;
  }
  __syncthreads();

  // ---- phase 2: one sweep over the rays ----
  double s_rho[3] = {0, 0, 0}, s_da[3] = {0, 0, 0}, s_ds[3] = {0, 0, 0};
  double s_sq = 0, s_sky = 0, s_err = 0, s_abs = 0;
  float minv[3] = {3.0e38f, 3.0e38f, 3.0e38f};
  int mini[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
  for (int n = tid; n < p.N; n += kLossThreads) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __ldg(p.rendered + 3 * n + c) - __ldg(p.gt + 3 * n + c);
      const float s = s_scale[c];
      const float zz = x / s;
      float rho, dra, drsq;
      barron_alpha<float>(zz * zz, s_alpha[c], rho, dra, drsq);
      s_rho[c] += (double)rho;
      s_da[c] += (double)dra;
      s_ds[c] += (double)(drsq * 2.f * zz * (-x / (s * s)));
      s_sq += (double)(x * x);
      const float sk = (__ldg(p.sky + 3 * n + c) - .5f) / .5f;
      const float rl = sk > 0.f ? sk : 0.f;
      s_sky += (double)(rl * rl);
      const float al = __ldg(p.albedo + 3 * n + c);
      if (al < minv[c]) minv[c] = al, mini[c] = n;
    }
    s_err += (double)__ldg(p.err + n);
    s_abs += (double)__ldg(p.absorb + n);
  }
  double tot_rho[3], tot_da[3], tot_ds[3];
  for (int c = 0; c < 3; ++c) {
    tot_rho[c] = block_sum(s_rho[c], scratch);
    tot_da[c] = block_sum(s_da[c], scratch);
    tot_ds[c] = block_sum(s_ds[c], scratch);
  }
  const double tot_sq = block_sum(s_sq, scratch), tot_sky = block_sum(s_sky, scratch);
  const double tot_err = block_sum(s_err, scratch), tot_abs = block_sum(s_abs, scratch);
  // per-channel minimum of the albedo with the smallest row index among equal values
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = minv[c];
      int ix = mini[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, ix, o);
        if (v2 < v || (v2 == v && i2 < ix)) v = v2, ix = i2;
      }
      if (lane == 0) s_minv[warp * 3 + c] = v, s_mini[warp * 3 + c] = ix;
    }
    __syncthreads();
  }

  // ---- phase 3: scalars ----
  if (tid == 0) {
    const float invN = 1.f / (float)p.N, inv3N = 1.f / (float)(3 * p.N);
    float mv[3];
    int mi[3];
    for (int c = 0; c < 3; ++c) {
      mv[c] = s_minv[c], mi[c] = s_mini[c];
      for (int w = 1; w < kLossThreads / 32; ++w) {
        const float v2 = s_minv[w * 3 + c];
        const int i2 = s_mini[w * 3 + c];
        if (v2 < mv[c] || (v2 == mv[c] && i2 < mi[c])) mv[c] = v2, mi[c] = i2;
      }
    }
    double nll = 0.0;
    for (int c = 0; c < 3; ++c) nll += tot_rho[c] + (double)p.N * ((double)logf(s_scale[c]) + (double)(float)s_logz[c]);
    const float color_ada = (float)(nll * (double)inv3N);
    const float color_alpha = (s_alpha[0] + s_alpha[1] + s_alpha[2]) / 3.f;
    const float color_width = (s_scale[0] + s_scale[1] + s_scale[2]) / 3.f;
    const float mse = (float)(tot_sq * (double)inv3N);
    const float err_m = (float)(tot_err * (double)invN), abs_m = (float)(tot_abs * (double)invN);
    const float sky_l = (float)(tot_sky * (double)inv3N);
    float alb_l = 0.f;
    for (int c = 0; c < 3; ++c)
      if (mv[c] < .2f) {
        const float d = 1.f - mv[c] / .2f;
        alb_l += d * d;
      }
    alb_l *= invN;
    const float scale_sq = color_width * color_width;
    const float w_s = p.sc_lambda / scale_sq;
    p.vals[0] = color_ada, p.vals[1] = color_alpha, p.vals[2] = color_width, p.vals[3] = mse;
    p.vals[4] = err_m, p.vals[5] = abs_m, p.vals[6] = sky_l, p.vals[7] = alb_l, p.vals[8] = scale_sq, p.vals[9] = w_s;
    // Eval_Tools_2.py:399-400,427-428: only the two solar-correction weights are divided by the squared colour-loss scale
    // (and the sky / albedo regularisers exist only without Solar_Type_2, :370)
    p.vals[10] = err_m * w_s + abs_m * w_s + (p.solar_type2 ? 0.f : sky_l * p.sc_lambda + alb_l * p.sc_lambda) + color_ada + color_alpha +
                 color_width + mse;
    for (int c = 0; c < 3; ++c) {
      p.aux[c] = (float)tot_da[c], p.aux[3 + c] = (float)tot_ds[c], p.aux[6 + c] = (float)s_dlogz[c];
      p.aux[9 + c] = mv[c], p.aux[12 + c] = __int_as_float(mi[c]);
    }
  }
}

struct LossTailBwdParams {
  int N;
  const float* rendered;
  const float* gt;
  const float* sky;
  const float* alpha;
  const float* scale;
  const float* aux;
  // upstream multipliers of the five differentiable terms (device scalars, may be null = 0) and of the weighted total
  const float* g_color;     // d / d Color_ada
  const float* g_err;       // d / d Solar_Correction
  const float* g_abs;       // d / d Solar_Correction_2 (solar type 2 only)
  const float* g_sky;       // d / d Sky_Color_Var
  const float* g_alb;       // d / d Albedo_Color
  const float* g_total;     // d / d total
  const float* vals;        // for the solar weight
  float sc_lambda;
  int solar_type2;
  float* d_rendered;        // [N,3]
  float* d_albedo;          // [N,3]
  float* d_sky;             // [N,3]
  float* d_err;             // [N]
  float* d_absorb;          // [N]
  float* d_alpha;           // [3]
  float* d_scale;           // [3]
};

__global__ void __launch_bounds__(256) loss_tail_bwd_kernel(const LossTailBwdParams p) {
  const float gt_ = p.g_total ? __ldg(p.g_total) : 0.f;
  const float w_s = __ldg(p.vals + 9);
  const float m_color = (p.g_color ? __ldg(p.g_color) : 0.f) + gt_;
  const float m_err = (p.g_err ? __ldg(p.g_err) : 0.f) + gt_ * w_s;
  const float m_abs = p.solar_type2 ? ((p.g_abs ? __ldg(p.g_abs) : 0.f) + gt_ * w_s) : 0.f;
  const float m_sky = (p.g_sky ? __ldg(p.g_sky) : 0.f) + (p.solar_type2 ? 0.f : gt_ * p.sc_lambda);
  const float m_alb = (p.g_alb ? __ldg(p.g_alb) : 0.f) + (p.solar_type2 ? 0.f : gt_ * p.sc_lambda);
  const float invN = 1.f / (float)p.N, inv3N = 1.f / (float)(3 * p.N);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < p.N) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __ldg(p.rendered + 3 * n + c) - __ldg(p.gt + 3 * n + c);
      const float s = __ldg(p.scale + c);
      const float zz = x / s;
      float rho, dra, drsq;
      barron_alpha<float>(zz * zz, __ldg(p.alpha + c), rho, dra, drsq);
      p.d_rendered[3 * n + c] = m_color * inv3N * (drsq * 2.f * zz / s);
      const float sk = (__ldg(p.sky + 3 * n + c) - .5f) / .5f;
      p.d_sky[3 * n + c] = sk > 0.f ? m_sky * inv3N * (2.f * sk / .5f) : 0.f;
      const float mv = __ldg(p.aux + 9 + c);
      const int mi = __float_as_int(__ldg(p.aux + 12 + c));
      p.d_albedo[3 * n + c] = (n == mi && mv < .2f) ? m_alb * invN * (2.f * (1.f - mv / .2f) * (-1.f / .2f)) : 0.f;
    }
    p.d_err[n] = m_err * invN;
    p.d_absorb[n] = m_abs * invN;
  }
  if (n < 3) {
    // mean over [N,3] of rho + log(scale) + logZ(alpha): the two broadcast terms contribute 1/3 per channel
    p.d_alpha[n] = m_color * (inv3N * __ldg(p.aux + n) + __ldg(p.aux + 6 + n) / 3.f);
    p.d_scale[n] = m_color * (inv3N * __ldg(p.aux + 3 + n) + 1.f / (3.f * __ldg(p.scale + n)));
  }
}

}  // namespace snb

using namespace snb;

extern "C" int snb_loss_tail_fwd(const float* rendered, const float* gt, const float* albedo, const float* sky, const float* err,
                                 const float* absorb, const float* alpha, const float* scale, const double* theta, const double* qw,
                                 int N, float sc_lambda, int solar_type2, float* vals, float* aux, void* stream) {
  SNB_CHECK_ARG(rendered && gt && albedo && sky && err && absorb && alpha && scale && theta && qw && vals && aux && N > 0);
  LossTailParams p;
  p.N = N, p.rendered = rendered, p.gt = gt, p.albedo = albedo, p.sky = sky, p.err = err, p.absorb = absorb, p.alpha = alpha,
  p.scale = scale, p.theta = theta, p.qw = qw, p.sc_lambda = sc_lambda, p.solar_type2 = solar_type2, p.vals = vals, p.aux = aux;
  loss_tail_fwd_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_loss_tail_bwd(const float* rendered, const float* gt, const float* sky, const float* alpha, const float* scale,
                                 const float* aux, const float* vals, const float* g_color, const float* g_err, const float* g_abs,
                                 const float* g_sky, const float* g_alb, const float* g_total, int N, float sc_lambda, int solar_type2,
                                 float* d_rendered, float* d_albedo, float* d_sky, float* d_err, float* d_absorb, float* d_alpha,
                                 float* d_scale, void* stream) {
  SNB_CHECK_ARG(rendered && gt && sky && alpha && scale && aux && vals && d_rendered && d_albedo && d_sky && d_err && d_absorb &&
                d_alpha && d_scale && N > 0);
  LossTailBwdParams p;
  p.N = N, p.rendered = rendered, p.gt = gt, p.sky = sky, p.alpha = alpha, p.scale = scale, p.aux = aux, p.vals = vals;
  p.g_color = g_color, p.g_err = g_err, p.g_abs = g_abs, p.g_sky = g_sky, p.g_alb = g_alb, p.g_total = g_total;
  p.solar_type2 = solar_type2, p.sc_lambda = sc_lambda;
  p.d_rendered = d_rendered, p.d_albedo = d_albedo, p.d_sky = d_sky, p.d_err = d_err, p.d_absorb = d_absorb, p.d_alpha = d_alpha,
  p.d_scale = d_scale;
  const int grid = (N + 255) / 256 < 1 ? 1 : (N + 255) / 256;
  loss_tail_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
