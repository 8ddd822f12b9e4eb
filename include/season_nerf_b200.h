/* season_nerf_b200 -- C ABI of the B200 (sm_100a) Season-NeRF render/train hot path.
 *
 * The upstream reference (EnterpriseCV-6/Season-NeRF) is pure Python/PyTorch and has no FFI; the
 * boundary it exposes for this path is its Python object API.  This header is
 * the C ABI that the drop-in Python classes in season_nerf_b200/ bind with ctypes: plain device
 * pointers, sizes and a cudaStream_t passed as void*.  Every entry point cites the reference
 * code (file:line, relative to the upstream repo root) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - matrices are row-major with an explicit leading dimension (elements);
 *   - dtype codes: SNB_F32 = 0, SNB_BF16 = 1;
 *   - return value 0 = success, >0 = cudaError_t, <0 = SNB_ERR_*; no entry point synchronises;
 *   - N = rays, S = samples per ray, M = N*S sample points.
 */
#ifndef SEASON_NERF_B200_H
#define SEASON_NERF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SNB_F32 0
#define SNB_BF16 1
#define SNB_F64 2

/* library / device info ------------------------------------------------------------------- */
int snb_version(void);
/* multiprocessor count of the current CUDA device (queried once per device): sizes every persistent grid */
int snb_num_sms(void);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
long long snb_launch_count(void);
const char* snb_error_string(int code);

/* ---- ray sampling: misc.py:234-247 sample_pt_coarse + misc.py:249-261 zero_invalid_pts ----
 * pts[n,s,:] = top[n]*(1-ts[s]) + bot[n]*ts[s]   (two rounded products, one rounded sum: bit-exact)
 * deltas[n,s] = ||top[n]-bot[n]||_2 / S           (0 where zero_oob and the point leaves [-1,1]^3)
 * ts is the length-S parameter vector the host builds exactly like the reference
 * (linspace (+ one shared jitter vector in train mode)). pts may be NULL. */
int snb_sample_rays(const float* top, const float* bot, const float* ts, int N, int S, int zero_oob,
                    float* pts, float* deltas, void* stream);

/* ---- shadow-march ray construction: mg_Img_Eval.py:57-64 / Eval_Tools_2.py:255-258 ----------
 * new_top = float32( double(p) + ((1-p_z)/sun_z) * double(sun) ) (float64 arithmetic as the reference's
 * numpy promotion does when f64 != 0; float32 arithmetic when f64 == 0, the engine variant). */
int snb_solar_tops(const float* pts, long long M, const double* sun3_host, int f64, float* tops, void* stream);

/* ---- compositing (engine convention): Eval_Tools_2.py:13-16 get_PV and :187-215 ------------
 * deltas: [N,S].  col: [N,S,3] activated colour.  vis: [N,S].  sky: [N,3] (sky_per_sample=0) or [N,S,3].
 * classic = args.Solar_Type_2.  Outputs PV,PE,PS [N,S] (each may be NULL), albedo [N,3], rendered [N,3],
 * vis_sum [N] = sum_s vis*PS (may be NULL). */
int snb_composite_fwd(const float* rho, const float* deltas, const float* col, const float* vis, const float* sky,
                      int sky_per_sample, int N, int S, int classic, float* PV, float* PE, float* PS,
                      float* albedo, float* rendered, float* vis_sum, void* stream);
/* backward of the above.  Incoming gradients (any may be NULL): d_rendered[N,3], d_albedo[N,3], dPE,dPV,dPS [N,S].
 * Outgoing: d_rho[N,S], d_col[N,S,3], d_sky (same layout as sky), d_vis[N,S] (written only when classic;
 * vis is detached otherwise, Eval_Tools_2.py:214). */
int snb_composite_bwd(const float* rho, const float* deltas, const float* col, const float* vis, const float* sky,
                      int sky_per_sample, int N, int S, int classic, const float* d_rendered, const float* d_albedo,
                      const float* dPE, const float* dPV, const float* dPS, float* d_rho, float* d_col, float* d_sky,
                      float* d_vis, void* stream);

/* transmittance of the shadow march: out[m] = exp(-sum_{k<S-1} rho[m,k]*deltas[m,k])  (mg_Img_Eval.py:68-70) */
int snb_march_transmittance(const float* rho, const float* deltas, long long M, int S, float* out, void* stream);

/* ---- CLI compositing: mg_Img_Eval.py:123-190 get_imgs_from_Img_Dict, float64 like the reference ------
 * rho,deltas,vis [N,S]; base [N,S,3] raw logits; adj [N,S,C,3] (element type in_dtype: SNB_F32 network outputs
 * promoted exactly, or SNB_F64); cls [C] float64.
 * base_img, season_img [N,3]; extreme [C,N,3]; raw_shadow [N] (all float64). exact_vis/raw_shadow_exact optional. */
int snb_cli_composite(const void* rho, const void* deltas, const void* base, const void* vis, const void* adj,
                      const double* cls, const void* exact_vis, int in_dtype, int N, int S, int C, double* base_img,
                      double* season_img, double* extreme, double* raw_shadow, double* raw_shadow_exact, void* stream);
/* classic shadows of get_imgs_from_Img_Dict (mg_Img_Eval.py:166-181): out[n,:] = sum_s PS * sigmoid(base + cls . adj) *
 * (vis + (1 - vis) * sky) with the per-sample sky colour sky [N,S,3]; `vis` is Est_Solar_Vis or Exact_Solar.  float64 sums. */
int snb_cli_classic_shadow(const void* rho, const void* deltas, const void* base, const void* vis, const void* adj,
                           const void* sky, const double* cls, int in_dtype, int N, int S, int C, double* out, void* stream);
/* year sweep: mg_Img_Eval.py:192-228 get_imgs_from_Img_Dict_t_step, fused over T class vectors.
 * cls [T,C] -> out [T,N,3] (float64) = season colour * shade[N,3] (the per-ray Shadow_Adjust factor of :214-226;
 * null = 1).  float32 components are recombined in float32 and reduced in float64 (|error| < 3e-7); float64
 * components keep float64 arithmetic throughout.  T*C <= 5120.  ps_weight [N,S] (element type in_dtype, optional)
 * multiplies PS per sample: the classic-shadow alignment sums PS * vis (mg_Img_Eval.py:448-449). */
int snb_year_sweep(const void* rho, const void* deltas, const void* base, const void* adj, const double* cls,
                   const double* shade, const void* ps_weight, int in_dtype, int N, int S, int C, int T, double* out,
                   void* stream);

/* ---- camera rays ("next" row of the scope table): P_img_Pinhole.invert_P, pre_NeRF/P_Img.py:133-147 -------------
 * For n pixels - explicit (rows[i], cols[i]) int32 device arrays, or the raster grid (i / W * ds, i % W * ds) when both
 * are null - the closed-form inversion of the 3x4 affine-approximated RPC projection P (HOST pointer, 12 doubles,
 * row-major, already normalised/scaled) at heights z_top and z_bot, in float64 in numpy's evaluation order:
 *   tops[i] = (x(z_top), y(z_top), z_top), bots[i] = (x(z_bot), y(z_bot), z_bot)     float32 (the reference's .float())
 *   xy64[i] = (x_top, y_top, x_bot, y_bot) float64 (optional), good[i] = all four inside bounds = (x_min, x_max, y_min,
 *   y_max) inclusive (HOST pointer; optional) - the filters of mg_Pt_holder.py:180-187 and mg_Img_Eval.py:83-84. */
int snb_camera_rays(const double* P, const int* rows, const int* cols, long long n, int W, int ds, double z_top,
                    double z_bot, const double* bounds, float* tops, float* bots, double* xy64, unsigned char* good,
                    void* stream);

/* ---- solar-ray generator: create_solor_rays_uniform.__call__, T_NeRF_Full_2/Eval_Tools_2.py:72-108 ----------------
 * n random solar rays from already drawn random numbers (device arrays): az_el [n,2] float64 = (azimuth in
 * [-180,180), elevation in [1,90)) degrees, u_xy [n,2] float32 uniforms, u_time [n,2] float32 uniforms (null with
 * times == null).  world_center (3 doubles) and W2L_H (4x4 row-major doubles) are HOST pointers.  Per ray, in float64
 * in numpy's evaluation order: v = world_angle_2_local_vec(el, az) (all_NeRF/mg_unit_converter.py:5-9,29-34,59-68),
 * delta = 2 v / v_z;  starts = (2 u_x - 1, 2 u_y - 1, 1) float32;  ends = float32(starts - delta);  vec = float32(v);
 * times = (cos f0, sin f0, cos f1, sin f1) with f = (u_time * 2) * fl32(pi) in float32.  Replaces the reference's
 * per-ray Python loop (89 us / ray) and the host->device copy of the four arrays. */
int snb_solar_rays(const double* world_center, const double* W2L_H, const double* az_el, const float* u_xy,
                   const float* u_time, int n, float* starts, float* ends, float* vec, float* times, void* stream);

/* ---- prior-DSM density: T_NeRF.Supervised_Sample, T_NeRF_Full_2/T_NeRF_net_v2.py:175-181 -----------------------------
 * pts [M,3], delta [M] float32; hm [H,W] float64 (the module keeps the map as a float64 tensor); out [M] float32 =
 * -log(1 - P) / delta with P = 0.99 where hm[long(((x,y)+1)/2*(shape-1))] >= z, else 0.  k_hit = -log(1 - fl32(0.99)) is
 * passed by the caller (evaluated once with the host's float32 log, so that the result is bit-identical to the reference's
 * CPU path). */
int snb_supervised_sample(const float* pts, const float* delta, const double* hm, int H, int W, float k_hit, long long M,
                          float* out, void* stream);

/* ---- positional encoding: misc.py:105-139 PE_Encode (extended) ------------------------------
 * out[m, col0 + ...] = [x (D), per dim: cos(k_j x) j<n, sin(k_j x) j<n], k_j = 2^j * fl32(pi/2);
 * width D*(2n+1), zero padded up to pad_to columns.  x: [M,D] float32, ldx elements. out dtype f32/bf16. */
int snb_pe_encode(const float* x, int ldx, long long M, int D, int n_freq, void* out, int out_dtype, int ldo,
                  int col0, int pad_to, void* stream);
/* d_x += d_enc * d enc / d x is never needed: inputs carry no gradient in the reference. */

/* ---- dense layers -----------------------------------------------------------------------------
 * C[M,N] = alpha * (A[M,K] . B[N,K]^T + bias[N]) (+ C if accumulate).  "TN" GEMM: both operands K-contiguous.
 * a_t / b_t select the transposed ("MN-major") operand forms used by the weight gradient:
 *   a_t=1: A is stored [K,M] (lda = row stride of that storage);  b_t=1: B is stored [K,N].
 * dtype SNB_F32 : CUDA-core fp32 validation path.  SNB_BF16: tcgen05/TMEM tensor-core path, fp32 accumulate.
 * out_dtype may differ from dtype (bf16 operands -> f32 result for weight gradients).
 * Replaces nn.Linear inside misc.py:188-189 and its autograd backward. */
int snb_gemm(const void* A, int lda, int a_t, const void* B, int ldb, int b_t, void* C, int ldc, const float* bias,
             float alpha, int accumulate, long long M, int N, int K, int dtype, int out_dtype, void* stream);

/* bf16 GEMM (both operands K-contiguous, bf16 result) with the train-mode BatchNorm statistics of the result fused
 * into the epilogue: stats[0..N) += column sums, stats[N..2N) += column sums of squares of the STORED bf16 values
 * (caller zeroes stats).  Returns SNB_ERR_UNSUPPORTED (-2) for shapes the CTA-pair kernel does not take
 * (M < 256, N < 128, N > 1024, unaligned C): the caller then uses snb_gemm + snb_col_stats.
 * Replaces nn.Linear + the batch-statistics pass of nn.BatchNorm1d in misc.py:169-170,188-189. */
int snb_gemm_stats(const void* A, int lda, const void* B, int ldb, void* C, int ldc, const float* bias, float alpha,
                   long long M, int N, int K, float* stats, void* stream);
/* snb_gemm_stats whose A operand is the SAVED PRE-ACTIVATION of the previous SIREN layer: C = alpha * (sin(xa[k] * Zprev[m,k]
 * + xc[k]) . B^T + bias) - the previous layer's activation sin(BatchNorm(.)) (misc.py:188-189, xa / xc = its folded affine)
 * is applied by transform warps to the TMA-landed operand tile in shared memory, before the tensor core reads it: the
 * stand-alone activation pass never runs and the activated [M,K] matrix is not read from HBM.  Y (optional, [M,K] bf16, row
 * pitch ldy) receives the activated operand for a later weight gradient; NULL = it is never written either.  Same outputs as
 * snb_gemm_stats (C bf16 + column sum / sum of squares); SNB_ERR_UNSUPPORTED for shapes the CTA-pair kernels do not take. */
int snb_gemm_stats_xf(const void* Zprev, int lda, const float* xa, const float* xc, const void* B, int ldb, void* C, int ldc,
                      const float* bias, float alpha, long long M, int N, int K, float* stats, void* Y, int ldy, void* stream);

/* Forward of a SIREN layer WITHOUT BatchNorm in one kernel (bf16 tcgen05 GEMM + activation epilogue):
 *   Z[M,N] = alpha*(A[M,K].B[N,K]^T + bias)  (bf16, kept for the backward),   Y[M,N] = sin(Z)  (bf16).
 * Replaces nn.Linear + torch.sin in misc.py:188-189 for layers whose norm is Identity.  Returns SNB_ERR_UNSUPPORTED (-2)
 * for shapes the CTA-pair kernel does not take (M < 256, N < 128, N > 512, N % 64, unaligned): use snb_gemm + snb_sine_fwd. */
int snb_gemm_sine_fwd(const void* A, int lda, const void* B, int ldb, void* Z, int ldz, void* Y, int ldy,
                      const float* bias, float alpha, long long M, int N, int K, void* stream);

/* Input-gradient GEMM of layer L+1 fused with the activation backward of layer L (autograd of misc.py:188-189):
 *   dY[M,N] = alpha * dZn[M,K] . W[K,N]      (W = weight of layer L+1 stored [out=K, in=N], row pitch ldw)
 *   G[M,N]  = dY * cos(a[n]*Z + c[n])        (Z = saved pre-activation of layer L, bf16; G bf16)
 *   stats[0..N) += sum_m G,  stats[N..2N) += sum_m G*(Z-mean[n])*invstd[n]     (caller zeroes stats; float32)
 * Without BatchNorm (a=1, c=0): G is dZ of layer L and stats[0..N) its bias gradient / alpha_L.  With train-mode
 * BatchNorm the two sums are the BatchNorm bias / weight gradients and snb_bn_bwd_apply finishes dZ.
 * SNB_ERR_UNSUPPORTED (-2) like snb_gemm_sine_fwd: the caller then runs snb_gemm + snb_sine_bwd_reduce/apply. */
int snb_gemm_sine_bwd(const void* dZn, int lda, const void* W, int ldw, void* G, int ldg, const void* Z, int ldz,
                      const float* a, const float* c, const float* mean, const float* invstd, float alpha,
                      long long M, int N, int K, float* stats, void* stream);

/* dZ = a[n]*(G - scale*k1[n] - (Z-mean[n])*invstd[n]*scale*k2[n])   (train-mode BatchNorm backward after snb_gemm_sine_bwd;
 * k1, k2 = the column sums it left, scale = 1/M; scale 0 = eval-mode BatchNorm; dZ may alias G) */
int snb_bn_bwd_apply(const void* G, int ldg, const void* Z, int ldz, const float* a, const float* mean,
                     const float* invstd, const float* k1, const float* k2, float scale, void* dZ, int ldo, long long M,
                     int N, int dtype, void* stream);

/* Train-mode nn.BatchNorm1d(momentum, eps) bookkeeping of one SineLayer in one launch (misc.py:169-170,189): from the
 * column sums (float32 or float64, stats_dtype) of the M = rows pre-activations: batch mean / biased variance, the
 * running_mean / running_var update (unbiased variance, in place), num_batches_tracked += 1 (may be null), and the
 * folded affine a = gamma*invstd, c = beta - mean*a plus mean, invstd for the backward (all float32 [N]). */
int snb_bn_finalize(const void* sum, const void* sumsq, int stats_dtype, long long rows, int N, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, long long* num_batches, float momentum,
                    float eps, float* a, float* c, float* mean, float* invstd, void* stream);

/* column statistics for train-mode BatchNorm1d (misc.py:169-170): sum[n], sumsq[n] over M rows (float64 out). */
int snb_col_stats(const void* Z, int dtype, int ldz, long long M, int N, double* sum, double* sumsq, void* stream);

/* Y = sin(a[n]*Z + c[n])  (a,c fold BatchNorm / identity).  Y dtype = dtype.  misc.py:189. */
int snb_sine_fwd(const void* Z, int ldz, const float* a, const float* c, void* Y, int ldy, long long M, int N,
                 int dtype, void* stream);
/* backward of sin(BN(Z)):  g = dY*cos(a Z + c).
 * pass 1 (reduce): sg[n] = sum_m g, sgx[n] = sum_m g*xhat, xhat = (Z-mean)*invstd   (float64 out)
 * pass 2 (apply):  dZ = a*(g - k1[n] - xhat*k2[n]); k1,k2 = sg/M, sgx/M in train-mode BN, 0 otherwise. */
int snb_sine_bwd_reduce(const void* dY, int ldd, const void* Z, int ldz, const float* a, const float* c,
                        const float* mean, const float* invstd, long long M, int N, int dtype, double* sg,
                        double* sgx, void* stream);
int snb_sine_bwd_apply(const void* dY, int ldd, const void* Z, int ldz, const float* a, const float* c,
                       const float* mean, const float* invstd, const float* k1, const float* k2, void* dZ, int ldo,
                       long long M, int N, int dtype, void* stream);

/* fp32 -> bf16 staging of n_seg <= 48 weight matrices in one launch (host arrays of device pointers / shapes / row pitches):
 * the bf16 operand copies of every nn.Linear a training pass uses (misc.py:148-194, G_NeRF.py, T_NeRF_net_v2.py). */
int snb_stage_weights(const void* const* src, void* const* dst, const int* rows, const int* cols, const int* lds,
                      const int* ldd, int n_seg, void* stream);

/* dtype conversion with leading dimensions (f32 <-> bf16), used to stage operands */
int snb_convert(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, long long M, int N,
                void* stream);

/* ---- output image of the render CLI from the RAW heads of a block of rays: T_NeRF_net_v2.py:139-151 activations +
 * mg_Img_Eval.py:123-160 float64 sums, main_run_Season_NeRF.py:90-92.  pos4 [M,4], vis_raw [M], adj [M,C,3] contiguous
 * float32, deltas [M] (zero outside the cube), cls [C] float64 (the image's class vector), exact_vis [M] or NULL
 * -> season_img [N,3], raw_shadow [N] (, raw_shadow_exact [N]) float64. */
int snb_render_composite_raw(const float* pos4, const float* vis_raw, const float* adj, const float* deltas,
                             const double* cls, const float* exact_vis, int N, int S, int C, double* season_img,
                             double* raw_shadow, double* raw_shadow_exact, void* stream);

/* year sweep (snb_year_sweep) fed by the RAW pos4 [M,4] = (sigma, base colour logits) of the network instead of activated
 * rho / base arrays: mg_Img_Eval.py:192-228 without materialising the component dict. */
int snb_year_sweep_raw(const float* pos4, const float* deltas, const float* adj, const double* cls, const double* shade, int N,
                       int S, int C, int T, double* out, void* stream);

/* ---- head activations fused with the compositing scan: T_NeRF_net_v2.py:87-98 + Eval_Tools_2.py:187-215 --------------
 * Reads the RAW network heads once - pos4 [M,4] = (sigma, colour logits), vis_raw [M], adj [M,C,3] (row pitches ld_* in
 * floats: 4 / 1 / 3C when contiguous), sky_raw [N,3], cls_logits [N,C], deltas [N,S] - applies softplus / softmax + class
 * mix / sigmoid in registers and composites: albedo [N,3], rendered [N,3] (classic = Solar_Type_2, else the Solar_Vis3
 * gate), sky_act [N,3] = sigmoid(sky_raw), vis_sum [N] = sum_s vis*PS, optional PV / PE / PS [N,S].  C <= 4, S <= 128. */
int snb_heads_composite_fwd(const float* pos4, int ld_pos, const float* vis_raw, int ld_vis, const float* adj, int ld_adj,
                            const float* sky_raw, const float* cls_logits, const float* deltas, int N, int S, int C,
                            int classic, float* albedo, float* rendered, float* sky_act, float* vis_sum, float* PV, float* PE,
                            float* PS, void* stream);
/* gradients of the raw heads given d_albedo, d_rendered, d_sky_act (any may be NULL): d_pos4 [M,4], d_adj [M,C,3],
 * d_sky_raw [N,3], d_cls_logits [N,C] (softmax Jacobian applied), d_vis_raw [M] only if classic (vis is detached in the
 * non-classic colour formula, Eval_Tools_2.py:214, while its PS weight is not). */
int snb_heads_composite_bwd(const float* pos4, int ld_pos, const float* vis_raw, int ld_vis, const float* adj, int ld_adj,
                            const float* sky_raw, const float* cls_logits, const float* deltas, int N, int S, int C,
                            int classic, const float* d_albedo, const float* d_rendered, const float* d_sky_act,
                            float* d_pos4, float* d_vis_raw, float* d_adj, float* d_sky_raw, float* d_cls_logits,
                            void* stream);
/* solar pass of get_loss (Eval_Tools_2.py:297-337, :353-368) from the raw heads of forward_Solar: per ray
 * err = sum_s (sigmoid(vis_raw) - PV)^2 and absorb = 1 - sum_s PE*PV*sigmoid(vis_raw), rho = softplus(rho_raw); PV and PE
 * are detached there, so the backward only has d_vis_raw [M] = (g_err*2(vis-PV) - g_abs*PE*PV) * vis(1-vis). */
int snb_solar_loss_fwd(const float* rho_raw, int ld_rho, const float* vis_raw, int ld_vis, const float* deltas, int N, int S,
                       float* err, float* absorb, void* stream);
int snb_solar_loss_bwd(const float* rho_raw, int ld_rho, const float* vis_raw, int ld_vis, const float* deltas, int N, int S,
                       const float* g_err, const float* g_abs, float* d_vis_raw, void* stream);

/* The O(N) loss terms of one training step in one kernel each way (Eval_Tools_2.py:353-443, default configuration: Barron
 * adaptive colour loss - lossfun / log-partition of season_nerf_b200/adaptive_loss.py -, solar correction terms, no prior):
 * vals[0..10] = Color_ada, Color_alpha, Color_width, Color (mse), Solar_Correction, Solar_Correction_2, Sky_Color_Var,
 * Albedo_Color, scale^2, sc_lambda / scale^2, weighted total; aux[0..14] = sums the backward kernel needs.  rendered, gt,
 * albedo, sky [N,3]; err, absorb [N]; alpha, scale [3] (computed by the caller from the latent parameters); theta, qw [768]
 * float64 Gauss-Legendre nodes / weights times pi/2. */
int snb_loss_tail_fwd(const float* rendered, const float* gt, const float* albedo, const float* sky, const float* err,
                      const float* absorb, const float* alpha, const float* scale, const double* theta, const double* qw,
                      int N, float sc_lambda, int solar_type2, float* vals, float* aux, void* stream);
/* Gradients of sum_k g_k * term_k + g_total * total (device scalars, null = 0) w.r.t. rendered, albedo, sky [N,3], err,
 * absorb [N], alpha, scale [3]. */
int snb_loss_tail_bwd(const float* rendered, const float* gt, const float* sky, const float* alpha, const float* scale,
                      const float* aux, const float* vals, const float* g_color, const float* g_err, const float* g_abs,
                      const float* g_sky, const float* g_alb, const float* g_total, int N, float sc_lambda, int solar_type2,
                      float* d_rendered, float* d_albedo, float* d_sky, float* d_err, float* d_absorb, float* d_alpha,
                      float* d_scale, void* stream);

/* ---- fused eval-mode network (render): T_NeRF_net_v2.py:75-105,131-151,169-170 + G_NeRF.py:74-133 --------------
 * One persistent tcgen05 kernel runs encoding -> trunk -> sigma/colour heads -> solar branch -> adjust branch with the
 * activations kept in shared memory / TMEM: two CTAs of a cluster render a tile of 256 sample points with
 * tcgen05.mma.cta_group::2 and stage half of every weight tile each.  `program` is the device image built by
 * season_nerf_b200/packing2.py: schedule tables, biases and the BatchNorm-folded bf16 weights as ONE row-major
 * [w_rows, 64] matrix (TMA applies the 128-byte swizzle); the *_off arguments are byte offsets of its sections,
 * n_mma / n_epi the table lengths.
 *   pts [M,3] f32; sun [ceil(M/S),3] f32 (one solar direction per S consecutive points; S=1: per point; may be
 *   NULL for a sigma-only program).
 * outputs (float32, any may be NULL): rho_raw [M], pos4 [M,4] = (sigma, colour[3]), vis_raw [M], adj [M,12]
 * (class-major [C,3]), all BEFORE softplus / sigmoid. */
int snb_fused_eval2(const void* program, unsigned n_mma, unsigned n_epi, unsigned mma_off, unsigned epi_off,
                    unsigned bias_off, unsigned w_off, unsigned w_rows, const float* pts, long long M, int S,
                    const float* sun, float* rho_raw, float* pos4, float* vis_raw, float* adj, void* stream);

#ifdef __cplusplus
}
#endif
#endif
