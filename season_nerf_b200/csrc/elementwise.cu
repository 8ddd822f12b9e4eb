// Bandwidth-bound per-element kernels around the dense layers: positional encoding, SIREN activation with a
// folded BatchNorm affine (forward / two-pass backward), column statistics, dtype staging.
//   reference: misc.py:105-139 (PE_Encode), misc.py:169-170,188-189 (SineLayer + BatchNorm1d).
#include "common.cuh"
#include "api.h"

namespace snb {

// ---------------------------------------------------------------------------------------------------
// PE_Encode, extended: out = [x | per dim: cos(k_j x) j<n , sin(k_j x) j<n], k_j = 2^j * fl32(pi/2).
// The argument is rounded to float32 exactly like the reference's tensor product (misc.py:127) and the
// accurate sinf/cosf are used: arguments reach ~800 rad, so no fast-math.
// One thread per (row, dim); a warp covers consecutive rows so stores of one column group are strided
// by the row pitch only (L2 write-combined).
template <typename TO>
__global__ void __launch_bounds__(256) pe_encode_kernel(const float* __restrict__ x, int ldx, long long M, int D, int n,
                                                        TO* __restrict__ out, int ldo, int col0, int pad_to) {
  const float kPiHalf = 1.57079637050628662109375f;  // float32(pi/2)
  const long long total = M * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / D;
    const int d = (int)(i - m * D);
    const float v = __ldg(x + m * ldx + d);
    TO* row = out + m * ldo + col0;
    row[d] = from_f32<TO>(v);
    TO* blk = row + D + d * 2 * n;
    float k = kPiHalf;
    for (int j = 0; j < n; ++j) {
      const float arg = __fmul_rn(k, v);
      float sv, cv;
      sincosf(arg, &sv, &cv);
      blk[j] = from_f32<TO>(cv);
      blk[n + j] = from_f32<TO>(sv);
      k = __fmul_rn(k, 2.0f);
    }
    if (d == 0)
      for (int c = D * (2 * n + 1); c < pad_to; ++c) row[c] = from_f32<TO>(0.f);
  }
}


// Warp-per-row variant for D = 3 with an output row of at most 64 columns (position: 63 -> 64, sun: 27 -> 32): lane l < 3n
// evaluates ONE sincosf (dimension l / n, frequency l % n), the row is assembled with two shuffles per output column pair
// and written as one contiguous 4-byte-per-lane store (bf16) - instead of 2-byte stores scattered over the row pitch.
template <typename TO>
__global__ void __launch_bounds__(256) pe_encode_row_kernel(const float* __restrict__ x, int ldx, long long M, int n,
                                                            TO* __restrict__ out, int ldo, int col0, int pad_to) {
  const float kPiHalf = 1.57079637050628662109375f;  // float32(pi/2)
  const int lane = threadIdx.x & 31;
  const int width = 3 * (2 * n + 1);
  // lane -> (dim, freq) it evaluates
  const int d_l = lane / n, j_l = lane - d_l * n;
  float k = kPiHalf;
  for (int j = 0; j < j_l; ++j) k = __fmul_rn(k, 2.0f);
  // output columns 2*lane, 2*lane+1 -> source lane and sin/cos selector (column c >= 3: d = (c-3)/(2n), r = (c-3)%(2n))
  int src[2], is_sin[2], kind[2];   // kind 0: raw x, 1: trig, 2: zero pad
  for (int h = 0; h < 2; ++h) {
    const int c = 2 * lane + h;
    if (c < 3) kind[h] = 0, src[h] = c, is_sin[h] = 0;
    else if (c < width) {
      const int d = (c - 3) / (2 * n), r = (c - 3) - d * 2 * n;
      kind[h] = 1, is_sin[h] = r >= n, src[h] = d * n + (r >= n ? r - n : r);
    } else kind[h] = 2, src[h] = 0, is_sin[h] = 0;
  }
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = warp0; m < M; m += nwarps) {
    const float xv = lane < 3 ? __ldg(x + m * ldx + lane) : 0.f;
    const float xd = __shfl_sync(0xffffffffu, xv, d_l < 3 ? d_l : 0);
    float sv = 0.f, cv = 0.f;
    if (lane < 3 * n) sincosf(__fmul_rn(k, xd), &sv, &cv);
    float o[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float s_ = __shfl_sync(0xffffffffu, sv, src[h]), c_ = __shfl_sync(0xffffffffu, cv, src[h]);
      const float r_ = __shfl_sync(0xffffffffu, xv, src[h] < 3 ? src[h] : 0);
      o[h] = kind[h] == 0 ? r_ : (kind[h] == 1 ? (is_sin[h] ? s_ : c_) : 0.f);
    }
    if (2 * lane < pad_to) {
      TO* dst = out + m * ldo + col0 + 2 * lane;
      if (2 * lane + 1 < pad_to) {
        if (sizeof(TO) == 2) {
          __nv_bfloat162 v2 = __floats2bfloat162_rn(o[0], o[1]);
          *reinterpret_cast<__nv_bfloat162*>(dst) = v2;
        } else {
          dst[0] = from_f32<TO>(o[0]), dst[1] = from_f32<TO>(o[1]);
        }
      } else {
        dst[0] = from_f32<TO>(o[0]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Train-mode BatchNorm1d bookkeeping of one layer in ONE launch (misc.py:169-170; nn.BatchNorm1d(momentum, eps)):
// batch mean / biased variance from the column sums, running-statistics update (unbiased variance), and the folded
// affine  y = a*z + c  with  a = gamma*invstd, c = beta - mean*a  that the activation kernels consume.
template <typename TS>
__global__ void bn_finalize_kernel(const TS* __restrict__ sum, const TS* __restrict__ sumsq, long long rows, int N,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                                   float* running_var, long long* num_batches, float momentum, float eps, float* __restrict__ a,
                                   float* __restrict__ c, float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_batches) *num_batches += 1;
  if (i >= N) return;
  const double mean = (double)sum[i] / (double)rows;
  double var = (double)sumsq[i] / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  running_mean[i] = running_mean[i] * (1.f - momentum) + (float)mean * momentum;
  const double unb = var * ((double)rows / (double)(rows > 1 ? rows - 1 : 1));
  running_var[i] = running_var[i] * (1.f - momentum) + (float)unb * momentum;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float meanf = (float)mean;
  const float av = gamma[i] * invstd;
  a[i] = av;
  c[i] = beta[i] - meanf * av;
  mean_out[i] = meanf;
  invstd_out[i] = invstd;
}

// ---------------------------------------------------------------------------------------------------
// Column sums over M rows (float64 results through double atomics; inner accumulation in float32 over
// at most 128 rows per flush).  Threads own column pairs -> 4/8-byte loads, fully coalesced per row.
template <typename T, typename F>
__device__ __forceinline__ void column_reduce(const T* __restrict__ base0, int ld0, const T* __restrict__ base1, int ld1,
                                              long long M, int N, double* __restrict__ o0, double* __restrict__ o1, F f) {
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  const int ty = threadIdx.y, ny = blockDim.y;
  for (int c = threadIdx.x * 2; c < N; c += blockDim.x * 2) {
    const bool two = c + 1 < N;
    double s0a = 0, s1a = 0, s0b = 0, s1b = 0;
    for (long long rb = r0 + ty; rb < r1; rb += (long long)ny * 128) {
      float f0a = 0, f1a = 0, f0b = 0, f1b = 0;
      long long rend = rb + (long long)ny * 128;
      if (rend > r1) rend = r1;
      for (long long r = rb; r < rend; r += ny) {
        float p0, p1, q0, q1;
        f(base0, ld0, base1, ld1, r, c, two, p0, p1, q0, q1);
        f0a += p0, f1a += p1, f0b += q0, f1b += q1;
      }
      s0a += f0a, s1a += f1a, s0b += f0b, s1b += f1b;
    }
    atomicAdd(o0 + c, s0a);
    atomicAdd(o1 + c, s1a);
    if (two) {
      atomicAdd(o0 + c + 1, s0b);
      atomicAdd(o1 + c + 1, s1b);
    }
  }
}

template <typename T>
__device__ __forceinline__ void load2(const T* p, bool two, float& a, float& b);
template <>
__device__ __forceinline__ void load2<float>(const float* p, bool two, float& a, float& b) {
  a = p[0];
  b = two ? p[1] : 0.f;
}
template <>
__device__ __forceinline__ void load2<__nv_bfloat16>(const __nv_bfloat16* p, bool two, float& a, float& b) {
  a = __bfloat162float(p[0]);
  b = two ? __bfloat162float(p[1]) : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256) col_stats_kernel(const T* __restrict__ Z, int ldz, long long M, int N,
                                                        double* __restrict__ sum, double* __restrict__ sumsq) {
  column_reduce<T>(Z, ldz, Z, ldz, M, N, sum, sumsq,
                   [] __device__(const T* z, int ld, const T*, int, long long r, int c, bool two, float& p0, float& p1,
                                 float& q0, float& q1) {
                     float a, b;
                     load2<T>(z + r * ld + c, two, a, b);
                     p0 = a, p1 = a * a, q0 = b, q1 = b * b;
                   });
}

struct AffineCols {
  const float* a;
  const float* c;
  const float* mean;
  const float* invstd;
};

template <typename T>
__global__ void __launch_bounds__(256) sine_bwd_reduce_kernel(const T* __restrict__ dY, int ldd, const T* __restrict__ Z,
                                                              int ldz, AffineCols p, long long M, int N,
                                                              double* __restrict__ sg, double* __restrict__ sgx) {
  column_reduce<T>(dY, ldd, Z, ldz, M, N, sg, sgx,
                   [p] __device__(const T* dy, int ld0, const T* z, int ld1, long long r, int c, bool two, float& p0,
                                  float& p1, float& q0, float& q1) {
                     float d0, d1, z0, z1;
                     load2<T>(dy + r * ld0 + c, two, d0, d1);
                     load2<T>(z + r * ld1 + c, two, z0, z1);
                     const float g0 = d0 * cosf(p.a[c] * z0 + p.c[c]);
                     p0 = g0, p1 = g0 * (z0 - p.mean[c]) * p.invstd[c];
                     q0 = q1 = 0.f;
                     if (two) {
                       const float g1 = d1 * cosf(p.a[c + 1] * z1 + p.c[c + 1]);
                       q0 = g1, q1 = g1 * (z1 - p.mean[c + 1]) * p.invstd[c + 1];
                     }
                   });
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) sine_fwd_kernel(const T* __restrict__ Z, int ldz, const float* __restrict__ a,
                                                       const float* __restrict__ c, T* __restrict__ Y, int ldy,
                                                       long long M, int N) {
  const int n2 = (N + 1) / 2;
  const long long total = M * n2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / n2;
    const int col = (int)(i - m * n2) * 2;
    const bool two = col + 1 < N;
    float z0, z1;
    load2<T>(Z + m * ldz + col, two, z0, z1);
    Y[m * ldy + col] = from_f32<T>(sinf(__ldg(a + col) * z0 + __ldg(c + col)));
    if (two) Y[m * ldy + col + 1] = from_f32<T>(sinf(__ldg(a + col + 1) * z1 + __ldg(c + col + 1)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
sine_bwd_apply_kernel(const T* __restrict__ dY, int ldd, const T* __restrict__ Z, int ldz, AffineCols p,
                      const float* __restrict__ k1, const float* __restrict__ k2, T* __restrict__ dZ, int ldo,
                      long long M, int N) {
  const long long total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N;
    const int col = (int)(i - m * N);
    const float z = to_f32<T>(Z[m * ldz + col]);
    const float a = __ldg(p.a + col);
    float g = to_f32<T>(dY[m * ldd + col]) * cosf(a * z + __ldg(p.c + col));
    if (k1) g -= __ldg(k1 + col) + (z - __ldg(p.mean + col)) * __ldg(p.invstd + col) * __ldg(k2 + col);
    dZ[m * ldo + col] = from_f32<T>(a * g);
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) convert_kernel(const TS* __restrict__ src, int lds, TD* __restrict__ dst, int ldd,
                                                      long long M, int N) {
  const long long total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N;
    const int col = (int)(i - m * N);
    dst[m * ldd + col] = from_f32<TD>(to_f32<TS>(src[m * lds + col]));
  }
}


// ===================================================================================================
// Vectorised variants (N % 8 == 0, 16-byte aligned rows): thread <-> 8 consecutive columns, so the per-column
// BatchNorm constants live in registers and every global access is a 16/32-byte vector; rows are strided over
// blockDim.y x gridDim.x.  bf16 build uses sin.approx / cos.approx (|argument| stays small after BatchNorm; the
// result is rounded to bf16 anyway), the fp32 validation build keeps the accurate sinf / cosf.
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
  static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 t2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&t2);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  static __device__ __forceinline__ float sin_(float x) { return __sinf(x); }
  static __device__ __forceinline__ float cos_(float x) { return __cosf(x); }
};

__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) { Vec8<float>::load(p, v); }

template <typename T>
__global__ void __launch_bounds__(256) sine_fwd_vec_kernel(const T* __restrict__ Z, int ldz, const float* __restrict__ a,
                                                           const float* __restrict__ c, T* __restrict__ Y, int ldy,
                                                           long long M, int N) {
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (col >= N) return;
  float av[8], cv[8];
  load8f(a + col, av);
  load8f(c + col, cv);
  for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < M; r += (long long)gridDim.x * blockDim.y) {
    float z[8];
    Vec8<T>::load(Z + r * ldz + col, z);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = Vec8<T>::sin_(fmaf(av[i], z[i], cv[i]));
    Vec8<T>::store(Y + r * ldy + col, z);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
sine_bwd_apply_vec_kernel(const T* __restrict__ dY, int ldd, const T* __restrict__ Z, int ldz, AffineCols p,
                          const float* __restrict__ k1, const float* __restrict__ k2, T* __restrict__ dZ, int ldo,
                          long long M, int N) {
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (col >= N) return;
  float av[8], cv[8], mv[8], iv[8], k1v[8], k2v[8];
  load8f(p.a + col, av);
  load8f(p.c + col, cv);
  const bool bn = k1 != nullptr;
  if (bn) {
    load8f(p.mean + col, mv);
    load8f(p.invstd + col, iv);
    load8f(k1 + col, k1v);
    load8f(k2 + col, k2v);
  }
  for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < M; r += (long long)gridDim.x * blockDim.y) {
    float z[8], d[8];
    Vec8<T>::load(Z + r * ldz + col, z);
    Vec8<T>::load(dY + r * ldd + col, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float g = d[i] * Vec8<T>::cos_(fmaf(av[i], z[i], cv[i]));
      if (bn) g -= k1v[i] + (z[i] - mv[i]) * iv[i] * k2v[i];
      d[i] = av[i] * g;
    }
    Vec8<T>::store(dZ + r * ldo + col, d);
  }
}

// dZ = a*(g - k1 - xhat*k2): second half of the train-mode BatchNorm backward when g = dY*cos(.) was already produced
// (and reduced) by the fused input-gradient GEMM epilogue.  In place allowed (dZ == G).
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(const T* G, int ldg, const T* __restrict__ Z, int ldz, const float* __restrict__ a,
                        const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ k1,
                        const float* __restrict__ k2, float scale, T* dZ, int ldo, long long M, int N) {
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (col >= N) return;
  float av[8], mv[8], iv[8], k1v[8], k2v[8];
  load8f(a + col, av);
  load8f(mean + col, mv);
  load8f(invstd + col, iv);
  load8f(k1 + col, k1v);
  load8f(k2 + col, k2v);
#pragma unroll
  for (int i = 0; i < 8; ++i) k1v[i] *= scale, k2v[i] *= iv[i] * scale;
  for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < M; r += (long long)gridDim.x * blockDim.y) {
    float z[8], g[8];
    Vec8<T>::load(Z + r * ldz + col, z);
    Vec8<T>::load(G + r * ldg + col, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = av[i] * (g[i] - k1v[i] - (z[i] - mv[i]) * k2v[i]);
    Vec8<T>::store(dZ + r * ldo + col, g);
  }
}

// column reductions: two quantities per column accumulated in fp32 registers over <= 64 rows, then in fp64, then one
// double atomicAdd per column per block-row.
template <typename T, bool kBwd>
__global__ void __launch_bounds__(256)
col_reduce_vec_kernel(const T* __restrict__ X0, int ld0, const T* __restrict__ X1, int ld1, AffineCols p, long long M, int N,
                      double* __restrict__ o0, double* __restrict__ o1) {
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (col >= N) return;
  float av[8], cv[8], mv[8], iv[8];
  if (kBwd) {
    load8f(p.a + col, av);
    load8f(p.c + col, cv);
    load8f(p.mean + col, mv);
    load8f(p.invstd + col, iv);
  }
  double s0[8], s1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.y;
  long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y;
  while (r < M) {
    float f0[8], f1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f0[i] = f1[i] = 0.f;
    for (int it = 0; it < 64 && r < M; ++it, r += stride) {
      float x[8];
      Vec8<T>::load(X0 + r * ld0 + col, x);
      if (kBwd) {
        float z[8];
        Vec8<T>::load(X1 + r * ld1 + col, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g = x[i] * Vec8<T>::cos_(fmaf(av[i], z[i], cv[i]));
          f0[i] += g;
          f1[i] = fmaf(g, (z[i] - mv[i]) * iv[i], f1[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f0[i] += x[i];
          f1[i] = fmaf(x[i], x[i], f1[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s0[i] += f0[i], s1[i] += f1[i];
  }
  // combine the blockDim.y row-lanes of this block through shared memory before touching global atomics
  __shared__ double red[2][8][256 + 1];
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i][t] = s0[i], red[1][i][t] = s1[i];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double a0 = 0, a1 = 0;
      for (int y = 0; y < blockDim.y; ++y) a0 += red[0][i][y * blockDim.x + threadIdx.x], a1 += red[1][i][y * blockDim.x + threadIdx.x];
      atomicAdd(o0 + col + i, a0);
      atomicAdd(o1 + col + i, a1);
    }
  }
}

struct VecLaunch {
  dim3 grid, block;
  bool ok;
};
static inline VecLaunch vec_launch(long long M, int N, int ld0, int ld1, int ld2, const void* p0, const void* p1, const void* p2,
                                   int elem_bytes, int max_waves) {
  VecLaunch v;
  v.ok = (N % 8 == 0) && (ld0 % 8 == 0) && (ld1 % 8 == 0) && (ld2 % 8 == 0) && ((((uintptr_t)p0) | ((uintptr_t)p1) | ((uintptr_t)p2)) % 16 == 0);
  int tx = N / 8;
  if (tx < 1) tx = 1;
  int by = 1;
  if (tx > 256) { by = (tx + 255) / 256; tx = 256; }
  int ty = 256 / tx;
  if (ty < 1) ty = 1;
  v.block = dim3(tx, ty);
  long long gx = (M + ty - 1) / ty;
  long long cap = (long long)num_sms() * max_waves;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  v.grid = dim3((unsigned)gx, by);
  (void)elem_bytes;
  return v;
}

}  // namespace snb

using namespace snb;
typedef __nv_bfloat16 bf16;

extern "C" int snb_pe_encode(const float* x, int ldx, long long M, int D, int n_freq, void* out, int out_dtype, int ldo,
                             int col0, int pad_to, void* stream) {
  SNB_CHECK_ARG(x && out && M >= 0 && D > 0 && n_freq >= 0 && ldx >= D && ldo >= col0 + D * (2 * n_freq + 1));
  SNB_CHECK_ARG(pad_to <= ldo - col0);
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 3 && n_freq >= 1 && 3 * n_freq <= 32 && pad_to <= 64 && pad_to >= D * (2 * n_freq + 1) && (ldo % 2) == 0 && (col0 % 2) == 0 &&
      ((uintptr_t)out % 4) == 0 && M >= 64) {
    const int grid_r = grid_for(M, 8, 16);
    if (out_dtype == SNB_F32) pe_encode_row_kernel<float><<<grid_r, 256, 0, st>>>(x, ldx, M, n_freq, (float*)out, ldo, col0, pad_to);
    else if (out_dtype == SNB_BF16) pe_encode_row_kernel<bf16><<<grid_r, 256, 0, st>>>(x, ldx, M, n_freq, (bf16*)out, ldo, col0, pad_to);
    else return SNB_ERR_ARG;
    count_launch();
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  const int grid = grid_for(M * D, 256, 16);
  if (out_dtype == SNB_F32) pe_encode_kernel<float><<<grid, 256, 0, st>>>(x, ldx, M, D, n_freq, (float*)out, ldo, col0, pad_to);
  else if (out_dtype == SNB_BF16) pe_encode_kernel<bf16><<<grid, 256, 0, st>>>(x, ldx, M, D, n_freq, (bf16*)out, ldo, col0, pad_to);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

static dim3 reduce_block(int N) {
  int tx = 32;
  while (tx * 2 < N && tx < 256) tx *= 2;
  return dim3(tx, 256 / tx);
}

extern "C" int snb_col_stats(const void* Z, int dtype, int ldz, long long M, int N, double* sum, double* sumsq,
                             void* stream) {
  SNB_CHECK_ARG(Z && sum && sumsq && M >= 0 && N > 0 && ldz >= N);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(double) * N, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(sumsq, 0, sizeof(double) * N, st);
  if (e != cudaSuccess) return (int)e;
  if (M == 0) return SNB_OK;
  {
    VecLaunch v = vec_launch(M, N, ldz, 8, 8, Z, nullptr, nullptr, 0, 4);
    if (v.ok && (N / 8) <= 256 && 256 % (N / 8) == 0) {
      AffineCols none = {nullptr, nullptr, nullptr, nullptr};
      if (dtype == SNB_F32) col_reduce_vec_kernel<float, false><<<v.grid, v.block, 0, st>>>((const float*)Z, ldz, nullptr, 0, none, M, N, sum, sumsq);
      else if (dtype == SNB_BF16) col_reduce_vec_kernel<bf16, false><<<v.grid, v.block, 0, st>>>((const bf16*)Z, ldz, nullptr, 0, none, M, N, sum, sumsq);
      else return SNB_ERR_ARG;
      count_launch();
      SNB_LAUNCH_CHECK();
      return SNB_OK;
    }
  }
  const int grid = grid_for(M, 512, 4);
  if (dtype == SNB_F32) col_stats_kernel<float><<<grid, reduce_block(N), 0, st>>>((const float*)Z, ldz, M, N, sum, sumsq);
  else if (dtype == SNB_BF16) col_stats_kernel<bf16><<<grid, reduce_block(N), 0, st>>>((const bf16*)Z, ldz, M, N, sum, sumsq);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_sine_fwd(const void* Z, int ldz, const float* a, const float* c, void* Y, int ldy, long long M, int N,
                            int dtype, void* stream) {
  SNB_CHECK_ARG(Z && a && c && Y && M >= 0 && N > 0 && ldz >= N && ldy >= N);
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {
    VecLaunch v = vec_launch(M, N, ldz, ldy, 8, Z, Y, nullptr, 0, 8);
    if (v.ok) {
      if (dtype == SNB_F32) sine_fwd_vec_kernel<float><<<v.grid, v.block, 0, st>>>((const float*)Z, ldz, a, c, (float*)Y, ldy, M, N);
      else if (dtype == SNB_BF16) sine_fwd_vec_kernel<bf16><<<v.grid, v.block, 0, st>>>((const bf16*)Z, ldz, a, c, (bf16*)Y, ldy, M, N);
      else return SNB_ERR_ARG;
      count_launch();
      SNB_LAUNCH_CHECK();
      return SNB_OK;
    }
  }
  const int grid = grid_for(M * ((N + 1) / 2), 256, 16);
  if (dtype == SNB_F32) sine_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)Z, ldz, a, c, (float*)Y, ldy, M, N);
  else if (dtype == SNB_BF16) sine_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)Z, ldz, a, c, (bf16*)Y, ldy, M, N);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_sine_bwd_reduce(const void* dY, int ldd, const void* Z, int ldz, const float* a, const float* c,
                                   const float* mean, const float* invstd, long long M, int N, int dtype, double* sg,
                                   double* sgx, void* stream) {
  SNB_CHECK_ARG(dY && Z && a && c && mean && invstd && sg && sgx && M >= 0 && N > 0);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sg, 0, sizeof(double) * N, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(sgx, 0, sizeof(double) * N, st);
  if (e != cudaSuccess) return (int)e;
  if (M == 0) return SNB_OK;
  AffineCols p = {a, c, mean, invstd};
  {
    VecLaunch v = vec_launch(M, N, ldd, ldz, 8, dY, Z, nullptr, 0, 4);
    if (v.ok && (N / 8) <= 256 && 256 % (N / 8) == 0) {
      if (dtype == SNB_F32) col_reduce_vec_kernel<float, true><<<v.grid, v.block, 0, st>>>((const float*)dY, ldd, (const float*)Z, ldz, p, M, N, sg, sgx);
      else if (dtype == SNB_BF16) col_reduce_vec_kernel<bf16, true><<<v.grid, v.block, 0, st>>>((const bf16*)dY, ldd, (const bf16*)Z, ldz, p, M, N, sg, sgx);
      else return SNB_ERR_ARG;
      count_launch();
      SNB_LAUNCH_CHECK();
      return SNB_OK;
    }
  }
  const int grid = grid_for(M, 512, 4);
  if (dtype == SNB_F32) sine_bwd_reduce_kernel<float><<<grid, reduce_block(N), 0, st>>>((const float*)dY, ldd, (const float*)Z, ldz, p, M, N, sg, sgx);
  else if (dtype == SNB_BF16) sine_bwd_reduce_kernel<bf16><<<grid, reduce_block(N), 0, st>>>((const bf16*)dY, ldd, (const bf16*)Z, ldz, p, M, N, sg, sgx);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_sine_bwd_apply(const void* dY, int ldd, const void* Z, int ldz, const float* a, const float* c,
                                  const float* mean, const float* invstd, const float* k1, const float* k2, void* dZ,
                                  int ldo, long long M, int N, int dtype, void* stream) {
  SNB_CHECK_ARG(dY && Z && a && c && dZ && M >= 0 && N > 0);
  SNB_CHECK_ARG((k1 == nullptr) == (k2 == nullptr));
  SNB_CHECK_ARG(!k1 || (mean && invstd));
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  AffineCols p = {a, c, mean, invstd};
  {
    VecLaunch v = vec_launch(M, N, ldd, ldz, ldo, dY, Z, dZ, 0, 8);
    if (v.ok) {
      if (dtype == SNB_F32) sine_bwd_apply_vec_kernel<float><<<v.grid, v.block, 0, st>>>((const float*)dY, ldd, (const float*)Z, ldz, p, k1, k2, (float*)dZ, ldo, M, N);
      else if (dtype == SNB_BF16) sine_bwd_apply_vec_kernel<bf16><<<v.grid, v.block, 0, st>>>((const bf16*)dY, ldd, (const bf16*)Z, ldz, p, k1, k2, (bf16*)dZ, ldo, M, N);
      else return SNB_ERR_ARG;
      count_launch();
      SNB_LAUNCH_CHECK();
      return SNB_OK;
    }
  }
  const int grid = grid_for(M * N, 256, 16);
  if (dtype == SNB_F32) sine_bwd_apply_kernel<float><<<grid, 256, 0, st>>>((const float*)dY, ldd, (const float*)Z, ldz, p, k1, k2, (float*)dZ, ldo, M, N);
  else if (dtype == SNB_BF16) sine_bwd_apply_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dY, ldd, (const bf16*)Z, ldz, p, k1, k2, (bf16*)dZ, ldo, M, N);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_bn_bwd_apply(const void* G, int ldg, const void* Z, int ldz, const float* a, const float* mean,
                                const float* invstd, const float* k1, const float* k2, float scale, void* dZ, int ldo,
                                long long M, int N, int dtype, void* stream) {
  SNB_CHECK_ARG(G && Z && a && mean && invstd && k1 && k2 && dZ && M >= 0 && N > 0);
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VecLaunch v = vec_launch(M, N, ldg, ldz, ldo, G, Z, dZ, 0, 8);
  if (!v.ok) return SNB_ERR_UNSUPPORTED;
  if (dtype == SNB_F32) bn_bwd_apply_vec_kernel<float><<<v.grid, v.block, 0, st>>>((const float*)G, ldg, (const float*)Z, ldz, a, mean, invstd, k1, k2, scale, (float*)dZ, ldo, M, N);
  else if (dtype == SNB_BF16) bn_bwd_apply_vec_kernel<bf16><<<v.grid, v.block, 0, st>>>((const bf16*)G, ldg, (const bf16*)Z, ldz, a, mean, invstd, k1, k2, scale, (bf16*)dZ, ldo, M, N);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_bn_finalize(const void* sum, const void* sumsq, int stats_dtype, long long rows, int N, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, long long* num_batches,
                               float momentum, float eps, float* a, float* c, float* mean, float* invstd, void* stream) {
  SNB_CHECK_ARG(sum && sumsq && gamma && beta && running_mean && running_var && a && c && mean && invstd && rows > 0 && N > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (N + 127) / 128;
  if (stats_dtype == SNB_F32)
    bn_finalize_kernel<float><<<grid, 128, 0, st>>>((const float*)sum, (const float*)sumsq, rows, N, gamma, beta, running_mean,
                                                    running_var, num_batches, momentum, eps, a, c, mean, invstd);
  else if (stats_dtype == SNB_F64)
    bn_finalize_kernel<double><<<grid, 128, 0, st>>>((const double*)sum, (const double*)sumsq, rows, N, gamma, beta, running_mean,
                                                     running_var, num_batches, momentum, eps, a, c, mean, invstd);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// fp32 -> bf16 staging of up to kStageMaxSeg weight matrices in ONE launch (the bf16 copies every layer's GEMMs read): the
// segment descriptors travel in the kernel parameters, blockIdx.y = segment
namespace snb {
constexpr int kStageMaxSeg = 48;
struct StageArgs {
  const float* src[kStageMaxSeg];
  __nv_bfloat16* dst[kStageMaxSeg];
  int rows[kStageMaxSeg], cols[kStageMaxSeg], lds[kStageMaxSeg], ldd[kStageMaxSeg];
};
__global__ void __launch_bounds__(256) stage_weights_kernel(const __grid_constant__ StageArgs a) {
  const int sg = blockIdx.y;
  const float* __restrict__ src = a.src[sg];
  __nv_bfloat16* __restrict__ dst = a.dst[sg];
  const int N = a.cols[sg], lds = a.lds[sg], ldd = a.ldd[sg];
  const int total = a.rows[sg] * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int m = i / N, col = i - m * N;
    dst[(long long)m * ldd + col] = __float2bfloat16_rn(src[(long long)m * lds + col]);
  }
}
}  // namespace snb

extern "C" int snb_stage_weights(const void* const* src, void* const* dst, const int* rows, const int* cols, const int* lds,
                                 const int* ldd, int n_seg, void* stream) {
  SNB_CHECK_ARG(src && dst && rows && cols && lds && ldd && n_seg >= 0 && n_seg <= snb::kStageMaxSeg);
  if (n_seg == 0) return SNB_OK;
  snb::StageArgs a;
  for (int i = 0; i < n_seg; ++i) {
    SNB_CHECK_ARG(src[i] && dst[i] && rows[i] > 0 && cols[i] > 0 && lds[i] >= cols[i] && ldd[i] >= cols[i]);
    a.src[i] = (const float*)src[i], a.dst[i] = (__nv_bfloat16*)dst[i];
    a.rows[i] = rows[i], a.cols[i] = cols[i], a.lds[i] = lds[i], a.ldd[i] = ldd[i];
  }
  snb::stage_weights_kernel<<<dim3(8, n_seg), 256, 0, (cudaStream_t)stream>>>(a);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_convert(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, long long M, int N,
                           void* stream) {
  SNB_CHECK_ARG(src && dst && M >= 0 && N > 0 && lds >= N && ldd >= N);
  if (M == 0) return SNB_OK;
  const int grid = grid_for(M * N, 256, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == SNB_F32 && dst_dtype == SNB_BF16) convert_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, M, N);
  else if (src_dtype == SNB_BF16 && dst_dtype == SNB_F32) convert_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, M, N);
  else if (src_dtype == SNB_F32 && dst_dtype == SNB_F32) convert_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, M, N);
  else if (src_dtype == SNB_BF16 && dst_dtype == SNB_BF16) convert_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, M, N);
  else return SNB_ERR_ARG;
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
