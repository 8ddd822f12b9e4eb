// fp32 CUDA-core GEMM: the "fp32 validation build" of the dense layers (1e-4 parity target).  Classic 128x64x16
// shared-memory tiling, 8x4 register micro-tile per thread, generic operand strides so that the same kernel
// serves forward (A.B^T), input gradients and weight gradients (transposed operands).  Not the production path:
// production uses the tcgen05 kernel in gemm_tc.cu.
//   reference: nn.Linear inside misc.py:188-189 and its autograd backward.
#include <cstdlib>
#include "common.cuh"
#include "api.h"

namespace snb {

constexpr int BM = 128, BN = 64, BK = 16;

// C[m,n] = alpha*(sum_k A(m,k)*B(n,k) + bias[n]) (+C).  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk].
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbn,
                long long sbk, float* __restrict__ C, int ldc, const float* __restrict__ bias, float alpha,
                int accumulate, long long M, int N, int K) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread tile 8 (m) x 4 (n)
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: 128 x 16 = 2048 elements, 8 per thread.  Orient the thread->element map along the contiguous axis.
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (sak == 1) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 127; kk = idx >> 7; }
      const long long m = m0 + mm;
      const int k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? __ldg(A + m * sam + k * sak) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int nn, kk;
      if (sbk == 1) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      const int n = n0 + nn;
      const int k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? __ldg(B + n * sbn + k * sbk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * (acc[i][j] + (bias ? __ldg(bias + n) : 0.f));
      float* c = C + m * ldc + n;
      if (accumulate) v += *c;
      *c = v;
    }
  }
}

}  // namespace snb

// implemented in gemm_tc.cu
int snb_gemm_bf16_tc(const void* A, int lda, int a_t, const void* B, int ldb, int b_t, void* C, int ldc,
                     const float* bias, float alpha, int accumulate, long long M, int N, int K, int out_dtype,
                     cudaStream_t st);
// implemented in gemm_tc2.cu (CTA-pair kernel; SNB_ERR_UNSUPPORTED = use the single-CTA kernel)
int snb_gemm_bf16_tc2(const void* A, int lda, int a_t, const void* B, int ldb, int b_t, void* C, int ldc,
                      const float* bias, float alpha, int accumulate, long long M, int N, int K, int out_dtype,
                      float* stats, cudaStream_t st, int epi = 0, const void* X = nullptr, int ldx = 0,
                      const float* ea = nullptr, const float* ec = nullptr, const float* emean = nullptr,
                      const float* einvstd = nullptr, const float* xa = nullptr, const float* xc = nullptr);

extern "C" int snb_gemm_sine_fwd(const void* A, int lda, const void* B, int ldb, void* Z, int ldz, void* Y, int ldy,
                                 const float* bias, float alpha, long long M, int N, int K, void* stream) {
  SNB_CHECK_ARG(A && B && Z && Y && M >= 0 && N > 0 && K > 0 && ldz >= N && ldy >= N);
  if (M == 0) return SNB_OK;
  return snb_gemm_bf16_tc2(A, lda, 0, B, ldb, 0, Z, ldz, bias, alpha, 0, M, N, K, SNB_BF16, nullptr, (cudaStream_t)stream, 1,
                           Y, ldy);
}

extern "C" int snb_gemm_sine_bwd(const void* dZn, int lda, const void* W, int ldw, void* G, int ldg, const void* Z, int ldz,
                                 const float* a, const float* c, const float* mean, const float* invstd, float alpha,
                                 long long M, int N, int K, float* stats, void* stream) {
  SNB_CHECK_ARG(dZn && W && G && Z && a && c && mean && invstd && stats && M >= 0 && N > 0 && K > 0 && ldg >= N && ldz >= N);
  if (M == 0) return SNB_OK;
  return snb_gemm_bf16_tc2(dZn, lda, 0, W, ldw, 1, G, ldg, nullptr, alpha, 0, M, N, K, SNB_BF16, stats, (cudaStream_t)stream, 2,
                           Z, ldz, a, c, mean, invstd);
}

extern "C" int snb_gemm_stats(const void* A, int lda, const void* B, int ldb, void* C, int ldc, const float* bias,
                              float alpha, long long M, int N, int K, float* stats, void* stream) {
  SNB_CHECK_ARG(A && B && C && stats && M >= 0 && N > 0 && K > 0 && ldc >= N);
  if (M == 0) return SNB_OK;
  return snb_gemm_bf16_tc2(A, lda, 0, B, ldb, 0, C, ldc, bias, alpha, 0, M, N, K, SNB_BF16, stats, (cudaStream_t)stream);
}

// implemented in gemm_tc3.cu (resident-A kernel: every operand element is activated once)
int snb_gemm_bf16_tc3(const void* Zprev, int lda, const float* xa, const float* xc, const void* B, int ldb, void* C, int ldc,
                      const float* bias, float alpha, long long M, int N, int K, float* stats, void* Y, int ldy,
                      cudaStream_t st);

extern "C" int snb_gemm_stats_xf(const void* Zprev, int lda, const float* xa, const float* xc, const void* B, int ldb, void* C,
                                 int ldc, const float* bias, float alpha, long long M, int N, int K, float* stats, void* Y,
                                 int ldy, void* stream) {
  SNB_CHECK_ARG(Zprev && xa && xc && B && C && stats && M >= 0 && N > 0 && K > 0 && ldc >= N && (!Y || ldy >= K));
  if (M == 0) return SNB_OK;
  static int force_tc2 = -1;
  if (force_tc2 < 0) {
    const char* e = getenv("SNB_XF_STREAMED");       // A/B switch: 1 = the streamed-A kernel of gemm_tc2.cu (activates twice)
    force_tc2 = (e && e[0] == '1') ? 1 : 0;
  }
  if (!force_tc2) {
    const int rc = snb_gemm_bf16_tc3(Zprev, lda, xa, xc, B, ldb, C, ldc, bias, alpha, M, N, K, stats, Y, ldy, (cudaStream_t)stream);
    if (rc != SNB_ERR_UNSUPPORTED) return rc;
  }
  if (Y) return SNB_ERR_UNSUPPORTED;                 // only the resident-A kernel writes the activated operand back
  return snb_gemm_bf16_tc2(Zprev, lda, 0, B, ldb, 0, C, ldc, bias, alpha, 0, M, N, K, SNB_BF16, stats, (cudaStream_t)stream, 0,
                           nullptr, 0, nullptr, nullptr, nullptr, nullptr, xa, xc);
}

extern "C" int snb_gemm(const void* A, int lda, int a_t, const void* B, int ldb, int b_t, void* C, int ldc,
                        const float* bias, float alpha, int accumulate, long long M, int N, int K, int dtype,
                        int out_dtype, void* stream) {
  SNB_CHECK_ARG(A && B && C && M >= 0 && N > 0 && K > 0 && ldc >= N);
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SNB_BF16) {
    const int rc = snb_gemm_bf16_tc2(A, lda, a_t, B, ldb, b_t, C, ldc, bias, alpha, accumulate, M, N, K, out_dtype, nullptr, st);
    if (rc != SNB_ERR_UNSUPPORTED) return rc;
    return snb_gemm_bf16_tc(A, lda, a_t, B, ldb, b_t, C, ldc, bias, alpha, accumulate, M, N, K, out_dtype, st);
  }

  if (dtype != SNB_F32 || out_dtype != SNB_F32) return SNB_ERR_UNSUPPORTED;
  const long long sam = a_t ? 1 : lda, sak = a_t ? lda : 1;
  const long long sbn = b_t ? 1 : ldb, sbk = b_t ? ldb : 1;
  dim3 grid((unsigned)((M + snb::BM - 1) / snb::BM), (unsigned)((N + snb::BN - 1) / snb::BN));
  snb::gemm_f32_kernel<<<grid, 256, 0, st>>>((const float*)A, sam, sak, (const float*)B, sbn, sbk, (float*)C, ldc, bias,
                                             alpha, accumulate, M, N, K);
  snb::count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
