// Shared helpers for the sm_100a kernels of the Season-NeRF hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define SNB_OK 0
#define SNB_ERR_ARG (-1)
#define SNB_ERR_UNSUPPORTED (-2)

#define SNB_CHECK_ARG(cond)                \
  do {                                     \
    if (!(cond)) return SNB_ERR_ARG;       \
  } while (0)

#define SNB_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

namespace snb {

// SM count of the current device, queried once per device (a full B200 has 148; a MIG slice or another SKU has fewer:
// every persistent grid and split-K factor is sized from this, never from a compile-time constant)
int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive warp prefix sum
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

inline int grid_for(long long work_items, int per_block, int max_waves = 32) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)num_sms() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace snb
