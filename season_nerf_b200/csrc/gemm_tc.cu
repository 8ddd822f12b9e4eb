// bf16 tensor-core GEMM for sm_100a: tcgen05.mma with TMEM accumulators, TMA-fed 4-stage shared-memory ring,
// persistent CTAs (one per SM) with two accumulator stages so that the epilogue of tile i overlaps the main
// loop of tile i+1.  Serves the per-layer dense work of the TRAINING path (forward, input gradient, weight
// gradient with split-K) where train-mode BatchNorm forces a grid-wide statistic between layers.
//
//   C[M,N] = alpha * (A[M,K] . B[N,K]^T + bias[N])  (+ C)       fp32 accumulate
//   operands bf16; each may be K-contiguous ("K-major") or stored transposed ("MN-major", used by the weight
//   gradient dW = dZ^T . X so that no transposed copies of the activations are ever materialised).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM -> registers -> global).   reference: nn.Linear inside misc.py:188-189 + autograd.
#include <mutex>
#include "common.cuh"
#include "tc_common.cuh"
#include "api.h"

namespace snb {
using namespace tc;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kMaxBlockN = 256;
constexpr int kGemmThreads = 192;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr uint32_t kBBytes = kMaxBlockN * kBlockK * 2;   // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kGemmSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  long long M;
  int N, K;
  int block_n;       // UMMA N (multiple of 16, <= 256)
  int tiles_m, tiles_n, splits, kb_per_split;
  void* C;
  int ldc;
  int out_bf16;
  int mode;          // 0 store, 1 accumulate (C += ...), 2 atomic add (split-K, fp32 only)
  const float* bias;
  float alpha;
};

template <bool kAT, bool kBT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  // barrier map: full[0..3], empty[4..7], tmem_full[8..9], tmem_empty[10..11], tmem ptr slot at +96
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + kStages * kStageBytes + 8u * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n * p.splits;
  const int num_kb_total = (p.K + kBlockK - 1) / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t b_box_bytes = kBT ? (uint32_t)(((p.block_n + 63) / 64) * 64 * kBlockK * 2) : (uint32_t)(p.block_n * kBlockK * 2);
  const uint32_t tx_bytes = kABytes + b_box_bytes;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t % p.splits;
        const int tn = (t / p.splits) % p.tiles_n;
        const int tm = t / (p.splits * p.tiles_n);
        const int m0 = tm * kBlockM, n0 = tn * p.block_n;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          mbar_expect_tx(full_bar(stage), tx_bytes);
          const int k0 = kb * kBlockK;
          if (!kAT) {
            tma_load_2d(sa, &tmapA, full_bar(stage), k0, m0);               // box {64 k, 128 m}
          } else {
            tma_load_2d(sa, &tmapA, full_bar(stage), m0, k0);               // box {64 m, 64 k}
            tma_load_2d(sa + 8192, &tmapA, full_bar(stage), m0 + 64, k0);
          }
          if (!kBT) {
            tma_load_2d(sb, &tmapB, full_bar(stage), k0, n0);               // box {64 k, block_n}
          } else {
            for (int j = 0; j * 64 < p.block_n; ++j) tma_load_2d(sb + j * 8192, &tmapB, full_bar(stage), n0 + 64 * j, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, kAT ? 1 : 0, kBT ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int ks = t % p.splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kMaxBlockN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = kAT ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = kBT ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0 && elect_one()) umma_commit(tfull_bar(acc));  // empty K range (defensive)
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ================= epilogue: TMEM -> registers -> global =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int tn = (t / p.splits) % p.tiles_n;
      const int tm = t / (p.splits * p.tiles_n);
      const long long row = (long long)tm * kBlockM + q * 32 + lane;
      const int n0 = tn * p.block_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kMaxBlockN;
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t r[32];
        if (p.block_n - c >= 32) {
          tmem_ld_32x32(t_addr + c, r);
        } else {  // block_n is a multiple of 16
          uint32_t h[16];
          tmem_ld_32x16(t_addr + c, h);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = h[i], r[16 + i] = 0;
        }
        tmem_ld_wait();
        if (row < p.M) {
          const int ncol = min(32, p.N - (n0 + c));   // valid columns of this chunk (may be <= 0)
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float b = (p.bias && i < ncol) ? __ldg(p.bias + n0 + c + i) : 0.f;
            v[i] = p.alpha * (__uint_as_float(r[i]) + b);
          }
          if (p.out_bf16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + n0 + c;
            if (ncol == 32 && (p.ldc & 7) == 0 && ((n0 + c) & 7) == 0 && p.mode == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 u = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                     pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
                *reinterpret_cast<uint4*>(dst + i) = u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncol) {
                  float o = v[i];
                  if (p.mode == 1) o += __bfloat162float(dst[i]);
                  dst[i] = __float2bfloat16_rn(o);
                }
            }
          } else {
            float* dst = reinterpret_cast<float*>(p.C) + row * p.ldc + n0 + c;
            if (p.mode == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncol) atomicAdd(dst + i, v[i]);
            } else if (ncol == 32 && (p.ldc & 3) == 0 && ((n0 + c) & 3) == 0 && p.mode == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncol) dst[i] = (p.mode == 1) ? dst[i] + v[i] : v[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] (cols contiguous, row pitch ld elements), box {box_cols, box_rows}, 128B swizzle
int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return SNB_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SNB_OK : SNB_ERR_ARG;
}

}  // namespace snb

using namespace snb;

int snb_gemm_bf16_tc(const void* A, int lda, int a_t, const void* B, int ldb, int b_t, void* C, int ldc,
                     const float* bias, float alpha, int accumulate, long long M, int N, int K, int out_dtype,
                     cudaStream_t st) {
  // TMA needs 16-byte aligned bases and row pitches
  SNB_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && (((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0);
  SNB_CHECK_ARG(out_dtype == SNB_F32 || out_dtype == SNB_BF16);
  SNB_CHECK_ARG(accumulate >= 0 && accumulate <= 2 && !(accumulate == 2 && out_dtype != SNB_F32));
  GemmParams p;
  p.M = M, p.N = N, p.K = K;
  int bn = N >= kMaxBlockN ? kMaxBlockN : ((N + 15) / 16) * 16;
  if (b_t) bn = ((bn + 63) / 64) * 64 > kMaxBlockN ? kMaxBlockN : ((bn + 63) / 64) * 64;
  p.block_n = bn;
  p.tiles_m = (int)((M + kBlockM - 1) / kBlockM);
  p.tiles_n = (N + bn - 1) / bn;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  int splits = 1;
  if (accumulate == 2) {  // split-K for the (few output tiles, huge K) weight-gradient shape
    const int tiles = p.tiles_m * p.tiles_n;
    splits = (2 * num_sms() + tiles - 1) / tiles;
    if (splits > num_kb) splits = num_kb;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.C = C, p.ldc = ldc, p.out_bf16 = out_dtype == SNB_BF16, p.mode = accumulate, p.bias = bias, p.alpha = alpha;

  CUtensorMap ta, tb;
  int rc;
  if (!a_t) rc = make_tmap_bf16(&ta, A, M, K, lda, kBlockK, kBlockM);
  else rc = make_tmap_bf16(&ta, A, K, M, lda, 64, kBlockK);
  if (rc) return rc;
  if (!b_t) rc = make_tmap_bf16(&tb, B, N, K, ldb, kBlockK, bn);
  else rc = make_tmap_bf16(&tb, B, K, N, ldb, 64, kBlockK);
  if (rc) return rc;

  const int total = p.tiles_m * p.tiles_n * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
#define SNB_LAUNCH_GEMM(AT, BT)                                                                             \
  do {                                                                                                      \
    static bool attr_set = false;                                                                           \
    if (!attr_set) {                                                                                        \
      cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<AT, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem); \
      if (e != cudaSuccess) return (int)e;                                                                  \
      attr_set = true;                                                                                      \
    }                                                                                                       \
    gemm_bf16_kernel<AT, BT><<<grid, kGemmThreads, kGemmSmem, st>>>(ta, tb, p);                             \
  } while (0)
  if (!a_t && !b_t) SNB_LAUNCH_GEMM(false, false);
  else if (!a_t && b_t) SNB_LAUNCH_GEMM(false, true);
  else if (a_t && !b_t) SNB_LAUNCH_GEMM(true, false);
  else SNB_LAUNCH_GEMM(true, true);
#undef SNB_LAUNCH_GEMM
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
