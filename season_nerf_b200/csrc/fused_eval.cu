// Fused eval-mode T_NeRF network for rendering (sm_100a): positional encoding generated in registers -> every dense
// layer as tcgen05.mma tiles with TMEM accumulators -> bias + sin epilogue written straight back to shared memory
// as the next layer's bf16 A-operand.  Activations of a 128-point tile never leave the SM; HBM traffic is the
// 12-byte point in and the 68 bytes of raw head outputs (sigma, colour, solar visibility, 12 seasonal adjusts) out.
// Weights (BatchNorm folded, bf16, pre-swizzled in HBM into the exact shared-memory image) are streamed by a
// producer warp with 1-D bulk-TMA copies (cp.async.bulk, 16 KB per stage) through a 3-stage mbarrier ring and
// stay L2-resident (5.8 MB).
//
// The kernel is a table-driven interpreter of the static program built by season_nerf_b200/packing.py:
//   MMA steps  : one weight tile [n<=128 x 64] x one activation chunk [128 x 64]  -> 4 tcgen05.mma (K=16 each)
//   epilogue   : encode inputs / drain one 128-column TMEM region (bias, sin, bf16, swizzled st.shared) / heads
// Roles (320 threads): warp 0 weight producer, warp 1 MMA issuer (+TMEM owner), warps 2-9 epilogue (two warps
// per TMEM lane quarter, each taking one 64-column chunk of the region being drained).
// Overlap: N-block 0 of a layer is drained while N-blocks 1..3 still run on the tensor core, and the next layer's
// MMAs start chunk by chunk as the remaining drains publish activation chunks (per-chunk mbarriers).
//   reference semantics: T_NeRF_net_v2.py:75-105,131-151,169-170; G_NeRF.py:74-133; misc.py:105-139,188-189.
#include "common.cuh"
#include "tc_common.cuh"
#include "api.h"

namespace snb {
using namespace tc;

constexpr int kSlots = 11;
constexpr int kWStages = 3;
constexpr uint32_t kSlotBytes = 128 * 64 * 2;      // 16 KB
constexpr uint32_t kWStageBytes = 128 * 64 * 2;    // 16 KB
constexpr int kFusedThreads = 320;
constexpr uint32_t kFusedSmem = kSlots * kSlotBytes + kWStages * kWStageBytes + 512 + 1024;

struct MmaStep {     // 16 bytes, mirrors packing.MMA_DT
  uint32_t w_off16;
  uint16_t w_bytes16;
  uint8_t a_slot, n_div8;
  uint16_t d_col;
  uint8_t flags, regions, ksteps, pad[3];
};
struct EpiStep {     // 16 bytes, mirrors packing.EPI_DT
  uint8_t kind, region, also_region, ncols16;
  uint16_t d_col;
  uint8_t dst0, dst1;
  uint32_t bias_off;
  uint8_t out_id, out_cols, pad[2];
};
struct ProgHeader {
  uint32_t magic, n_mma, n_epi, mma_off, epi_off, bias_off, w_off, total;
};
static_assert(sizeof(MmaStep) == 16 && sizeof(EpiStep) == 16, "table layout");

enum { F_ACC = 1, F_WAIT_CHUNK = 2, F_WAIT_EMPTY = 4, F_COMMIT = 8 };
enum { K_ENC_POS = 0, K_ENC_SUN = 1, K_SINE = 2, K_HEAD = 3 };

struct FusedParams {
  const MmaStep* mma;
  const EpiStep* epi;
  const float* bias;
  const uint8_t* weights;
  uint32_t n_mma, n_epi;
  const float* pts;      // [M,3]
  const float* sun;      // [ceil(M/S),3]
  long long M;
  int S;
  int num_tiles;
  float* rho_raw;        // [M]
  float* pos4;           // [M,4] = (sigma, colour[3]) raw
  float* vis_raw;        // [M]
  float* adj;            // [M,12]
};

__device__ __forceinline__ uint4 ld_step(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// write 64 consecutive columns (8 x 16-byte units) of one row into a swizzled 128x64 bf16 chunk
__device__ __forceinline__ void store_row_unit(uint32_t slot_addr, int row, int unit, uint4 v) {
  const uint32_t a = slot_addr + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// PE_Encode (misc.py:105-139), extended: [x (D) | per dim: cos(k_j x) j<n, sin(k_j x) j<n], k_j = 2^j fl32(pi/2).
template <int kFreq>
__device__ __forceinline__ void encode_row(float x0, float x1, float x2, uint32_t slot_addr, int row) {
  constexpr int kW = 3 * (2 * kFreq + 1);
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.f;
  v[0] = x0, v[1] = x1, v[2] = x2;
  const float xs[3] = {x0, x1, x2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float k = 1.57079637050628662109375f;
#pragma unroll
    for (int j = 0; j < kFreq; ++j) {
      float s, c;
      sincosf(__fmul_rn(k, xs[d]), &s, &c);
      v[3 + d * 2 * kFreq + j] = c;
      v[3 + d * 2 * kFreq + kFreq + j] = s;
      k *= 2.0f;
    }
  }
  static_assert(kW <= 64, "encoding wider than one chunk");
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    uint4 q = make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                         pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
    store_row_unit(slot_addr, row, u, q);
  }
}

__global__ void __launch_bounds__(kFusedThreads, 1) fused_eval_kernel(const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t slots_base = smem_base;
  const uint32_t wst_base = smem_base + kSlots * kSlotBytes;
  const uint32_t bar_base = wst_base + kWStages * kWStageBytes;
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (kWStages + s); };
  auto acc_full = [&](int r) { return bar_base + 8u * (2 * kWStages + r); };
  auto acc_empty = [&](int r) { return bar_base + 8u * (2 * kWStages + 4 + r); };
  auto chunk_ready = [&](int s) { return bar_base + 8u * (2 * kWStages + 8 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWStages + 8 + kSlots);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_al + kSlots * kSlotBytes + kWStages * kWStageBytes + 8u * (2 * kWStages + 8 + kSlots));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int r = 0; r < 4; ++r) {
      mbar_init(acc_full(r), 1);
      mbar_init(acc_empty(r), 8);
    }
    for (int s = 0; s < kSlots; ++s) mbar_init(chunk_ready(s), 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== weight producer: 1-D bulk-TMA copies of pre-swizzled tiles =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (uint32_t i = 0; i < p.n_mma; ++i) {
          const uint4 raw = ld_step(p.mma + i);
          const uint32_t off16 = raw.x;
          const uint32_t bytes = (raw.y & 0xFFFFu) << 4;
          mbar_wait(w_empty(stage), phase ^ 1);
          mbar_expect_tx(w_full(stage), bytes);
          bulk_load(wst_base + stage * kWStageBytes, p.weights + ((size_t)off16 << 4), bytes, w_full(stage));
          if (++stage == kWStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t chunk_par = 0;        // parity to wait for, per slot
    uint32_t empty_par = 0xF;      // first use of every region passes immediately
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (uint32_t i = 0; i < p.n_mma; ++i) {
        const uint4 raw = ld_step(p.mma + i);
        const uint32_t a_slot = (raw.y >> 16) & 0xFF, n = ((raw.y >> 24) & 0xFF) * 8;
        const uint32_t d_col = raw.z & 0xFFFF, flags = (raw.z >> 16) & 0xFF, regions = (raw.z >> 24) & 0xFF;
        if (flags & F_WAIT_CHUNK) {
          mbar_wait(chunk_ready(a_slot), (chunk_par >> a_slot) & 1);
          chunk_par ^= 1u << a_slot;
        }
        if (flags & F_WAIT_EMPTY) {
          const uint32_t r = regions & 15;
          mbar_wait(acc_empty(r), (empty_par >> r) & 1);
          empty_par ^= 1u << r;
        }
        mbar_wait(w_full(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = make_idesc_bf16(128, (int)n, 0, 0);
          const uint32_t sa = slots_base + a_slot * kSlotBytes;
          const uint32_t sb = wst_base + stage * kWStageBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + d_col, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                     ((flags & F_ACC) || k > 0) ? 1u : 0u);
          umma_commit(w_empty(stage));
          if (flags & F_COMMIT) umma_commit(acc_full(regions >> 4));
        }
        __syncwarp();
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue / activation warps =====================
    const int wq = warp & 3;                 // TMEM lane quarter
    const int h = (warp - 2) >> 2;           // which 64-column chunk of a region this warp drains
    const int row = wq * 32 + lane;
    uint32_t full_par = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long m = (long long)tile * 128 + row;
      const bool valid = m < p.M;
      for (uint32_t i = 0; i < p.n_epi; ++i) {
        const uint4 raw = ld_step(p.epi + i);
        const uint32_t kind = raw.x & 0xFF, region = (raw.x >> 8) & 0xFF, also = (raw.x >> 16) & 0xFF;
        const uint32_t d_col = raw.y & 0xFFFF, dst0 = (raw.y >> 16) & 0xFF, dst1 = (raw.y >> 24) & 0xFF;
        const uint32_t bias_off = raw.z, out_id = raw.w & 0xFF, out_cols = (raw.w >> 8) & 0xFF;
        if (kind == K_ENC_POS) {
          if (h == 0) {
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (valid) x0 = __ldg(p.pts + 3 * m), x1 = __ldg(p.pts + 3 * m + 1), x2 = __ldg(p.pts + 3 * m + 2);
            encode_row<10>(x0, x1, x2, slots_base + dst0 * kSlotBytes, row);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(chunk_ready(dst0));
          }
        } else if (kind == K_ENC_SUN) {
          if (h == 1) {
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (valid) {
              const long long ray = m / p.S;
              x0 = __ldg(p.sun + 3 * ray), x1 = __ldg(p.sun + 3 * ray + 1), x2 = __ldg(p.sun + 3 * ray + 2);
            }
            encode_row<4>(x0, x1, x2, slots_base + dst0 * kSlotBytes, row);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(chunk_ready(dst0));
          }
        } else {
          // bias values do not depend on the accumulator: fetch them BEFORE blocking on the commit barrier
          float bv[64];
          if (kind == K_SINE) {
            const float4* bp4 = reinterpret_cast<const float4*>(p.bias + bias_off + 64 * h);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float4 q = __ldg(bp4 + e);
              bv[4 * e] = q.x, bv[4 * e + 1] = q.y, bv[4 * e + 2] = q.z, bv[4 * e + 3] = q.w;
            }
          }
          mbar_wait(acc_full(region), (full_par >> region) & 1);
          full_par ^= 1u << region;
          if (also != 0xFF) mbar_wait(acc_full(also), (full_par >> also) & 1);
          tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + d_col;
          if (kind == K_SINE) {
            const uint32_t slot_addr = slots_base + (h ? dst1 : dst0) * kSlotBytes;
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(t_row + 64 * h, r0);
            tmem_ld_32x32(t_row + 64 * h + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t a = u < 4 ? r0[8 * u + e] : r1[8 * (u - 4) + e];
                v[e] = __sinf(__uint_as_float(a) + bv[8 * u + e]);
              }
              store_row_unit(slot_addr, row, u,
                             make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(chunk_ready(h ? dst1 : dst0));
              mbar_arrive(acc_empty(region));
            }
          } else {  // K_HEAD: 16 accumulator columns -> raw outputs in global memory
            if (h == 0) {
              uint32_t r[16];
              tmem_ld_32x16(t_row, r);
              tmem_ld_wait();
              if (valid) {
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) + __ldg(p.bias + bias_off + e);
                if (out_id == 0) {
                  if (p.rho_raw) p.rho_raw[m] = v[0];
                  if (out_cols >= 4 && p.pos4) *reinterpret_cast<float4*>(p.pos4 + 4 * m) = make_float4(v[0], v[1], v[2], v[3]);
                } else if (out_id == 1) {
                  if (p.vis_raw) p.vis_raw[m] = v[0];
                } else if (p.adj) {
                  float4* o = reinterpret_cast<float4*>(p.adj + 12 * m);
                  o[0] = make_float4(v[0], v[1], v[2], v[3]);
                  o[1] = make_float4(v[4], v[5], v[6], v[7]);
                  o[2] = make_float4(v[8], v[9], v[10], v[11]);
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(region));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace snb

using namespace snb;

extern "C" int snb_fused_eval(const void* program, unsigned n_mma, unsigned n_epi, unsigned mma_off, unsigned epi_off,
                              unsigned bias_off, unsigned w_off, const float* pts, long long M, int S, const float* sun,
                              float* rho_raw, float* pos4, float* vis_raw, float* adj, void* stream) {
  SNB_CHECK_ARG(program && pts && M >= 0 && S >= 1 && n_mma > 0 && n_epi > 0);
  SNB_CHECK_ARG((((uintptr_t)program) & 127) == 0 && (mma_off & 15) == 0 && (epi_off & 15) == 0 && (w_off & 127) == 0);
  SNB_CHECK_ARG((!adj || (((uintptr_t)adj) & 15) == 0) && (!pos4 || (((uintptr_t)pos4) & 15) == 0));
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(program);
  FusedParams p;
  p.mma = reinterpret_cast<const MmaStep*>(base + mma_off);
  p.epi = reinterpret_cast<const EpiStep*>(base + epi_off);
  p.bias = reinterpret_cast<const float*>(base + bias_off);
  p.weights = base + w_off;
  p.n_mma = n_mma, p.n_epi = n_epi;
  p.pts = pts, p.sun = sun, p.M = M, p.S = S;
  p.num_tiles = (int)((M + 127) / 128);
  p.rho_raw = rho_raw, p.pos4 = pos4, p.vis_raw = vis_raw, p.adj = adj;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fused_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  fused_eval_kernel<<<grid, kFusedThreads, kFusedSmem, st>>>(p);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
