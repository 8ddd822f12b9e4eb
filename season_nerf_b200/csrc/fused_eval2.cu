// Fused eval-mode T_NeRF network for rendering (sm_100a), CTA-pair edition: two CTAs of a cluster (one TPC) render
// a tile of 256 sample points, 128 rows each, with tcgen05.mma.cta_group::2 (M = 256, N = 256 per instruction).
// Positional encodings are generated in registers, every dense layer is a sequence of MMA steps with TMEM
// accumulators, and the bias + sin epilogue writes bf16 activations straight back into shared memory as the next
// layer's A operand: activations never leave the SM; HBM traffic is the 12-byte point in and the 68 bytes of raw
// head outputs out.  Weights (BatchNorm folded, bf16, 5.8 MB, L2 resident) are streamed by TMA; each CTA stages only
// HALF of every weight tile (its 128 of the 256 N-rows), so the tensor core reads 8 KB of shared memory per K=16
// step and CTA instead of the 12..16 KB of a single-CTA tile, and L2->SMEM traffic per point halves.
//
// Table-driven interpreter of the static program built by season_nerf_b200/packing2.py (schedule, slot plan and
// hazard rules are documented there).  Roles (320 threads per CTA): warp 0 weight producer (both CTAs), warp 1 MMA
// issuer (ONE thread of the leader CTA runs the whole issue loop; TMEM allocation in both), warps 2-9 epilogue: two
// warps per TMEM lane quarter, warp half h drains columns [128h, 128h+128) of a 256-column region = two 64-column
// activation chunks.  A warp hands its part of the TMEM region back to the MMA issuer as soon as the accumulators of its
// last chunk are in registers - before the sin / store work - so the next layer's first MMAs overlap the drain.
// SNB_FUSED_EPI_WARPS=16 selects four warps per lane quarter (one chunk each): measured slower on B200 (1.08 vs 1.14
// PFLOP/s algorithmic sustained; the drain is bound by MUFU.SIN + TMEM reads, not by latency hiding).
// Barriers (same offsets in both CTAs):
//   w_full[5]       leader; expect_tx by the leader's producer, TMA bytes of both CTAs
//   w_empty[5]      each CTA; tcgen05.commit multicast
//   acc_full[2]     each CTA; tcgen05.commit multicast (region complete)
//   acc_empty[2]    leader; 2 x kEpiWarps arrivals (epilogue warps of both CTAs, remote arrive from the peer)
//   chunk_ready[9]  leader; 8 arrivals (4 lane-quarter warps x 2 CTAs)
//   reference semantics: T_NeRF_net_v2.py:75-105,131-151,169-170; G_NeRF.py:74-133; misc.py:105-139,188-189.
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "api.h"

namespace snb {
using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows);

namespace f2 {

constexpr int kSlots = 9;
constexpr int kWStages = 5;
constexpr uint32_t kSlotBytes = 128 * 64 * 2;      // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr uint32_t kWStageBytes = 128 * 64 * 2;    // 16 KB: this CTA's half (<= 128 rows) of a weight tile

constexpr uint32_t kBarBytes = 512;
constexpr uint32_t kSmem = 1024 + kSlots * kSlotBytes + kWStages * kWStageBytes + kBarBytes;

enum { F_ACC = 1, F_WAIT_CHUNK = 2, F_WAIT_EMPTY = 4, F_COMMIT = 8 };
enum { K_ENC_POS = 0, K_ENC_SUN = 1, K_SINE = 2, K_HEAD = 3 };

struct Params {
  const uint4* mma;      // packing2.MMA_DT records
  const uint4* epi;      // packing2.EPI_DT records
  const float* bias;
  uint32_t n_mma, n_epi;
  const float* pts;      // [M,3]
  const float* sun;      // [ceil(M/S),3]
  long long M;
  int S;
  int num_tiles;         // tiles of 256 points
  float* rho_raw;        // [M]
  float* pos4;           // [M,4] = (sigma, colour[3]) raw
  float* vis_raw;        // [M]
  float* adj;            // [M,12]
};

__device__ __forceinline__ uint4 ld_step(const uint4* p) { return __ldg(p); }

__device__ __forceinline__ void store_row_unit(uint32_t slot_addr, int row, int unit, uint4 v) {
  const uint32_t a = slot_addr + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// bias + sin -> bf16 of 32 accumulator columns (four 16-byte units of a 128-byte swizzled activation row)
__device__ __forceinline__ void sine_units(const uint32_t (&r)[32], const float4* __restrict__ bp4, uint32_t slot_addr, int row,
                                           int unit0) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float4 b0 = __ldg(bp4 + 2 * u), b1 = __ldg(bp4 + 2 * u + 1);
    const float v0 = __sinf(__uint_as_float(r[8 * u + 0]) + b0.x), v1 = __sinf(__uint_as_float(r[8 * u + 1]) + b0.y);
    const float v2 = __sinf(__uint_as_float(r[8 * u + 2]) + b0.z), v3 = __sinf(__uint_as_float(r[8 * u + 3]) + b0.w);
    const float v4 = __sinf(__uint_as_float(r[8 * u + 4]) + b1.x), v5 = __sinf(__uint_as_float(r[8 * u + 5]) + b1.y);
    const float v6 = __sinf(__uint_as_float(r[8 * u + 6]) + b1.z), v7 = __sinf(__uint_as_float(r[8 * u + 7]) + b1.w);
    store_row_unit(slot_addr, row, unit0 + u,
                   make_uint4(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3), pack_bf16x2(v4, v5), pack_bf16x2(v6, v7)));
  }
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

// kEpiWarps: 8 (default: two warps per lane quarter, two chunks each) or 16 (one chunk each, SNB_FUSED_EPI_WARPS=16)
template <int kEpiWarps>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (2 + kEpiWarps), 1)
fused_eval2_kernel(const __grid_constant__ CUtensorMap tmapW128, const __grid_constant__ CUtensorMap tmapW8, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t slots_base = smem_base;
  const uint32_t wst_base = smem_base + kSlots * kSlotBytes;
  const uint32_t bar_base = wst_base + kWStages * kWStageBytes;
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (kWStages + s); };
  auto acc_full = [&](int r) { return bar_base + 8u * (2 * kWStages + r); };
  auto acc_empty = [&](int r) { return bar_base + 8u * (2 * kWStages + 2 + r); };
  auto chunk_ready = [&](int s) { return bar_base + 8u * (2 * kWStages + 4 + s); };
  constexpr uint32_t kTmemSlotOff = 8u * (2 * kWStages + 4 + kSlots);
  const uint32_t tmem_slot = bar_base + kTmemSlotOff;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_al + kSlots * kSlotBytes + kWStages * kWStageBytes + kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(acc_full(r), 1);
      mbar_init(acc_empty(r), 2 * kEpiWarps);
    }
    for (int s = 0; s < kSlots; ++s) mbar_init(chunk_ready(s), 8);
    fence_barrier_init();
    tma_prefetch_desc(&tmapW128);
    tma_prefetch_desc(&tmapW8);
  }
  if (warp == 1) tmem_alloc_cg2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t leader_bar_base = mapa_shared(bar_base, 0);     // shared::cluster address of the leader's barrier block

  if (warp == 0) {
    // ===================== weight producer: each CTA loads its half of every weight tile =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
        uint4 raw = ld_step(p.mma);
        for (uint32_t i = 0; i < p.n_mma; ++i) {
          const uint4 nxt = ld_step(p.mma + (i + 1 < p.n_mma ? i + 1 : 0));
          const uint32_t w_row = raw.x;
          const uint32_t n = raw.y & 0xFFFFu;
          const uint32_t half = n >> 1;
          mbar_wait(w_empty(stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx(w_full(stage), n * 128u);
          const uint32_t fb = leader_bar_base + 8u * stage;
          tma_load_2d_cg2(wst_base + stage * kWStageBytes, half == 128 ? &tmapW128 : &tmapW8, fb, 0, (int)(w_row + rank * half));
          if (++stage == kWStages) { stage = 0; phase ^= 1; }
          raw = nxt;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    // The whole loop runs in ONE elected thread (no per-step elect / warp re-convergence): the time between two steps'
    // issue has to stay well below the 512 tensor clocks one step keeps the pipe busy.
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t chunk_par = 0;        // parity to wait for, per slot
      uint32_t empty_par = 0x3;      // first use of every region passes immediately
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
        uint4 raw = ld_step(p.mma);
        for (uint32_t i = 0; i < p.n_mma; ++i) {
          const uint4 nxt = ld_step(p.mma + (i + 1 < p.n_mma ? i + 1 : 0));
          const uint32_t n = raw.y & 0xFFFFu, a_slot = (raw.y >> 16) & 0xFF, flags = (raw.y >> 24) & 0xFF;
          const uint32_t d_col = raw.z & 0xFFFFu, regions = (raw.z >> 16) & 0xFF;
          if (flags & F_WAIT_CHUNK) {
            mbar_wait_cluster(chunk_ready(a_slot), (chunk_par >> a_slot) & 1);
            chunk_par ^= 1u << a_slot;
          }
          if (flags & F_WAIT_EMPTY) {
            const uint32_t r = regions & 15;
            mbar_wait_cluster(acc_empty(r), (empty_par >> r) & 1);
            empty_par ^= 1u << r;
          }
          mbar_wait(w_full(stage), phase);
          tc_fence_after();
          {
            const uint32_t idesc = make_idesc_bf16(256, (int)n, 0, 0);
            const uint32_t sa = slots_base + a_slot * kSlotBytes;
            const uint32_t sb = wst_base + stage * kWStageBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_cg2(tmem_base + d_col, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                           ((flags & F_ACC) || k > 0) ? 1u : 0u);
            umma_commit_cg2_mc(w_empty(stage), 3);
            if (flags & F_COMMIT) umma_commit_cg2_mc(acc_full(regions >> 4), 3);
          }
          if (++stage == kWStages) { stage = 0; phase ^= 1; }
          raw = nxt;
        }
      }
    }
  } else {
    // ===================== epilogue / activation warps (both CTAs) =====================
    const int wq = warp & 3;                 // TMEM lane quarter
    constexpr int kChunksPerWarp = 16 / kEpiWarps;
    const int h = (warp - 2) >> 2;           // which 64-column chunk(s) of a region this warp drains
    const int row = wq * 32 + lane;
    const uint32_t l_acc_empty = leader_bar_base + 8u * (2 * kWStages + 2);
    const uint32_t l_chunk_ready = leader_bar_base + 8u * (2 * kWStages + 4);
    uint32_t full_par = 0;
    for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
      const long long m = (long long)tile * 256 + (long long)rank * 128 + row;
      const bool valid = m < p.M;
      uint4 raw = ld_step(p.epi);
      for (uint32_t i = 0; i < p.n_epi; ++i) {
        const uint4 nxt = ld_step(p.epi + (i + 1 < p.n_epi ? i + 1 : 0));
        const uint32_t kind = raw.x & 0xFF, region = (raw.x >> 8) & 0xFF, also = (raw.x >> 16) & 0xFF;
        const uint32_t d_col = raw.y & 0xFFFFu, out_id = (raw.y >> 16) & 0xFF, out_cols = (raw.y >> 24) & 0xFF;
        const uint32_t bias_off = raw.z;
        const uint32_t dstw = raw.w;
        if (kind == K_ENC_POS) {
          if (h == 0) {
            const uint32_t dst0 = dstw & 0xFF;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (valid) x0 = __ldg(p.pts + 3 * m), x1 = __ldg(p.pts + 3 * m + 1), x2 = __ldg(p.pts + 3 * m + 2);
            encode_row<10>(x0, x1, x2, slots_base + dst0 * kSlotBytes, row);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_chunk_ready + 8u * dst0);
          }
        } else if (kind == K_ENC_SUN) {
          if (h == 1) {
            const uint32_t dst0 = dstw & 0xFF;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (valid) {
              const long long ray = m / p.S;
              x0 = __ldg(p.sun + 3 * ray), x1 = __ldg(p.sun + 3 * ray + 1), x2 = __ldg(p.sun + 3 * ray + 2);
            }
            encode_row<4>(x0, x1, x2, slots_base + dst0 * kSlotBytes, row);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_chunk_ready + 8u * dst0);
          }
        } else {
          mbar_wait(acc_full(region), (full_par >> region) & 1);
          full_par ^= 1u << region;
          if (also != 0xFF) mbar_wait(acc_full(also), (full_par >> also) & 1);
          tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + d_col;
          if (kind == K_SINE) {
#pragma unroll 1
            for (int cc = 0; cc < kChunksPerWarp; ++cc) {
              const int chunk = kChunksPerWarp * h + cc;           // 64-column chunk of the region
              const uint32_t dst = (dstw >> (8 * chunk)) & 0xFF;
              const uint32_t slot_addr = slots_base + dst * kSlotBytes;
              const float4* bp4 = reinterpret_cast<const float4*>(p.bias + bias_off + 64 * chunk);
              uint32_t r0[32], r1[32];
              tmem_ld_32x32(t_row + 64 * chunk, r0);
              tmem_ld_32x32(t_row + 64 * chunk + 32, r1);
              tmem_ld_wait();
              if (cc == kChunksPerWarp - 1) {
                // the accumulators of this warp's part of the region are in registers: hand the TMEM region back to the
                // MMA issuer BEFORE the sin / store work, so that the next layer's first MMAs overlap this drain
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(l_acc_empty + 8u * region);
              }
              sine_units(r0, bp4, slot_addr, row, 0);
              sine_units(r1, bp4 + 8, slot_addr, row, 4);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(l_chunk_ready + 8u * dst);
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
            if (lane == 0) mbar_arrive_cluster(l_acc_empty + 8u * region);
          }
        }
        raw = nxt;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer must not exit while the leader's MMAs / commits still target its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

}  // namespace f2
}  // namespace snb

using namespace snb;

extern "C" int snb_fused_eval2(const void* program, unsigned n_mma, unsigned n_epi, unsigned mma_off, unsigned epi_off,
                               unsigned bias_off, unsigned w_off, unsigned w_rows, const float* pts, long long M, int S,
                               const float* sun, float* rho_raw, float* pos4, float* vis_raw, float* adj, void* stream) {
  SNB_CHECK_ARG(program && pts && M >= 0 && S >= 1 && n_mma > 0 && n_epi > 0 && w_rows > 0);
  SNB_CHECK_ARG((((uintptr_t)program) & 127) == 0 && (mma_off & 15) == 0 && (epi_off & 15) == 0 && (w_off & 127) == 0);
  SNB_CHECK_ARG((!adj || (((uintptr_t)adj) & 15) == 0) && (!pos4 || (((uintptr_t)pos4) & 15) == 0));
  if (M == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(program);
  f2::Params p;
  p.mma = reinterpret_cast<const uint4*>(base + mma_off);
  p.epi = reinterpret_cast<const uint4*>(base + epi_off);
  p.bias = reinterpret_cast<const float*>(base + bias_off);
  p.n_mma = n_mma, p.n_epi = n_epi;
  p.pts = pts, p.sun = sun, p.M = M, p.S = S;
  p.num_tiles = (int)((M + 255) / 256);
  p.rho_raw = rho_raw, p.pos4 = pos4, p.vis_raw = vis_raw, p.adj = adj;
  CUtensorMap t128, t8;
  int rc = make_tmap_bf16(&t128, base + w_off, w_rows, 64, 64, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&t8, base + w_off, w_rows, 64, 64, 64, 8);
  if (rc) return rc;
  static int epi_warps = 0;
  if (!epi_warps) {
    const char* ev = getenv("SNB_FUSED_EPI_WARPS");
    epi_warps = (ev && atoi(ev) == 16) ? 16 : 8;
    cudaError_t e = cudaFuncSetAttribute(f2::fused_eval2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, f2::kSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(f2::fused_eval2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, f2::kSmem);
    if (e != cudaSuccess) {
      epi_warps = 0;
      return (int)e;
    }
  }
  const int num_pairs = num_sms() / 2;
  const int grid = 2 * (p.num_tiles < num_pairs ? p.num_tiles : num_pairs);
  if (epi_warps == 8)
    f2::fused_eval2_kernel<8><<<grid, 32 * 10, f2::kSmem, st>>>(t128, t8, p);
  else
    f2::fused_eval2_kernel<16><<<grid, 32 * 18, f2::kSmem, st>>>(t128, t8, p);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
