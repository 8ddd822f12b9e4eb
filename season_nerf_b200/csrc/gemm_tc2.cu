// CTA-pair bf16 GEMM for sm_100a (the large per-layer GEMMs of the TRAINING path: forward, input gradient, weight
// gradient).  Two CTAs of a cluster (one TPC) cooperate on a 256 x 256 output tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N-rows), so that per CTA the
// tensor core reads 8 KB of shared memory per K=16 step instead of 12 KB and L2->SMEM traffic per flop halves
// with respect to the single-CTA 128 x 256 tile of gemm_tc.cu.  5-stage TMA ring (32 KB / stage / CTA), two TMEM
// accumulator stages (2 x 256 columns) so that the epilogue of tile i overlaps the main loop of tile i+1.
//
//   C[M,N] = alpha * (A[M,K] . B[N,K]^T + bias[N])   (+ C | atomically added for split-K)
//
// Epilogue (4 warps per CTA, warp q owns TMEM lanes 32q..32q+31 = 32 output rows):
//   * bf16 store mode: TMEM -> registers -> (alpha, bias) -> bf16 -> 128B-swizzled shared staging (4 KB per warp,
//     double buffered) -> TMA tensor store (fully coalesced 128-byte rows, M/N tails clipped by the tensor map).
//     Optionally fused: per-column sum and sum of squares of the STORED (bf16-rounded) values - the train-mode
//     BatchNorm statistics of the layer (misc.py:169-170,189) - accumulated in shared memory over all tiles of the
//     CTA and flushed with one atomicAdd per column per CTA at the end.
//   * direct mode (fp32 output, accumulate, split-K atomics): registers -> global, vectorised (red.global.add.v4.f32).
//
// Where the time goes (r02, in-kernel timeline scripts/tc2_timeline.py -> profiles/r02_tc2_timeline_*.txt, trunk shape): every
// variant moves 50-63 bytes per clock through the SM's port to the L2 fabric, reads and writes together (forward + statistics:
// 256 KB in + 128 KB out per 256 x 256 x 512 item in 6.1 k clocks; sin epilogue: 512 KB in 9.2 k; input gradient: 448 KB in
// 8.7 k; 4.1 k clocks of MMA each).  A reloaded stage lands 3-5 k clocks after its request whether or not the rows were
// prefetched into L2 (measured: an L2 prefetch of the next items' A rows made every variant 4-6 % slower) - the requests
// queue at the port, they do not wait for HBM.  The remaining lever is bytes per item, not latency hiding: the resident-A
// kernel (gemm_tc3.cu) reads A once for both 256-column halves.
//
// Barrier protocol per CTA pair (leader = cluster rank 0):
//   full[s]   leader only; 1 arrival (leader's expect_tx) + the TMA bytes of BOTH CTAs
//   empty[s]  each CTA; released by tcgen05.commit multicast from the leader's MMA thread
//   tfull[a]  each CTA; accumulator stage complete (commit multicast)
//   tempty[a] leader only; 8 arrivals = 4 epilogue warps x 2 CTAs (remote mbarrier.arrive from the peer)
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "api.h"

namespace snb {
using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows);

constexpr int k2BM = 128;                       // rows of A / C per CTA (pair tile: 256)
constexpr int k2BK = 64;
// warps: 0 = TMA producer, 1 = MMA issuer, 2.. = epilogue (4 warps; 8 for the input-gradient epilogue kEpi 2, whose
// per-element work (cos, BatchNorm-backward sums) would otherwise outlast the main loop: two warps per TMEM lane quarter,
// each taking half of the tile's 64-column chunks)
// kEpi 3 = kEpi 0 with 8 epilogue warps: the forward GEMM with fused BatchNorm statistics (measured on the trunk shape:
// 205 -> 197 us inside a training step; the same change for the sin epilogue, kEpi 1, measured 248 -> 253 us and was
// dropped.  For reference, cuBLAS and the plain-store kernel both take 179 us on this shape, scripts/gemm_probe.py.)
__host__ __device__ constexpr int k2_epi_warps(int epi) { return epi >= 2 ? 8 : 4; }
__host__ __device__ constexpr int k2_base_epi(int epi) { return epi == 3 ? 0 : epi; }
// kXf = 1: four TRANSFORM warps (2..5) sit between the TMA producer and the MMA issuer: the A operand arrives as the saved
// pre-activation Z of the previous SIREN layer and is rewritten in shared memory as Y = sin(xa[k] * z + xc[k]) (the folded
// BatchNorm affine of that layer, per K column) before the tensor core reads it - the consumer-side activation: the
// stand-alone sin pass over the [M,K] matrix and the Y stream through HBM disappear (misc.py:188-189 for the NEXT layer's input).
__host__ __device__ constexpr int k2_xf_warps(int xf) { return xf ? 4 : 0; }
__host__ __device__ constexpr int k2_threads(int epi, int xf = 0) { return 64 + 32 * k2_xf_warps(xf) + 32 * k2_epi_warps(epi); }
constexpr int k2MaxXfK = 1024;
constexpr uint32_t k2XfBytes = 2 * k2MaxXfK * 4;              // xa[K], xc[K]
constexpr int k2MaxBN = 256;
constexpr uint32_t k2ABytes = k2BM * k2BK * 2;                 // 16 KB
constexpr uint32_t k2BBytes = (k2MaxBN / 2) * k2BK * 2;        // 16 KB (this CTA's half of the B tile)
constexpr uint32_t k2StageBytes = k2ABytes + k2BBytes;         // 32 KB
constexpr uint32_t k2CWarpBytes = 2 * 32 * 128;                // two 32-row x 128-byte staging buffers per epilogue warp
constexpr uint32_t k2CBytes = 4 * k2CWarpBytes;                // 32 KB
constexpr int k2MaxStatN = 1024;
constexpr uint32_t k2StatBytes = 2 * k2MaxStatN * 4;           // 8 KB
// the fused SIREN epilogues need 64 KB of staging (kEpi 1: Z and Y tiles of 4 warps; kEpi 2: G tiles of 8 warps):
// one ring stage less
__host__ __device__ constexpr int k2_stages(int epi) { return epi ? 4 : 5; }
__host__ __device__ constexpr uint32_t k2_cbytes(int epi) { return (uint32_t)k2_epi_warps(epi) * k2CWarpBytes; }
__host__ __device__ constexpr uint32_t k2_xbytes(int epi) { return epi == 1 ? k2CBytes : 0u; }
__host__ __device__ constexpr uint32_t k2_smem(int epi, int xf = 0) {
  return 1024 + k2_stages(epi) * k2StageBytes + k2_cbytes(epi) + k2_xbytes(epi) + k2StatBytes + (xf ? k2XfBytes : 0u) + 256;
}

struct Gemm2Params {
  long long M;
  int N, K;
  int block_n;                 // UMMA N of the pair (multiple of 32; 128 or 256 if B is MN-major)
  int tiles_m, tiles_n, splits, kb_per_split;
  void* C;
  int ldc;
  int out_bf16;
  int mode;                    // 0 store, 1 accumulate, 2 atomic add (fp32)
  int tma_store;               // bf16 store through shared memory + TMA
  const float* bias;
  float alpha;
  float* stats;                // [2*N] column sum / sum of squares of the stored values, or null
  // fused SIREN epilogues (tmapX = 4th tensor map, same shape and box as C):
  //  kEpi 1 (forward, layer without BatchNorm):  C = z = alpha*(acc+bias) (bf16),  X = sin(z)              misc.py:189
  //  kEpi 2 (input-gradient GEMM feeding a sine layer whose pre-activation Z = X[M,N] (row pitch ldx) was saved):
  //          C = g = alpha*acc * cos(ea*z + ec);  stats[0..N) += sum g,  stats[N..2N) += sum g*(z-emean)*einvstd
  const __nv_bfloat16* X;
  int ldx;
  const float* ea;
  const float* ec;
  const float* emean;
  const float* einvstd;
  // kXf: A[m,k] <- sin(xa[k] * A[m,k] + xc[k]) in shared memory before the MMA
  const float* xa;
  const float* xc;
  long long* dbg;              // clock stamps of CTA 0 (SNB_TC2_TIMELINE, scripts/tc2_timeline.py); nullptr in production
};

// timeline of the first kDbg2Items work items of CTA 0: producer [it*8+kb] = stage free / loads issued; MMA thread
// 200 + it*24 + kb*3 + {0: operands landed, 1: MMAs issued}, 200 + it*24 + 23 = accumulator stage free; epilogue warp 0
// 600 + it*2 + {0: accumulator complete, 1: epilogue done}
constexpr int kDbg2Items = 12;
#define SNB_TL2(cond, idx)                                                      \
  do {                                                                          \
    if (p.dbg != nullptr && blockIdx.x == 0 && (cond)) p.dbg[(idx)] = clock64(); \
  } while (0)

template <bool kAT, bool kBT, int kEpiT, int kXf = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2_threads(kEpiT, kXf), 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                  const __grid_constant__ CUtensorMap tmapC, const __grid_constant__ CUtensorMap tmapX,
                  const Gemm2Params p) {
  constexpr int kEpi = k2_base_epi(kEpiT);          // epilogue arithmetic; kEpiT also selects the warp count / staging
  constexpr int k2Stages = k2_stages(kEpiT);
  constexpr int kEpiWarps = k2_epi_warps(kEpiT);
  constexpr int kXfWarps = k2_xf_warps(kXf);
  constexpr int kEpiWarp0 = 2 + kXfWarps;                     // first epilogue warp
  constexpr int k2Threads = k2_threads(kEpiT, kXf);
  static_assert(kXf == 0 || (!kAT && !kBT), "the A transform is written for K-major operands");
  constexpr uint32_t kCBytes = k2_cbytes(kEpiT);
  constexpr uint32_t kXBytes = k2_xbytes(kEpiT);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t cstage_base = smem_base + k2Stages * k2StageBytes;
  const uint32_t xstage_base = cstage_base + kCBytes;
  float* stat_smem = reinterpret_cast<float*>(smem_al + k2Stages * k2StageBytes + kCBytes + kXBytes);
  float* xf_smem = stat_smem + 2 * k2MaxStatN;                 // kXf: xa[0..K), xc at +k2MaxXfK
  const uint32_t bar_base = cstage_base + kCBytes + kXBytes + k2StatBytes + (kXf ? k2XfBytes : 0u);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (k2Stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * k2Stages + 4);
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + 5 + s); };      // kXf: this CTA's A tile has landed
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_al + k2Stages * k2StageBytes + kCBytes + kXBytes + k2StatBytes + (kXf ? k2XfBytes : 0u) + 8u * (2 * k2Stages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_items = p.tiles_m * p.tiles_n * p.splits;
  const int num_kb_total = (p.K + k2BK - 1) / k2BK;
  const int half_n = p.block_n >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < k2Stages; ++s) {
      // kXf: the leader's full barrier also collects one arrival per transform warp of BOTH CTAs (operand rewritten)
      mbar_init(full_bar(s), kXf ? 1 + 2 * kXfWarps : 1);
      mbar_init(empty_bar(s), 1);
      if (kXf) mbar_init(afull_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * kEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    if (p.tma_store) tma_prefetch_desc(&tmapC);
    if (kEpi == 1) tma_prefetch_desc(&tmapX);
  }
  if (p.stats) {
    for (int i = threadIdx.x; i < 2 * k2MaxStatN; i += k2Threads) stat_smem[i] = 0.f;
  }
  if (kXf) {
    for (int i = threadIdx.x; i < k2MaxXfK; i += k2Threads) {
      xf_smem[i] = i < p.K ? __ldg(p.xa + i) : 0.f;
      xf_smem[k2MaxXfK + i] = i < p.K ? __ldg(p.xc + i) : 0.f;
    }
  }
  if (warp == 1) tmem_alloc_cg2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // work item -> (tile_m, tile_n, split).  Without split-K: tile_n fastest, so that the pair reuses its A rows from L2.
  // Split-K (weight gradient: few output tiles, K = all rows of the batch): the output tile is the fastest index, so that
  // the pairs that run at the same time share one K range and both operand slabs are fetched from HBM once (the four
  // 256 x 256 tiles of a 512 x 512 weight gradient would otherwise read each operand twice).
  auto decode = [&](int t, int& tm, int& tn, int& ks) {
    if (p.splits > 1) {
      const int tiles = p.tiles_m * p.tiles_n;
      ks = t / tiles;
      const int tile = t - ks * tiles;
      tn = tile % p.tiles_n;
      tm = tile / p.tiles_n;
    } else {
      ks = 0;
      tn = t % p.tiles_n;
      tm = t / p.tiles_n;
    }
  };

  if (warp == 0) {
    // ================= TMA producer (both CTAs; bytes are credited to the leader's full barrier) =================
    if (elect_one()) {
      const uint32_t b_bytes = kBT ? (uint32_t)(((half_n + 63) / 64) * 64 * k2BK * 2) : (uint32_t)(half_n * k2BK * 2);
      const uint32_t tx_pair = kXf ? 2u * b_bytes : 2u * (k2ABytes + b_bytes);
      int stage = 0;
      uint32_t phase = 0;
      int itp = 0;
      for (int t = pair; t < total_items; t += num_pairs, ++itp) {
        int tm, tn, ks;
        decode(t, tm, tn, ks);
        const int m0 = tm * (2 * k2BM) + (int)rank * k2BM;
        const int n0 = tn * p.block_n + (int)rank * half_n;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          SNB_TL2(itp < kDbg2Items && kb - kb0 < 8, itp * 8 + (kb - kb0));
          const uint32_t sa = smem_base + stage * k2StageBytes;
          const uint32_t sb = sa + k2ABytes;
          const uint32_t fb = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_expect_tx(full_bar(stage), tx_pair);
          const int k0 = kb * k2BK;
          if (kXf) {
            // the A tile is rewritten by this CTA's transform warps first: it completes on a LOCAL barrier
            mbar_expect_tx(afull_bar(stage), k2ABytes);
            tma_load_2d(sa, &tmapA, afull_bar(stage), k0, m0);
          } else if (!kAT) {
            tma_load_2d_cg2(sa, &tmapA, fb, k0, m0);                 // box {64 k, 128 m}
          } else {
            tma_load_2d_cg2(sa, &tmapA, fb, m0, k0);                 // box {64 m, 64 k} x 2
            tma_load_2d_cg2(sa + 8192, &tmapA, fb, m0 + 64, k0);
          }
          if (!kBT) {
            tma_load_2d_cg2(sb, &tmapB, fb, k0, n0);                 // box {64 k, half_n}
          } else {
            for (int j = 0; j * 64 < half_n; ++j) tma_load_2d_cg2(sb + j * 8192, &tmapB, fb, n0 + 64 * j, k0);
          }
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * k2BM, p.block_n, kAT ? 1 : 0, kBT ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int itm = 0;
      for (int t = pair; t < total_items; t += num_pairs, ++itm) {
        int tm, tn, ks;
        decode(t, tm, tn, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        SNB_TL2(itm < kDbg2Items && lane == 0, 200 + itm * 24 + 23);
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * k2MaxBN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          SNB_TL2(itm < kDbg2Items && lane == 0 && kb - kb0 < 7, 200 + itm * 24 + (kb - kb0) * 3);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * k2StageBytes;
            const uint32_t sb = sa + k2ABytes;
#pragma unroll
            for (int k = 0; k < k2BK / 16; ++k) {
              const uint64_t adesc = kAT ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t bdesc = kBT ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
              umma_f16_cg2(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_cg2_mc(empty_bar(stage), 3);                   // frees the stage in both CTAs
            if (kb == kb1 - 1) umma_commit_cg2_mc(tfull_bar(acc), 3);  // accumulator complete -> both epilogues
          }
          __syncwarp();
          SNB_TL2(itm < kDbg2Items && lane == 0 && kb - kb0 < 7, 200 + itm * 24 + (kb - kb0) * 3 + 1);
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (kXf != 0 && warp < kEpiWarp0) {
    // ================= A-operand transform (both CTAs): Z tile -> sin(xa * z + xc), in place =================
    // The A stage is 128 rows of 128 bytes (64 bf16), SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7).
    // Thread tt owns chunk j = tt & 7 of rows r0 + 16 i (r0 = tt >> 3): its 8 K-columns - and their xa / xc - stay the same
    // for all 8 rows, and the 32 lanes of a warp cover 4 whole rows per access (conflict-free 16-byte accesses).
    const int tt = threadIdx.x - 64;
    const int j = tt & 7, r0 = tt >> 3;
    const uint32_t toff = (uint32_t)r0 * 128u + (uint32_t)((j ^ (r0 & 7)) << 4);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < total_items; t += num_pairs) {
      int tm, tn, ks;
      decode(t, tm, tn, ks);
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
      for (int kb = kb0; kb < kb1; ++kb) {
        const float4* xa4 = reinterpret_cast<const float4*>(xf_smem + kb * k2BK + 8 * j);
        const float4* xc4 = reinterpret_cast<const float4*>(xf_smem + k2MaxXfK + kb * k2BK + 8 * j);
        const float4 a_lo = xa4[0], a_hi = xa4[1], c_lo = xc4[0], c_hi = xc4[1];
        const float xa[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
        const float xc[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
        mbar_wait(afull_bar(stage), phase);
        const uint32_t sa = smem_base + stage * k2StageBytes + toff;
        uint32_t w[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w[i][0]), "=r"(w[i][1]), "=r"(w[i][2]), "=r"(w[i][3])
                       : "r"(sa + (uint32_t)i * 2048u));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z0 = __uint_as_float(w[i][e] << 16), z1 = __uint_as_float(w[i][e] & 0xFFFF0000u);
            w[i][e] = pack_bf16x2(__sinf(fmaf(xa[2 * e], z0, xc[2 * e])), __sinf(fmaf(xa[2 * e + 1], z1, xc[2 * e + 1])));
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + (uint32_t)i * 2048u), "r"(w[i][0]), "r"(w[i][1]),
                       "r"(w[i][2]), "r"(w[i][3])
                       : "memory");
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(full_bar(stage), 0));
        if (++stage == k2Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const uint32_t tempty_leader0 = mapa_shared(tempty_bar(0), 0);
    const uint32_t tempty_leader1 = mapa_shared(tempty_bar(1), 0);
    const int ew = warp - kEpiWarp0;        // epilogue warp index
    const int eh = ew >> 2;                 // which half of the tile's chunks (kEpiWarps == 8), else 0
    const uint32_t cbuf = cstage_base + (uint32_t)ew * k2CWarpBytes;
    const uint32_t xbuf = xstage_base + (uint32_t)(ew & 3) * k2CWarpBytes;     // kEpi 1: Y tiles out
    const int c_begin = kEpiWarps == 8 ? eh * (p.block_n >> 1) : 0;
    const int c_end = kEpiWarps == 8 ? c_begin + (p.block_n >> 1) : p.block_n;
    uint32_t cpar = 0;
    // kEpi 2: swizzled shared-memory offsets of this lane's column pair (2*lane, 2*lane+1) in rows r = k (mod 8)
    uint32_t sw_off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sw_off[k] = (uint32_t)((lane & 3) << 2) + (uint32_t)(((lane >> 2) ^ k) << 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    float st_acc[2][4][4];       // [tile column][64-column chunk][sum c0, sum c1, sumsq c0, sumsq c1]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) st_acc[i][j][k] = 0.f;
    int ite = -1;
    for (int t = pair; t < total_items; t += num_pairs) {
      ++ite;
      int tm, tn, ks;
      decode(t, tm, tn, ks);
      const long long row0 = (long long)tm * (2 * k2BM) + (long long)rank * k2BM + q * 32;
      const long long row = row0 + lane;
      const int n0 = tn * p.block_n;
      int rows_valid = 0;
      if (row0 < p.M) rows_valid = (int)min((long long)32, p.M - row0);
      // kEpi 2: the saved pre-activation Z in the column-pair layout (lane j <-> columns 2j, 2j+1; register r <-> row r):
      // 32 coalesced 128-byte row reads per chunk, issued one chunk ahead - the first one while the MMAs still run
      uint32_t zc[32];
      auto load_z = [&](int c, uint32_t (&z)[32]) {
        const uint32_t* zp = reinterpret_cast<const uint32_t*>(p.X + row0 * p.ldx + n0 + c + 2 * lane);
        const long long ldw = p.ldx >> 1;
        if (rows_valid == 32 && n0 + c + 64 <= p.N) {
#pragma unroll
          for (int r = 0; r < 32; ++r) z[r] = __ldg(zp + r * ldw);
        } else {
#pragma unroll
          for (int r = 0; r < 32; ++r) z[r] = (r < rows_valid && n0 + c + 2 * lane < p.N) ? __ldg(zp + r * ldw) : 0u;
        }
      };
      if (kEpi == 2) {
        load_z(c_begin, zc);
        // pull the Z lines of this warp's chunks of the NEXT work item into L2 now (one 128-byte line per lane and
        // chunk): by the time the register loads above are issued for that item they no longer pay the HBM latency,
        // which one chunk of lookahead cannot hide when the epilogue is the critical path
        const int tn_ = t + num_pairs;
        if (tn_ < total_items) {
          int tm2, tn2, ks2;
          decode(tn_, tm2, tn2, ks2);
          const long long r2 = (long long)tm2 * (2 * k2BM) + (long long)rank * k2BM + q * 32 + lane;
          if (r2 < p.M) {
            for (int c = c_begin; c < c_end && tn2 * p.block_n + c < p.N; c += 64) {
              const __nv_bfloat16* zl = p.X + r2 * p.ldx + tn2 * p.block_n + c;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(zl));
            }
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      SNB_TL2(ite < kDbg2Items && ew == 0 && lane == 0, 600 + ite * 2);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * k2MaxBN;
      if (p.tma_store) {
        for (int c = c_begin; c < c_end; c += 64) {
          const int cols_valid = min(64, p.N - (n0 + c));
          if (cols_valid <= 0) break;
          // bias of the 64 columns of this chunk: independent of the accumulator, issue before the TMEM load wait
          float ab[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) ab[i] = 0.f;
          if (kEpi != 2 && p.bias) {
            if (cols_valid == 64) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float4 v = __ldg(b4 + i);
                ab[4 * i] = v.x, ab[4 * i + 1] = v.y, ab[4 * i + 2] = v.z, ab[4 * i + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (i < cols_valid) ab[i] = __ldg(p.bias + n0 + c + i);
            }
          }
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_addr + c, r0);
          tmem_ld_32x32(t_addr + c + 32, r1);
          const uint32_t buf = cbuf + (cpar ? 4096u : 0u);
          if (lane == 0) bulk_wait_group_read<1>();      // the store that last read this buffer has drained it
          __syncwarp();
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t a = u < 4 ? r0[8 * u + e] : r1[8 * (u - 4) + e];
              v[e] = kEpi == 2 ? p.alpha * __uint_as_float(a) : p.alpha * (__uint_as_float(a) + ab[8 * u + e]);
            }
            const uint32_t soff = (uint32_t)lane * 128u + (uint32_t)((u ^ (lane & 7)) << 4);
            const uint32_t addr = buf + soff;
            uint32_t w4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w4[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3])
                         : "memory");
            if (kEpi == 1) {
              // Y = sin(z) of the STORED (bf16-rounded) pre-activation, like the stand-alone activation kernel
              uint32_t y4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                y4[e] = pack_bf16x2(__sinf(__uint_as_float(w4[e] << 16)), __sinf(__uint_as_float(w4[e] & 0xFFFF0000u)));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xbuf + (cpar ? 4096u : 0u) + soff), "r"(y4[0]),
                           "r"(y4[1]), "r"(y4[2]), "r"(y4[3])
                           : "memory");
            }
          }
          if (kEpi == 2) {
            // g = dY * cos(a*z + c) in the column-pair layout: the per-column constants live in registers, and the
            // column sums  sum g, sum g*xhat  (BatchNorm backward; sum g alone is the bias gradient of a layer without
            // BatchNorm) accumulate like the forward statistics.  Rows beyond M contribute zeros: their accumulator
            // rows are zero (TMA zero-fills A) and there is no bias.
            uint32_t zn[32];
            const bool more = c + 64 < c_end && n0 + c + 64 < p.N;
            if (more) load_z(c + 64, zn);
            __syncwarp();
            const int cg = n0 + c + 2 * lane;
            const float2 a2 = __ldg(reinterpret_cast<const float2*>(p.ea + cg));
            const float2 c2 = __ldg(reinterpret_cast<const float2*>(p.ec + cg));
            const float2 m2 = __ldg(reinterpret_cast<const float2*>(p.emean + cg));
            const float2 i2 = __ldg(reinterpret_cast<const float2*>(p.einvstd + cg));
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;     // sum g, sum g*z per column
            // three phases (all loads, all arithmetic, all stores): the volatile shared-memory accesses keep their
            // program order, so interleaving them with the arithmetic would serialise the 32 rows
            uint32_t dw[32];
#pragma unroll
            for (int r = 0; r < 32; ++r)
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(dw[r]) : "r"(buf + (uint32_t)r * 128u + sw_off[r & 7]));
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float z0 = __uint_as_float(zc[r] << 16), z1 = __uint_as_float(zc[r] & 0xFFFF0000u);
              const float g0 = __uint_as_float(dw[r] << 16) * __cosf(fmaf(a2.x, z0, c2.x));
              const float g1 = __uint_as_float(dw[r] & 0xFFFF0000u) * __cosf(fmaf(a2.y, z1, c2.y));
              dw[r] = pack_bf16x2(g0, g1);
              s0 += g0, s1 += g1;
              q0 = fmaf(g0, z0, q0), q1 = fmaf(g1, z1, q1);
            }
#pragma unroll
            for (int r = 0; r < 32; ++r)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(buf + (uint32_t)r * 128u + sw_off[r & 7]), "r"(dw[r]) : "memory");
            // sum g*xhat = invstd * (sum g*z - mean * sum g)
            q0 = i2.x * (q0 - m2.x * s0), q1 = i2.y * (q1 - m2.y * s1);
#pragma unroll
            for (int tnn = 0; tnn < 2; ++tnn)
#pragma unroll
              for (int cc = 0; cc < 2; ++cc)
                if (tn == tnn && c == c_begin + 64 * cc) {
                  st_acc[tnn][cc][0] += s0, st_acc[tnn][cc][1] += s1;
                  st_acc[tnn][cc][2] += q0, st_acc[tnn][cc][3] += q1;
                }
            if (more) {
#pragma unroll
              for (int r = 0; r < 32; ++r) zc[r] = zn[r];
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows_valid > 0) {
            if (p.mode == 1) tma_reduce_add_2d(&tmapC, buf, n0 + c, (int)row0);
            else tma_store_2d(&tmapC, buf, n0 + c, (int)row0);
            if (kEpi == 1) tma_store_2d(&tmapX, xbuf + (cpar ? 4096u : 0u), n0 + c, (int)row0);
            bulk_commit_group();
          }
          if (kEpi != 2 && p.stats) {
            // column statistics of the stored bf16 values: lane j owns columns (2j, 2j+1) of the chunk and keeps
            // their partial sums in registers across all tiles of the CTA (N <= 512: 2 tile columns x 4 chunks)
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            const uint32_t lbase = buf + (uint32_t)((lane & 3) << 2);
            if (rows_valid == 32) {
#pragma unroll
              for (int r8 = 0; r8 < 32; r8 += 8) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int r = r8 + i;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"(lbase + (uint32_t)r * 128u + (uint32_t)(((lane >> 2) ^ (r & 7)) << 4)));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xFFFF0000u);
                  s0 += lo, s1 += hi;
                  q0 = fmaf(lo, lo, q0), q1 = fmaf(hi, hi, q1);
                }
              }
            } else {
              for (int r = 0; r < rows_valid; ++r) {
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(lbase + (uint32_t)r * 128u + (uint32_t)(((lane >> 2) ^ (r & 7)) << 4)));
                const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xFFFF0000u);
                s0 += lo, s1 += hi;
                q0 = fmaf(lo, lo, q0), q1 = fmaf(hi, hi, q1);
              }
            }
#pragma unroll
            for (int tnn = 0; tnn < 2; ++tnn)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc)
                if (tn == tnn && c == c_begin + 64 * cc) {
                  st_acc[tnn][cc][0] += s0, st_acc[tnn][cc][1] += s1;
                  st_acc[tnn][cc][2] += q0, st_acc[tnn][cc][3] += q1;
                }
            __syncwarp();
          }
          cpar ^= 1;
        }
      } else {
        for (int c = 0; c < p.block_n; c += 32) {
          const int ncol = min(32, p.N - (n0 + c));
          if (ncol <= 0) break;
          float ab[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) ab[i] = (p.bias && i < ncol) ? __ldg(p.bias + n0 + c + i) : 0.f;
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c, r);
          tmem_ld_wait();
          if (row < p.M) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = p.alpha * (__uint_as_float(r[i]) + ab[i]);
            if (p.out_bf16) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + n0 + c;
              const bool vec = ncol == 32 && (p.ldc & 7) == 0 && ((n0 + c) & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
              if (vec) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  if (p.mode == 1) {
                    const uint4 o = *reinterpret_cast<const uint4*>(dst + i);
                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      v[i + 2 * e] += __uint_as_float(ow[e] << 16);
                      v[i + 2 * e + 1] += __uint_as_float(ow[e] & 0xFFFF0000u);
                    }
                  }
                  *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                                                   pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
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
              const bool vec = ncol == 32 && (p.ldc & 3) == 0 && ((n0 + c) & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
              if (p.mode == 2) {
                if (vec) {
#pragma unroll
                  for (int i = 0; i < 32; i += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                                 "f"(v[i + 2]), "f"(v[i + 3])
                                 : "memory");
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (i < ncol) atomicAdd(dst + i, v[i]);
                }
              } else if (vec) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                  if (p.mode == 1) {
                    const float4 old = *reinterpret_cast<const float4*>(dst + i);
                    o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
                  }
                  *reinterpret_cast<float4*>(dst + i) = o;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < ncol) dst[i] = (p.mode == 1) ? dst[i] + v[i] : v[i];
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
      SNB_TL2(ite < kDbg2Items && ew == 0 && lane == 0, 600 + ite * 2 + 1);
      if (p.dbg != nullptr && blockIdx.x < 2 && ite < kDbg2Items && lane == 0) p.dbg[700 + (ite * 2 + blockIdx.x) * 8 + ew] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.tma_store && lane == 0) bulk_wait_group<0>();
    if (p.stats) {
      // flush the CTA's column statistics: combine the four warps in shared memory, then one atomic per column
#pragma unroll
      for (int tnn = 0; tnn < 2; ++tnn)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int col = tnn * p.block_n + c_begin + 64 * cc + 2 * lane;
          if (c_begin + 64 * cc < c_end && col + 1 < k2MaxStatN) {
            atomicAdd(stat_smem + col, st_acc[tnn][cc][0]);
            atomicAdd(stat_smem + col + 1, st_acc[tnn][cc][1]);
            atomicAdd(stat_smem + k2MaxStatN + col, st_acc[tnn][cc][2]);
            atomicAdd(stat_smem + k2MaxStatN + col + 1, st_acc[tnn][cc][3]);
          }
        }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      const int te = threadIdx.x - 32 * kEpiWarp0;
      const int ncols = p.N < k2MaxStatN ? p.N : k2MaxStatN;
      for (int i = te; i < ncols; i += 32 * kEpiWarps) {
        const float s = stat_smem[i], ss = stat_smem[k2MaxStatN + i];
        if (s != 0.f || ss != 0.f) {
          atomicAdd(p.stats + i, s);
          atomicAdd(p.stats + p.N + i, ss);
        }
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

}  // namespace snb

using namespace snb;

// returns SNB_ERR_UNSUPPORTED when the shape is better served by the single-CTA kernel (small M / N)
// epi: 0 plain; 1 forward SIREN epilogue (X = second output Y [M,N] bf16); 2 input-gradient epilogue (X = saved Z [M,N] bf16)
int snb_gemm_bf16_tc2(const void* A, int lda, int a_t, const void* B, int ldb, int b_t, void* C, int ldc,
                      const float* bias, float alpha, int accumulate, long long M, int N, int K, int out_dtype,
                      float* stats, cudaStream_t st, int epi, const void* X, int ldx, const float* ea, const float* ec,
                      const float* emean, const float* einvstd, const float* xa, const float* xc) {
  if (M < 256 || N < 128) return SNB_ERR_UNSUPPORTED;
  const bool xf = xa != nullptr;
  if (xf && (a_t || b_t || !xc || K > k2MaxXfK || accumulate != 0 || epi != 0 || !stats)) return SNB_ERR_UNSUPPORTED;
  if (stats && (N > 512 || accumulate != 0 || out_dtype != SNB_BF16)) return SNB_ERR_UNSUPPORTED;
  if (epi) {
    if (accumulate != 0 || out_dtype != SNB_BF16 || (N % 64) != 0 || N > 512 || !X || (ldx % 8) != 0 || (((uintptr_t)X) & 15) != 0)
      return SNB_ERR_UNSUPPORTED;
    if (epi == 1 && (a_t || b_t || stats)) return SNB_ERR_UNSUPPORTED;
    if (epi == 2 && (a_t || !b_t || !stats || bias || !ea || !ec || !emean || !einvstd)) return SNB_ERR_UNSUPPORTED;
    if (epi != 1 && epi != 2) return SNB_ERR_ARG;
  }
  SNB_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && (((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0);
  SNB_CHECK_ARG(out_dtype == SNB_F32 || out_dtype == SNB_BF16);
  SNB_CHECK_ARG(accumulate >= 0 && accumulate <= 2 && !(accumulate == 2 && out_dtype != SNB_F32));
  Gemm2Params p;
  p.M = M, p.N = N, p.K = K;
  p.block_n = N >= 192 ? 256 : 128;
  p.tiles_m = (int)((M + 2 * k2BM - 1) / (2 * k2BM));
  p.tiles_n = (N + p.block_n - 1) / p.block_n;
  const int num_pairs = num_sms() / 2;
  const int num_kb = (K + k2BK - 1) / k2BK;
  int splits = 1;
  if (accumulate == 2) {  // split-K for the (few output tiles, huge K) weight-gradient shape
    const int tiles = p.tiles_m * p.tiles_n;
    splits = (2 * num_pairs + tiles - 1) / tiles;
    if (splits > num_kb) splits = num_kb;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.C = C, p.ldc = ldc, p.out_bf16 = out_dtype == SNB_BF16, p.mode = accumulate, p.bias = bias, p.alpha = alpha;
  p.stats = stats;
  p.ea = ea, p.ec = ec, p.emean = emean, p.einvstd = einvstd;
  p.xa = xa, p.xc = xc;
  p.X = reinterpret_cast<const __nv_bfloat16*>(X), p.ldx = ldx;
  const char* tl = getenv("SNB_TC2_TIMELINE");      // device pointer (decimal) of >= 1024 int64: debugging only
  p.dbg = tl ? reinterpret_cast<long long*>(strtoull(tl, nullptr, 10)) : nullptr;
  p.tma_store = (out_dtype == SNB_BF16 && accumulate <= 1 && (ldc % 8) == 0 && (((uintptr_t)C) & 15) == 0) ? 1 : 0;
  if ((stats || epi) && !p.tma_store) return SNB_ERR_UNSUPPORTED;

  CUtensorMap ta, tb, tcm, tx;
  int rc;
  const int half_n = p.block_n / 2;
  if (!a_t) rc = make_tmap_bf16(&ta, A, M, K, lda, k2BK, k2BM);
  else rc = make_tmap_bf16(&ta, A, K, M, lda, 64, k2BK);
  if (rc) return rc;
  if (!b_t) rc = make_tmap_bf16(&tb, B, N, K, ldb, k2BK, half_n);
  else rc = make_tmap_bf16(&tb, B, K, N, ldb, 64, k2BK);
  if (rc) return rc;
  if (p.tma_store) {
    rc = make_tmap_bf16(&tcm, C, M, N, ldc, 64, 32);
    if (rc) return rc;
  } else {
    tcm = ta;
  }
  if (epi == 1) {
    rc = make_tmap_bf16(&tx, X, M, N, ldx, 64, 32);
    if (rc) return rc;
  } else {
    tx = ta;
  }
  const int total = p.tiles_m * p.tiles_n * p.splits;
  const int grid = 2 * (total < num_pairs ? total : num_pairs);
#define SNB_LAUNCH_GEMM2X(AT, BT, EPI, XF)                                                                   \
  do {                                                                                                       \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<AT, BT, EPI, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           k2_smem(EPI, XF));                                                \
      if (e != cudaSuccess) return (int)e;                                                                   \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    gemm2_bf16_kernel<AT, BT, EPI, XF><<<grid, k2_threads(EPI, XF), k2_smem(EPI, XF), st>>>(ta, tb, tcm, tx, p); \
  } while (0)
#define SNB_LAUNCH_GEMM2(AT, BT, EPI) SNB_LAUNCH_GEMM2X(AT, BT, EPI, 0)
  if (xf) SNB_LAUNCH_GEMM2X(false, false, 3, 1);          // forward + BatchNorm statistics, A = sin(xa * Z_prev + xc)
  else if (epi == 1) SNB_LAUNCH_GEMM2(false, false, 1);
  else if (epi == 2) SNB_LAUNCH_GEMM2(false, true, 2);
  else if (!a_t && !b_t && stats) SNB_LAUNCH_GEMM2(false, false, 3);      // forward + BatchNorm statistics: 8 epilogue warps
  else if (!a_t && !b_t) SNB_LAUNCH_GEMM2(false, false, 0);
  else if (!a_t && b_t) SNB_LAUNCH_GEMM2(false, true, 0);
  else if (a_t && !b_t) SNB_LAUNCH_GEMM2(true, false, 0);
  else SNB_LAUNCH_GEMM2(true, true, 0);
#undef SNB_LAUNCH_GEMM2
#undef SNB_LAUNCH_GEMM2X
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
