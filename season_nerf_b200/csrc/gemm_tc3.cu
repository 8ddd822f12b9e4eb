// Forward GEMM of a SIREN layer whose INPUT activation is applied on the consumer side (training path, sm_100a):
//
//   C[M,N] = alpha * ( sin(xa[k] * Zprev[m,k] + xc[k]) . B[N,K]^T + bias[N] )       + column sum / sum of squares of C
//
// Zprev is the saved pre-activation of the previous layer and xa / xc its folded train-mode BatchNorm affine
// (misc.py:169-170,188-189): the activated matrix Y = sin(BN(Zprev)) is never read from HBM - and, unless a backward pass
// needs it, never written either.  The stand-alone activation pass (403 MB in + 403 MB out per 512-wide layer at 393 216
// rows) disappears from the step.
//
// Layout of the work (CTA pair = cluster of 2, tcgen05.mma.cta_group::2, 256 x 256 output tiles like gemm_tc2.cu):
//   * A is RESIDENT: the 128 x K operand block of a CTA (K <= 512: 8 slots of 128 x 64 bf16, 128 KB) lands once per 256-row
//     tile, is rewritten in place by eight TRANSFORM warps (sin.approx on the MUFU pipe) and then feeds BOTH 256-column passes
//     of the tile - every element is activated exactly once, so the MUFU pipe runs at ~50 % of the tensor time instead
//     of 100 % (transforming a streamed A stage once per N tile made the kernel MUFU bound: 320 us vs 204 us, gemm_tc2.cu kXf).
//     Eight warps = two per scheduler: with four, the in-kernel timeline (scripts/tc3_timeline.py) showed the MMA warp
//     waiting on `aready` while each transform warp spent 1.7 k clocks per slot at an issue rate of 0.16.
//   * B (the weight, L2 resident) streams through a 3-stage ring, 128 rows x 64 k per CTA and stage.
//   * two TMEM accumulator stages (2 x 256 columns): pass 0 / pass 1 of a tile, so the epilogue of one pass overlaps the
//     MMAs of the next; slot kb of the NEXT tile is loaded and transformed as soon as pass 1 has consumed it.
//   * store_y: the transform warps also write the activated values to HBM, straight from their registers (st.global, 16
//     bytes per thread, a quarter warp = one whole 128-byte line; a TMA store would read the slot once more through the
//     shared-memory pipe the tensor core and the TMA fills already saturate) - the image pass keeps Y for the weight
//     gradient of this layer; the no-grad solar pass does not.
//
// Warps (640 threads, setmaxnreg moves registers from the data-movement warps to the epilogue):
//   0 A producer | 1 MMA issuer (leader CTA) | 2 B producer | 3 idle | 4..11 transform | 12..19 epilogue
//
// Measured and NOT kept (r02, profiles/r02_tc3_experiments/): a fourth B stage paid for by 4 epilogue warps (-4 %: the
// epilogue then outlasts a pass); an L2 prefetch of the next tile's operand, per slot or whole (0 / -10 % with the Y
// store); B multicast over a 4-CTA cluster (correct, -8 %: the limit is per SM - the tile period does not change when only
// 36 of the 148 SMs run); feeding A through registers (ld.global two slots ahead, st.shared; -8 %).  The timeline says why:
// per tile 8.2 k clocks of MMA take 14.5 k; the slots of the next tile can only be requested as the last pass drains them,
// they arrive one per ~1.4 k clocks (the SM ingests ~42 B/clk through TMA, shared with the B stream of that pass), and with
// 20 resident warps the transform, the epilogue and the MMA issue loop also compete for issue slots.
// Barriers per CTA (leader = cluster rank 0):
//   afull[kb]   each CTA   A slot landed (own TMA bytes)
//   aready[kb]  leader     16 arrivals = transform warps of both CTAs (operand rewritten, fenced for the async proxy)
//   aempty[kb]  each CTA   slot consumed by the last pass (commit multicast)
//   bfull[s]    leader     B bytes of both CTAs;   bempty[s] each CTA (commit multicast)
//   tfull[a]    each CTA   accumulator complete;   tempty[a] leader, 16 arrivals = 8 epilogue warps x 2 CTAs
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "api.h"

namespace snb {
using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows);

constexpr int k3BM = 128;
constexpr int k3BK = 64;
constexpr int k3MaxKB = 8;                         // K <= 512
constexpr int k3BStages = 3;
constexpr int k3EpiWarps = 8;
constexpr int k3XfWarps = 8;
constexpr int k3Threads = 32 * (4 + k3XfWarps + k3EpiWarps);          // 640
constexpr uint32_t k3SlotBytes = k3BM * k3BK * 2;                      // 16 KB
constexpr uint32_t k3ABytes = k3MaxKB * k3SlotBytes;                   // 128 KB
constexpr uint32_t k3BBytes = k3BStages * k3SlotBytes;                 // 48 KB
constexpr uint32_t k3CWarpBytes = 32 * 128;                            // one 32-row x 128-byte staging buffer per epilogue warp
constexpr uint32_t k3CBytes = k3EpiWarps * k3CWarpBytes;               // 32 KB
constexpr int k3MaxStatN = 512;
constexpr uint32_t k3StatBytes = 2 * k3MaxStatN * 4;                   // 4 KB
constexpr int k3MaxXfK = k3MaxKB * k3BK;                               // 512
constexpr uint32_t k3XfBytes = 2 * k3MaxXfK * 4;                       // 4 KB
constexpr uint32_t k3BiasBytes = k3MaxStatN * 4;                       // 2 KB
constexpr uint32_t k3BarBytes = 512;
constexpr uint32_t k3Smem = 1024 + k3ABytes + k3BBytes + k3CBytes + k3StatBytes + k3XfBytes + k3BiasBytes + k3BarBytes;
static_assert(k3Smem <= 232448, "shared memory budget");

struct Gemm3Params {
  long long M;
  int N, K;
  int tiles_m, n_passes, nkb;
  const float* bias;
  float alpha;
  float* stats;              // [2*N]
  const float* xa;           // [K]
  const float* xc;           // [K]
  __nv_bfloat16* Y;          // activated operand out [M, K] (row pitch ldy) or nullptr
  long long ldy;
  long long* dbg;            // timeline of CTA 0 (SNB_TC3_TIMELINE, scripts/tc3_timeline.py); nullptr in production
};

// clock stamps of the first kDbgTiles tiles of CTA 0: MMA thread [((it*2+tn)*8+kb)*3 + {aready, bfull, issued}], accumulator
// free 300 + it*2+tn, transform thread 0 [400 + (it*8+kb)*2 + {A landed, slot done}], epilogue warp 0 [600 + (it*2+tn)*2 +
// {tfull, done}], A producer 700 + it*8+kb (slot free, load issued)
constexpr int kDbgTiles = 6;
#define SNB_TL(cond, idx)                                                      \
  do {                                                                         \
    if (p.dbg != nullptr && blockIdx.x == 0 && (cond)) p.dbg[(idx)] = clock64(); \
  } while (0)

template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k3Threads, 1)
gemm3_xf_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                const __grid_constant__ CUtensorMap tmapC, const Gemm3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + k3ABytes;
  const uint32_t c_base = b_base + k3BBytes;
  float* stat_smem = reinterpret_cast<float*>(smem_al + k3ABytes + k3BBytes + k3CBytes);
  float* xf_smem = stat_smem + 2 * k3MaxStatN;
  float* bias_smem = xf_smem + 2 * k3MaxXfK;
  const uint32_t bar_base = c_base + k3CBytes + k3StatBytes + k3XfBytes + k3BiasBytes;
  auto afull = [&](int k) { return bar_base + 8u * k; };
  auto aready = [&](int k) { return bar_base + 8u * (k3MaxKB + k); };
  auto aempty = [&](int k) { return bar_base + 8u * (2 * k3MaxKB + k); };
  auto bfull = [&](int s) { return bar_base + 8u * (3 * k3MaxKB + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (3 * k3MaxKB + k3BStages + s); };
  auto tfull = [&](int a) { return bar_base + 8u * (3 * k3MaxKB + 2 * k3BStages + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (3 * k3MaxKB + 2 * k3BStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * k3MaxKB + 2 * k3BStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_al + k3ABytes + k3BBytes + k3CBytes + k3StatBytes + k3XfBytes + k3BiasBytes + 8u * (3 * k3MaxKB + 2 * k3BStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int nkb = p.nkb;

  if (threadIdx.x == 0) {
    for (int k = 0; k < k3MaxKB; ++k) {
      mbar_init(afull(k), 1);
      mbar_init(aready(k), 2 * k3XfWarps);
      mbar_init(aempty(k), 1);
    }
    for (int s = 0; s < k3BStages; ++s) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), 2 * k3EpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    tma_prefetch_desc(&tmapC);
  }
  for (int i = threadIdx.x; i < 2 * k3MaxStatN; i += k3Threads) stat_smem[i] = 0.f;
  for (int i = threadIdx.x; i < k3MaxStatN; i += k3Threads) bias_smem[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
  // operand-transform constants as float4 planes: [k-block][plane: xa 0-3, xa 4-7, xc 0-3, xc 4-7][chunk j][4] - the 8 lanes
  // of a quarter warp (j = 0..7) read 128 contiguous bytes per plane (conflict-free)
  for (int i = threadIdx.x; i < k3MaxXfK; i += k3Threads) {
    const int kb_ = i >> 6, j_ = (i & 63) >> 3, e_ = i & 7;
    const int o = (((kb_ * 4 + (e_ >> 2)) * 8 + j_) << 2) + (e_ & 3);
    xf_smem[o] = i < p.K ? __ldg(p.xa + i) : 0.f;
    xf_smem[o + 64] = i < p.K ? __ldg(p.xc + i) : 0.f;
  }
  if (warp == 1) tmem_alloc_cg2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0) {
      // ================= A producer (both CTAs): one 128 x K block per tile, slot by slot =================
      if (elect_one()) {
        int it = 0;
        for (int tm = pair; tm < p.tiles_m; tm += num_pairs, ++it) {
          const int m0 = tm * (2 * k3BM) + (int)rank * k3BM;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(aempty(kb), (uint32_t)((it & 1) ^ 1));
            SNB_TL(it < kDbgTiles, 700 + it * 8 + kb);
            mbar_expect_tx(afull(kb), k3SlotBytes);
            tma_load_2d(a_base + kb * k3SlotBytes, &tmapA, afull(kb), kb * k3BK, m0);
          }
        }
      }
    } else if (warp == 2) {
      // ================= B producer (both CTAs; bytes credited to the leader's barrier) =================
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int tm = pair; tm < p.tiles_m; tm += num_pairs) {
          for (int tn = 0; tn < p.n_passes; ++tn) {
            const int n0 = tn * 256 + (int)rank * 128;
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(bempty(stage), phase ^ 1);
              const uint32_t fb = mapa_shared(bfull(stage), 0);
              if (rank == 0) mbar_expect_tx(bfull(stage), 2u * k3SlotBytes);
              tma_load_2d_cg2(b_base + stage * k3SlotBytes, &tmapB, fb, kb * k3BK, n0);
              if (++stage == k3BStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer (leader CTA only) =================
      if (rank == 0) {
        const uint32_t idesc = make_idesc_bf16(2 * k3BM, 256, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int item = 0, it = 0;
        for (int tm = pair; tm < p.tiles_m; tm += num_pairs, ++it) {
          for (int tn = 0; tn < p.n_passes; ++tn, ++item) {
            const int acc = item & 1;
            const uint32_t acc_phase = (uint32_t)((item >> 1) & 1);
            mbar_wait_cluster(tempty(acc), acc_phase ^ 1);
            tc_fence_after();
            SNB_TL(it < kDbgTiles && lane == 0, 300 + it * 2 + tn);
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait_cluster(aready(kb), (uint32_t)(it & 1));
              SNB_TL(it < kDbgTiles && lane == 0, ((it * 2 + tn) * 8 + kb) * 3);
              mbar_wait(bfull(stage), phase);
              SNB_TL(it < kDbgTiles && lane == 0, ((it * 2 + tn) * 8 + kb) * 3 + 1);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t sa = a_base + kb * k3SlotBytes;
                const uint32_t sb = b_base + stage * k3SlotBytes;
#pragma unroll
                for (int k = 0; k < k3BK / 16; ++k) {
                  const uint64_t adesc = make_smem_desc(sa + k * 32, 16, 1024);
                  const uint64_t bdesc = make_smem_desc(sb + k * 32, 16, 1024);
                  umma_f16_cg2(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit_cg2_mc(bempty(stage), 3);
                if (tn == p.n_passes - 1) umma_commit_cg2_mc(aempty(kb), 3);      // last pass: the slot may be reloaded
                if (kb == nkb - 1) umma_commit_cg2_mc(tfull(acc), 3);
              }
              __syncwarp();
              SNB_TL(it < kDbgTiles && lane == 0, ((it * 2 + tn) * 8 + kb) * 3 + 2);
              if (++stage == k3BStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp < 4 + k3XfWarps) {
    reg_dec<64>();
    // ================= A-operand transform (both CTAs): Z slot -> sin(xa * z + xc), in place =================
    // slot = 128 rows of 128 bytes (64 bf16), SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7).  Thread tt
    // owns chunk j = tt & 7 of rows r0 + 32 i (r0 = tt >> 3, i < 4): its 8 K-columns and their xa / xc are the same for all
    // its rows; the 32 lanes of a warp cover 4 whole rows per access (conflict-free 16-byte accesses).
    constexpr int kRS = 4 * k3XfWarps;    // row stride between a thread's rows
    constexpr int kRI = k3BM / kRS;       // rows per thread and slot
    const int tt = threadIdx.x - 128;
    const int j = tt & 7, r0 = tt >> 3;
    const uint32_t toff = (uint32_t)r0 * 128u + (uint32_t)((j ^ (r0 & 7)) << 4);
    int it = 0;
    for (int tm = pair; tm < p.tiles_m; tm += num_pairs, ++it) {
      const int m0 = tm * (2 * k3BM) + (int)rank * k3BM;
      for (int kb = 0; kb < nkb; ++kb) {
        const float4* xp = reinterpret_cast<const float4*>(xf_smem) + kb * 32 + j;
        const float4 a_lo = xp[0], a_hi = xp[8], c_lo = xp[16], c_hi = xp[24];
        const float xa[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
        const float xc[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
        mbar_wait(afull(kb), (uint32_t)(it & 1));
        SNB_TL(it < kDbgTiles && tt == 0, 400 + (it * 8 + kb) * 2);
        const uint32_t sa = a_base + kb * k3SlotBytes + toff;
        uint32_t w[kRI][4];
#pragma unroll
        for (int i = 0; i < kRI; ++i)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w[i][0]), "=r"(w[i][1]), "=r"(w[i][2]), "=r"(w[i][3])
                       : "r"(sa + (uint32_t)i * (uint32_t)(kRS * 128)));
#pragma unroll
        for (int i = 0; i < kRI; ++i) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z0 = __uint_as_float(w[i][e] << 16), z1 = __uint_as_float(w[i][e] & 0xFFFF0000u);
            w[i][e] = pack_bf16x2(__sinf(fmaf(xa[2 * e], z0, xc[2 * e])), __sinf(fmaf(xa[2 * e + 1], z1, xc[2 * e + 1])));
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + (uint32_t)i * (uint32_t)(kRS * 128)), "r"(w[i][0]), "r"(w[i][1]),
                       "r"(w[i][2]), "r"(w[i][3])
                       : "memory");
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the async proxy (tensor core)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(aready(kb), 0));
        SNB_TL(it < kDbgTiles && tt == 0, 400 + (it * 8 + kb) * 2 + 1);
        if (p.Y != nullptr) {
          // the activated slot is the next layer's input matrix Y (needed by this layer's weight gradient)
          const long long row = (long long)m0 + r0;
          __nv_bfloat16* dst = p.Y + row * p.ldy + kb * k3BK + 8 * j;
#pragma unroll
          for (int i = 0; i < kRI; ++i)
            if (row + kRS * i < p.M)
              asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (long long)(kRS * i) * p.ldy), "r"(w[i][0]), "r"(w[i][1]),
                           "r"(w[i][2]), "r"(w[i][3])
                           : "memory");
        }
      }
    }
  } else {
    reg_inc<152>();
    // ================= epilogue: TMEM -> (alpha, bias) -> bf16 -> swizzled staging -> TMA store; column statistics =========
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const uint32_t tempty_leader0 = mapa_shared(tempty(0), 0);
    const uint32_t tempty_leader1 = mapa_shared(tempty(1), 0);
    const int ew = warp - (4 + k3XfWarps);
    const int eh = ew >> 2;                 // which half of the pass's 256 columns
    const uint32_t cbuf = c_base + (uint32_t)ew * k3CWarpBytes;
    const int c_begin = eh * 128;
    float st_acc[2][2][4];       // [pass][64-column chunk][sum c0, sum c1, sumsq c0, sumsq c1]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int jx = 0; jx < 2; ++jx)
#pragma unroll
        for (int k = 0; k < 4; ++k) st_acc[i][jx][k] = 0.f;
    int item = 0;
    for (int tm = pair; tm < p.tiles_m; tm += num_pairs) {
      for (int tn = 0; tn < p.n_passes; ++tn, ++item) {
        const int acc = item & 1;
        const uint32_t acc_phase = (uint32_t)((item >> 1) & 1);
        const long long row0 = (long long)tm * (2 * k3BM) + (long long)rank * k3BM + q * 32;
        const int n0 = tn * 256;
        int rows_valid = 0;
        if (row0 < p.M) rows_valid = (int)min((long long)32, p.M - row0);
        mbar_wait(tfull(acc), acc_phase);
        tc_fence_after();
        SNB_TL(item < 2 * kDbgTiles && ew == 0 && lane == 0, 600 + item * 2);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = c_begin + 64 * cc;
          const float4* b4 = reinterpret_cast<const float4*>(bias_smem + n0 + c);      // broadcast reads
          uint32_t r0v[32], r1v[32];
          tmem_ld_32x32(t_addr + c, r0v);
          tmem_ld_32x32(t_addr + c + 32, r1v);
          if (lane == 0) bulk_wait_group_read<0>();      // the store that last read this staging buffer has drained it
          __syncwarp();
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float v[8];
            const float4 b_lo = b4[2 * u], b_hi = b4[2 * u + 1];
            const float ab[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t a = u < 4 ? r0v[8 * u + e] : r1v[8 * (u - 4) + e];
              v[e] = p.alpha * (__uint_as_float(a) + ab[e]);
            }
            const uint32_t addr = cbuf + (uint32_t)lane * 128u + (uint32_t)((u ^ (lane & 7)) << 4);
            uint32_t w4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w4[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows_valid > 0) {
            tma_store_2d(&tmapC, cbuf, n0 + c, (int)row0);
            bulk_commit_group();
          }
          // column statistics of the stored bf16 values: lane j owns columns (2j, 2j+1) of the chunk
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint32_t lbase = cbuf + (uint32_t)((lane & 3) << 2);
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
            if (tn == tnn) {
              st_acc[tnn][cc][0] += s0, st_acc[tnn][cc][1] += s1;
              st_acc[tnn][cc][2] += q0, st_acc[tnn][cc][3] += q1;
            }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
        SNB_TL(item < 2 * kDbgTiles && ew == 0 && lane == 0, 600 + item * 2 + 1);
      }
    }
    if (lane == 0) bulk_wait_group<0>();
    // flush the CTA's column statistics: combine the warps in shared memory, then one atomic per column
#pragma unroll
    for (int tnn = 0; tnn < 2; ++tnn)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col = tnn * 256 + c_begin + 64 * cc + 2 * lane;
        if (tnn < p.n_passes) {
          atomicAdd(stat_smem + col, st_acc[tnn][cc][0]);
          atomicAdd(stat_smem + col + 1, st_acc[tnn][cc][1]);
          atomicAdd(stat_smem + k3MaxStatN + col, st_acc[tnn][cc][2]);
          atomicAdd(stat_smem + k3MaxStatN + col + 1, st_acc[tnn][cc][3]);
        }
      }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * k3EpiWarps) : "memory");
    const int te = threadIdx.x - 32 * (4 + k3XfWarps);
    for (int i = te; i < p.N; i += 32 * k3EpiWarps) {
      const float s = stat_smem[i], ss = stat_smem[k3MaxStatN + i];
      if (s != 0.f || ss != 0.f) {
        atomicAdd(p.stats + i, s);
        atomicAdd(p.stats + p.N + i, ss);
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

// C = alpha * (sin(xa * Zprev + xc) . B^T + bias) with column statistics; Y (optional) receives the activated operand.
// SNB_ERR_UNSUPPORTED for shapes outside the resident-A design (the caller falls back to the stand-alone activation pass).
int snb_gemm_bf16_tc3(const void* Zprev, int lda, const float* xa, const float* xc, const void* B, int ldb, void* C, int ldc,
                      const float* bias, float alpha, long long M, int N, int K, float* stats, void* Y, int ldy,
                      cudaStream_t st) {
  if (M < 256 || (N != 256 && N != 512) || K < 64 || K > k3MaxKB * k3BK || (K % k3BK) != 0) return SNB_ERR_UNSUPPORTED;
  SNB_CHECK_ARG(Zprev && xa && xc && B && C && stats);
  SNB_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && (ldc % 8) == 0 && (((uintptr_t)Zprev) & 15) == 0 && (((uintptr_t)B) & 15) == 0 &&
                (((uintptr_t)C) & 15) == 0);
  if (Y) SNB_CHECK_ARG((ldy % 8) == 0 && (((uintptr_t)Y) & 15) == 0);
  if (bias) SNB_CHECK_ARG((((uintptr_t)bias) & 15) == 0);
  Gemm3Params p;
  p.M = M, p.N = N, p.K = K;
  p.tiles_m = (int)((M + 2 * k3BM - 1) / (2 * k3BM));
  p.n_passes = N / 256;
  p.nkb = K / k3BK;
  p.bias = bias, p.alpha = alpha, p.stats = stats, p.xa = xa, p.xc = xc;
  p.Y = reinterpret_cast<__nv_bfloat16*>(Y), p.ldy = ldy;
  const char* tl = getenv("SNB_TC3_TIMELINE");      // device pointer (decimal) of >= 1024 int64: debugging only
  p.dbg = tl ? reinterpret_cast<long long*>(strtoull(tl, nullptr, 10)) : nullptr;
  CUtensorMap ta, tb, tcm;
  int rc = make_tmap_bf16(&ta, Zprev, M, K, lda, k3BK, k3BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tb, B, N, K, ldb, k3BK, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tcm, C, M, N, ldc, 64, 32);
  if (rc) return rc;
  const int num_pairs = num_sms() / 2;
  const int grid = 2 * (p.tiles_m < num_pairs ? p.tiles_m : num_pairs);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_xf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k3Smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  gemm3_xf_kernel<<<grid, k3Threads, k3Smem, st>>>(ta, tb, tcm, p);
  count_launch();
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
