// 2-CTA variant of the split-bf16 GEMM engine: a cluster of two CTAs (one TPC) computes a 256 x BN tile
// with tcgen05.mma.cta_group::2.  Each CTA stages only ITS 128 rows of A and ITS half (BN/2 rows) of B,
// so the operand bytes every SM pulls through L2 -> SMEM per MMA drop by 1.33x (BN=128) / 1.5x (BN=256)
// versus the 1-CTA kernel -- and the freed shared memory deepens the TMA ring (4 resp. 3 stages).
// The GEMM engine is L2->SMEM / latency bound on this path, so that is the lever.
//
// Roles per CTA (320 threads): warp 0 TMA producer (both CTAs; completions are counted on the LEADER's
// full barrier), warp 1 MMA issuer (leader CTA only; commits are multicast to both CTAs' barriers),
// warps 2..9 epilogue (each CTA drains its own 128 TMEM lanes; two warps per lane quadrant split the columns).
#include <cstring>

#include "common.cuh"
#include "gemm.cuh"
#include "gemm_epilogue.cuh"

namespace s3r {

namespace g2 {
constexpr int BM = 128, BK = 64;
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kSmemRing = 192 * 1024;
}  // namespace g2

template <int BN>
struct Gemm2Cfg {
  static constexpr int A_TILE = g2::BM * g2::BK * 2;      // one plane, this CTA's 128 rows
  static constexpr int B_TILE = (BN / 2) * g2::BK * 2;    // one plane, this CTA's half of the BN rows
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = g2::kSmemRing / STAGE;
  static constexpr int COLV = 2 * 2 * BN * 4;   // per accumulator stage: staged bias + LN-fold column sums of the tile
  static constexpr int SW = 16;                 // epilogue staging width: 8 warps x 32 x 32 floats would not fit
  static constexpr int STG = g2::kEpiWarps * Stg<SW>::WARP_BYTES;
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256 + COLV + STG;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

// ---- cluster / 2-CTA PTX ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma2_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// raw shared-address variants for the warp-uniform producer (bar: this CTA's address of the barrier)
__device__ __forceinline__ void tma2_load_4d_u(uint32_t smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                               int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d_u(uint32_t smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the same-offset mbarrier of CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far -> arrive(1) on the same-offset mbarrier in BOTH CTAs when they retire
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::kThreads, 1)
    gemm2_bf16x3_kernel(const __grid_constant__ GemmArgs args) {
  using Cfg = Gemm2Cfg<BN>;
  using namespace g2;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round trip would lose the address
  // space and turn every access through `smem` into a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;                       // used in the leader only (count 2: both producers)
  uint64_t* empty_bar = bars + Cfg::STAGES;        // per CTA, multicast commit from the leader
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;    // per CTA, multicast commit from the leader
  uint64_t* tmem_empty = tmem_full + 2;            // used in the leader only (count 2 * kEpiWarps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* colv = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE + 256);   // [2 stages][bias | colsum][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // optional timeline of CTA 0 (tools/trace_gemm2.py): [0] entry, [1] prologue done, [2] dependency wait done, [3] exit, then per
  // tile it < 4 at [4 + 6 it]: first operands landed, all MMAs issued, accumulator seen by the epilogue, first chunk done,
  // epilogue done, last k-block's loads issued
  unsigned long long* const trace = (blockIdx.x == 0) ? args.trace : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = globaltimer_ns();

  const int n_tiles = (args.N + BN - 1) / BN;
  const int m_tiles = args.tiles_w * args.tiles_h * args.NB;   // even (host guarantees)
  const int m_pairs = m_tiles / 2;
  const int pairs_per_group = n_tiles * m_pairs;
  const int total_pairs = pairs_per_group * args.groups;
  const int num_kb = args.taps * args.kpt;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmA_hi);
    tma_prefetch_desc(&args.tmA_lo);
    tma_prefetch_desc(&args.tmB_hi);
    tma_prefetch_desc(&args.tmB_lo);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  if (trace && threadIdx.x == 0) trace[1] = globaltimer_ns();
  // B tiles of the first ring pass before the dependency wait (see gemm.cu): weights do not depend on the previous kernel
  int pre_b = 0;
  if (warp == 0 && args.b_static && cluster_id < total_pairs) {
    const int pt = cluster_id;
    const int g = pt / pairs_per_group;
    const int nt = (pt - g * pairs_per_group) % n_tiles;
    const int brow = g * args.b_group_rows + nt * BN + (int)rank * (BN / 2);
    pre_b = num_kb < Cfg::STAGES ? num_kb : Cfg::STAGES;
    if (lane == 0) {
      for (int kb = 0; kb < pre_b; ++kb) {
        const int tap = kb / args.kpt, kc = (kb - tap * args.kpt) * BK;
        const uint32_t s = smem_u32(smem) + kb * Cfg::STAGE;
        const uint32_t fb = smem_u32(&full_bar[kb]);
        if (leader) mbar_arrive_expect_tx_u(fb, 2 * Cfg::STAGE);
        else mbar_arrive_remote(&full_bar[kb], 0);
        tma2_load_3d_u(s + 2 * Cfg::A_TILE, &args.tmB_hi, fb, kc, tap, brow);
        tma2_load_3d_u(s + 2 * Cfg::A_TILE + Cfg::B_TILE, &args.tmB_lo, fb, kc, tap, brow);
      }
    }
    __syncwarp();
  }
  pdl_wait();
  if (trace && threadIdx.x == 0) trace[2] = globaltimer_ns();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // whole warp in uniform control flow, one elected lane issues; incremental tap / channel coordinates (see gemm.cu)
    {
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t full_u = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
      int stage = 0;
      uint32_t phase = 0;
      int pit = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++pit) {
        const int g = pt / pairs_per_group;
        const int rem = pt - g * pairs_per_group;
        const int mp = rem / n_tiles;
        const int nt = rem - mp * n_tiles;
        const int mt = 2 * mp + (int)rank;
        const int tw = mt % args.tiles_w;
        const int th = (mt / args.tiles_w) % args.tiles_h;
        const int nb = mt / (args.tiles_w * args.tiles_h);
        const int w0 = tw * args.bw, h0 = th * args.bh;
        const int img = ((args.a_swap && nt * BN >= args.swap_col0) ? (args.groups - 1 - g) : g) * args.NB + nb;
        const int brow = g * args.b_group_rows + nt * BN + (int)rank * (BN / 2);
        int tap = 0, kc = 0, dx = (args.taps == 9) ? -1 : 0, dy = dx;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const bool b_done = (pt == cluster_id) && (kb < pre_b);   // B (and the arrive) went out before the wait
          if (elect_one()) {
            const uint32_t s = smem_u + stage * Cfg::STAGE;
            const uint32_t fb = full_u + stage * 8;
            if (!b_done) {
              if (leader) mbar_arrive_expect_tx_u(fb, 2 * Cfg::STAGE);
              else mbar_arrive_remote(&full_bar[stage], 0);
            }
            tma2_load_4d_u(s, &args.tmA_hi, fb, kc, w0 + dx, h0 + dy, img);
            tma2_load_4d_u(s + Cfg::A_TILE, &args.tmA_lo, fb, kc, w0 + dx, h0 + dy, img);
            if (!b_done) {
              tma2_load_3d_u(s + 2 * Cfg::A_TILE, &args.tmB_hi, fb, kc, tap, brow);
              tma2_load_3d_u(s + 2 * Cfg::A_TILE + Cfg::B_TILE, &args.tmB_lo, fb, kc, tap, brow);
            }
          }
          __syncwarp();
          kc += BK;
          if (kc >= args.kpt * BK) {
            kc = 0;
            ++tap;
            if (++dx > 1) {
              dx = -1;
              ++dy;
            }
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (trace && pit < 4 && lane == 0) trace[4 + 6 * pit + 5] = globaltimer_ns();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // whole warp in uniform control flow, one elected lane issues, bases made provably uniform (see gemm.cu)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc(kFmtBF16, 2 * BM, BN);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_u + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (trace && kb == 0 && it < 4 && lane == 0) trace[4 + 6 * it] = globaltimer_ns();
          if (elect_one()) {
            const uint32_t sa = smem_u + stage * Cfg::STAGE;
            const uint64_t da_hi = umma_desc_sw128_kmajor(sa);
            const uint64_t da_lo = umma_desc_sw128_kmajor(sa + Cfg::A_TILE);
            const uint64_t db_hi = umma_desc_sw128_kmajor(sa + 2 * Cfg::A_TILE);
            const uint64_t db_lo = umma_desc_sw128_kmajor(sa + 2 * Cfg::A_TILE + Cfg::B_TILE);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t ko = (uint64_t)(kk * 32 >> 4);
              umma2_bf16(tmem_d, da_hi + ko, db_lo + ko, idesc, (kb | kk) != 0);
              umma2_bf16(tmem_d, da_lo + ko, db_hi + ko, idesc, 1);
              umma2_bf16(tmem_d, da_hi + ko, db_hi + ko, idesc, 1);
            }
            umma2_commit_mc(&empty_bar[stage]);   // frees this smem slot in BOTH CTAs
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_mc(&tmem_full[as]);        // accumulator ready in BOTH CTAs
        __syncwarp();
        if (trace && it < 4 && lane == 0) trace[4 + 6 * it + 1] = globaltimer_ns();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9), own 128 rows
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;            // column half handled by this warp
    int it = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int g = pt / pairs_per_group;
      const int rem = pt - g * pairs_per_group;
      const int mp = rem / n_tiles;
      const int nt = rem - mp * n_tiles;
      const int mt = 2 * mp + (int)rank;
      const int tw = mt % args.tiles_w;
      const int th = (mt / args.tiles_w) % args.tiles_h;
      const int nb = mt / (args.tiles_w * args.tiles_h);

      // requested while the main loop of this tile still runs (see gemm.cu): staged bias / colsum columns, the rows'
      // LayerNorm statistics, RoPE positions and output addresses, the first chunk's residual values
      constexpr int CH = BN / 64;  // 32-column chunks per warp
      constexpr int SW = Cfg::SW;
      float* sb = colv + as * 2 * BN;
      float* scs = sb + BN;
      float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(colv) + Cfg::COLV + (warp - 2) * Stg<SW>::WARP_BYTES);
      epi_stage_cols<EPI, BN>(args, sb, scs, g, nt, (int)threadIdx.x - 64, 32 * kEpiWarps);
      const TileGeom tg = make_geom(args, g, nb, th, tw, nt * BN);
      EpiRow er;
      EpiTRows tr;
      epi_tile_pre<EPI, SW>(args, tg, quad, lane, er, tr);
      float4 rcur[8], rnxt[8];
      const int cfirst = nt * BN + half * CH * 32;
      if (cfirst < args.N) epi_prefetch_res<EPI, SW>(args, tr, rcur, cfirst, lane);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // staged columns visible to the 8 epilogue warps

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after_sync();
      if (trace && it < 4 && threadIdx.x == 64) trace[4 + 6 * it + 2] = globaltimer_ns();
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
      float ht_acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int cc = 0; cc < CH; ++cc) {
        const int c = half * CH + cc;
        const int col0 = nt * BN + c * 32;
        if (col0 >= args.N) break;
        uint32_t raw[32];
        tmem_ld_32x32(tbase + c * 32, raw);
        if (cc + 1 < CH && col0 + 32 < args.N) epi_prefetch_res<EPI, SW>(args, tr, rnxt, col0 + 32, lane);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        epi_chunk<EPI, SW>(args, v, sb + c * 32, scs + c * 32, stg, tg, er, tr, rcur, col0, lane, ht_acc);
        if (trace && it < 4 && cc == 0 && threadIdx.x == 64) trace[4 + 6 * it + 3] = globaltimer_ns();
#pragma unroll
        for (int q = 0; q < 8; ++q) rcur[q] = rnxt[q];
      }
      if (trace && it < 4 && threadIdx.x == 64) trace[4 + 6 * it + 4] = globaltimer_ns();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[as]);
        else mbar_arrive_remote(&tmem_empty[as], 0);
      }
      if constexpr (EPI == EPI_HEADTAIL) {
        // the row's 4 dot products are split over the two column halves: half 1 hands its partial sums to half 0
        // through its staging tile (two alternating slots), 64-thread named barrier per lane quadrant (as in gemm.cu)
        uint8_t* stg_base = reinterpret_cast<uint8_t*>(colv) + Cfg::COLV;
        float4* slot = reinterpret_cast<float4*>(stg_base + (4 + quad) * Stg<SW>::WARP_BYTES) + (it & 1) * 32;
        if (half == 1) slot[lane] = make_float4(ht_acc[0], ht_acc[1], ht_acc[2], ht_acc[3]);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
        if (half == 0) {
          const float4 p = slot[lane];
          ht_acc[0] += p.x; ht_acc[1] += p.y; ht_acc[2] += p.z; ht_acc[3] += p.w;
          epi_headtail_finish(args, tg, er, ht_acc);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading our smem through its MMAs / receiving our arrives
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
  }
  if (trace && threadIdx.x == 0) trace[3] = globaltimer_ns();
}

// ------------------------------------------------------------------------------------------------
template <int BN, int EPI>
static int launch2_bn(const GemmPlan& plan, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  static PerDeviceOnce once;
  bool& attr_set = once.cur();
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm2_bf16x3_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm2 smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return -5;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = plan.grid;
  cfg.blockDim = dim3(g2::kThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool use_pdl = (getenv("S3R_NO_PDL") == nullptr);
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_kernel<BN, EPI>, plan.args);
  if (e != cudaSuccess) {
    set_error("gemm2 launch failed: %s", cudaGetErrorString(e));
    return -6;
  }
  return 0;
}

template <int BN>
static int launch2_epi(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.args.epi) {
    case EPI_PLAIN: return launch2_bn<BN, EPI_PLAIN>(plan, stream);
    case EPI_PIXSHUF: return launch2_bn<BN, EPI_PIXSHUF>(plan, stream);
    case EPI_QKV: return launch2_bn<BN, EPI_QKV>(plan, stream);
    case EPI_HEADTAIL:
      if constexpr (BN == 128) return launch2_bn<128, EPI_HEADTAIL>(plan, stream);
      break;
  }
  set_error("gemm2_launch: epilogue mode %d is not available on the 2-CTA kernel at bn %d", plan.args.epi, BN);
  return -1;
}

int gemm2_launch(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.bn) {
    case 64: return launch2_epi<64>(plan, stream);   // experiment (S3R_GEMM2_64=1 / force_bn 2064): 256 x 64 pair tiles
    case 128: return launch2_epi<128>(plan, stream);
    case 256: return launch2_epi<256>(plan, stream);
  }
  set_error("gemm2_launch: bad bn %d", plan.bn);
  return -1;
}

}  // namespace s3r
