// Persistent CHAIN of dependent split-bf16 GEMMs in ONE launch (tcgen05 cta_group::2 pairs, as gemm2.cu).
//
// Between two attention launches a transformer block is a pure GEMM chain -- proj -> fc1 -> fc2 -> (next block's) qkv in a
// ViT block (croco/models/blocks.py:127-130), proj -> q and cproj -> fc1 -> fc2 -> qkv' in a DecoderBlock (:186-191) -- whose
// only cross-CTA dependency is per 128-row block: a tile of phase p+1 needs every n-tile of the SAME rows of phase p (its A
// planes, the LayerNorm statistics folded into its epilogue, its residual rows).  As separate launches each boundary costs a
// grid-wide completion + dependent-launch release (~1 us), the first-operand latency (~2 us, weights cold) and the tail of the
// slowest CTA (profiles/r1_trace_gemm_after.txt).  Here the CTA pairs walk a static list of (phase, tile) items:
//   * a finished tile bumps the counter of its (phase, group, 128-row block) after its stores (every storing thread fences
//     the generic -> async proxy hand-over, bar.sync among the epilogue warps, then ONE red.release.gpu);
//   * the TMA producer of a dependent tile issues its WEIGHT (B) tiles first, then spins on the counter (ld.acquire.gpu,
//     bounded: a broken dependency traps instead of hanging the box) and issues the A tiles;
//   * the epilogue warps take the same acquire before they read statistics / residual rows (through L2: ld.global.cg);
//   * the last CTA to finish re-zeroes the counters, so a cached plan replays without host work.
// No deadlock: a tile only waits on tiles of EARLIER phases, every CTA walks the phases in order, and all CTAs are
// co-resident (grid <= SM count; the next kernel of the stream cannot take SMs before every CTA here has started, because
// programmatic dependents launch only after all CTAs executed griddepcontrol.launch_dependents).
// Tile widths (64 / 128 / 256) and epilogue modes (plain / QKV + RoPE head split) change from phase to phase at run time.
#include <cstring>

#include "common.cuh"
#include "gemm.cuh"
#include "gemm_epilogue.cuh"

namespace s3r {

namespace gc {
constexpr int BM = 128, BK = 64;
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kSmemRing = 192 * 1024;
constexpr int A_TILE = BM * BK * 2;      // one plane, this CTA's 128 rows
constexpr int kMaxStages = 4;
constexpr int SW = 16;
constexpr int STG = kEpiWarps * Stg<SW>::WARP_BYTES;
constexpr int COLV = 2 * 2 * 256 * 4;    // [2 accumulator stages][bias | colsum][<= 256 columns]
constexpr int SMEM = kSmemRing + 1024 + 256 + COLV + STG;
}  // namespace gc

// ---- cluster / 2-CTA PTX (same forms as gemm2.cu) ----
__device__ __forceinline__ uint32_t c_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void c_cluster_sync() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void c_tma2_4d(uint32_t smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void c_tma2_3d(uint32_t smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void c_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void c_tmem_alloc2(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void c_tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void c_umma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void c_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// ---- inter-CTA dependency counters ----
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Wait until the row block is complete.  Bounded: ~2^25 probes (seconds) then trap -- surfaces as a CUDA error.
__device__ __forceinline__ void dep_wait(const uint32_t* ctr, uint32_t need) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(ctr) < need) {
    if (++spins > (1u << 24)) { asm volatile("trap;"); }   // each probe is an L2 round trip (~0.7 us): seconds
  }
}

struct TileIdx {
  int g, mt, nt, ga;
};
__device__ __forceinline__ TileIdx chain_tile(const GemmArgs& a, int bn, int pt, int rank) {
  const int n_tiles = (a.N + bn - 1) / bn;
  const int m_pairs = (a.tiles_w * a.tiles_h * a.NB) >> 1;
  const int ppg = n_tiles * m_pairs;
  TileIdx t;
  t.g = pt / ppg;
  const int rem = pt - t.g * ppg;
  const int mp = rem / n_tiles;
  t.nt = rem - mp * n_tiles;
  t.mt = 2 * mp + rank;
  t.ga = (a.a_swap && t.nt * bn >= a.swap_col0) ? (a.groups - 1 - t.g) : t.g;
  return t;
}
__device__ __forceinline__ int chain_pairs(const GemmArgs& a, int bn) {
  return ((a.N + bn - 1) / bn) * ((a.tiles_w * a.tiles_h * a.NB) >> 1) * a.groups;
}

// One tile's epilogue (this CTA's 128 rows), the body of gemm2.cu's epilogue loop with a runtime tile width.
template <int EPI>
__device__ __forceinline__ void chain_epilogue_tile(const GemmArgs& args, int bn, const TileIdx& t, int as, uint32_t aphase,
                                                    float* colv, uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base,
                                                    int bn_max, int warp, int lane, bool leader) {
  using namespace gc;
  const int quad = warp & 3;
  const int half = (warp - 2) >> 2;
  const int ch = bn >> 6;   // 32-column chunks per warp
  float* sb = colv + as * 2 * 256;
  float* scs = sb + 256;
  float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(colv) + COLV + (warp - 2) * Stg<SW>::WARP_BYTES);
  epi_stage_cols_rt<EPI>(args, sb, scs, t.g, t.nt, bn, (int)threadIdx.x - 64, 32 * kEpiWarps);
  const int tw = t.mt % args.tiles_w;
  const int th = (t.mt / args.tiles_w) % args.tiles_h;
  const int nb = t.mt / (args.tiles_w * args.tiles_h);
  const TileGeom tg = make_geom(args, t.g, nb, th, tw, t.nt * bn);
  EpiRow er;
  EpiTRows tr;
  epi_tile_pre<EPI, SW, true>(args, tg, quad, lane, er, tr);
  float4 rcur[8], rnxt[8];
  const int cfirst = t.nt * bn + half * ch * 32;
  if (cfirst < args.N) epi_prefetch_res<EPI, SW, true>(args, tr, rcur, cfirst, lane);
  asm volatile("bar.sync 1, 256;" ::: "memory");   // staged columns visible to the 8 epilogue warps

  mbar_wait(&tmem_full[as], aphase);
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + as * bn_max;
  float ht_acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int cc = 0; cc < ch; ++cc) {
    const int c = half * ch + cc;
    const int col0 = t.nt * bn + c * 32;
    if (col0 >= args.N) break;
    uint32_t raw[32];
    tmem_ld_32x32(tbase + c * 32, raw);
    if (cc + 1 < ch && col0 + 32 < args.N) epi_prefetch_res<EPI, SW, true>(args, tr, rnxt, col0 + 32, lane);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    epi_chunk<EPI, SW, true>(args, v, sb + c * 32, scs + c * 32, stg, tg, er, tr, rcur, col0, lane, ht_acc);
#pragma unroll
    for (int q = 0; q < 8; ++q) rcur[q] = rnxt[q];
  }
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) {
    if (leader) mbar_arrive(&tmem_empty[as]);
    else c_arrive_remote(&tmem_empty[as], 0);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gc::kThreads, 1)
    gemm2_chain_kernel(const __grid_constant__ ChainParams P) {
  // The phase table lives in the kernel's parameter space (constant bank): the epilogue reads its fields like the standalone
  // kernels read their GemmArgs, and the tensor maps are addressed there.  (A first version kept the table in shared memory:
  // every field use became an LDS that the staging-tile stores forced the compiler to re-issue.)
  // ctr[0] = ticket of finished CTAs; ctr[ctr_base + ...] = row-block completion counters (ctr_base >= 1)
  const ChainPhase* const ph = P.ph;
  const int nph = P.nph, bn_max = P.bn_max, n_ctr = P.n_ctr;
  uint32_t* const ctr = P.ctr;
  using namespace gc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemRing);
  uint64_t* full_bar = bars;                       // leader only (count 2: both producers)
  uint64_t* empty_bar = bars + kMaxStages;         // per CTA, multicast commit from the leader
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // per CTA, multicast commit from the leader
  uint64_t* tmem_empty = tmem_full + 2;            // leader only (count 2 * kEpiWarps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* colv = reinterpret_cast<float*>(smem + kSmemRing + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = c_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int b_tile_max = (bn_max / 2) * BK * 2;
  const int stage_bytes = 2 * A_TILE + 2 * b_tile_max;      // one slot size for every phase
  const int stages = kSmemRing / stage_bytes;               // 3 (bn_max 256) or 4
  const uint32_t tmem_cols = 2u * (uint32_t)bn_max;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < nph; ++p) {
      tma_prefetch_desc(&ph[p].args.tmA_hi);
      tma_prefetch_desc(&ph[p].args.tmA_lo);
      tma_prefetch_desc(&ph[p].args.tmB_hi);
      tma_prefetch_desc(&ph[p].args.tmB_lo);
    }
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) c_tmem_alloc2(tmem_ptr_smem, tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  c_cluster_sync();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t full_u = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int p = 0; p < nph; ++p) {
      const GemmArgs& a = ph[p].args;
      const GemmArgs* ag = &ph[p].args;           // tensor maps: addressed in the parameter space
      const int bn = ph[p].bn;
      const int b_tile = (bn / 2) * BK * 2;
      const int num_kb = a.kpt;                          // linear layers only (taps == 1)
      const int total = chain_pairs(a, bn);
      const uint32_t tx = 2u * (uint32_t)(2 * A_TILE + 2 * b_tile);
      for (int pt = cluster_id; pt < total; pt += num_clusters) {
        const TileIdx t = chain_tile(a, bn, pt, (int)rank);
        const int row0 = t.mt * BM;                      // W coordinate of this CTA's rows (H = NB = 1)
        const int brow = t.g * a.b_group_rows + t.nt * bn + (int)rank * (bn / 2);
        const uint32_t* dep = (ph[p].dep >= 0) ? ctr + ph[ph[p].dep].ctr_base + t.ga * (a.tiles_w * a.tiles_h * a.NB) + t.mt
                                               : nullptr;
        // Per k-block: wait for the slot, issue the WEIGHT (B) tiles at once, and issue the A tiles as soon as the row block
        // this tile depends on is complete.  One probe up front: in a many-wave phase the dependency is long satisfied and
        // the loop is exactly the standalone kernel's (A and B together, slot by slot -- a first version issued a whole ring
        // pass of B before any A and so drained the ring at EVERY tile boundary: +2.5 us per tile).  While it is not
        // satisfied (single-wave phases) B runs ahead by up to one ring pass, then the producer blocks on the counter.
        bool ready = true;
        if (dep) {
          uint32_t seen = 0;
          if (lane == 0) seen = ld_acquire_gpu(dep);
          ready = __shfl_sync(0xffffffffu, seen, 0) >= (uint32_t)ph[p].dep_need;
        }
        int a_next = 0, a_stage = stage;                   // next k-block whose A tiles are still to be issued, and its slot
        const int run_ahead = num_kb < stages ? num_kb : stages;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t s = smem_u + stage * stage_bytes;
            const uint32_t fb = full_u + stage * 8;
            if (leader) mbar_arrive_expect_tx_u(fb, tx);
            else c_arrive_remote(&full_bar[stage], 0);
            c_tma2_3d(s + 2 * A_TILE, &ag->tmB_hi, fb, kb * BK, 0, brow);
            c_tma2_3d(s + 2 * A_TILE + b_tile, &ag->tmB_lo, fb, kb * BK, 0, brow);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (!ready) {
            if (kb + 1 - a_next >= run_ahead || kb + 1 == num_kb) {   // the ring is full of B tiles (or all are out): block
              if (lane == 0) dep_wait(dep, (uint32_t)ph[p].dep_need);
              __syncwarp();
              ready = true;
            } else {
              uint32_t seen = 0;
              if (lane == 0) seen = ld_acquire_gpu(dep);
              ready = __shfl_sync(0xffffffffu, seen, 0) >= (uint32_t)ph[p].dep_need;
            }
          }
          if (ready) {
            for (; a_next <= kb; ++a_next) {
              if (elect_one()) {
                const uint32_t s = smem_u + a_stage * stage_bytes;
                const uint32_t fb = full_u + a_stage * 8;
                c_tma2_4d(s, &ag->tmA_hi, fb, a_next * BK, row0, 0, t.ga);
                c_tma2_4d(s + A_TILE, &ag->tmA_lo, fb, a_next * BK, row0, 0, t.ga);
              }
              __syncwarp();
              if (++a_stage == stages) a_stage = 0;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (leader) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int p = 0; p < nph; ++p) {
        const GemmArgs& a = ph[p].args;
        const int bn = ph[p].bn;
        const int b_tile = (bn / 2) * BK * 2;
        const uint32_t idesc = umma_idesc(kFmtBF16, 2 * BM, (uint32_t)bn);
        const int num_kb = a.kpt;
        const int total = chain_pairs(a, bn);
        for (int pt = cluster_id; pt < total; pt += num_clusters, ++it) {
          const int as = it & 1;
          const uint32_t aphase = (it >> 1) & 1;
          mbar_wait(&tmem_empty[as], aphase ^ 1);
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_u + as * bn_max;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t sa = smem_u + stage * stage_bytes;
              const uint64_t da_hi = umma_desc_sw128_kmajor(sa);
              const uint64_t da_lo = umma_desc_sw128_kmajor(sa + A_TILE);
              const uint64_t db_hi = umma_desc_sw128_kmajor(sa + 2 * A_TILE);
              const uint64_t db_lo = umma_desc_sw128_kmajor(sa + 2 * A_TILE + b_tile);
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk) {
                const uint64_t ko = (uint64_t)(kk * 32 >> 4);
                c_umma2(tmem_d, da_hi + ko, db_lo + ko, idesc, (kb | kk) != 0);
                c_umma2(tmem_d, da_lo + ko, db_hi + ko, idesc, 1);
                c_umma2(tmem_d, da_hi + ko, db_hi + ko, idesc, 1);
              }
              c_commit_mc(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) c_commit_mc(&tmem_full[as]);
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9), own 128 rows
    int it = 0;
    for (int p = 0; p < nph; ++p) {
      const GemmArgs& a = ph[p].args;
      const int bn = ph[p].bn;
      const int total = chain_pairs(a, bn);
      const int m_tiles = a.tiles_w * a.tiles_h * a.NB;
      for (int pt = cluster_id; pt < total; pt += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const TileIdx t = chain_tile(a, bn, pt, (int)rank);
        if (ph[p].dep >= 0) {   // statistics / residual rows of an earlier phase: same acquire as the producer warp
          if (lane == 0) dep_wait(ctr + ph[ph[p].dep].ctr_base + t.ga * m_tiles + t.mt, (uint32_t)ph[p].dep_need);
          __syncwarp();
        }
        if (a.epi == EPI_QKV)
          chain_epilogue_tile<EPI_QKV>(a, bn, t, as, aphase, colv, tmem_full, tmem_empty, tmem_base, bn_max, warp, lane, leader);
        else
          chain_epilogue_tile<EPI_PLAIN>(a, bn, t, as, aphase, colv, tmem_full, tmem_empty, tmem_base, bn_max, warp, lane, leader);
        // this CTA's 128 rows x bn columns are stored: publish.  Writer side of the generic -> async proxy hand-over (the
        // consumer's TMA reads these rows), then a barrier among the 8 epilogue warps, then ONE release by one thread.
        asm volatile("fence.proxy.async.global;" ::: "memory");
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (threadIdx.x == 64) red_release_gpu_add(ctr + ph[p].ctr_base + t.g * m_tiles + t.mt, 1u);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  c_cluster_sync();   // the peer may still be reading our smem through its MMAs / receiving our arrives
  if (warp == 1) {
    tc_fence_after_sync();
    c_tmem_dealloc2(tmem_base, tmem_cols);
  }
  // self-cleaning: every CTA is past all of its waits here; the last one zeroes the counters for the next replay
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t old = atomicAdd(ctr, 1u);
    if (old == gridDim.x - 1) {
      for (int i = 1; i < n_ctr; ++i) ctr[i] = 0u;
      __threadfence();
      ctr[0] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int chain_plan_init(ChainPlan* cp, const GemmPlan* const* plans, int nph, uint32_t* dev_counters, int counters_cap) {
  if (nph < 1 || nph > kChainMaxPhases) {
    set_error("chain_plan_init: %d phases (1..%d supported)", nph, kChainMaxPhases);
    return -1;
  }
  memset(cp, 0, sizeof(*cp));
  ChainPhase* host = cp->params.ph;
  int ctr_used = 1;   // counters[0] = the "done" ticket
  int bn_max = 64;
  long long max_pairs = 0;
  double flops = 0;
  for (int p = 0; p < nph; ++p) {
    const GemmPlan& g = *plans[p];
    const GemmArgs& a = g.args;
    if (!g.two_cta || a.taps != 1 || a.H != 1 || a.NB != 1 || a.bw != 128 || (a.epi != EPI_PLAIN && a.epi != EPI_QKV) ||
        (g.bn != 64 && g.bn != 128 && g.bn != 256)) {
      set_error("chain_plan_init: phase %d is not a CTA-pair linear layer with a plain / QKV epilogue (two_cta=%d taps=%d H=%d "
                "NB=%d epi=%d bn=%d)", p, g.two_cta, a.taps, a.H, a.NB, a.epi, g.bn);
      return -1;
    }
    const int m_tiles = a.tiles_w * a.tiles_h * a.NB;
    if (p > 0) {
      const GemmArgs& prev = plans[p - 1]->args;
      if (prev.tiles_w * prev.tiles_h * prev.NB != m_tiles || prev.groups != a.groups) {
        set_error("chain_plan_init: phases %d and %d differ in row tiling / groups", p - 1, p);
        return -1;
      }
    }
    host[p].args = a;
    host[p].args.b_static = 1;
    host[p].args.trace = nullptr;
    host[p].bn = g.bn;
    host[p].dep = p - 1;
    host[p].dep_need = p > 0 ? (plans[p - 1]->args.N + plans[p - 1]->bn - 1) / plans[p - 1]->bn : 0;
    host[p].ctr_base = ctr_used;
    ctr_used += a.groups * m_tiles;
    if (g.bn > bn_max) bn_max = g.bn;
    const long long pairs = (long long)((a.N + g.bn - 1) / g.bn) * (m_tiles / 2) * a.groups;
    if (pairs > max_pairs) max_pairs = pairs;
    flops += g.flops;
  }
  if (ctr_used > counters_cap) {
    set_error("chain_plan_init: needs %d counters, %d available", ctr_used, counters_cap);
    return -1;
  }
  if (cudaMemset(dev_counters, 0, sizeof(uint32_t) * ctr_used) != cudaSuccess) {
    set_error("chain_plan_init: counter reset failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -7;
  }
  cp->params.ctr = dev_counters;
  cp->params.nph = nph;
  cp->params.n_ctr = ctr_used;
  cp->params.bn_max = bn_max;
  const long long clusters = max_pairs < num_sms() / 2 ? max_pairs : num_sms() / 2;
  cp->grid = dim3((unsigned)(2 * clusters), 1, 1);
  cp->flops = flops;
  return 0;
}

int chain_launch(const ChainPlan& cp, cudaStream_t stream) {
  static PerDeviceOnce once;
  bool& attr_set = once.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gc::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(chain smem=%d): %s", gc::SMEM, cudaGetErrorString(e));
      return -5;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = cp.grid;
  cfg.blockDim = dim3(gc::kThreads);
  cfg.dynamicSmemBytes = gc::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm2_chain_kernel, cp.params);
  if (e != cudaSuccess) {
    set_error("chain launch failed: %s", cudaGetErrorString(e));
    return -6;
  }
  return 0;
}

}  // namespace s3r
