// Fused multi-head attention core for head dim 64 on tcgen05 (kind::tf32), sm_100a.
//
//   O[b, q, h*64:(h+1)*64] = softmax_k( Q[b,h,q,:] . K[b,h,k,:] ) V[b,h,k,:]
//
// Replaces the materialised `q @ k.T -> softmax -> @ v` of croco/models/blocks.py:106-110 (self
// attention of Block / DecoderBlock) and :162-166 (CrossAttention).  Q and K arrive already rotated
// (2-D RoPE), Q pre-scaled by 64^-0.5, V transposed, all rounded to tf32 by the QKV-projection GEMM
// epilogue (gemm.cu, EPI_QKV), so RoPE never exists as a separate op or tensor.
//
// One CTA per (128-query tile, batch*head); 192 threads:
//   warp 0     TMA: Q tile once, then per 128-key block K (2 boxes) and V^T (4 boxes) into a 2-stage ring
//   warp 1     MMA issuer: S_j = Q K_j^T  (M128 N128 K64, 8 x tcgen05.mma kind::tf32) into TMEM (double
//              buffered), then O_j = P_j V_j (M128 N64 K128, 16 MMAs) into a second TMEM region
//   warps 2-5  softmax, thread = query row: tcgen05.ld S (two passes: max, exp), running max / sum in
//              registers, P_j written to smem as the tf32 A operand (128B-swizzled by hand), and the
//              per-block O_j folded into a register accumulator  o = o*alpha + O_j  (no TMEM rescale pass).
// Output: split-bf16 planes (A operand of the following projection GEMM) and/or fp32.
#include "common.cuh"
#include "kernels.cuh"

#include <cstring>

namespace s3r {

namespace attn {
constexpr int BQ = 128;   // queries per CTA
constexpr int BKV = 128;  // keys per block
constexpr int D = 64;
constexpr int Q_BYTES = BQ * D * 4;       // 32 KB (2 swizzle atoms of [128 x 32 f32])
constexpr int K_BYTES = BKV * D * 4;      // 32 KB
constexpr int V_BYTES = D * BKV * 4;      // 32 KB (4 atoms of [64 x 32 f32])
constexpr int P_BYTES = BQ * BKV * 4;     // 64 KB (4 atoms of [128 x 32 f32])
constexpr int KV_STAGES = 2;
constexpr int SMEM = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + P_BYTES + 1024 + 128;
constexpr uint32_t TMEM_COLS = 512;       // S0 [0,128) S1 [128,256) O [256,320)
constexpr int kThreads = 192;
}  // namespace attn


__global__ void __launch_bounds__(attn::kThreads, 1) attention_kernel(const __grid_constant__ AttnArgs args) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Q_BYTES;                            // stage s: K at s*(K+V), V after K
  uint8_t* sP = sKV + KV_STAGES * (K_BYTES + V_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // 2
  uint64_t* kv_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;        // 2
  uint64_t* p_full = bars + 7;        // 1 (128 arrivals)
  uint64_t* o_full = bars + 8;        // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int bh = blockIdx.y;
  const int nblk = (args.nk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmQ);
    tma_prefetch_desc(&args.tmK);
    tma_prefetch_desc(&args.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_3d(sQ, &args.tmQ, q_full, 0, q0, bh);
      tma_load_3d(sQ + Q_BYTES / 2, &args.tmQ, q_full, 32, q0, bh);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        uint8_t* k = sKV + st * (K_BYTES + V_BYTES);
        uint8_t* v = k + K_BYTES;
        mbar_arrive_expect_tx(&kv_full[st], K_BYTES + V_BYTES);
        tma_load_3d(k, &args.tmK, &kv_full[st], 0, j * BKV, bh);
        tma_load_3d(k + K_BYTES / 2, &args.tmK, &kv_full[st], 32, j * BKV, bh);
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_load_3d(v + a * (V_BYTES / 4), &args.tmV, &kv_full[st], j * BKV + a * 32, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(kFmtTF32, BQ, BKV);
      constexpr uint32_t idesc_o = umma_idesc(kFmtTF32, BQ, D);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      const uint32_t tmem_o = tmem_base + 256;
      mbar_wait(q_full, 0);
      for (int j = 0; j <= nblk; ++j) {
        if (j < nblk) {
          const int st = j % KV_STAGES;
          mbar_wait(&kv_full[st], (j / KV_STAGES) & 1);
          tc_fence_after_sync();
          const uint32_t aK = smem_u32(sKV + st * (K_BYTES + V_BYTES));
          const uint32_t tmem_s = tmem_base + (j & 1) * BKV;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const uint64_t dq = umma_desc_sw128_kmajor(aQ + a * (Q_BYTES / 2));
            const uint64_t dk = umma_desc_sw128_kmajor(aK + a * (K_BYTES / 2));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)  // 8 tf32 = 32 bytes per K step
              umma_tf32(tmem_s, dq + 2 * kk, dk + 2 * kk, idesc_s, (a | kk) != 0);
          }
          umma_commit(&s_full[j & 1]);
        }
        if (j >= 1) {
          const int jj = j - 1;
          const int st = jj % KV_STAGES;
          mbar_wait(p_full, jj & 1);
          tc_fence_after_sync();
          const uint32_t aV = smem_u32(sKV + st * (K_BYTES + V_BYTES) + K_BYTES);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const uint64_t dp = umma_desc_sw128_kmajor(aP + a * (P_BYTES / 4));
            const uint64_t dv = umma_desc_sw128_kmajor(aV + a * (V_BYTES / 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_tf32(tmem_o, dp + 2 * kk, dv + 2 * kk, idesc_o, (a | kk) != 0);
          }
          umma_commit(o_full);
          umma_commit(&kv_empty[st]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps, thread = query row
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t ts = tmem_base + lane_sel + (j & 1) * BKV;
      const int kbase = j * BKV;
      // pass 1: block max
      float bmax = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(ts + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (kbase + c * 32 + i < args.nk) ? __uint_as_float(raw[i]) : -INFINITY;
          bmax = fmaxf(bmax, s);
        }
      }
      const float m_new = fmaxf(m, bmax);
      const float alpha = __expf(m - m_new);  // first block: exp(-inf) = 0
      // retire the previous block's P.V into the register accumulator (also proves sP is free again)
      if (j >= 1) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + lane_sel + 256 + c * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(raw[i]));
        }
      }
      // pass 2: p = exp(s - m_new), row sum, P tile (tf32) into swizzled smem
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(ts + c * 32, raw);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(raw[i]);
          float e = (kbase + c * 32 + i < args.nk) ? __expf(s - m_new) : 0.f;
          e = to_tf32(e);
          psum += e;
          p[i] = e;
        }
        uint8_t* pa = prow + c * (P_BYTES / 4);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(pa + ((q ^ sw) << 4)) = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
      }
      l = l * alpha + psum;
      m = m_new;
      alpha_prev = alpha;
      tc_fence_before_sync();   // our tcgen05.ld of S_j are done before the MMA warp may overwrite the buffer
      fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(p_full);
    }
    // last block
    mbar_wait(o_full, (nblk - 1) & 1);
    tc_fence_after_sync();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + lane_sel + 256 + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(raw[i]));
    }
    const int q = q0 + r;
    if (q < args.nq) {
      const float inv = 1.0f / l;
      const int b = bh / args.heads, h = bh - b * args.heads;
      const long long off = ((long long)b * args.nq + q) * args.ldo + h * D;
      if (args.o_f32) {
#pragma unroll
        for (int i = 0; i < D; i += 4)
          st_f4(args.o_f32 + off + i, o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
      if (args.o_hi) {
        uint32_t ph[32], pl[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          __nv_bfloat16 ah, al, bh2, bl;
          split_bf16(o[2 * i] * inv, ah, al);
          split_bf16(o[2 * i + 1] * inv, bh2, bl);
          ph[i] = pack_bf16(ah, bh2);
          pl[i] = pack_bf16(al, bl);
        }
        uint4* hp = reinterpret_cast<uint4*>(args.o_hi + off);
        uint4* lp = reinterpret_cast<uint4*>(args.o_lo + off);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          hp[i] = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
          lp[i] = make_uint4(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3]);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<attn::TMEM_COLS>(tmem_base);
  }
}

int attn_plan_init(AttnPlan* plan, const float* q, const float* k, const float* vt, int BH, int heads, int nq, int nk,
                   int nk_pad) {
  using namespace attn;
  memset(plan, 0, sizeof(*plan));
  if (nk_pad % 4 != 0 || nk_pad < nk) {
    set_error("attention: nk_pad=%d must be >= nk=%d and a multiple of 4", nk_pad, nk);
    return -1;
  }
  AttnArgs& a = plan->args;
  {
    uint64_t dims[3] = {64, (uint64_t)nq, (uint64_t)BH};
    uint64_t str[2] = {64 * 4, (uint64_t)nq * 64 * 4};
    uint32_t box[3] = {32, (uint32_t)BQ, 1};
    int r = encode_tmap(&a.tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, q, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)nk, (uint64_t)BH};
    uint64_t str[2] = {64 * 4, (uint64_t)nk * 64 * 4};
    uint32_t box[3] = {32, (uint32_t)BKV, 1};
    int r = encode_tmap(&a.tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, k, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[3] = {(uint64_t)nk, 64, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)nk_pad * 4, (uint64_t)nk_pad * 64 * 4};
    uint32_t box[3] = {32, 64, 1};
    int r = encode_tmap(&a.tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vt, dims, str, box);
    if (r) return r;
  }
  a.nq = nq; a.nk = nk; a.heads = heads;
  plan->grid = dim3((nq + BQ - 1) / BQ, BH);
  plan->flops = 4.0 * BH * (double)nq * nk * 64;
  return 0;
}

int attn_launch(const AttnPlan& plan, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, float* o_f32, long long ldo,
                cudaStream_t st) {
  using namespace attn;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute(smem=%d): %s", SMEM, cudaGetErrorString(e));
      return -5;
    }
    attr_set = true;
  }
  AttnArgs a = plan.args;
  a.o_hi = o_hi; a.o_lo = o_lo; a.o_f32 = o_f32; a.ldo = ldo;
  cudaError_t e = launch_pdl(attention_kernel, plan.grid, dim3(kThreads), SMEM, st, a);
  if (e != cudaSuccess) {
    set_error("attention launch failed: %s", cudaGetErrorString(e));
    return -6;
  }
  return 0;
}

int launch_attention(const float* q, const float* k, const float* vt, int BH, int heads, int nq, int nk, int nk_pad,
                     __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, float* o_f32, long long ldo, cudaStream_t st) {
  if (nq <= 0 || nk <= 0 || BH <= 0) return 0;
  AttnPlan plan;
  int r = attn_plan_init(&plan, q, k, vt, BH, heads, nq, nk, nk_pad);
  if (r) return r;
  return attn_launch(plan, o_hi, o_lo, o_f32, ldo, st);
}

}  // namespace s3r
