// Split-bf16 ("bf16x3") GEMM / implicit-GEMM 3x3 convolution on tcgen05 tensor cores, sm_100a.
//
//   D[pixel, n] = sum_{tap, k} A[pixel shifted by tap, k] * Wt[n, tap, k]      (fp32 result)
//
// Replaces, on the Spann3R forward path, every nn.Linear (croco/models/blocks.py:73-79,94-112,
// 149-169; dust3r/model.py:189-190; spann3r/model.py:250-261,310), the patch-embedding conv
// (dust3r/patch_embed.py:19-29, after an im2col kernel), and every Conv2d / ConvTranspose2d of the
// DPT head (croco/models/dpt_block.py:33-75,121-142,189-218,318-324,356-410).
//
// Structure: persistent, warp-specialised, one CTA per SM.
//   warp 0      TMA producer: per k-block four SWIZZLE_128B boxes (A_hi, A_lo: 128 pixels x 64 ch;
//               B_hi, B_lo: BN x 64) into a STAGES-deep smem ring; a 3x3 conv is 9 taps whose A box is
//               the same 4-D tensor map at (w+dx, h+dy) -- out-of-bounds pixels are zero-filled by TMA,
//               which is exactly the conv's zero padding.
//   warp 1      MMA issuer (one thread): per k-block 4 K-steps x 3 MMAs (hi*lo, lo*hi, hi*hi) into a
//               128 x BN fp32 accumulator in TMEM; two accumulator stages so the epilogue of tile i
//               overlaps the main loop of tile i+1.
//   warps 2..5  epilogue: tcgen05.ld 32 columns at a time (thread = tile row), fused bias / GELU / ReLU /
//               residual adds / split-bf16 re-encode for the next GEMM / RoPE + head split for attention /
//               ConvTranspose pixel shuffle / DPT head tail + postprocess.
#include "gemm.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace s3r {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kNumThreads = 192;
static constexpr int kSmemRing = 192 * 1024;

template <int BN>
struct GemmCfg {
  static constexpr int A_TILE = BM * BK * 2;  // bytes, one plane
  static constexpr int B_TILE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = kSmemRing / STAGE;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1) gemm_bf16x3_kernel(const __grid_constant__ GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles = (args.N + BN - 1) / BN;
  const int m_tiles = args.tiles_w * args.tiles_h * args.NB;
  const int tiles_per_group = n_tiles * m_tiles;
  const int total_tiles = tiles_per_group * args.groups;
  const int num_kb = args.taps * args.kpt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmA_hi);
    tma_prefetch_desc(&args.tmA_lo);
    tma_prefetch_desc(&args.tmB_hi);
    tma_prefetch_desc(&args.tmB_lo);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int g = tile / tiles_per_group;
        const int rem = tile - g * tiles_per_group;
        const int mt = rem / n_tiles;
        const int nt = rem - mt * n_tiles;
        const int tw = mt % args.tiles_w;
        const int th = (mt / args.tiles_w) % args.tiles_h;
        const int nb = mt / (args.tiles_w * args.tiles_h);
        const int w0 = tw * args.bw, h0 = th * args.bh;
        const int img = g * args.NB + nb;
        const int brow = g * args.b_group_rows + nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / args.kpt;
          const int kc = (kb - tap * args.kpt) * BK;
          int dx = 0, dy = 0;
          if (args.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + stage * Cfg::STAGE;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE);
          tma_load_4d(s, &args.tmA_hi, &full_bar[stage], kc, w0 + dx, h0 + dy, img);
          tma_load_4d(s + Cfg::A_TILE, &args.tmA_lo, &full_bar[stage], kc, w0 + dx, h0 + dy, img);
          tma_load_3d(s + 2 * Cfg::A_TILE, &args.tmB_hi, &full_bar[stage], kc, tap, brow);
          tma_load_3d(s + 2 * Cfg::A_TILE + Cfg::B_TILE, &args.tmB_lo, &full_bar[stage], kc, tap, brow);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kFmtBF16, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
          const uint64_t da_hi = umma_desc_sw128_kmajor(sa);
          const uint64_t da_lo = umma_desc_sw128_kmajor(sa + Cfg::A_TILE);
          const uint64_t db_hi = umma_desc_sw128_kmajor(sa + 2 * Cfg::A_TILE);
          const uint64_t db_lo = umma_desc_sw128_kmajor(sa + 2 * Cfg::A_TILE + Cfg::B_TILE);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ko = (uint64_t)(kk * 32 >> 4);  // 16 bf16 = 32 bytes along K inside the swizzle span
            umma_bf16(tmem_d, da_hi + ko, db_lo + ko, idesc, (kb | kk) != 0);
            umma_bf16(tmem_d, da_lo + ko, db_hi + ko, idesc, 1);
            umma_bf16(tmem_d, da_hi + ko, db_hi + ko, idesc, 1);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;
    const int dh = r / args.bw, dw = r - dh * args.bw;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int g = tile / tiles_per_group;
      const int rem = tile - g * tiles_per_group;
      const int mt = rem / n_tiles;
      const int nt = rem - mt * n_tiles;
      const int tw = mt % args.tiles_w;
      const int th = (mt / args.tiles_w) % args.tiles_h;
      const int nb = mt / (args.tiles_w * args.tiles_h);
      const int h = th * args.bh + dh, w = tw * args.bw + dw;
      const bool valid = (h < args.H) && (w < args.W);
      const long long pix = ((long long)nb * args.H + h) * args.W + w;  // row inside the group
      const long long grow = (long long)g * args.out_group_rows + pix;  // global output row (PLAIN)

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after_sync();
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;

      float ht_acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = nt * BN + c * 32;
        if (col0 >= args.N) break;  // warp-uniform
        uint32_t raw[32];
        tmem_ld_32x32(tbase + c * 32, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

        // bias
        if (args.bias != nullptr) {
          const int bcol = (args.epi == EPI_PIXSHUF) ? (col0 % args.ps_cout) : col0;
          const int bstride = (args.epi == EPI_PIXSHUF) ? args.ps_cout : args.N;
          const float4* bp = reinterpret_cast<const float4*>(args.bias + (long long)g * bstride + bcol);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = __ldg(bp + q);
            v[4 * q + 0] += b.x;
            v[4 * q + 1] += b.y;
            v[4 * q + 2] += b.z;
            v[4 * q + 3] += b.w;
          }
        }
        if (args.act != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], args.act);
        }

        if (args.epi == EPI_PLAIN || args.epi == EPI_PIXSHUF) {
          long long orow = grow;
          int ocol = col0;
          if (args.epi == EPI_PIXSHUF) {
            const int ij = col0 / args.ps_cout;
            ocol = col0 - ij * args.ps_cout;
            const int s = args.ps_s;
            const int i = ij / s, j = ij - i * s;
            orow = (long long)g * args.out_group_rows +
                   ((long long)nb * (args.H * s) + (h * s + i)) * (args.W * s) + (w * s + j);
          }
          if (valid) {
            if (args.res1 != nullptr) {
              const float4* rp = reinterpret_cast<const float4*>(args.res1 + orow * args.ldr1 + ocol);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = rp[q];
                v[4 * q + 0] += t.x;
                v[4 * q + 1] += t.y;
                v[4 * q + 2] += t.z;
                v[4 * q + 3] += t.w;
              }
            }
            if (args.res2 != nullptr) {
              const float4* rp = reinterpret_cast<const float4*>(args.res2 + orow * args.ldr2 + ocol);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = rp[q];
                v[4 * q + 0] += t.x;
                v[4 * q + 1] += t.y;
                v[4 * q + 2] += t.z;
                v[4 * q + 3] += t.w;
              }
            }
            if (args.out_f32 != nullptr) {
              float* op = args.out_f32 + orow * args.ldo + ocol;
#pragma unroll
              for (int q = 0; q < 8; ++q) st_f4(op + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (args.out_hi != nullptr) {
              uint32_t ph[16], pl[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                float a = v[2 * q], b = v[2 * q + 1];
                if (args.plane_relu) {
                  a = fmaxf(a, 0.f);
                  b = fmaxf(b, 0.f);
                }
                __nv_bfloat16 ah, al, bh, bl;
                split_bf16(a, ah, al);
                split_bf16(b, bh, bl);
                ph[q] = pack_bf16(ah, bh);
                pl[q] = pack_bf16(al, bl);
              }
              const long long po = orow * args.ldp + args.plane_col0 + ocol;
              uint4* hp = reinterpret_cast<uint4*>(args.out_hi + po);
              uint4* lp = reinterpret_cast<uint4*>(args.out_lo + po);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                hp[q] = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
                lp[q] = make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
              }
            }
          }
        } else if (args.epi == EPI_QKV) {
          // croco/models/blocks.py:97-104 (self) / :154-160 (cross) + RoPE2D (pos_embed.py:112-159,
          // curope/kernels.cu:18-81): head dim 64 = [y half | x half], each half = 16 (u, v) pairs
          // (j, j+16) rotated by pos * 100^(-j/16).
          const int role = args.q_role_base + col0 / args.q_C;  // 0 q, 1 k, 2 v
          const int cc = col0 % args.q_C;
          const int head = cc >> 6;
          const int d0 = cc & 63;  // 0 or 32
          const int heads = args.q_C >> 6;
          const int bidx = (int)(pix / args.q_ntok);
          const int t = (int)(pix - (long long)bidx * args.q_ntok);
          const long long gb = (long long)g * args.q_nb + bidx;
          if (valid) {
            if (role <= 1 && args.q_rope) {
              const int p = args.q_pos[(grow) * 2 + (d0 >> 5)];
              const float2* cs = args.q_cs + p * 16;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 t2 = __ldg(cs + j);
                const float u = v[j], x = v[j + 16];
                v[j] = u * t2.x - x * t2.y;
                v[j + 16] = x * t2.x + u * t2.y;
              }
            }
            if (role == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= args.q_scale;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = to_tf32(v[j]);
            if (role <= 1) {
              float* op = (role == 0 ? args.q_out : args.k_out) + ((gb * heads + head) * args.q_ntok + t) * 64 + d0;
#pragma unroll
              for (int q = 0; q < 8; ++q) st_f4(op + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              float* op = args.vt_out + ((gb * heads + head) * 64 + d0) * (long long)args.q_ntok_pad + t;
#pragma unroll
              for (int j = 0; j < 32; ++j) op[(long long)j * args.q_ntok_pad] = v[j];
            }
          }
        } else {  // EPI_HEADTAIL: dpt_block.py:318-324 (ReLU, 1x1 conv) + heads/postprocess.py:10-58
          const float* wt = args.ht_w + (long long)g * 4 * 128 + col0;
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const float4* wp = reinterpret_cast<const float4*>(wt + o * 128);
            float acc = ht_acc[o];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = __ldg(wp + q);
              acc = fmaf(v[4 * q + 0], t.x, acc);
              acc = fmaf(v[4 * q + 1], t.y, acc);
              acc = fmaf(v[4 * q + 2], t.z, acc);
              acc = fmaf(v[4 * q + 3], t.w, acc);
            }
            ht_acc[o] = acc;
          }
        }
      }
      // accumulator fully read -> hand the TMEM stage back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);

      if (args.epi == EPI_HEADTAIL && valid) {
        const float* b4 = args.ht_b + g * 4;
        const float x = ht_acc[0] + b4[0], y = ht_acc[1] + b4[1], z = ht_acc[2] + b4[2], cf = ht_acc[3] + b4[3];
        const float d = sqrtf(x * x + y * y + z * z);
        const float sc = expm1f(d) / fmaxf(d, 1e-8f);
        float* pp = args.ht_pts + grow * 3;
        pp[0] = x * sc;
        pp[1] = y * sc;
        pp[2] = z * sc;
        args.ht_conf[grow] = 1.0f + expf(cf);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// dims/box innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1). Always SWIZZLE_128B.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -3;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[4];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu %llu box %u %u %u %u base %p)", (int)r,
              rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
              (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
              rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0, base);
    return -4;
  }
  return 0;
}

static int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int gemm_plan_init(GemmPlan* plan, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo,
                   const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, int groups, int NB, int H, int W, int Kc,
                   int taps, int N, int force_bn, long long lda, long long ldb, long long b_group_rows) {
  memset(plan, 0, sizeof(*plan));
  GemmArgs& a = plan->args;
  if (lda == 0) lda = Kc;
  if (ldb == 0) ldb = (long long)Kc * taps;
  if (b_group_rows == 0) b_group_rows = N;
  if (lda % 8 != 0 || ldb % 8 != 0 || (taps != 1 && taps != 9) || (taps == 9 && Kc % 8 != 0)) {
    set_error("gemm_plan_init: unsupported shape Kc=%d N=%d taps=%d lda=%lld ldb=%lld (row strides must be multiples "
              "of 8 elements, taps in {1,9})", Kc, N, taps, lda, ldb);
    return -1;
  }
  a.b_group_rows = (int)b_group_rows;
  a.W = W; a.H = H; a.NB = NB; a.N = N; a.Kc = Kc; a.taps = taps;
  a.kpt = (Kc + BK - 1) / BK;
  a.bw = W >= 128 ? 128 : next_pow2(W);
  a.bh = 128 / a.bw;
  a.tiles_w = (W + a.bw - 1) / a.bw;
  a.tiles_h = (H + a.bh - 1) / a.bh;
  a.out_group_rows = (long long)NB * H * W;
  // tile width: keep >= ~1 wave of CTAs where the problem allows it
  const long long m_tiles = (long long)a.tiles_w * a.tiles_h * NB * groups;
  int bn = 128;
  if (N >= 256 && m_tiles * ((N + 255) / 256) >= 2 * num_sms()) bn = 256;
  else if (m_tiles * ((N + 127) / 128) < num_sms() && N >= 64) bn = 64;
  if (N <= 64) bn = 64;
  if (force_bn) bn = force_bn;
  plan->bn = bn;

  const uint64_t esz = 2;
  {
    uint64_t dims[4] = {(uint64_t)Kc, (uint64_t)W, (uint64_t)H, (uint64_t)NB * groups};
    uint64_t str[3] = {(uint64_t)lda * esz, (uint64_t)lda * W * esz, (uint64_t)lda * W * H * esz};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)a.bw, (uint32_t)a.bh, 1};
    int r;
    if ((r = encode_tmap(&a.tmA_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a_hi, dims, str, box))) return r;
    if ((r = encode_tmap(&a.tmA_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a_lo, dims, str, box))) return r;
  }
  {
    uint64_t dims[3] = {(uint64_t)Kc, (uint64_t)taps, (uint64_t)(b_group_rows * (groups - 1) + N)};
    // with a single tap the tap stride is never used, but must still be a multiple of 16 bytes
    uint64_t str[2] = {(uint64_t)(taps == 1 ? ldb : Kc) * esz, (uint64_t)ldb * esz};
    uint32_t box[3] = {(uint32_t)BK, 1, (uint32_t)bn};
    int r;
    if ((r = encode_tmap(&a.tmB_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, b_hi, dims, str, box))) return r;
    if ((r = encode_tmap(&a.tmB_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, b_lo, dims, str, box))) return r;
  }
  const long long n_tiles = (N + bn - 1) / bn;
  const long long total = m_tiles * n_tiles;
  a.groups = groups;
  plan->grid = dim3((unsigned)((total < num_sms()) ? total : num_sms()), 1, 1);  // persistent: <= 1 CTA per SM
  plan->flops = 2.0 * (double)NB * H * W * groups * (double)N * (double)Kc * taps;
  return 0;
}

template <int BN>
static int launch_bn(const GemmPlan& plan, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return -5;
    }
    attr_set = true;
  }
  gemm_bf16x3_kernel<BN><<<plan.grid, kNumThreads, Cfg::SMEM, stream>>>(plan.args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm launch failed: %s", cudaGetErrorString(e));
    return -6;
  }
  return 0;
}

int gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.bn) {
    case 64: return launch_bn<64>(plan, stream);
    case 128: return launch_bn<128>(plan, stream);
    case 256: return launch_bn<256>(plan, stream);
  }
  set_error("gemm_launch: bad bn %d", plan.bn);
  return -1;
}

}  // namespace s3r
