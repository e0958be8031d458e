// Fused epilogue of the split-bf16 GEMM engine, shared by the 1-CTA and the 2-CTA (cta_group::2) kernels:
// one call handles 32 consecutive accumulator columns of one tile row (thread = row).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace s3r {

__device__ __forceinline__ void apply_act(float (&v)[32], int act) {
  if (act == ACT_GELU) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = gelu_erf4(make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      v[4 * q] = t.x;
      v[4 * q + 1] = t.y;
      v[4 * q + 2] = t.z;
      v[4 * q + 3] = t.w;
    }
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
}

// Per-tile column vectors staged in shared memory by the epilogue warps BEFORE they wait for the accumulator (the
// loads overlap the main loop): sb[j] = bias of tile column j (0 when there is none), scs[j] = column sum of the
// LayerNorm-folded weights (ln_cs, see GemmArgs).  Call with every epilogue thread, then barrier among them.
template <int EPI, int BN>
__device__ __forceinline__ void epi_stage_cols(const GemmArgs& args, float* sb, float* scs, int g, int nt, int tid_e,
                                               int nthr_e) {
  for (int j = tid_e; j < BN; j += nthr_e) {
    const int col = nt * BN + j;
    float b = 0.f, c = 0.f;
    if (col < args.N) {
      if (args.bias != nullptr)
        b = __ldg(args.bias + ((EPI == EPI_PIXSHUF) ? (long long)g * args.ps_cout + (col % args.ps_cout)
                                                    : (long long)g * args.N + col));
      if (args.ln_cs != nullptr) c = __ldg(args.ln_cs + (long long)g * args.N + col);
    }
    sb[j] = b;
    scs[j] = c;
  }
}

// Per-row state of one tile, loaded before the accumulator wait: LayerNorm statistics of the A row (folded
// LayerNorm: the GEMM ran on the raw x planes with gamma folded into the weights, the epilogue applies
// rstd * (acc - mean * colsum) -- identical to LN(x) W^T up to fp32 rounding) and the RoPE position of the row.
struct EpiRow {
  float rstd = 1.f, rm = 0.f;   // rm = rstd * mean
  int py = 0, px = 0;
};
template <int EPI>
__device__ __forceinline__ void epi_row_init(const GemmArgs& args, EpiRow& er, int g, long long pix, long long grow,
                                             bool valid) {
  if (args.ln_stats != nullptr && valid) {
    const int ga = args.a_swap ? (args.groups - 1 - g) : g;
    // ln_np <= 32 chunk pairs (C <= 1024, even count): all loads are issued before the first add (a rolled loop
    // would serialise ln_np dependent L2 round trips); fixed summation order
    const float4* sp = reinterpret_cast<const float4*>(args.ln_stats + ((long long)ga * args.out_group_rows + pix) * args.ln_np);
    const int np2 = args.ln_np >> 1;
    float4 t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = (i < np2) ? sp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      s1 += t[i].x + t[i].z;
      s2 += t[i].y + t[i].w;
    }
    const float inv_c = 1.0f / (float)(args.ln_np * 32);
    const float mean = s1 * inv_c;
    const float var = fmaxf(s2 * inv_c - mean * mean, 0.f);
    er.rstd = rsqrtf(var + args.ln_eps);
    er.rm = er.rstd * mean;
  }
  if constexpr (EPI == EPI_QKV) {
    if (args.q_rope && valid) {
      er.py = args.q_pos[grow * 2];
      er.px = args.q_pos[grow * 2 + 1];
    }
  }
}

// Residual rows of the NEXT 32-column chunk, requested while the current chunk is processed (EPI_PLAIN only).
template <int EPI>
__device__ __forceinline__ void epi_prefetch_res(const GemmArgs& args, float4 (&rp)[8], long long grow, bool valid,
                                                 int col0) {
  if constexpr (EPI == EPI_PLAIN) {
    if (args.res1 != nullptr && valid) {
      const float4* p = reinterpret_cast<const float4*>(args.res1 + grow * args.ldr1 + col0);
#pragma unroll
      for (int q = 0; q < 8; ++q) rp[q] = p[q];
    }
  }
}

// v: 32 accumulator values (columns col0 .. col0+31 of group g); (nb, h, w): the row's pixel; pix: row index inside the
// group; grow: global output row for EPI_PLAIN; sb / scs: this chunk's 32 staged bias / colsum values (shared memory);
// rp: the prefetched res1 values of this chunk (EPI_PLAIN); ht_acc: running 1x1-conv dot products of EPI_HEADTAIL.
template <int EPI>
__device__ __forceinline__ void epi_chunk(const GemmArgs& args, float (&v)[32], const float* sb, const float* scs,
                                          const EpiRow& er, const float4 (&rp)[8], int g, int nb, int h, int w,
                                          bool valid, long long pix, long long grow, int col0, float (&ht_acc)[4]) {
  if (args.ln_stats != nullptr) {
    const float4* c4 = reinterpret_cast<const float4*>(scs);   // 128-byte aligned chunk of the staged columns
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 c = c4[q];
      v[4 * q + 0] = fmaf(er.rstd, v[4 * q + 0], -er.rm * c.x);
      v[4 * q + 1] = fmaf(er.rstd, v[4 * q + 1], -er.rm * c.y);
      v[4 * q + 2] = fmaf(er.rstd, v[4 * q + 2], -er.rm * c.z);
      v[4 * q + 3] = fmaf(er.rstd, v[4 * q + 3], -er.rm * c.w);
    }
  }
  {
    const float4* b4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = b4[q];
      v[4 * q + 0] += b.x;
      v[4 * q + 1] += b.y;
      v[4 * q + 2] += b.z;
      v[4 * q + 3] += b.w;
    }
  }
  if (args.act != ACT_NONE) apply_act(v, args.act);

  if constexpr (EPI == EPI_PLAIN || EPI == EPI_PIXSHUF) {
    long long orow = grow;
    int ocol = col0;
    if constexpr (EPI == EPI_PIXSHUF) {
      const int ij = col0 / args.ps_cout;
      ocol = col0 - ij * args.ps_cout;
      const int s = args.ps_s;
      const int i = ij / s, j = ij - i * s;
      orow = (long long)g * args.out_group_rows +
             ((long long)nb * (args.H * s) + (h * s + i)) * (args.W * s) + (w * s + j);
    }
    if (valid) {
      if (args.res1 != nullptr) {
        if constexpr (EPI == EPI_PLAIN) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[4 * q + 0] += rp[q].x;
            v[4 * q + 1] += rp[q].y;
            v[4 * q + 2] += rp[q].z;
            v[4 * q + 3] += rp[q].w;
          }
        } else {
          const float4* r1 = reinterpret_cast<const float4*>(args.res1 + orow * args.ldr1 + ocol);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = r1[q];
            v[4 * q + 0] += t.x;
            v[4 * q + 1] += t.y;
            v[4 * q + 2] += t.z;
            v[4 * q + 3] += t.w;
          }
        }
      }
      if (args.res2 != nullptr) {
        const float4* r2 = reinterpret_cast<const float4*>(args.res2 + orow * args.ldr2 + ocol);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = r2[q];
          v[4 * q + 0] += t.x;
          v[4 * q + 1] += t.y;
          v[4 * q + 2] += t.z;
          v[4 * q + 3] += t.w;
        }
      }
      if constexpr (EPI == EPI_PLAIN) {
        if (args.stats_out != nullptr) {   // (sum, sum of squares) of this row over the chunk: LayerNorm statistics
          float s1 = 0.f, s2 = 0.f;         // of the residual stream for the NEXT GEMM's folded LayerNorm
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s1 += v[j];
            s2 = fmaf(v[j], v[j], s2);
          }
          args.stats_out[orow * (long long)(args.N >> 5) + (col0 >> 5)] = make_float2(s1, s2);
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
          split2_bf16(a, b, ph[q], pl[q]);
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
  } else if constexpr (EPI == EPI_QKV) {
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
        const int p = (d0 >> 5) ? er.px : er.py;
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

}  // namespace s3r
