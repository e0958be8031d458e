// Fused epilogue of the split-bf16 GEMM engine, shared by the 1-CTA and the 2-CTA (cta_group::2) kernels:
// one call handles 32 consecutive accumulator columns of one tile row (thread = row).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace s3r {

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// v: 32 accumulator values (columns col0 .. col0+31 of group g); (nb, h, w): the row's pixel; pix: row index inside the
// group; grow: global output row for EPI_PLAIN; ht_acc: running 1x1-conv dot products of EPI_HEADTAIL.
template <int EPI>
__device__ __forceinline__ void epi_chunk(const GemmArgs& args, float (&v)[32], int g, int nb, int h, int w, bool valid,
                                          long long pix, long long grow, int col0, float (&ht_acc)[4]) {
  // bias
  if (args.bias != nullptr) {
    const int bcol = (EPI == EPI_PIXSHUF) ? (col0 % args.ps_cout) : col0;
    const int bstride = (EPI == EPI_PIXSHUF) ? args.ps_cout : args.N;
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

}  // namespace s3r
