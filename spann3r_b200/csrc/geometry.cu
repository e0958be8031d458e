// Post-path geometry on the GPU (first step): focal length of the first camera from its
// pointmap, the reference's `estimate_focal_knowing_depth(..., focal_mode='weiszfeld')`
// (dust3r/post_process.py:12-60, called by demo.py:148-150 on a CPU copy of preds[0]['pts3d']).
//
//   a = (x/z, y/z) with non-finite values -> 0,  p = (i - ppx, j - ppy)
//   f0 = sum(a.p) / sum(a.a);   10 x:  w = 1 / max(|p - f a|, 1e-8),  f = sum(w a.p) / sum(w a.a)
//
// (the reference's means cancel in the ratio).  Each iteration is two fixed-shape launches -- 148 x 256-thread partial
// sums per image in a fixed order, then one block per image -- so the result is deterministic; the pointmap
// (2.4 MB at 512 x 384) stays in L2 across the 11 passes and never crosses PCIe.
#include "kernels.cuh"

#include "common.cuh"

namespace s3r {

constexpr int kFocalBlocks = 148;

__global__ void __launch_bounds__(256) focal_partial_kernel(const float* __restrict__ pts, int H, int W, float ppx,
                                                            float ppy, const float* __restrict__ focal, int first,
                                                            float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float r0[256], r1[256];
  const int b = blockIdx.y;
  const long long n = (long long)H * W;
  const float* p = pts + (long long)b * n * 3;
  const float f = first ? 0.f : focal[b];
  float s_px = 0.f, s_xx = 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * kFocalBlocks) {
    const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    float ax = x / z, ay = y / z;
    if (!isfinite(ax)) ax = 0.f;     // nan_to_num(posinf=0, neginf=0): nan -> 0 as well
    if (!isfinite(ay)) ay = 0.f;
    const float u = (float)(i % W) - ppx, v = (float)(i / W) - ppy;
    const float d_px = ax * u + ay * v, d_xx = ax * ax + ay * ay;
    float w = 1.f;
    if (!first) {
      const float du = u - f * ax, dv = v - f * ay;
      w = 1.0f / fmaxf(sqrtf(du * du + dv * dv), 1e-8f);
    }
    s_px += w * d_px;
    s_xx += w * d_xx;
  }
  r0[threadIdx.x] = s_px;
  r1[threadIdx.x] = s_xx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      r0[threadIdx.x] += r0[threadIdx.x + o];
      r1[threadIdx.x] += r1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[((long long)b * kFocalBlocks + blockIdx.x) * 2] = r0[0];
    part[((long long)b * kFocalBlocks + blockIdx.x) * 2 + 1] = r1[0];
  }
}

__global__ void __launch_bounds__(256) focal_final_kernel(const float* __restrict__ part, float lo, float hi, int last,
                                                          float* __restrict__ focal) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float r0[256], r1[256];
  const int b = blockIdx.x;
  r0[threadIdx.x] = threadIdx.x < kFocalBlocks ? part[((long long)b * kFocalBlocks + threadIdx.x) * 2] : 0.f;
  r1[threadIdx.x] = threadIdx.x < kFocalBlocks ? part[((long long)b * kFocalBlocks + threadIdx.x) * 2 + 1] : 0.f;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      r0[threadIdx.x] += r0[threadIdx.x + o];
      r1[threadIdx.x] += r1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float f = r0[0] / r1[0];
    if (last) f = fminf(fmaxf(f, lo), hi);   // focal.clip(min_focal * base, max_focal * base)
    focal[b] = f;
  }
}

int launch_focal_weiszfeld(const float* pts3d, int B, int H, int W, float ppx, float ppy, int iters, float lo, float hi,
                           float* scratch, float* focal, cudaStream_t st) {
  if (B <= 0 || H <= 0 || W <= 0 || iters < 0) {
    set_error("focal_weiszfeld: bad arguments");
    return -1;
  }
  for (int it = 0; it <= iters; ++it) {
    launch_pdl(focal_partial_kernel, dim3(kFocalBlocks, B), dim3(256), 0, st, pts3d, H, W, ppx, ppy, (const float*)focal,
               it == 0 ? 1 : 0, scratch);
    launch_pdl(focal_final_kernel, dim3(B), dim3(256), 0, st, (const float*)scratch, lo, hi, it == iters ? 1 : 0, focal);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace s3r
