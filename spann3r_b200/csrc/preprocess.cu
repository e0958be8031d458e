// Input adapter on the GPU: the reference's per-frame preprocessing between the decoded
// RGB image and the network input -- centre crop, PIL Lanczos down-scale, centred crop, ToTensor + Normalize(0.5, 0.5)
// (spann3r/datasets/demo.py:57-86 -> dust3r/datasets/base/base_stereo_view_dataset.py:143-194 ->
// dust3r/datasets/utils/cropping.py:55-124 -> dust3r/utils/image.py:23).  The down-scale is Pillow's 8-bit separable
// resampler (src/libImaging/Resample.c: ImagingResampleHorizontal_8bpc / Vertical_8bpc): integer arithmetic, 22-bit
// fixed-point coefficients, uint8 intermediate -- reproduced BIT-EXACTLY here; the coefficient tables come from the host
// (spann3r_b200/preprocess.py, Pillow's precompute_coeffs).  Both crops are folded into the passes: only the rows /
// columns that survive the final crop are ever computed.
//
// HBM-bound byte work: one read of the crop of the source image, one small uint8 intermediate, one fp32 write.
#include "kernels.cuh"

#include "common.cuh"

namespace s3r {

constexpr int kPrec = 32 - 8 - 2;   // PRECISION_BITS of Resample.c

__device__ __forceinline__ int clip8(int v) {   // clip8(): (in >> PRECISION_BITS) clamped to [0, 255]
  v >>= kPrec;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass.  Block = one source row x 128 output columns; the source span those columns need is staged in
// shared memory with coalesced byte loads, then thread t computes output column x0 + t (3 channels).
// src: RGB rows of `row_stride` bytes, first needed row / column already applied by the caller through the pointer and
// the bounds; bounds[x] = (first source column, tap count), kk[x][ksize] fixed-point taps.
__global__ void __launch_bounds__(128) resample_h_u8_kernel(const uint8_t* __restrict__ src, long long row_stride,
                                                            int out_cols, const int* __restrict__ bounds,
                                                            const int* __restrict__ kk, int ksize,
                                                            uint8_t* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t span[];
  const int row = blockIdx.y;
  const int x0 = blockIdx.x * 128;
  const int xl = min(x0 + 127, out_cols - 1);
  const int s0 = bounds[2 * x0];                                  // first source column of the block's span
  const int s1 = bounds[2 * xl] + bounds[2 * xl + 1];             // one past the last
  const uint8_t* srow = src + (long long)row * row_stride + 3LL * s0;
  const int nbytes = 3 * (s1 - s0);
  for (int i = threadIdx.x; i < nbytes; i += 128) span[i] = srow[i];
  __syncthreads();
  const int x = x0 + threadIdx.x;
  if (x >= out_cols) return;
  const int b0 = bounds[2 * x] - s0, n = bounds[2 * x + 1];
  const int* k = kk + (long long)x * ksize;
  int a0 = 1 << (kPrec - 1), a1 = a0, a2 = a0;
  for (int i = 0; i < n; ++i) {
    const int c = __ldg(k + i);
    const uint8_t* p = span + 3 * (b0 + i);
    a0 += p[0] * c;
    a1 += p[1] * c;
    a2 += p[2] * c;
  }
  uint8_t* o = dst + ((long long)row * out_cols + x) * 3;
  o[0] = (uint8_t)clip8(a0);
  o[1] = (uint8_t)clip8(a1);
  o[2] = (uint8_t)clip8(a2);
}

// Vertical pass + ToTensor + Normalize: thread = one byte column of the intermediate (x * 3 + c, coalesced across the
// warp), loops over the output rows of its block.  dst [3, out_rows, cols] fp32 = ((v / 255) - 0.5) / 0.5 in fp32, the
// operation order of torchvision's ToTensor / Normalize.
__global__ void __launch_bounds__(256) resample_v_u8_norm_kernel(const uint8_t* __restrict__ tmp, int cols,
                                                                 int out_rows, const int* __restrict__ bounds,
                                                                 const int* __restrict__ kk, int ksize,
                                                                 float* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int j = blockIdx.x * 256 + threadIdx.x;   // byte column
  if (j >= cols * 3) return;
  const int x = j / 3, c = j - 3 * x;
  const int y = blockIdx.y;
  const int y0 = bounds[2 * y], n = bounds[2 * y + 1];
  const int* k = kk + (long long)y * ksize;
  const uint8_t* p = tmp + (long long)y0 * cols * 3 + j;
  int a = 1 << (kPrec - 1);
  for (int i = 0; i < n; ++i) a += (int)p[(long long)i * cols * 3] * __ldg(k + i);
  const float v = (float)clip8(a) / 255.0f;
  dst[((long long)c * out_rows + y) * cols + x] = (v - 0.5f) / 0.5f;
}

int launch_resample_h_u8(const uint8_t* src, long long row_stride, int rows, int out_cols, const int* bounds,
                         const int* kk, int ksize, int max_span, uint8_t* dst, cudaStream_t st) {
  if (rows <= 0 || out_cols <= 0) return 0;
  const size_t smem = (size_t)3 * max_span;
  if (smem > 160 * 1024) {
    set_error("resample_h: source span of %d pixels per 128 output columns is too large", max_span);
    return -1;
  }
  static PerDeviceOnce once;
  if (smem > 48 * 1024 && !once.cur()) {
    cudaFuncSetAttribute(resample_h_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    once.cur() = true;
  }
  launch_pdl(resample_h_u8_kernel, dim3((out_cols + 127) / 128, rows), dim3(128), smem, st, src, row_stride, out_cols, bounds,
             kk, ksize, dst);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

int launch_resample_v_u8_norm(const uint8_t* tmp, int cols, int out_rows, const int* bounds, const int* kk, int ksize,
                              float* dst, cudaStream_t st) {
  if (out_rows <= 0 || cols <= 0) return 0;
  launch_pdl(resample_v_u8_norm_kernel, dim3((cols * 3 + 255) / 256, out_rows), dim3(256), 0, st, tmp, cols, out_rows,
             bounds, kk, ksize, dst);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace s3r
