// TEST HARNESS (not a product path): the median focal estimate run sequentially on the CPU with the product's device math
// header (spann3r_b200/csrc/focal_math.cuh) and the same 4 x 8-bit radix select csrc/geometry.cu performs on the GPU.
#include <vector>

#include "../../spann3r_b200/csrc/focal_math.cuh"

using namespace s3r::focal;

extern "C" float focal_median_host(const float* pts, int H, int W, float ppx, float ppy) {
  const long long hw = (long long)H * W, n2 = 2 * hw;
  uint32_t prefix = 0;
  long long k = 0, n = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    long long hist[256] = {0};
    for (long long j = 0; j < n2; ++j) {
      const float f = vote(pts, j, hw, W, ppx, ppy);
      if (f != f) continue;
      const uint32_t key = order_key(f);
      if (pass > 0 && (key >> (shift + 8)) != prefix) continue;
      ++hist[(key >> shift) & 255];
    }
    if (pass == 0) {
      for (int b = 0; b < 256; ++b) n += hist[b];
      if (n == 0) return nanf("");
      k = (n - 1) / 2;   // torch.nanmedian: the lower of the two middle values
    }
    int b = 0;
    while (k >= hist[b]) k -= hist[b++];
    prefix = (prefix << 8) | (uint32_t)b;
  }
  return key_value(prefix);
}
