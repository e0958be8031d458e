// TEST HARNESS (not a product path): compiles the product's device math header, spann3r_b200/csrc/pnp_math.cuh,
// with g++ and runs the same pipeline csrc/pnp.cu runs on the GPU -- hypotheses -> inlier counts -> best -> mask ->
// damped Gauss-Newton -- sequentially on the CPU, so tests/test_pnp.py can pin the arithmetic of every
// __host__ __device__ function against cv2.solvePnPRansac golden poses without a GPU.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../spann3r_b200/csrc/pnp_math.cuh"

using namespace s3r::pnp;

extern "C" int pnp_host_check(const float* pts, const float* img, long long n, int width, double fx, double fy, double cx,
                              double cy, double reproj_err, int n_samples, int iters, unsigned long long seed, double* out18,
                              unsigned char* mask) {
  const Cam k{fx, fy, cx, cy};
  const double thr2 = reproj_err * reproj_err;
  std::vector<Pose> hyps;
  for (int m = 0; m < n_samples; ++m) {
    Pose h[4];
    const int c = sample_hypotheses(seed, m, pts, img, n, width, k, h);
    for (int j = 0; j < c; ++j) hyps.push_back(h[j]);
  }
  long long best = -1, best_count = -1;
  for (size_t h = 0; h < hyps.size(); ++h) {
    long long count = 0;
    for (long long i = 0; i < n; ++i) {
      double X[3], u, v;
      if (!load_point(pts, img, i, width, X, u, v)) continue;
      if (reproj_err2(hyps[h], k, X, u, v) < thr2) ++count;
    }
    if (count > best_count) {
      best_count = count;
      best = (long long)h;
    }
  }
  std::memset(out18, 0, 18 * sizeof(double));
  if (best < 0 || best_count < 4) return 0;
  Pose T = hyps[best];
  std::vector<unsigned char> m(n, 0);
  for (long long i = 0; i < n; ++i) {
    double X[3], u, v;
    if (load_point(pts, img, i, width, X, u, v) && reproj_err2(T, k, X, u, v) < thr2) m[i] = 1;
  }
  if (mask) std::memcpy(mask, m.data(), n);
  Pose good = T;
  double good_acc[kAcc], good_cost = 0, lambda = 1e-4;
  bool have_good = false;
  for (int it = 0; it <= iters; ++it) {
    double acc[kAcc];
    for (int j = 0; j < kAcc; ++j) acc[j] = 0;
    for (long long i = 0; i < n; ++i) {
      if (!m[i]) continue;
      double X[3], u, v;
      load_point(pts, img, i, width, X, u, v);
      gn_accumulate(T, k, X, u, v, acc);
    }
    if (!have_good || acc[27] <= good_cost) {
      good = T;
      std::memcpy(good_acc, acc, sizeof(acc));
      good_cost = acc[27];
      have_good = true;
      lambda = lambda * 0.1 > 1e-9 ? lambda * 0.1 : 1e-9;
    } else {
      lambda = lambda * 10 < 1e6 ? lambda * 10 : 1e6;
    }
    T = good;
    if (it < iters) gn_step(good_acc, lambda, T);
  }
  for (int i = 0; i < 9; ++i) out18[i] = good.R[i];
  for (int i = 0; i < 3; ++i) out18[9 + i] = good.t[i];
  so3_log(good.R, out18 + 12);
  out18[15] = (double)best_count;
  out18[16] = good_acc[28] > 0 ? sqrt(good_cost / good_acc[28]) : 0;
  out18[17] = 1.0;
  return 1;
}

// p3p alone, for the unit test: returns the number of solutions, poses as 12 doubles each
extern "C" int p3p_host(const double* P9, const double* f9, double* out48) {
  double P[3][3], f[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      P[i][j] = P9[3 * i + j];
      f[i][j] = f9[3 * i + j];
    }
  Pose o[4];
  const int n = p3p(P, f, o);
  for (int k = 0; k < n; ++k) {
    for (int i = 0; i < 9; ++i) out48[12 * k + i] = o[k].R[i];
    for (int i = 0; i < 3; ++i) out48[12 * k + 9 + i] = o[k].t[i];
  }
  return n;
}
