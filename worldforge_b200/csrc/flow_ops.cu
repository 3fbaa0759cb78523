// worldforge_b200 - dense optical flow (Farneback) and the flow-similarity metrics of the FLF channel selector.
//
// The reference scores every latent channel by comparing OpenCV Farneback flows of two uint8 clips
// (scheduling_unipc_multistep_clean.py:156-248 -> cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0),
// :497-607 for the metrics) on the CPU: 640 calls + 32 device-to-host copies per guided step, on the critical path.  These
// kernels restate the algorithm of OpenCV's modules/video/src/optflowgf.cpp for the two cases the selector hits - frames of
// 60 x 104 (480p) whose pyramid has ONE level (the driver loop stops when the smaller side would drop under 32 pixels) and
// frames of 90 x 160 (720p; 88 x 160 for LongCat's refine pass) with TWO: the half-resolution level is
// GaussianBlur(3 x 3, sigma 0.5) + an exact 2 x 2 mean (cv::resize INTER_LINEAR by 1/2), its flow returns through a bilinear
// doubling times 2 - as oracle/farneback.py does in numpy (pinned to OpenCV to 4e-6 px / 1e-4 px, tests/test_farneback_oracle.py):
//
//   poly_exp      GaussianBlur(3x3, sigma 0) = [1/4 1/2 1/4] reflect-101, then FarnebackPolyExp(n = 5, sigma = 1.2): per pixel
//                 the 5 coefficients (y, x, yy, xx, xy) of the local quadratic; vertical pass in float, horizontal in double
//   matrices      FarnebackUpdateMatrices: bilinear sample of the second frame's coefficients at x + flow, the 2x2 system
//                 (g11, g12, g22, h1, h2) per pixel, border down-weighting
//   box + solve   FarnebackUpdateFlow_Blur: 15 x 15 box sums with replicated borders in double (separable), then the solve
//
// Three iterations, matrices refreshed after the first two.  One thread per pixel; every clip pair of a call runs in the
// same launches (640 pairs x 6240 pixels), intermediates stay in L2-sized scratch.  Arithmetic follows the oracle's order;
// nvcc may contract a*b+c into FMAs where OpenCV's build does not, which moves flows by ~1e-6 px - far inside the
// selector's decision margins (the GPU test compares flows to 1e-4 px and the selections exactly).
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

constexpr int FB_N = 5;                 // poly_n
constexpr int FB_BORDER = 5;

struct FbConst {
  float g[2 * FB_N + 1], xg[2 * FB_N + 1], xxg[2 * FB_N + 1];
  double ig11, ig03, ig33, ig55;
};

// ---------------------------------------------------------------- 0. the half-resolution pyramid level (two-level pyramids)
// frames: uint8 [nf][H][W] (H, W even) -> half: float [nf][H/2][W/2] = 2x2 mean of GaussianBlur(3x3, k) with reflect-101 borders
__global__ void fb_blur_half_kernel(const unsigned char* __restrict__ frames, float* __restrict__ half, int nf, int H, int W, float k0,
                                    float k1, float k2) {
  const int H2 = H / 2, W2 = W / 2;
  const size_t total = static_cast<size_t>(nf) * H2 * W2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x2 = static_cast<int>(i % W2);
    const int y2 = static_cast<int>((i / W2) % H2);
    const unsigned char* img = frames + (i / (static_cast<size_t>(H2) * W2)) * H * W;
    auto refl = [](int p, int n) { return p < 0 ? -p : (p >= n ? 2 * n - 2 - p : p); };      // BORDER_REFLECT_101
    auto blurred = [&](int y, int x) {       // rows, then columns, float
      float col[3];
#pragma unroll
      for (int d = -1; d <= 1; ++d) {
        const unsigned char* row = img + static_cast<size_t>(refl(y + d, H)) * W;
        float r = __fadd_rn(__fmul_rn(static_cast<float>(row[refl(x - 1, W)]), k0), __fmul_rn(static_cast<float>(row[x]), k1));
        col[d + 1] = __fadd_rn(r, __fmul_rn(static_cast<float>(row[refl(x + 1, W)]), k2));
      }
      return __fadd_rn(__fadd_rn(__fmul_rn(col[0], k0), __fmul_rn(col[1], k1)), __fmul_rn(col[2], k2));
    };
    const float a = blurred(2 * y2, 2 * x2), b = blurred(2 * y2, 2 * x2 + 1), c = blurred(2 * y2 + 1, 2 * x2), d = blurred(2 * y2 + 1, 2 * x2 + 1);
    half[i] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 0.25f);
  }
}

// vertical pass of the polynomial expansion on a float image (the half-resolution level is not smoothed again)
__global__ void fb_vertical_f32_kernel(const float* __restrict__ imgs, float* __restrict__ vert, int nf, int H, int W, FbConst c) {
  const size_t total = static_cast<size_t>(nf) * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const float* img = imgs + (i / (static_cast<size_t>(H) * W)) * H * W;
    auto at = [&](int yy) { return img[static_cast<size_t>(yy) * W + x]; };
    float r0 = __fmul_rn(at(y), c.g[FB_N]), r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 1; k <= FB_N; ++k) {
      const float up = at(max(y - k, 0)), dn = at(min(y + k, H - 1));
      const float p = __fadd_rn(up, dn);
      r0 = __fadd_rn(r0, __fmul_rn(c.g[FB_N + k], p));
      r1 = __fadd_rn(r1, __fmul_rn(c.xg[FB_N + k], __fsub_rn(dn, up)));
      r2 = __fadd_rn(r2, __fmul_rn(c.xxg[FB_N + k], p));
    }
    float* o = vert + i * 3;
    o[0] = r0; o[1] = r1; o[2] = r2;
  }
}

// flow of the half-resolution level -> initial flow of the full level: cv::resize(INTER_LINEAR) by 2 (source coordinate
// (dst + 0.5) / 2 - 0.5, taps clamped) times 1 / pyr_scale = 2.  half [pairs][H/2][W/2][2] -> flow [pairs][H][W][2]
__global__ void fb_double_flow_kernel(const float* __restrict__ half, float* __restrict__ flow, int pairs, int H, int W) {
  const int H2 = H / 2, W2 = W / 2;
  const size_t total = static_cast<size_t>(pairs) * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const float* src = half + (i / (static_cast<size_t>(H) * W)) * H2 * W2 * 2;
    // (d + 0.5) / 2 - 0.5 = d/2 - 0.25: floor and fraction are exact in integers (fraction 0.75 for even d, 0.25 for odd d)
    const int ix = (x & 1) ? (x - 1) / 2 : x / 2 - 1, iy = (y & 1) ? (y - 1) / 2 : y / 2 - 1;
    const float wx = (x & 1) ? 0.25f : 0.75f, wy = (y & 1) ? 0.25f : 0.75f;
    const int x0 = min(max(ix, 0), W2 - 1), x1 = min(max(ix + 1, 0), W2 - 1);
    const int y0 = min(max(iy, 0), H2 - 1), y1 = min(max(iy + 1, 0), H2 - 1);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float p00 = src[(static_cast<size_t>(y0) * W2 + x0) * 2 + ch], p01 = src[(static_cast<size_t>(y0) * W2 + x1) * 2 + ch];
      const float p10 = src[(static_cast<size_t>(y1) * W2 + x0) * 2 + ch], p11 = src[(static_cast<size_t>(y1) * W2 + x1) * 2 + ch];
      const float top = __fadd_rn(__fmul_rn(p00, __fsub_rn(1.f, wx)), __fmul_rn(p01, wx));
      const float bot = __fadd_rn(__fmul_rn(p10, __fsub_rn(1.f, wx)), __fmul_rn(p11, wx));
      flow[i * 2 + ch] = __fmul_rn(__fadd_rn(__fmul_rn(top, __fsub_rn(1.f, wy)), __fmul_rn(bot, wy)), 2.0f);
    }
  }
}

// ---------------------------------------------------------------- 1. blur + vertical pass of the polynomial expansion
// frames: uint8 [nf][H][W] -> vert: float [nf][H][W][3]
__global__ void fb_vertical_kernel(const unsigned char* __restrict__ frames, float* __restrict__ vert, int nf, int H, int W, FbConst c) {
  const size_t total = static_cast<size_t>(nf) * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const unsigned char* img = frames + (i / (static_cast<size_t>(H) * W)) * H * W;
    auto refl = [](int p, int n) { return p < 0 ? -p : (p >= n ? 2 * n - 2 - p : p); };      // BORDER_REFLECT_101
    auto blurred = [&](int yy) {       // GaussianBlur 3x3, rows then columns, float
      float col[3];
#pragma unroll
      for (int d = -1; d <= 1; ++d) {
        const unsigned char* row = img + static_cast<size_t>(refl(yy + d, H)) * W;
        float r = __fadd_rn(__fmul_rn(static_cast<float>(row[refl(x - 1, W)]), 0.25f), __fmul_rn(static_cast<float>(row[x]), 0.5f));
        col[d + 1] = __fadd_rn(r, __fmul_rn(static_cast<float>(row[refl(x + 1, W)]), 0.25f));
      }
      return __fadd_rn(__fadd_rn(__fmul_rn(col[0], 0.25f), __fmul_rn(col[1], 0.5f)), __fmul_rn(col[2], 0.25f));
    };
    float r0 = __fmul_rn(blurred(y), c.g[FB_N]), r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 1; k <= FB_N; ++k) {
      const float up = blurred(max(y - k, 0)), dn = blurred(min(y + k, H - 1));
      const float p = __fadd_rn(up, dn);
      r0 = __fadd_rn(r0, __fmul_rn(c.g[FB_N + k], p));
      r1 = __fadd_rn(r1, __fmul_rn(c.xg[FB_N + k], __fsub_rn(dn, up)));
      r2 = __fadd_rn(r2, __fmul_rn(c.xxg[FB_N + k], p));
    }
    float* o = vert + i * 3;
    o[0] = r0; o[1] = r1; o[2] = r2;
  }
}

// ---------------------------------------------------------------- 2. horizontal pass -> R [nf][H][W][5]
__global__ void fb_horizontal_kernel(const float* __restrict__ vert, float* __restrict__ R, int nf, int H, int W, FbConst c) {
  const size_t total = static_cast<size_t>(nf) * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const float* row = vert + (i - x) * 3;
    double b1 = static_cast<double>(row[x * 3]) * c.g[FB_N], b2 = 0, b3 = static_cast<double>(row[x * 3 + 1]) * c.g[FB_N], b4 = 0,
           b5 = static_cast<double>(row[x * 3 + 2]) * c.g[FB_N], b6 = 0;
#pragma unroll
    for (int k = 1; k <= FB_N; ++k) {
      const float* rp = row + min(x + k, W - 1) * 3;
      const float* rm = row + max(x - k, 0) * 3;
      const double tg = static_cast<double>(__fadd_rn(rp[0], rm[0]));
      b1 += tg * c.g[FB_N + k];
      b4 += tg * c.xxg[FB_N + k];
      b2 += static_cast<double>(__fsub_rn(rp[0], rm[0])) * c.xg[FB_N + k];
      b3 += static_cast<double>(__fadd_rn(rp[1], rm[1])) * c.g[FB_N + k];
      b6 += static_cast<double>(__fsub_rn(rp[1], rm[1])) * c.xg[FB_N + k];
      b5 += static_cast<double>(__fadd_rn(rp[2], rm[2])) * c.g[FB_N + k];
    }
    float* o = R + i * 5;
    o[1] = static_cast<float>(b2 * c.ig11);
    o[0] = static_cast<float>(b3 * c.ig11);
    o[3] = static_cast<float>(b1 * c.ig03 + b4 * c.ig33);
    o[2] = static_cast<float>(b1 * c.ig03 + b5 * c.ig33);
    o[4] = static_cast<float>(b6 * c.ig55);
  }
}

// ---------------------------------------------------------------- 3. FarnebackUpdateMatrices
// pair p = (clip, t): R0 = R[clip][t], R1 = R[clip][t+1]; flow [pairs][H][W][2] (ignored when zero_flow); M [pairs][H][W][5]
__global__ void fb_matrices_kernel(const float* __restrict__ R, const float* __restrict__ flow, float* __restrict__ M, int clips, int T,
                                   int H, int W, int zero_flow) {
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(clips) * (T - 1) * hw;
  const float border[FB_BORDER] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const size_t pair = i / hw;
    const size_t clip = pair / (T - 1), t = pair % (T - 1);
    const float* R0 = R + ((clip * T + t) * hw + static_cast<size_t>(y) * W + x) * 5;
    const float* R1 = R + (clip * T + t + 1) * hw * 5;
    const float dx = zero_flow ? 0.f : flow[i * 2], dy = zero_flow ? 0.f : flow[i * 2 + 1];
    float fx = __fadd_rn(static_cast<float>(x), dx), fy = __fadd_rn(static_cast<float>(y), dy);
    const int x1 = static_cast<int>(floorf(fx)), y1 = static_cast<int>(floorf(fy));
    fx = __fsub_rn(fx, static_cast<float>(x1));
    fy = __fsub_rn(fy, static_cast<float>(y1));
    float r2, r3, r4, r5, r6;
    if (static_cast<unsigned>(x1) < static_cast<unsigned>(W - 1) && static_cast<unsigned>(y1) < static_cast<unsigned>(H - 1)) {
      const float a00 = __fmul_rn(__fsub_rn(1.f, fx), __fsub_rn(1.f, fy)), a01 = __fmul_rn(fx, __fsub_rn(1.f, fy));
      const float a10 = __fmul_rn(__fsub_rn(1.f, fx), fy), a11 = __fmul_rn(fx, fy);
      const float* p00 = R1 + (static_cast<size_t>(y1) * W + x1) * 5;
      const float* p10 = p00 + static_cast<size_t>(W) * 5;
      auto bil = [&](int ch) {
        float v = __fmul_rn(a00, p00[ch]);
        v = __fadd_rn(v, __fmul_rn(a01, p00[5 + ch]));
        v = __fadd_rn(v, __fmul_rn(a10, p10[ch]));
        return __fadd_rn(v, __fmul_rn(a11, p10[5 + ch]));
      };
      r2 = bil(0); r3 = bil(1);
      r4 = __fmul_rn(__fadd_rn(R0[2], bil(2)), 0.5f);
      r5 = __fmul_rn(__fadd_rn(R0[3], bil(3)), 0.5f);
      r6 = __fmul_rn(__fadd_rn(R0[4], bil(4)), 0.25f);
    } else {
      r2 = r3 = 0.f;
      r4 = R0[2]; r5 = R0[3]; r6 = __fmul_rn(R0[4], 0.5f);
    }
    r2 = __fmul_rn(__fsub_rn(R0[0], r2), 0.5f);
    r3 = __fmul_rn(__fsub_rn(R0[1], r3), 0.5f);
    r2 = __fadd_rn(r2, __fadd_rn(__fmul_rn(r4, dy), __fmul_rn(r6, dx)));
    r3 = __fadd_rn(r3, __fadd_rn(__fmul_rn(r6, dy), __fmul_rn(r5, dx)));
    if (static_cast<unsigned>(x - FB_BORDER) >= static_cast<unsigned>(W - FB_BORDER * 2) ||
        static_cast<unsigned>(y - FB_BORDER) >= static_cast<unsigned>(H - FB_BORDER * 2)) {
      float s = x < FB_BORDER ? border[x] : 1.f;
      s = __fmul_rn(s, x >= W - FB_BORDER ? border[W - x - 1] : 1.f);
      s = __fmul_rn(s, y < FB_BORDER ? border[y] : 1.f);
      s = __fmul_rn(s, y >= H - FB_BORDER ? border[H - y - 1] : 1.f);
      r2 = __fmul_rn(r2, s); r3 = __fmul_rn(r3, s); r4 = __fmul_rn(r4, s); r5 = __fmul_rn(r5, s); r6 = __fmul_rn(r6, s);
    }
    float* o = M + i * 5;
    o[0] = __fadd_rn(__fmul_rn(r4, r4), __fmul_rn(r6, r6));
    o[1] = __fmul_rn(__fadd_rn(r4, r5), r6);
    o[2] = __fadd_rn(__fmul_rn(r5, r5), __fmul_rn(r6, r6));
    o[3] = __fadd_rn(__fmul_rn(r4, r2), __fmul_rn(r6, r3));
    o[4] = __fadd_rn(__fmul_rn(r6, r2), __fmul_rn(r5, r3));
  }
}

// ---------------------------------------------------------------- 4. box sums (separable, double) and the 2x2 solve
__global__ void fb_box_vertical_kernel(const float* __restrict__ M, double* __restrict__ V, int pairs, int H, int W, int m) {
  const size_t total = static_cast<size_t>(pairs) * H * W * 5;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row_elems = static_cast<size_t>(W) * 5;
    const int y = static_cast<int>((i / row_elems) % H);
    const size_t base = i - static_cast<size_t>(y) * row_elems;      // element (pair, y = 0, x, ch)
    double s = 0.0;
    for (int d = -m; d <= m; ++d) s += static_cast<double>(M[base + static_cast<size_t>(min(max(y + d, 0), H - 1)) * row_elems]);
    V[i] = s;
  }
}

__global__ void fb_box_solve_kernel(const double* __restrict__ V, float* __restrict__ flow, int pairs, int H, int W, int m, double scale) {
  const size_t total = static_cast<size_t>(pairs) * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const double* row = V + (i - x) * 5;
    double s[5] = {0, 0, 0, 0, 0};
    for (int d = -m; d <= m; ++d) {
      const double* p = row + static_cast<size_t>(min(max(x + d, 0), W - 1)) * 5;
#pragma unroll
      for (int ch = 0; ch < 5; ++ch) s[ch] += p[ch];
    }
    const double g11 = s[0] * scale, g12 = s[1] * scale, g22 = s[2] * scale, h1 = s[3] * scale, h2 = s[4] * scale;
    const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
    flow[i * 2] = static_cast<float>((g11 * h2 - g12 * h1) * idet);
    flow[i * 2 + 1] = static_cast<float>((g22 * h1 - g12 * h2) * idet);
  }
}

// ---------------------------------------------------------------- 5. flow similarity metrics (scheduler :541-604)
// flows [channels][pairs_per_channel][H][W][2] (dx, dy); out [channels][3] = mean EPE, mean outlier, mean angle (degrees)
constexpr int FM_THREADS = 1024;
__global__ void __launch_bounds__(FM_THREADS) flow_metrics_kernel(const float* __restrict__ ref, const float* __restrict__ cand,
                                                                  float* __restrict__ out, long long per_channel, int outlier_or) {
  __shared__ double red[3][FM_THREADS / 32];
  const float* r = ref + static_cast<size_t>(blockIdx.x) * per_channel * 2;
  const float* c = cand + static_cast<size_t>(blockIdx.x) * per_channel * 2;
  double s_epe = 0, s_out = 0, s_ang = 0;
  for (long long i = threadIdx.x; i < per_channel; i += FM_THREADS) {
    const float rx = r[2 * i], ry = r[2 * i + 1], cx = c[2 * i], cy = c[2 * i + 1];
    const float dx = __fsub_rn(rx, cx), dy = __fsub_rn(ry, cy);
    const float epe = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1e-8f));
    const float dot = __fadd_rn(__fmul_rn(rx, cx), __fmul_rn(ry, cy));
    const float rn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), 1e-8f));
    const float cn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), 1e-8f));
    float cosv = __fdiv_rn(dot, __fadd_rn(__fmul_rn(rn, cn), 1e-8f));
    cosv = fminf(fmaxf(cosv, -1.0f), 1.0f);
    const float ang = __fdiv_rn(__fmul_rn(acosf(cosv), 180.0f), 3.14159265358979323846f);
    s_epe += epe;
    s_ang += ang;
    // Wan: epe > 3 AND epe > 0.05 |ref| (scheduling_unipc_multistep_clean.py:557); LongCat: OR (scheduling_flow_match_euler_discrete.py:226)
    const bool a = epe > 3.0f, b = epe > __fmul_rn(rn, 0.05f);
    s_out += (outlier_or ? (a || b) : (a && b)) ? 1.0 : 0.0;
  }
  double v[3] = {s_epe, s_out, s_ang};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int w = 0; w < FM_THREADS / 32; ++w) t += red[threadIdx.x][w];
    out[blockIdx.x * 3 + threadIdx.x] = static_cast<float>(t / static_cast<double>(per_channel));
  }
}

static FbConst fb_constants(double sigma) {
  // FarnebackPrepareGaussian(n = 5, sigma): weights in float, the four used entries of G^-1 in double
  FbConst c{};
  const int n = FB_N;
  double s = 0;
  for (int x = -n; x <= n; ++x) { c.g[x + n] = static_cast<float>(std::exp(-x * x / (2 * sigma * sigma))); s += c.g[x + n]; }
  s = 1. / s;
  for (int x = -n; x <= n; ++x) {
    c.g[x + n] = static_cast<float>(c.g[x + n] * s);
    c.xg[x + n] = static_cast<float>(x * c.g[x + n]);
    c.xxg[x + n] = static_cast<float>(x * x * c.g[x + n]);
  }
  // G is block structured: [G00 .. G03 G04; G11; G22 = G11; G03 G33 G34; G04 G34 G44; G55] with G03 = G04 = G11, G44 = G33, G34 = G55
  double G00 = 0, G11 = 0, G33 = 0, G55 = 0;
  for (int y = -n; y <= n; ++y)
    for (int x = -n; x <= n; ++x) {
      const double w = static_cast<double>(c.g[y + n]) * c.g[x + n];
      G00 += w; G11 += w * x * x; G33 += w * x * x * x * x; G55 += w * x * x * y * y;
    }
  // inverse of the 3x3 block over (1, x^2, y^2): A = [[G00, G11, G11], [G11, G33, G55], [G11, G55, G33]]
  const double a = G00, b = G11, d = G33, e = G55;
  const double det = a * (d * d - e * e) - b * (b * d - b * e) + b * (b * e - b * d);
  c.ig03 = -(b * d - b * e) / det;              // cofactor C01 / det (A symmetric)
  c.ig33 = (a * d - b * b) / det;               // C11 / det
  c.ig11 = 1.0 / G11;
  c.ig55 = 1.0 / G55;
  return c;
}

// cv::getGaussianKernel(3, sigma > 0, CV_32F): [e s, s, e s] with e = exp(-1 / (2 sigma^2)), s = 1 / (1 + 2 e)
static void fb_gauss3(double sigma, float k[3]) {
  const double e = std::exp(-1.0 / (2.0 * sigma * sigma)), s = 1.0 / (1.0 + 2.0 * e);
  k[0] = k[2] = static_cast<float>(e * s);
  k[1] = static_cast<float>(s);
}

// number of pyramid levels BELOW full resolution the OpenCV driver runs for (pyr_scale 0.5, levels 3)
static int fb_extra_levels(int H, int W) {
  int k = 0;
  double scale = 1.0;
  for (; k < 3; ++k) {
    scale *= 0.5;
    if (W * scale < 32 || H * scale < 32) break;
  }
  return k;
}

}  // namespace wf

using namespace wf;

static int flow_grid(size_t n) {
  return static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(sm_count()) * 16));
}

extern "C" long long wf_farneback_workspace_bytes(int clips, int T, int H, int W) {
  const long long hw = static_cast<long long>(H) * W;
  const long long frames = static_cast<long long>(clips) * T, pairs = static_cast<long long>(clips) * (T - 1);
  // vert (frames*hw*3 f32) | R (frames*hw*5 f32) | M (pairs*hw*5 f32) | V (pairs*hw*5 f64) | half-resolution images
  // (frames*hw/4 f32) | half-resolution flow (pairs*hw/4*2 f32), each rounded to 256 B; the first four serve both levels
  auto r = [](long long b) { return (b + 255) / 256 * 256; };
  return r(frames * hw * 3 * 4) + r(frames * hw * 5 * 4) + r(pairs * hw * 5 * 4) + r(pairs * hw * 5 * 8) + r(frames * (hw / 4) * 4) +
         r(pairs * (hw / 4) * 2 * 4);
}

extern "C" int wf_farneback_u8(const unsigned char* clips_u8, int clips, int T, int H, int W, int winsize, int iterations,
                               float* flow, void* workspace, void* stream) {
  WF_REQUIRE(clips_u8 && flow && workspace, "wf_farneback_u8: null pointer");
  WF_REQUIRE(clips > 0 && T >= 2 && H >= 2 * FB_BORDER && W >= 2 * FB_BORDER, "wf_farneback_u8: empty clips or frames smaller than 10x10");
  const int extra = fb_extra_levels(H, W);
  WF_REQUIRE(extra == 0 || (extra == 1 && H % 2 == 0 && W % 2 == 0),
             "wf_farneback_u8: frames whose Farneback pyramid has one level (min side < 64) or two with even sides (min side < 128)");
  WF_REQUIRE(winsize >= 3 && (winsize & 1) && iterations >= 1, "wf_farneback_u8: odd window >= 3, at least one iteration");
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t frames = static_cast<size_t>(clips) * T, pairs = static_cast<size_t>(clips) * (T - 1);
  auto r = [](size_t b) { return (b + 255) / 256 * 256; };
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* vert = reinterpret_cast<float*>(ws); ws += r(frames * hw * 3 * 4);
  float* R = reinterpret_cast<float*>(ws); ws += r(frames * hw * 5 * 4);
  float* M = reinterpret_cast<float*>(ws); ws += r(pairs * hw * 5 * 4);
  double* V = reinterpret_cast<double*>(ws); ws += r(pairs * hw * 5 * 8);
  float* half_img = reinterpret_cast<float*>(ws); ws += r(frames * (hw / 4) * 4);
  float* half_flow = reinterpret_cast<float*>(ws);
  static const FbConst c = fb_constants(1.2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int m = winsize / 2;
  const double box_scale = 1.0 / (static_cast<double>(winsize) * winsize);
  auto iterate = [&](float* fl, int h, int w, bool zero_start) {
    const size_t n = pairs * static_cast<size_t>(h) * w;
    for (int it = 0; it < iterations; ++it) {
      fb_matrices_kernel<<<flow_grid(n), 256, 0, st>>>(R, fl, M, clips, T, h, w, (it == 0 && zero_start) ? 1 : 0);
      fb_box_vertical_kernel<<<flow_grid(n * 5), 256, 0, st>>>(M, V, static_cast<int>(pairs), h, w, m);
      fb_box_solve_kernel<<<flow_grid(n), 256, 0, st>>>(V, fl, static_cast<int>(pairs), h, w, m, box_scale);
    }
  };
  if (extra == 1) {
    // half-resolution level first: sigma = (1 / 0.5 - 1) * 0.5 = 0.5, 3 x 3 kernel, exact 2 x 2 mean; flow starts at zero
    const int h2 = H / 2, w2 = W / 2;
    float k3[3];
    fb_gauss3(0.5, k3);
    fb_blur_half_kernel<<<flow_grid(frames * h2 * w2), 256, 0, st>>>(clips_u8, half_img, static_cast<int>(frames), H, W, k3[0], k3[1], k3[2]);
    fb_vertical_f32_kernel<<<flow_grid(frames * h2 * w2), 256, 0, st>>>(half_img, vert, static_cast<int>(frames), h2, w2, c);
    fb_horizontal_kernel<<<flow_grid(frames * h2 * w2), 256, 0, st>>>(vert, R, static_cast<int>(frames), h2, w2, c);
    iterate(half_flow, h2, w2, true);
    fb_double_flow_kernel<<<flow_grid(pairs * hw), 256, 0, st>>>(half_flow, flow, static_cast<int>(pairs), H, W);
  }
  fb_vertical_kernel<<<flow_grid(frames * hw), 256, 0, st>>>(clips_u8, vert, static_cast<int>(frames), H, W, c);
  fb_horizontal_kernel<<<flow_grid(frames * hw), 256, 0, st>>>(vert, R, static_cast<int>(frames), H, W, c);
  iterate(flow, H, W, extra == 0);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_flow_metrics(const float* flow_ref, const float* flow_cand, float* out3, int channels, long long per_channel,
                               int outlier_or, void* stream) {
  WF_REQUIRE(flow_ref && flow_cand && out3 && channels > 0 && per_channel > 0, "wf_flow_metrics: bad arguments");
  flow_metrics_kernel<<<channels, FM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(flow_ref, flow_cand, out3, per_channel, outlier_or);
  WF_LAUNCH_OK();
  return WF_OK;
}
