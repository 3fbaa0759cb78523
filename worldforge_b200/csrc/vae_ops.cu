// worldforge_b200 - HBM-bound kernels of the Wan 3D-VAE on channels-last fp32 activations [T][H][W][C].
//
//   rms-norm (+SiLU)   F.normalize(x, dim=channel) * sqrt(C) * gamma, then SiLU     vae.py:51-54, :195-197
//   layout conversion  planar [C][T][H][W] <-> channels-last (zero-padded channels)  (VAE boundary)
//   space-to-depth     [T][H][W][C] -> [T][H/2][W/2][4C] for the stride-2 convs       vae.py:87-96
//   softmax rows       single-head spatial attention of the mid block                vae.py:252-256
//   transpose          V^T for the PV product of that attention
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per pixel; C <= 32*RMS_MAXK
constexpr int VR_MAXK = 24;   // up to 768 channels
__global__ void __launch_bounds__(256) rms_silu_cl_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                          const float* __restrict__ gamma, size_t pixels, int C, int ldx,
                                                          int ldo, int silu, int round_tf32) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const size_t nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
  const float scale = sqrtf(static_cast<float>(C));
  for (size_t pix = warp0; pix < pixels; pix += nwarps) {
    const float* xr = x + pix * ldx;
    float v[VR_MAXK];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VR_MAXK; ++k) {
      const int c = lane + 32 * k;
      v[k] = (c < C) ? xr[c] : 0.f;
      ss += v[k] * v[k];
    }
    ss = wsum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    float* orow = out + pix * ldo;
#pragma unroll
    for (int k = 0; k < VR_MAXK; ++k) {
      const int c = lane + 32 * k;
      if (c < C) {
        float y = v[k] / denom * scale * gamma[c];
        if (silu) y = y / (1.0f + expf(-y));
        orow[c] = round_tf32 ? tf32_round(y) : y;
      }
    }
  }
}

// planar [C][N] -> channels-last [N][Cp] with channels C..Cp-1 zero
__global__ void planar_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t N, int C, int Cp, int round_tf32) {
  const size_t total = N * Cp;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cp);
    const size_t n = i / Cp;
    const float v = (c < C) ? src[static_cast<size_t>(c) * N + n] : 0.f;
    dst[i] = round_tf32 ? tf32_round(v) : v;
  }
}
// dst = src rounded to tf32 (the raw residual stream read by a 1x1 shortcut convolution, vae.py:192-193)
__global__ void round_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 v = src[i];
    v.x = tf32_round(v.x); v.y = tf32_round(v.y); v.z = tf32_round(v.z); v.w = tf32_round(v.w);
    dst[i] = v;
  }
}
// channels-last [N][ld] (first C channels) -> planar [C][N]
__global__ void cl_to_planar_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t N, int C, int ld) {
  const size_t total = N * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i % N;
    const int c = static_cast<int>(i / N);
    dst[i] = src[n * ld + c];
  }
}

// [T][H][W][C] -> [T][H/2][W/2][(p*2+q)*C + c] = src[t][2y+p][2x+q][c]
__global__ void space_to_depth_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int T, int H, int W, int C4) {
  const int H2 = H / 2, W2 = W / 2;
  const size_t total = static_cast<size_t>(T) * H2 * W2 * 4 * C4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    size_t r = i / C4;
    const int pq = static_cast<int>(r % 4); r /= 4;
    const int x = static_cast<int>(r % W2); r /= W2;
    const int y = static_cast<int>(r % H2);
    const int t = static_cast<int>(r / H2);
    const int p = pq >> 1, q = pq & 1;
    dst[i] = src[((static_cast<size_t>(t) * H + 2 * y + p) * W + 2 * x + q) * C4 + c];
  }
}

// in-place softmax(scale * x) over each row of x [rows][ld]
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int cols, int ld, float scale) {
  __shared__ float red[8];
  __shared__ float bc;
  float* row = x + static_cast<size_t>(blockIdx.x) * ld;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += 256) m = fmaxf(m, row[i]);
  m = wmax(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { float t = red[0]; for (int k = 1; k < 8; ++k) t = fmaxf(t, red[k]); bc = t; }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += 256) { const float e = expf((row[i] - m) * scale); row[i] = e; s += e; }
  s = wsum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += red[k]; bc = t; }
  __syncthreads();
  const float inv = 1.0f / bc;
  for (int i = threadIdx.x; i < cols; i += 256) row[i] *= inv;
}

// dst[c][r] = src[r][c] for r < R, c < C  (src row stride lds, dst row stride ldd)
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C, int lds, int ldd) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? src[static_cast<size_t>(r) * lds + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[static_cast<size_t>(c) * ldd + r] = tile[threadIdx.x][j];
  }
}

// hi = x with the 13 low mantissa bits cleared (exactly what a kind::tf32 tcgen05.mma reads of x), lo = x - hi (exact in
// fp32).  A . B = A_hi.B_hi + A_hi.B_lo + A_lo.B_hi + O(2^-22): three tf32 products accumulate to fp32 accuracy - used for
// the mid-block attention of the VAE, whose matmuls the reference computes in fp32 (vae.py:252-256; torch keeps
// allow_tf32 OFF for matmuls, unlike cuDNN convolutions).
__global__ void split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    hi[i] = h; lo[i] = l;
  }
}

static int grid_for_n(size_t n, int waves = 8) {
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>((n + 255) / 256, static_cast<size_t>(sm_count()) * waves)));
}

}  // namespace wf

using namespace wf;
#define WF_STREAM static_cast<cudaStream_t>(stream)

extern "C" int wf_rms_norm_cl(const float* x, int ldx, float* out, int ldo, const float* gamma, long long pixels, int C,
                              int silu, int round_tf32, void* stream) {
  WF_REQUIRE(x && out && gamma && pixels > 0, "wf_rms_norm_cl: bad arguments");
  WF_REQUIRE(C > 0 && C <= 32 * VR_MAXK, "wf_rms_norm_cl: 1..768 channels");
  const size_t warps = static_cast<size_t>(pixels);
  const int blocks = static_cast<int>(std::min<size_t>((warps + 7) / 8, static_cast<size_t>(sm_count()) * 16));
  rms_silu_cl_kernel<<<blocks, 256, 0, WF_STREAM>>>(x, out, gamma, static_cast<size_t>(pixels), C, ldx, ldo, silu, round_tf32);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_planar_to_cl(const float* src, float* dst, long long n, int C, int Cp, int round_tf32, void* stream) {
  WF_REQUIRE(src && dst && n > 0 && C > 0 && Cp >= C, "wf_planar_to_cl: bad arguments");
  planar_to_cl_kernel<<<grid_for_n(static_cast<size_t>(n) * Cp), 256, 0, WF_STREAM>>>(src, dst, static_cast<size_t>(n), C, Cp, round_tf32);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_cl_to_planar(const float* src, float* dst, long long n, int C, int ld, void* stream) {
  WF_REQUIRE(src && dst && n > 0 && C > 0 && ld >= C, "wf_cl_to_planar: bad arguments");
  cl_to_planar_kernel<<<grid_for_n(static_cast<size_t>(n) * C), 256, 0, WF_STREAM>>>(src, dst, static_cast<size_t>(n), C, ld);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_space_to_depth(const float* src, float* dst, int T, int H, int W, int C, void* stream) {
  WF_REQUIRE(src && dst && T > 0, "wf_space_to_depth: bad arguments");
  WF_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "wf_space_to_depth: H, W even and C a multiple of 4");
  const size_t total = static_cast<size_t>(T) * H * W * C / 4;
  space_to_depth_kernel<<<grid_for_n(total), 256, 0, WF_STREAM>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), T, H, W, C / 4);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_softmax_rows(float* x, int rows, int cols, int ld, float scale, void* stream) {
  WF_REQUIRE(x && rows > 0 && cols > 0 && ld >= cols, "wf_softmax_rows: bad arguments");
  softmax_rows_kernel<<<rows, 256, 0, WF_STREAM>>>(x, cols, ld, scale);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_transpose_f32(const float* src, float* dst, int R, int C, int lds, int ldd, void* stream) {
  WF_REQUIRE(src && dst && R > 0 && C > 0, "wf_transpose_f32: bad arguments");
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_f32_kernel<<<grid, block, 0, WF_STREAM>>>(src, dst, R, C, lds, ldd);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream) {
  WF_REQUIRE(x && hi && lo && n > 0 && n % 4 == 0, "wf_split_tf32: bad arguments (n must be a multiple of 4)");
  WF_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) % 16 == 0,
             "wf_split_tf32: pointers must be 16-byte aligned");
  split_tf32_kernel<<<grid_for_n(static_cast<size_t>(n) / 4), 256, 0, WF_STREAM>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), static_cast<size_t>(n) / 4);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_round_tf32(const float* src, float* dst, long long n, void* stream) {
  WF_REQUIRE(src && dst && n > 0 && n % 4 == 0, "wf_round_tf32: bad arguments (n must be a multiple of 4)");
  WF_REQUIRE((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0, "wf_round_tf32: pointers must be 16-byte aligned");
  round_tf32_kernel<<<grid_for_n(static_cast<size_t>(n) / 4), 256, 0, WF_STREAM>>>(reinterpret_cast<const float4*>(src),
                                                                                    reinterpret_cast<float4*>(dst), static_cast<size_t>(n) / 4);
  WF_LAUNCH_OK();
  return WF_OK;
}
