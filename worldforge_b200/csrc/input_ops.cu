// worldforge_b200 - input preparation of the entry script on the device.
//
//   soften_mask (wan_for_worldforge/infer_worldforge.py:105-150): every frame of the warp mask gets a smooth 1 -> 0 ramp
//   inside its "one" region: a pixel that is set and lies within `transition_distance` of an unset pixel becomes
//   smooth(d / transition_distance), d = its exact Euclidean distance to the nearest unset pixel
//   (scipy.ndimage.distance_transform_edt).  Only distances <= transition_distance matter, so the exact distance is a
//   windowed search: d^2 = min over the (2R+1)^2 neighbourhood of dx^2 + dy^2 at unset pixels - an integer - and the
//   ramp value is looked up in a table the host computed with the reference's own float64 numpy expression for every
//   possible d^2 (at most R^2 + 1 entries): the result is bit-identical to the reference by construction.
//
//   clip_from_u8 (infer_worldforge.py:232-238): uint8 frames [F][H][W][3] -> planar fp32 [3][F][H][W] / 255.
#include "common.cuh"
#include "host_util.h"
#include <algorithm>

namespace wf {

constexpr int SM_TILE = 32;
constexpr int SM_MAX_R = 31;

__global__ void __launch_bounds__(SM_TILE * SM_TILE)
soften_mask_kernel(const float* __restrict__ mask, float* __restrict__ out, int H, int W, int R, int max_d2,
                   const float* __restrict__ lut) {
  extern __shared__ unsigned char set[];                   // (TILE + 2R)^2: 1 = set, 0 = unset, 2 = outside the image
  const int span = SM_TILE + 2 * R;
  const size_t frame = static_cast<size_t>(blockIdx.z) * H * W;
  const int x0 = blockIdx.x * SM_TILE - R, y0 = blockIdx.y * SM_TILE - R;
  for (int i = threadIdx.x; i < span * span; i += blockDim.x) {
    const int yy = y0 + i / span, xx = x0 + i % span;
    set[i] = (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 2 : (mask[frame + static_cast<size_t>(yy) * W + xx] != 0.0f ? 1 : 0);
  }
  __syncthreads();
  const int tx = threadIdx.x % SM_TILE, ty = threadIdx.x / SM_TILE;
  const int x = blockIdx.x * SM_TILE + tx, y = blockIdx.y * SM_TILE + ty;
  if (x >= W || y >= H) return;
  const size_t at = frame + static_cast<size_t>(y) * W + x;
  const float m = mask[at];
  float v = m;
  if (m != 0.0f) {
    int best = max_d2 + 1;
    for (int dy = -R; dy <= R; ++dy) {
      const int row = (ty + R + dy) * span + tx + R;
      const int dy2 = dy * dy;
      if (dy2 >= best) continue;
      for (int dx = -R; dx <= R; ++dx) {
        const int d2 = dy2 + dx * dx;
        if (d2 < best && set[row + dx] == 0) best = d2;
      }
    }
    if (best <= max_d2) v = lut[best];
  }
  out[at] = v;
}

__global__ void clip_from_u8_kernel(const unsigned char* __restrict__ frames, float* __restrict__ out, size_t pixels) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < pixels; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const unsigned char* p = frames + 3 * i;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c * pixels + i] = static_cast<float>(p[c]) / 255.0f;
  }
}

}  // namespace wf

using namespace wf;

extern "C" int wf_soften_mask(const float* mask, float* out, int frames, int H, int W, int radius, int max_d2, const float* lut,
                              void* stream) {
  WF_REQUIRE(mask && out && lut, "wf_soften_mask: null pointer");
  WF_REQUIRE(frames > 0 && H > 0 && W > 0, "wf_soften_mask: empty mask");
  WF_REQUIRE(radius >= 0 && radius <= SM_MAX_R && max_d2 >= 0 && max_d2 <= radius * radius + 2 * radius,
             "wf_soften_mask: transition distance out of range (0..31 pixels)");
  dim3 grid((W + SM_TILE - 1) / SM_TILE, (H + SM_TILE - 1) / SM_TILE, frames);
  const int span = SM_TILE + 2 * radius;
  static PerDeviceOnce once;
  const int rc = once.run([] {
    WF_CUDA_OK(cudaFuncSetAttribute(soften_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (SM_TILE + 2 * SM_MAX_R) * (SM_TILE + 2 * SM_MAX_R)));
    return static_cast<int>(WF_OK);
  });
  if (rc) return rc;
  soften_mask_kernel<<<grid, SM_TILE * SM_TILE, span * span, static_cast<cudaStream_t>(stream)>>>(mask, out, H, W, radius, max_d2, lut);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_clip_from_u8(const unsigned char* frames, float* out, long long pixels, void* stream) {
  WF_REQUIRE(frames && out && pixels > 0, "wf_clip_from_u8: bad arguments");
  const int grid = static_cast<int>(std::min<long long>((pixels + 255) / 256, 148LL * 16));
  clip_from_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(frames, out, static_cast<size_t>(pixels));
  WF_LAUNCH_OK();
  return WF_OK;
}
