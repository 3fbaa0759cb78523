"""The device Farneback kernels (worldforge_b200/csrc/flow_ops.cu) compiled FOR THE HOST and run on the CPU: the kernel
source itself - grid-stride loops, indexing, arithmetic order - is executed with one 'thread' (blockDim = gridDim = 1) behind
a few shims for the CUDA intrinsics, and must reproduce oracle/farneback.py BIT FOR BIT (the _rn intrinsics map to plain IEEE
operations, g++ runs with -ffp-contract=off).  This is the no-GPU half of the kernel's parity evidence; tests/test_flow_gpu.py
is the other half."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import farneback as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r'''
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <vector>
struct dim3e { unsigned x; };
static dim3e blockIdx{0}, blockDim{1}, gridDim{1}, threadIdx{0};
#define __global__
#define __restrict__
#define __launch_bounds__(x)
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
using std::min; using std::max;
'''
MAIN = r'''
using namespace wf;
int main(int argc, char** argv) {
  int clips = atoi(argv[1]), T = atoi(argv[2]), H = atoi(argv[3]), W = atoi(argv[4]);
  size_t hw = (size_t)H * W, frames = (size_t)clips * T, pairs = (size_t)clips * (T - 1);
  std::vector<unsigned char> u8(frames * hw);
  FILE* f = fopen(argv[5], "rb"); if (!f || fread(u8.data(), 1, u8.size(), f) != u8.size()) return 2; fclose(f);
  std::vector<float> vert(frames * hw * 3), R(frames * hw * 5), M(pairs * hw * 5), flow(pairs * hw * 2);
  std::vector<double> V(pairs * hw * 5);
  FbConst c = fb_constants(1.2);
  const int extra = fb_extra_levels(H, W);
  if (extra > 1 || (extra == 1 && (H % 2 || W % 2))) return 3;
  if (extra == 1) {      // the launch sequence of wf_farneback_u8 for a two-level pyramid
    const int h2 = H / 2, w2 = W / 2;
    std::vector<float> half_img(frames * h2 * w2), half_flow(pairs * h2 * w2 * 2);
    float k3[3]; fb_gauss3(0.5, k3);
    fb_blur_half_kernel(u8.data(), half_img.data(), (int)frames, H, W, k3[0], k3[1], k3[2]);
    fb_vertical_f32_kernel(half_img.data(), vert.data(), (int)frames, h2, w2, c);
    fb_horizontal_kernel(vert.data(), R.data(), (int)frames, h2, w2, c);
    for (int it = 0; it < 3; ++it) {
      fb_matrices_kernel(R.data(), half_flow.data(), M.data(), clips, T, h2, w2, it == 0);
      fb_box_vertical_kernel(M.data(), V.data(), (int)pairs, h2, w2, 7);
      fb_box_solve_kernel(V.data(), half_flow.data(), (int)pairs, h2, w2, 7, 1.0 / 225.0);
    }
    fb_double_flow_kernel(half_flow.data(), flow.data(), (int)pairs, H, W);
  }
  fb_vertical_kernel(u8.data(), vert.data(), (int)frames, H, W, c);
  fb_horizontal_kernel(vert.data(), R.data(), (int)frames, H, W, c);
  for (int it = 0; it < 3; ++it) {
    fb_matrices_kernel(R.data(), flow.data(), M.data(), clips, T, H, W, it == 0 && extra == 0);
    fb_box_vertical_kernel(M.data(), V.data(), (int)pairs, H, W, 7);
    fb_box_solve_kernel(V.data(), flow.data(), (int)pairs, H, W, 7, 1.0 / 225.0);
  }
  f = fopen(argv[6], "wb"); fwrite(flow.data(), 4, flow.size(), f); fclose(f);
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("H,W", [(60, 104), (90, 160)])           # 480p latent frames: one pyramid level; 720p: two
def test_flow_kernel_source_reproduces_the_oracle_on_the_host(tmp_path, H, W):
    src = open(os.path.join(ROOT, "worldforge_b200", "csrc", "flow_ops.cu")).read()
    i = src.index("namespace wf {")
    j = src.index("// ---------------------------------------------------------------- 5. flow similarity metrics")
    k = src.index("static FbConst fb_constants(double sigma) {")
    k2 = src.index("}  // namespace wf")
    cpp = tmp_path / "emu.cpp"
    cpp.write_text(SHIM + src[i:j] + src[k:k2] + "}\n" + MAIN)
    exe = tmp_path / "emu"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", str(cpp), "-o", str(exe)], check=True)
    rng = np.random.default_rng(5)
    clips, T = 3, 3
    yy, xx = np.mgrid[0:H, 0:W]
    u8 = np.zeros((clips, T, H, W), np.uint8)
    for t in range(T):
        u8[0, t] = (127 + 100 * np.sin((xx - 2 * t) / 7.0) * np.cos(yy / 5.0)).astype(np.uint8)          # moves right
        u8[1, t] = (127 + 90 * np.sin((xx + t) / 9.0 + (yy - t) / 6.0)).astype(np.uint8)                 # moves diagonally
    u8[2] = (rng.random((T, H, W)) * 255).astype(np.uint8)                                                # noise
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    u8.tofile(fin)
    subprocess.run([str(exe), str(clips), str(T), str(H), str(W), str(fin), str(fout)], check=True)
    got = np.fromfile(fout, np.float32).reshape(clips, T - 1, H, W, 2)
    for c in range(clips):
        for t in range(T - 1):
            assert np.array_equal(got[c, t], fb.farneback(u8[c, t], u8[c, t + 1])), (c, t)
