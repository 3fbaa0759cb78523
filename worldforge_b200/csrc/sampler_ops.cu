// worldforge_b200 - the WorldForge inference-time ops as fused, 128-bit vectorised HBM kernels.
//
//   CFG combine            v = v_c + s (v_c - v_u)                         pipeline_wan_i2v_clean.py:611
//   x0 conversion          x0 = x - sigma v                                scheduling_unipc...:952-958
//   UniP-bh2 update        x' = (s_t/s_0) x - a_t E m0 - a_t B (m1-m0)/(2 r)        :1084-1099
//   IRR re-noise           x = (1-sigma) x0 + sigma n                      :1584
//   DSG                    v* = g + w sin(th) (g - rho cos(th) w)          pipeline...:664-681
//   FLF pixel blend        F = (2 ref - 1) m + D (1 - m)                   scheduling...:1375-1381
//   latent (de)normalise   z = x0/inv_std + mean ; E = (enc-mean) inv_std  :1281,:1385
//   channel replace        E[:,c] = x0[:,c] for the FLF-selected channels  :1410-1412
//   uint8 quantisation     min-max normalise, *255, truncate               :376-378,:175-176
//
// Rounding contract.  The reference evaluates each of these as a chain of separate torch ops
// on tensors that are bf16 from the first DSG step on (pipeline...:708), so every
// intermediate is rounded to bf16 and type promotion decides which are not.  The kernels
// carry a (value, is_bf16) pair through the same chain: one pass over HBM, but bit-identical
// to the op-by-op evaluation for any mix of fp32 / bf16 operands.  Products and sums use the
// _rn intrinsics so the compiler cannot contract them into FMAs.
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

struct TV { float v; bool b; };   // a torch tensor element: value and "dtype is bf16"

__device__ __forceinline__ TV rnd(float v, bool b) { return TV{b ? bf16_round(v) : v, b}; }
__device__ __forceinline__ TV t_mul_s(float s, TV a) { return rnd(__fmul_rn(s, a.v), a.b); }        // 0-dim tensor / python scalar * tensor
__device__ __forceinline__ TV t_div_s(TV a, float s) { return rnd(__fdiv_rn(a.v, s), a.b); }
__device__ __forceinline__ TV t_add(TV a, TV c) { return rnd(__fadd_rn(a.v, c.v), a.b && c.b); }
__device__ __forceinline__ TV t_sub(TV a, TV c) { return rnd(__fsub_rn(a.v, c.v), a.b && c.b); }
__device__ __forceinline__ TV t_mul(TV a, TV c) { return rnd(__fmul_rn(a.v, c.v), a.b && c.b); }
__device__ __forceinline__ TV t_div(TV a, TV c) { return rnd(__fdiv_rn(a.v, c.v), a.b && c.b); }

__device__ __forceinline__ TV t_load(const void* p, bool is_bf16, size_t i) {
  return is_bf16 ? TV{__bfloat162float(static_cast<const bf16*>(p)[i]), true} : TV{static_cast<const float*>(p)[i], false};
}
__device__ __forceinline__ void t_store(void* p, bool is_bf16, size_t i, float v) {
  if (is_bf16) static_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
  else static_cast<float*>(p)[i] = v;
}

// 4 consecutive elements per thread: 16-byte fp32 / 8-byte bf16 accesses
struct Vec4 { float v[4]; };
__device__ __forceinline__ Vec4 ld4(const void* p, bool is_bf16, size_t i4) {
  Vec4 r;
  if (is_bf16) {
    const uint2 raw = static_cast<const uint2*>(p)[i4];
    float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = c.x; r.v[3] = c.y;
  } else {
    const float4 f = static_cast<const float4*>(p)[i4];
    r.v[0] = f.x; r.v[1] = f.y; r.v[2] = f.z; r.v[3] = f.w;
  }
  return r;
}
__device__ __forceinline__ void st4(void* p, bool is_bf16, size_t i4, const float (&v)[4]) {
  if (is_bf16) static_cast<uint2*>(p)[i4] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  else static_cast<float4*>(p)[i4] = make_float4(v[0], v[1], v[2], v[3]);
}

#define WF_GRID_STRIDE(i, n) \
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < (n); i += static_cast<size_t>(gridDim.x) * blockDim.x)

// ------------------------------------------------------------------------------- CFG combine
__global__ void cfg_combine_kernel(const void* c, const void* u, void* out, bool bf, float s, size_t n4) {
  WF_GRID_STRIDE(i, n4) {
    Vec4 a = ld4(c, bf, i), b = ld4(u, bf, i);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      TV vc{a.v[k], bf}, vu{b.v[k], bf};
      o[k] = t_add(vc, t_mul_s(s, t_sub(vc, vu))).v;
    }
    st4(out, bf, i, o);
  }
}

// ------------------------------------------------------------------------------ x0 conversion
__global__ void x0_convert_kernel(const void* x, bool x_bf, const void* v, bool v_bf, void* out, float sigma, size_t n4) {
  const bool o_bf = x_bf && v_bf;
  WF_GRID_STRIDE(i, n4) {
    Vec4 a = ld4(x, x_bf, i), b = ld4(v, v_bf, i);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = t_sub(TV{a.v[k], x_bf}, t_mul_s(sigma, TV{b.v[k], v_bf})).v;
    st4(out, o_bf, i, o);
  }
}

// ----------------------------------------------------------------------------- UniP-bh2 update
struct UnipArgs {
  const void* x; const void* m0; const void* m1; void* out;
  int x_bf, m0_bf, m1_bf, order, rk_is_reciprocal;
  float c_x, c_m0, rk, c_res;
  size_t n4;
};
__global__ void unip_update_kernel(UnipArgs p) {
  WF_GRID_STRIDE(i, p.n4) {
    Vec4 xs = ld4(p.x, p.x_bf, i), a = ld4(p.m0, p.m0_bf, i), b;
    if (p.order == 2) b = ld4(p.m1, p.m1_bf, i);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      TV x{xs.v[k], p.x_bf != 0}, m0{a.v[k], p.m0_bf != 0};
      TV r = t_sub(t_mul_s(p.c_x, x), t_mul_s(p.c_m0, m0));
      if (p.order == 2) {
        // tensor / r_1: torch's CUDA kernel multiplies by the fp32 reciprocal when r_1 is a host scalar and
        // divides when it is a device scalar (the resample tables live on the device)
        TV diff = t_sub(TV{b.v[k], p.m1_bf != 0}, m0);
        TV d1 = p.rk_is_reciprocal ? t_mul_s(p.rk, diff) : t_div_s(diff, p.rk);
        TV pred{0.5f * d1.v, d1.b && x.b};                    // einsum with rhos_p = [0.5] in x.dtype
        r = t_sub(r, t_mul_s(p.c_res, pred));
      }
      o[k] = p.x_bf ? bf16_round(r.v) : r.v;                   // .to(x.dtype)
    }
    st4(p.out, p.x_bf != 0, i, o);
  }
}

// -------------------------------------------------------------------------------- IRR re-noise
// out = (1 - sigma) * x0 + sigma * noise, with sigma held in x0's dtype as a 1-element tensor
__global__ void renoise_kernel(const void* x0, bool x_bf, const float* noise, void* out, float one_minus_sigma, float sigma,
                               size_t n4) {
  WF_GRID_STRIDE(i, n4) {
    Vec4 a = ld4(x0, x_bf, i), nz = ld4(noise, false, i);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      TV l = t_mul(TV{one_minus_sigma, x_bf}, TV{a.v[k], x_bf});
      TV r = t_mul(TV{sigma, x_bf}, TV{nz.v[k], false});
      o[k] = t_add(l, r).v;
    }
    st4(out, false, i, o);
  }
}

// ----------------------------------------------------------------------------------------- DSG
constexpr int DSG_THREADS = 256;
constexpr int DSG_MAX_BLOCKS = 1024;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// partial[k*gridDim + block] for k = {sum g*w, sum g*g, sum w*w}; element products rounded like the tensors
__global__ void __launch_bounds__(DSG_THREADS) dsg_reduce_kernel(const void* g, const void* w, bool bf, size_t n4, float* partial) {
  __shared__ float red[3][DSG_THREADS / 32];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  WF_GRID_STRIDE(i, n4) {
    Vec4 a = ld4(g, bf, i), b = ld4(w, bf, i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      TV x{a.v[k], bf}, y{b.v[k], bf};
      s0 += t_mul(x, y).v; s1 += t_mul(x, x).v; s2 += t_mul(y, y).v;
    }
  }
  s0 = warp_sum_f(s0); s1 = warp_sum_f(s1); s2 = warp_sum_f(s2);
  const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (ln == 0) { red[0][wp] = s0; red[1][wp] = s1; red[2][wp] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int k = 0; k < DSG_THREADS / 32; ++k) t += red[threadIdx.x][k];
    partial[threadIdx.x * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(DSG_THREADS) dsg_apply_kernel(const void* g, const void* w, void* out, bool bf, size_t n4,
                                                                const float* partial, int nparts, float omega, float* stats) {
  __shared__ float tot[3];
  if (threadIdx.x < 96) {
    const int k = threadIdx.x >> 5, ln = threadIdx.x & 31;
    float t = 0.f;
    for (int j = ln; j < nparts; j += 32) t += partial[k * nparts + j];
    t = warp_sum_f(t);
    if (ln == 0) tot[k] = t;
  }
  __syncthreads();
  // scalar chain on 1-element tensors of the noise-pred dtype (pipeline...:669-676)
  TV dot = rnd(tot[0], bf);
  TV ng = rnd(sqrtf(rnd(tot[1], bf).v), bf), nw = rnd(sqrtf(rnd(tot[2], bf).v), bf);
  TV cosv = t_div(dot, rnd(__fadd_rn(t_mul(ng, nw).v, 1e-8f), bf));
  TV ang = rnd(acosf(fminf(fmaxf(cosv.v, -1.0f), 1.0f)), bf);
  TV sinv = rnd(sinf(ang.v), bf);
  TV ratio = t_div(ng, rnd(__fadd_rn(nw.v, 1e-8f), bf));
  TV rc = t_mul(ratio, cosv);
  TV os = t_mul_s(omega, sinv);
  if (stats && blockIdx.x == 0 && threadIdx.x == 0) { stats[0] = cosv.v; stats[1] = sinv.v; stats[2] = ratio.v; }
  WF_GRID_STRIDE(i, n4) {
    Vec4 a = ld4(g, bf, i), b = ld4(w, bf, i);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      TV x{a.v[k], bf}, y{b.v[k], bf};
      o[k] = t_add(x, t_mul(os, t_sub(x, t_mul(rc, y)))).v;
    }
    st4(out, bf, i, o);
  }
}

// ------------------------------------------------------------------------------ FLF pixel blend
// dec, ref: [3, F*H*W]; mask [F*H*W]; all fp32
__global__ void flf_blend_kernel(const float4* __restrict__ dec, const float4* __restrict__ ref, const float4* __restrict__ mask,
                                 float4* __restrict__ out, size_t plane4, int channels) {
  const size_t total = plane4 * channels;
  WF_GRID_STRIDE(i, total) {
    const float4 d = dec[i], r = ref[i], m = mask[i % plane4];
    auto f = [](float dv, float rv, float mv) {
      const float t = __fsub_rn(__fmul_rn(2.0f, rv), 1.0f);
      return __fadd_rn(__fmul_rn(t, mv), __fmul_rn(dv, __fsub_rn(1.0f, mv)));
    };
    out[i] = make_float4(f(d.x, r.x, m.x), f(d.y, r.y, m.y), f(d.z, r.z, m.z), f(d.w, r.w, m.w));
  }
}

// ---------------------------------------------------------------------- refine-pass input upsampling
// generate_refine step 5 (longcat_video/pipeline_longcat_video.py:1393-1419) in one pass over the OUTPUT:
//   uint8 [F,H0,W0,3] -> bf16 -> bilinear (align_corners) to [H,W] -> /255 -> trilinear (align_corners) to F2 frames (the spatial
//   part of that second resize is the identity) -> *2-1, every intermediate rounded to bf16 as the reference's bf16 tensors
//   are, then pad_front copies of the first frame and pad_back copies of the last.  out: planar fp32 [3][F_out][H][W].
// torch's CUDA upsample kernels (UpSampleBilinear2d.cu / UpSampleTrilinear3d.cu) compute in fp32: src = dst*(in-1)/(out-1),
// lambda1 = src - floor(src), val = h0*(w0*a + w1*b) + h1*(w0*c + w1*d).
__device__ __forceinline__ float bilerp_u8(const unsigned char* __restrict__ fr, int H0, int W0, int c, float hr, float wr) {
  const int h1 = static_cast<int>(hr), w1 = static_cast<int>(wr);
  const int hp = h1 < H0 - 1 ? 1 : 0, wp = w1 < W0 - 1 ? 1 : 0;
  const float h1l = hr - h1, h0l = 1.f - h1l, w1l = wr - w1, w0l = 1.f - w1l;
  auto px = [&](int y, int x) { return static_cast<float>(fr[(static_cast<size_t>(y) * W0 + x) * 3 + c]); };
  const float v = h0l * (w0l * px(h1, w1) + w1l * px(h1, w1 + wp)) + h1l * (w0l * px(h1 + hp, w1) + w1l * px(h1 + hp, w1 + wp));
  // bf16 result of the resize, then "/ 255.0" on a bf16 tensor: fp32 multiply by the reciprocal, rounded to bf16
  return bf16_round(__fmul_rn(bf16_round(v), 1.0f / 255.0f));
}

__global__ void refine_upsample_kernel(const unsigned char* __restrict__ video, int F, int H0, int W0, float* __restrict__ out, int F2,
                                       int H, int W, int pad_front, int pad_back) {
  const int F_out = pad_front + F2 + pad_back;
  const size_t total = static_cast<size_t>(3) * F_out * H * W;
  const float rh = H > 1 ? static_cast<float>(H0 - 1) / (H - 1) : 0.f;
  const float rw = W > 1 ? static_cast<float>(W0 - 1) / (W - 1) : 0.f;
  const float rt = F2 > 1 ? static_cast<float>(F - 1) / (F2 - 1) : 0.f;
  const size_t frame = static_cast<size_t>(H0) * W0 * 3;
  WF_GRID_STRIDE(i, total) {
    const int x = static_cast<int>(i % W), y = static_cast<int>((i / W) % H);
    const int fo = static_cast<int>((i / (static_cast<size_t>(W) * H)) % F_out), c = static_cast<int>(i / (static_cast<size_t>(W) * H * F_out));
    const int f2 = min(max(fo - pad_front, 0), F2 - 1);
    const float tr = rt * f2;
    const int t1 = static_cast<int>(tr);
    const int tp = t1 < F - 1 ? 1 : 0;
    const float t1l = tr - t1, t0l = 1.f - t1l;
    const float hr = rh * y, wr = rw * x;
    const float a = bilerp_u8(video + t1 * frame, H0, W0, c, hr, wr);
    float v = a;
    if (F2 != F) {
      const float b = bilerp_u8(video + (t1 + tp) * frame, H0, W0, c, hr, wr);
      v = bf16_round(t0l * a + t1l * b);
    }
    out[i] = bf16_round(__fsub_rn(bf16_round(__fmul_rn(v, 2.0f)), 1.0f));
  }
}

// ---------------------------------------------------------------------- latent (de)normalisation
struct LatentStats { float mean[16]; float inv_std[16]; };   // values already rounded to the latent dtype by the host

// z(fp32) = x0 / inv_std + mean, evaluated in x0's dtype
__global__ void latent_denorm_kernel(const void* x0, bool bf, float* out, LatentStats st, size_t per_channel4, int channels) {
  const size_t total = per_channel4 * channels;
  WF_GRID_STRIDE(i, total) {
    const int c = static_cast<int>(i / per_channel4);
    Vec4 a = ld4(x0, bf, i);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = t_add(t_div(TV{a.v[k], bf}, TV{st.inv_std[c], bf}), TV{st.mean[c], bf}).v;
    st4(out, false, i, o);
  }
}
// out(dtype of x0) = replace[c] ? x0 : (enc - mean) * inv_std
__global__ void latent_norm_replace_kernel(const float* enc, const void* x0, bool bf, void* out, LatentStats st, uint32_t replace_mask,
                                           size_t per_channel4, int channels) {
  const size_t total = per_channel4 * channels;
  WF_GRID_STRIDE(i, total) {
    const int c = static_cast<int>(i / per_channel4);
    float o[4];
    if ((replace_mask >> c) & 1u) {
      Vec4 a = ld4(x0, bf, i);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = a.v[k];
    } else {
      Vec4 e = ld4(enc, false, i);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = __fmul_rn(__fsub_rn(e.v[k], st.mean[c]), st.inv_std[c]);
    }
    st4(out, bf, i, o);
  }
}

// ------------------------------------------------------------------------ uint8 quantisation
__global__ void __launch_bounds__(256) minmax_kernel(const void* x, bool bf, size_t n, float* partial) {
  __shared__ float smin[8], smax[8];
  float lo = INFINITY, hi = -INFINITY;
  WF_GRID_STRIDE(i, n) { const float v = t_load(x, bf, i).v; lo = fminf(lo, v); hi = fmaxf(hi, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { lo = fminf(lo, smin[k]); hi = fmaxf(hi, smax[k]); }
    lo = fminf(lo, smin[0]); hi = fmaxf(hi, smax[0]);
    partial[blockIdx.x] = lo; partial[gridDim.x + blockIdx.x] = hi;
  }
}
__global__ void __launch_bounds__(256) quantise_u8_kernel(const void* x, bool bf, size_t n, const float* partial, int nparts, uint8_t* out, int mode) {
  __shared__ float s_lo, s_rng;
  if (threadIdx.x < 32) {
    float lo = INFINITY, hi = -INFINITY;
    for (int j = threadIdx.x; j < nparts; j += 32) { lo = fminf(lo, partial[j]); hi = fmaxf(hi, partial[nparts + j]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (threadIdx.x == 0) { s_lo = lo; s_rng = __fadd_rn(__fsub_rn(hi, lo), 1e-8f); }
  }
  __syncthreads();
  const float lo = s_lo, rng = s_rng;
  WF_GRID_STRIDE(i, n) {
    const float v = t_load(x, bf, i).v;
    const float nrm = __fdiv_rn(__fsub_rn(v, lo), rng);
    // mode 0 (Wan selector): n*255, truncated.  mode 1 (LongCat selector): the [0,1] value goes through the "[-1,1] video"
    // branch, ((n + 1) * 127.5).clip(0, 255) truncated - i.e. onto the upper half of the uint8 range
    float q = mode == 0 ? __fmul_rn(nrm, 255.0f) : fminf(fmaxf(__fmul_rn(__fadd_rn(nrm, 1.0f), 127.5f), 0.0f), 255.0f);
    out[i] = static_cast<uint8_t>(static_cast<int>(q));
  }
}

// ------------------------------------------------------------------------- CFG-zero (LongCat)
// st* = <c,u> / (|u|^2 + 1e-8);  out = -(u*st + s*(c - u*st))   (pipeline_longcat_video.py:374-383, 875-888), fp32
__global__ void __launch_bounds__(DSG_THREADS) cfgz_reduce_kernel(const float4* c, const float4* u, size_t n4, float* partial) {
  __shared__ float red[2][DSG_THREADS / 32];
  float s0 = 0.f, s1 = 0.f;
  WF_GRID_STRIDE(i, n4) {
    const float4 a = c[i], b = u[i];
    s0 += __fmul_rn(a.x, b.x) + __fmul_rn(a.y, b.y) + __fmul_rn(a.z, b.z) + __fmul_rn(a.w, b.w);
    s1 += __fmul_rn(b.x, b.x) + __fmul_rn(b.y, b.y) + __fmul_rn(b.z, b.z) + __fmul_rn(b.w, b.w);
  }
  s0 = warp_sum_f(s0); s1 = warp_sum_f(s1);
  const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (ln == 0) { red[0][wp] = s0; red[1][wp] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int k = 0; k < DSG_THREADS / 32; ++k) t += red[threadIdx.x][k];
    partial[threadIdx.x * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(DSG_THREADS) cfgz_apply_kernel(const float4* c, const float4* u, float4* out, size_t n4,
                                                                 const float* partial, int nparts, float scale, float* stats) {
  __shared__ float tot[2];
  if (threadIdx.x < 64) {
    const int k = threadIdx.x >> 5, ln = threadIdx.x & 31;
    float t = 0.f;
    for (int j = ln; j < nparts; j += 32) t += partial[k * nparts + j];
    t = warp_sum_f(t);
    if (ln == 0) tot[k] = t;
  }
  __syncthreads();
  const float st = __fdiv_rn(tot[0], __fadd_rn(tot[1], 1e-8f));
  if (stats && blockIdx.x == 0 && threadIdx.x == 0) stats[0] = st;
  auto f = [&](float cv, float uv) {
    const float t1 = __fmul_rn(uv, st);
    return -__fadd_rn(t1, __fmul_rn(scale, __fsub_rn(cv, t1)));
  };
  WF_GRID_STRIDE(i, n4) {
    const float4 a = c[i], b = u[i];
    out[i] = make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
  }
}

static int grid_for(size_t n, int threads = 256, int waves = 8) {
  const size_t want = (n + threads - 1) / threads;
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, static_cast<size_t>(sm_count()) * waves)));
}

}  // namespace wf

using namespace wf;
#define WF_STREAM static_cast<cudaStream_t>(stream)
#define WF_VEC4_OK(n, what) WF_REQUIRE((n) > 0 && (n) % 4 == 0, what ": element count must be a positive multiple of 4")

extern "C" int wf_cfg_combine(const void* cond, const void* uncond, void* out, int is_bf16, float scale, long long n, void* stream) {
  WF_REQUIRE(cond && uncond && out, "wf_cfg_combine: null pointer");
  WF_VEC4_OK(n, "wf_cfg_combine");
  cfg_combine_kernel<<<grid_for(n / 4), 256, 0, WF_STREAM>>>(cond, uncond, out, is_bf16 != 0, scale, static_cast<size_t>(n / 4));
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_x0_convert(const void* sample, int sample_bf16, const void* v, int v_bf16, void* out, float sigma, long long n, void* stream) {
  WF_REQUIRE(sample && v && out, "wf_x0_convert: null pointer");
  WF_VEC4_OK(n, "wf_x0_convert");
  x0_convert_kernel<<<grid_for(n / 4), 256, 0, WF_STREAM>>>(sample, sample_bf16 != 0, v, v_bf16 != 0, out, sigma, static_cast<size_t>(n / 4));
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_unip_update(const void* x, int x_bf16, const void* m0, int m0_bf16, const void* m1, int m1_bf16, void* out, int order,
                              float c_x, float c_m0, float rk, int rk_is_reciprocal, float c_res, long long n, void* stream) {
  WF_REQUIRE(x && m0 && out, "wf_unip_update: null pointer");
  WF_REQUIRE(order == 1 || (order == 2 && m1), "wf_unip_update: order must be 1, or 2 with a previous model output");
  WF_VEC4_OK(n, "wf_unip_update");
  UnipArgs a{x, m0, m1, out, x_bf16, m0_bf16, m1_bf16, order, rk_is_reciprocal, c_x, c_m0, rk, c_res, static_cast<size_t>(n / 4)};
  unip_update_kernel<<<grid_for(n / 4), 256, 0, WF_STREAM>>>(a);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_renoise(const void* x0, int x0_bf16, const float* noise, float* out, float one_minus_sigma, float sigma, long long n, void* stream) {
  WF_REQUIRE(x0 && noise && out, "wf_renoise: null pointer");
  WF_VEC4_OK(n, "wf_renoise");
  renoise_kernel<<<grid_for(n / 4), 256, 0, WF_STREAM>>>(x0, x0_bf16 != 0, noise, out, one_minus_sigma, sigma, static_cast<size_t>(n / 4));
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" long long wf_dsg_workspace_bytes(void) { return 3ll * DSG_MAX_BLOCKS * sizeof(float); }

extern "C" int wf_dsg(const void* g, const void* w, void* out, int is_bf16, float omega, long long n, void* workspace, float* stats, void* stream) {
  WF_REQUIRE(g && w && out && workspace, "wf_dsg: null pointer");
  WF_VEC4_OK(n, "wf_dsg");
  const int blocks = std::min(grid_for(n / 4, DSG_THREADS, 4), DSG_MAX_BLOCKS);
  float* partial = static_cast<float*>(workspace);
  dsg_reduce_kernel<<<blocks, DSG_THREADS, 0, WF_STREAM>>>(g, w, is_bf16 != 0, static_cast<size_t>(n / 4), partial);
  WF_LAUNCH_OK();
  dsg_apply_kernel<<<grid_for(n / 4, DSG_THREADS), DSG_THREADS, 0, WF_STREAM>>>(g, w, out, is_bf16 != 0, static_cast<size_t>(n / 4), partial, blocks, omega, stats);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_flf_blend(const float* decoded, const float* ref, const float* mask, float* out, int channels, long long plane, void* stream) {
  WF_REQUIRE(decoded && ref && mask && out && channels > 0, "wf_flf_blend: bad arguments");
  WF_VEC4_OK(plane, "wf_flf_blend");
  flf_blend_kernel<<<grid_for(plane / 4 * channels), 256, 0, WF_STREAM>>>(reinterpret_cast<const float4*>(decoded), reinterpret_cast<const float4*>(ref),
                                                                         reinterpret_cast<const float4*>(mask), reinterpret_cast<float4*>(out),
                                                                         static_cast<size_t>(plane / 4), channels);
  WF_LAUNCH_OK();
  return WF_OK;
}

static int fill_stats(LatentStats* st, const float* mean, const float* inv_std, int channels) {
  if (channels < 1 || channels > 16) return fail(WF_EINVAL, "latent stats: 1..16 channels");
  for (int c = 0; c < 16; ++c) { st->mean[c] = c < channels ? mean[c] : 0.f; st->inv_std[c] = c < channels ? inv_std[c] : 1.f; }
  return WF_OK;
}

extern "C" int wf_latent_denorm(const void* x0, int is_bf16, float* out, const float* mean, const float* inv_std, int channels, long long per_channel, void* stream) {
  WF_REQUIRE(x0 && out && mean && inv_std, "wf_latent_denorm: null pointer");
  WF_VEC4_OK(per_channel, "wf_latent_denorm");
  LatentStats st; int rc = fill_stats(&st, mean, inv_std, channels); if (rc) return rc;
  latent_denorm_kernel<<<grid_for(per_channel / 4 * channels), 256, 0, WF_STREAM>>>(x0, is_bf16 != 0, out, st, static_cast<size_t>(per_channel / 4), channels);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_latent_norm_replace(const float* enc, const void* x0, int is_bf16, void* out, const float* mean, const float* inv_std,
                                      unsigned replace_mask, int channels, long long per_channel, void* stream) {
  WF_REQUIRE(enc && x0 && out && mean && inv_std, "wf_latent_norm_replace: null pointer");
  WF_VEC4_OK(per_channel, "wf_latent_norm_replace");
  LatentStats st; int rc = fill_stats(&st, mean, inv_std, channels); if (rc) return rc;
  latent_norm_replace_kernel<<<grid_for(per_channel / 4 * channels), 256, 0, WF_STREAM>>>(enc, x0, is_bf16 != 0, out, st, replace_mask,
                                                                                         static_cast<size_t>(per_channel / 4), channels);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" long long wf_quantise_workspace_bytes(void) { return 2ll * 1024 * sizeof(float); }

extern "C" int wf_quantise_u8(const void* x, int is_bf16, unsigned char* out, long long n, int mode, void* workspace, void* stream) {
  WF_REQUIRE(x && out && workspace && n > 0, "wf_quantise_u8: bad arguments");
  const int blocks = std::min(grid_for(n, 256, 4), 1024);
  float* partial = static_cast<float*>(workspace);
  minmax_kernel<<<blocks, 256, 0, WF_STREAM>>>(x, is_bf16 != 0, static_cast<size_t>(n), partial);
  WF_LAUNCH_OK();
  WF_REQUIRE(mode == 0 || mode == 1, "wf_quantise_u8: mode 0 (Wan) or 1 (LongCat)");
  quantise_u8_kernel<<<grid_for(n), 256, 0, WF_STREAM>>>(x, is_bf16 != 0, static_cast<size_t>(n), partial, blocks, out, mode);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_cfg_zero(const float* cond, const float* uncond, float* out, float scale, long long n, void* workspace,
                           float* stats, void* stream) {
  WF_REQUIRE(cond && uncond && out && workspace, "wf_cfg_zero: null pointer");
  WF_VEC4_OK(n, "wf_cfg_zero");
  const int blocks = std::min(grid_for(n / 4, DSG_THREADS, 4), DSG_MAX_BLOCKS);
  float* partial = static_cast<float*>(workspace);
  cfgz_reduce_kernel<<<blocks, DSG_THREADS, 0, WF_STREAM>>>(reinterpret_cast<const float4*>(cond), reinterpret_cast<const float4*>(uncond),
                                                           static_cast<size_t>(n / 4), partial);
  WF_LAUNCH_OK();
  cfgz_apply_kernel<<<grid_for(n / 4, DSG_THREADS), DSG_THREADS, 0, WF_STREAM>>>(reinterpret_cast<const float4*>(cond), reinterpret_cast<const float4*>(uncond),
                                                                               reinterpret_cast<float4*>(out), static_cast<size_t>(n / 4), partial, blocks, scale, stats);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_refine_upsample(const unsigned char* video, int F, int H0, int W0, float* out, int F2, int H, int W, int pad_front,
                                  int pad_back, void* stream) {
  WF_REQUIRE(video && out, "wf_refine_upsample: null pointer");
  WF_REQUIRE(F > 0 && H0 > 0 && W0 > 0 && F2 > 0 && H > 0 && W > 0 && pad_front >= 0 && pad_back >= 0, "wf_refine_upsample: bad sizes");
  const long long total = 3ll * (pad_front + F2 + pad_back) * H * W;
  refine_upsample_kernel<<<grid_for(total), 256, 0, WF_STREAM>>>(video, F, H0, W0, out, F2, H, W, pad_front, pad_back);
  WF_LAUNCH_OK();
  return WF_OK;
}
