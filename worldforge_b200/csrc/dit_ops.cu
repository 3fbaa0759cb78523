// worldforge_b200 - the HBM-bound kernels around the tensor-core GEMMs / attention of the Wan DiT block.
//
// Each kernel is one fused pass over a [tokens, dim] activation; the reference spends ~10
// separate elementwise passes per block on these.  Rounding points follow
// wan/modules/model.py so results match the oracle value for value:
//   layer-norm + modulate   norm1/norm2: LN(x.float()).type_as(x).float()*(1+e_scale)+e_shift  (:303,:311)
//                           norm3:       LN with affine weight/bias                             (:262-264,:310)
//   rms-norm + RoPE         (x.float()*rsqrt(mean(x^2)+eps)).type_as(x) * w                     (:86-89)
//                           complex rotation in float64, rounded to fp32                        (:55-70)
//   patchify                Conv3d(k=s=(1,2,2)) written as im2col rows for the GEMM             (:534-537)
//   head                    LN, modulate, fp32 Linear(dim -> 64), un-patchify                   (:337-347, :600-607)
//   gemv (fp32)             the time-embedding MLPs, batch 1, fp32 weights                      (:546-550)
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < THREADS / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// ------------------------------------------------------------------ layer norm (+ modulate / affine)
constexpr int LN_THREADS = 256;
constexpr int LN_MAXV = 8;   // float4 per thread cached in registers: dim <= 8192

struct LnArgs {
  const void* x; int ldx; int x_is_bf16;
  void* out; int ldo; int out_is_bf16;
  const float* scale;   // modulate: y*(1+scale)+shift ; may be null
  const float* shift;
  const float* weight;  // affine:   y*weight+bias      ; may be null
  const float* bias;
  int D; float eps; int round_norm_bf16;
  int rows_per_group;   // > 0: scale/shift are [groups, D] tables, row r uses group r / rows_per_group (per-frame adaLN)
};

// One row per iteration.  RING = false: one CTA per row, the row is loaded straight into registers.  RING = true: a
// persistent CTA walks rows blockIdx.x, blockIdx.x + gridDim.x, ... and one thread keeps the next NORM_RING - 1 rows in
// flight with cp.async.bulk into a shared-memory ring: a row's CTA spends more than half of its life in the two block
// reductions and the stores, with no loads outstanding, and register pressure allows only four CTAs per SM - too few bytes
// in flight to keep HBM busy (the one-row-per-CTA form measured 3.0 TB/s of 6.4).  Same arithmetic, same element-to-thread
// mapping, same reduction order: bit-identical results.
constexpr int NORM_RING = 3;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool RING>
__global__ void __launch_bounds__(LN_THREADS) layer_norm_kernel(LnArgs p, int rows) {
  extern __shared__ __align__(128) uint8_t norm_ring[];
  __shared__ float red[32];
  __shared__ uint64_t full[NORM_RING];
  const int nvec = p.D >> 2;
  const uint32_t row_bytes = static_cast<uint32_t>(p.D) * (p.x_is_bf16 ? 2 : 4);
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  const size_t src_pitch = static_cast<size_t>(p.ldx) * (p.x_is_bf16 ? 2 : 4);
  auto issue = [&](int i) {
    const long long r = blockIdx.x + static_cast<long long>(i) * gridDim.x;
    if (r < rows) {
      const int sl = i % NORM_RING;
      mbar_arrive_expect_tx(&full[sl], row_bytes);
      bulk_g2s(norm_ring + sl * slot_bytes, static_cast<const uint8_t*>(p.x) + r * src_pitch, row_bytes, &full[sl]);
    }
  };
  if (RING) {
    if (threadIdx.x == 0) {
      for (int sl = 0; sl < NORM_RING; ++sl) mbar_init(&full[sl], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) for (int i = 0; i < NORM_RING - 1; ++i) issue(i);
  }
  for (int it = 0;; ++it) {
    const long long row_ll = blockIdx.x + static_cast<long long>(it) * gridDim.x;
    if (row_ll >= rows) break;
    const int row = static_cast<int>(row_ll);
    const uint8_t* src = static_cast<const uint8_t*>(p.x) + row * src_pitch;
    if (RING) {
      // slot (it - 1) % NORM_RING was read by the previous row, whose block reductions every thread has passed
      if (threadIdx.x == 0) issue(it + NORM_RING - 1);
      mbar_wait(&full[it % NORM_RING], (it / NORM_RING) & 1);
      src = norm_ring + (it % NORM_RING) * slot_bytes;
    }
    const size_t goff = p.rows_per_group > 0 ? static_cast<size_t>(row / p.rows_per_group) * p.D : 0;
    float4 v[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int idx = threadIdx.x + i * LN_THREADS;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < nvec) {
        if (p.x_is_bf16) {
          const uint2 raw = *reinterpret_cast<const uint2*>(src + static_cast<size_t>(idx) * 8);
          float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
          float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
          v[i] = make_float4(a.x, a.y, b.x, b.y);
        } else {
          v[i] = *reinterpret_cast<const float4*>(src + static_cast<size_t>(idx) * 16);
        }
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = block_sum<LN_THREADS>(sum, red) / p.D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int idx = threadIdx.x + i * LN_THREADS;
      if (idx < nvec) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(block_sum<LN_THREADS>(sq, red) / p.D + p.eps);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int idx = threadIdx.x + i * LN_THREADS;
      if (idx < nvec) {
        float y[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
        const int c = idx * 4;
        float4 wv, bv, scv, shv;
        if (p.weight) { wv = *reinterpret_cast<const float4*>(p.weight + c); bv = *reinterpret_cast<const float4*>(p.bias + c); }
        if (p.scale) { scv = *reinterpret_cast<const float4*>(p.scale + goff + c); shv = *reinterpret_cast<const float4*>(p.shift + goff + c); }
        const float w4[4] = {wv.x, wv.y, wv.z, wv.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
        const float sc4[4] = {scv.x, scv.y, scv.z, scv.w}, sh4[4] = {shv.x, shv.y, shv.z, shv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (p.weight) y[u] = y[u] * w4[u] + b4[u];
          if (p.round_norm_bf16) y[u] = bf16_round(y[u]);
          if (p.scale) y[u] = __fadd_rn(__fmul_rn(y[u], __fadd_rn(1.0f, sc4[u])), sh4[u]);
        }
        if (p.out_is_bf16) {
          uint2 o = make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
          *reinterpret_cast<uint2*>(static_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + c) = o;
        } else {
          *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + c) = make_float4(y[0], y[1], y[2], y[3]);
        }
      }
    }
    if (!RING) break;
  }
}

// --------------------------------------------------------------------------- rms norm (+ RoPE)
constexpr int RMS_THREADS = 256;
constexpr int RMS_MAXV = 4;   // 8 bf16 per vector: dim <= 8192

struct RmsArgs {
  bf16* x; int ldx;             // in place: [rows, D] slice of a wider matrix
  const float* weight;          // [D]
  const double* rope;           // [rows, 64, 2] (cos, sin) per token and complex pair, or null
  int D; float eps;
};

// RING: persistent CTAs with the shared-memory row ring of layer_norm_kernel (see there); results are bit-identical.
template <bool RING>
__global__ void __launch_bounds__(RMS_THREADS) rms_norm_rope_kernel(RmsArgs p, int rows) {
  extern __shared__ __align__(128) uint8_t norm_ring[];
  __shared__ float red[32];
  __shared__ uint64_t full[NORM_RING];
  const int nvec = p.D >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(p.D) * 2;
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  auto issue = [&](int i) {
    const long long r = blockIdx.x + static_cast<long long>(i) * gridDim.x;
    if (r < rows) {
      const int sl = i % NORM_RING;
      mbar_arrive_expect_tx(&full[sl], row_bytes);
      bulk_g2s(norm_ring + sl * slot_bytes, p.x + r * p.ldx, row_bytes, &full[sl]);
    }
  };
  if (RING) {
    if (threadIdx.x == 0) {
      for (int sl = 0; sl < NORM_RING; ++sl) mbar_init(&full[sl], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) for (int i = 0; i < NORM_RING - 1; ++i) issue(i);
  }
  for (int it = 0;; ++it) {
    const long long row_ll = blockIdx.x + static_cast<long long>(it) * gridDim.x;
    if (row_ll >= rows) break;
    const int row = static_cast<int>(row_ll);
    bf16* xr = p.x + static_cast<size_t>(row) * p.ldx;
    const bf16* src = xr;
    if (RING) {
      if (threadIdx.x == 0) issue(it + NORM_RING - 1);
      mbar_wait(&full[it % NORM_RING], (it / NORM_RING) & 1);
      src = reinterpret_cast<const bf16*>(norm_ring + (it % NORM_RING) * slot_bytes);
    }
    uint4 raw[RMS_MAXV];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < RMS_MAXV; ++i) {
      const int idx = threadIdx.x + i * RMS_THREADS;
      if (idx < nvec) {
        raw[i] = *reinterpret_cast<const uint4*>(src + idx * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 f = __bfloat1622float2(h[u]);
          sq += f.x * f.x + f.y * f.y;
        }
      }
    }
    const float rstd = rsqrtf(block_sum<RMS_THREADS>(sq, red) / p.D + p.eps);
#pragma unroll
    for (int i = 0; i < RMS_MAXV; ++i) {
      const int idx = threadIdx.x + i * RMS_THREADS;
      if (idx < nvec) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
        const int c = idx * 8;
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 f = __bfloat1622float2(h[u]);
          // (x.float() * rstd).type_as(x) * weight
          float a = __fmul_rn(bf16_round(__fmul_rn(f.x, rstd)), p.weight[c + 2 * u]);
          float b = __fmul_rn(bf16_round(__fmul_rn(f.y, rstd)), p.weight[c + 2 * u + 1]);
          if (p.rope) {
            const int pair = ((c + 2 * u) & 127) >> 1;                 // complex pair index inside the head
            const double2 cs = *reinterpret_cast<const double2*>(p.rope + (static_cast<size_t>(row) * 64 + pair) * 2);
            const double da = a, db = b;
            const float re = static_cast<float>(da * cs.x - db * cs.y);
            const float im = static_cast<float>(da * cs.y + db * cs.x);
            a = re; b = im;
          }
          o[u] = pack_bf16x2(a, b);
        }
        *reinterpret_cast<uint4*>(xr + c) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (!RING) break;
  }
}


// ---------------------------------------------------------- q|k|v: RMSNorm + RoPE + scatter to the peers (Ulysses)
// Same per-element arithmetic as rms_norm_rope_kernel (the results are bit-identical); instead of writing in place the
// 8-element vectors go to the rank that owns their head: blockIdx.y = 0 (q), 1 (k), 2 (v, copied).
constexpr int QS_MAX_PEERS = 8;

struct QkvScatterArgs {
  const bf16* qkv; int ldx;
  const float* weight_q; const float* weight_k;
  const double* rope;
  int D; float eps;
  bf16* dst[QS_MAX_PEERS]; int n_peers; int ld_dst; int row0;
};

__global__ void __launch_bounds__(RMS_THREADS) qkv_norm_rope_scatter_kernel(QkvScatterArgs p) {
  __shared__ float red[32];
  const int row = blockIdx.x, which = blockIdx.y;
  const int nvec = p.D >> 3;
  const bf16* xr = p.qkv + static_cast<size_t>(row) * p.ldx + static_cast<size_t>(which) * p.D;
  const float* weight = which == 0 ? p.weight_q : p.weight_k;
  const int hpp = (p.D >> 7) / p.n_peers;                        // heads per peer
  uint4 raw[RMS_MAXV];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < RMS_MAXV; ++i) {
    const int idx = threadIdx.x + i * RMS_THREADS;
    if (idx < nvec) {
      raw[i] = *reinterpret_cast<const uint4*>(xr + idx * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float2 f = __bfloat1622float2(h[u]);
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  float rstd = 0.f;
  if (which < 2) rstd = rsqrtf(block_sum<RMS_THREADS>(sq, red) / p.D + p.eps);
#pragma unroll
  for (int i = 0; i < RMS_MAXV; ++i) {
    const int idx = threadIdx.x + i * RMS_THREADS;
    if (idx < nvec) {
      const int c = idx * 8;
      uint4 outv = raw[i];
      if (which < 2) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 f = __bfloat1622float2(h[u]);
          float a = __fmul_rn(bf16_round(__fmul_rn(f.x, rstd)), weight[c + 2 * u]);
          float b = __fmul_rn(bf16_round(__fmul_rn(f.y, rstd)), weight[c + 2 * u + 1]);
          const int pair = ((c + 2 * u) & 127) >> 1;
          const double2 cs = *reinterpret_cast<const double2*>(p.rope + (static_cast<size_t>(row) * 64 + pair) * 2);
          const double da = a, db = b;
          const float re = static_cast<float>(da * cs.x - db * cs.y);
          const float im = static_cast<float>(da * cs.y + db * cs.x);
          o[u] = pack_bf16x2(re, im);
        }
        outv = make_uint4(o[0], o[1], o[2], o[3]);
      }
      const int head = c >> 7;
      const int peer = head / hpp;
      bf16* d = p.dst[peer] + static_cast<size_t>(p.row0 + row) * p.ld_dst + (which * hpp + head % hpp) * 128 + (c & 127);
      *reinterpret_cast<uint4*>(d) = outv;
    }
  }
}

// ----------------------------------------------------------------------------------- patchify
// hidden [C, F, H, W] bf16 -> rows [F*(H/2)*(W/2), C*4] bf16, column = c*4 + ph*2 + pw
__global__ void patchify_kernel(const bf16* __restrict__ x, bf16* __restrict__ cols, int C, int F, int H, int W) {
  const int gh = H >> 1, gw = W >> 1;
  const size_t total = static_cast<size_t>(F) * gh * gw * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t tok = i / C;
    const int w = static_cast<int>(tok % gw), h = static_cast<int>((tok / gw) % gh), f = static_cast<int>(tok / (static_cast<size_t>(gw) * gh));
    const bf16* src = x + ((static_cast<size_t>(c) * F + f) * H + 2 * h) * W + 2 * w;
    const __nv_bfloat162 top = *reinterpret_cast<const __nv_bfloat162*>(src);
    const __nv_bfloat162 bot = *reinterpret_cast<const __nv_bfloat162*>(src + W);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&top);
    o.y = *reinterpret_cast<const uint32_t*>(&bot);
    *reinterpret_cast<uint2*>(cols + tok * (static_cast<size_t>(C) * 4) + c * 4) = o;
  }
}

// --------------------------------------------------------------------------------------- head
// out[c, f, 2h+ph, 2w+pw] = sum_k W[(ph*2+pw)*Cout + c, k] * mod(LN(x[tok,:]))[k] + b
constexpr int HEAD_THREADS = 256;
constexpr int HEAD_TOK = 4;

struct HeadArgs {
  const void* x; int x_is_bf16; int ldx; int L; int D;
  const float* scale; const float* shift;   // [D] (or [groups, D])
  const float* w; const float* b;           // [NO, D], [NO]
  int NO;                                   // 4 * Cout
  float* out; int Cout, F, GH, GW;          // out [Cout, F, 2*GH, 2*GW]
  float eps;
  int tok_offset;                           // global index of x's first row (sequence-parallel shards)
  int rows_per_group;                       // > 0: scale/shift are [groups, D] tables (per-frame modulation)
  int round_bf16;                           // 1: the modulated row is rounded to bf16 before the fp32 projection
};

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(HeadArgs p) {
  extern __shared__ float sh[];             // HEAD_TOK * D normalised + modulated rows
  __shared__ float red[32];
  const int tok0 = blockIdx.x * HEAD_TOK;
  for (int t = 0; t < HEAD_TOK; ++t) {
    const int tok = tok0 + t;
    float* row = sh + t * p.D;
    if (tok >= p.L) { for (int i = threadIdx.x; i < p.D; i += HEAD_THREADS) row[i] = 0.f; continue; }
    float s = 0.f;
    for (int i = threadIdx.x; i < p.D; i += HEAD_THREADS) {
      const size_t at = static_cast<size_t>(tok) * p.ldx + i;
      float v = p.x_is_bf16 ? __bfloat162float(static_cast<const bf16*>(p.x)[at]) : static_cast<const float*>(p.x)[at];
      row[i] = v; s += v;
    }
    const float mean = block_sum<HEAD_THREADS>(s, red) / p.D;
    float sq = 0.f;
    for (int i = threadIdx.x; i < p.D; i += HEAD_THREADS) { float d = row[i] - mean; sq += d * d; }
    const float rstd = rsqrtf(block_sum<HEAD_THREADS>(sq, red) / p.D + p.eps);
    const size_t goff = p.rows_per_group > 0 ? static_cast<size_t>((tok + p.tok_offset) / p.rows_per_group) * p.D : 0;
    for (int i = threadIdx.x; i < p.D; i += HEAD_THREADS) {
      float y = (row[i] - mean) * rstd;
      y = __fadd_rn(__fmul_rn(y, __fadd_rn(1.0f, p.scale[goff + i])), p.shift[goff + i]);
      row[i] = p.round_bf16 ? bf16_round(y) : y;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < p.NO; o += HEAD_THREADS / 32) {
    const float* wr = p.w + static_cast<size_t>(o) * p.D;
    float acc[HEAD_TOK];
#pragma unroll
    for (int t = 0; t < HEAD_TOK; ++t) acc[t] = 0.f;
    for (int k = lane * 4; k < p.D; k += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
      for (int t = 0; t < HEAD_TOK; ++t) {
        const float4 xv = *reinterpret_cast<const float4*>(sh + t * p.D + k);
        acc[t] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
      }
    }
#pragma unroll
    for (int t = 0; t < HEAD_TOK; ++t) {
      float v = warp_sum(acc[t]);
      const int tok = tok0 + t;
      if (lane == 0 && tok < p.L) {
        v += p.b[o];
        const int c = o % p.Cout, pq = o / p.Cout, ph = pq >> 1, pw = pq & 1;
        const int gt = tok + p.tok_offset;
        const int gw = gt % p.GW, gh = (gt / p.GW) % p.GH, f = gt / (p.GW * p.GH);
        p.out[((static_cast<size_t>(c) * p.F + f) * (2 * p.GH) + 2 * gh + ph) * (2 * p.GW) + 2 * gw + pw] = v;
      }
    }
  }
}

// --------------------------------------------------------------------------------- fp32 GEMV
// out[n] = act_out( sum_k W[n,k] * act_in(x[k]) + b[n] );  act: 0 none, 1 SiLU
__global__ void __launch_bounds__(256) gemv_f32_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                       const float* __restrict__ b, float* __restrict__ out, int N,
                                                       int K, int silu_in, int silu_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* wr = w + static_cast<size_t>(warp) * K;
  float acc = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 wv = *reinterpret_cast<const float4*>(wr + k);
    float4 xv = *reinterpret_cast<const float4*>(x + k);
    if (silu_in) {
      xv.x = xv.x / (1.0f + expf(-xv.x)); xv.y = xv.y / (1.0f + expf(-xv.y));
      xv.z = xv.z / (1.0f + expf(-xv.z)); xv.w = xv.w / (1.0f + expf(-xv.w));
    }
    acc += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (b ? b[warp] : 0.f);
    if (silu_out) v = v / (1.0f + expf(-v));
    out[warp] = v;
  }
}

// ------------------------------------------------------------- bf16 elementwise GELU (erf form)
__global__ void gelu_erf_bf16_kernel(bf16* x, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = __bfloat162float(x[i]);
    x[i] = __float2bfloat16_rn(0.5f * v * (1.0f + erff(v * 0.7071067811865476f)));
  }
}


// ------------------------------------------------- sinusoidal timestep embedding (model.py:18-28)
// out[0:half] = cos(t * 10000^(-i/half)), out[half:] = sin(...), evaluated in float64 like the reference
__global__ void time_sinusoid_kernel(const long long* __restrict__ t, float* __restrict__ out, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const double pos = static_cast<double>(t[0]);
  const double ang = pos * pow(10000.0, -static_cast<double>(i) / half);
  out[i] = static_cast<float>(cos(ang));
  out[half + i] = static_cast<float>(sin(ang));
}

// out[r, :] = a[r, :] + b[:]  (modulation tables + time projection, model.py:298,345)
__global__ void add_bcast_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                     size_t rows, int inner) {
  const size_t total = rows * inner;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __fadd_rn(a[i], b[i % inner]);
}


// ------------------------------------------------ per-head RMSNorm (+ fp32 RoPE): LongCat's q_norm / k_norm + rope_3d
// x: bf16 [rows, heads*128] slice (leading dimension ldx); gain bf16 [128]; rope fp32 [rows, 64, 2] (cos, sin) or null.
// RMSNorm_FP32 (blocks.py:46-52): (x.float()*rsqrt(mean(x^2)+eps)).type_as(x) * weight  -> bf16 (bf16 gain);
// RotaryPositionalEmbedding.forward (rope_3d.py:99-119): q*cos + rotate_half(q)*sin in fp32, back to bf16.
__global__ void __launch_bounds__(256) rms_head_rope_kernel(bf16* x, int ldx, const bf16* __restrict__ gain,
                                                            const float* __restrict__ rope, size_t rows, int heads, float eps) {
  const int lane = threadIdx.x & 31;
  const size_t unit = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;      // (row, head)
  if (unit >= rows * heads) return;
  const size_t row = unit / heads;
  const int head = static_cast<int>(unit % heads);
  bf16* px = x + row * ldx + head * 128 + lane * 4;
  const uint2 raw = *reinterpret_cast<const uint2*>(px);
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  float ss = a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss * (1.0f / 128.0f) + eps);
  const uint2 graw = *reinterpret_cast<const uint2*>(gain + lane * 4);
  float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&graw.x));
  float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&graw.y));
  float v[4] = {bf16_round(__fmul_rn(bf16_round(__fmul_rn(a.x, rstd)), g0.x)), bf16_round(__fmul_rn(bf16_round(__fmul_rn(a.y, rstd)), g0.y)),
                bf16_round(__fmul_rn(bf16_round(__fmul_rn(b.x, rstd)), g1.x)), bf16_round(__fmul_rn(bf16_round(__fmul_rn(b.y, rstd)), g1.y))};
  if (rope) {
    const float4 cs = *reinterpret_cast<const float4*>(rope + (row * 64 + lane * 2) * 2);   // (cos0, sin0, cos1, sin1)
    const float r0 = __fadd_rn(__fmul_rn(v[0], cs.x), __fmul_rn(-v[1], cs.y));
    const float i0 = __fadd_rn(__fmul_rn(v[1], cs.x), __fmul_rn(v[0], cs.y));
    const float r1 = __fadd_rn(__fmul_rn(v[2], cs.z), __fmul_rn(-v[3], cs.w));
    const float i1 = __fadd_rn(__fmul_rn(v[3], cs.z), __fmul_rn(v[2], cs.w));
    v[0] = r0; v[1] = i0; v[2] = r1; v[3] = i1;
  }
  *reinterpret_cast<uint2*>(px) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// ------------------------------------------------------------------- SwiGLU gate: silu(w1 x) * (w3 x), bf16 (blocks.py:39)
// in bf16 [rows, 2F] = [w1 x | w3 x]; out bf16 [rows, F]
__global__ void swiglu_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, size_t rows, int F) {
  const int f8 = F >> 3;
  const size_t total = rows * f8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / f8; const int c = static_cast<int>(i % f8) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(in + r * 2 * F + c);
    const uint4 b = *reinterpret_cast<const uint4*>(in + r * 2 * F + F + c);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
    uint32_t o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 x = __bfloat1622float2(a2[u]), y = __bfloat1622float2(b2[u]);
      const float s0 = bf16_round(x.x / (1.0f + expf(-x.x))), s1 = bf16_round(x.y / (1.0f + expf(-x.y)));
      o[u] = pack_bf16x2(__fmul_rn(s0, y.x), __fmul_rn(s1, y.y));
    }
    *reinterpret_cast<uint4*>(out + r * F + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------- TimestepEmbedder.timestep_embedding (blocks.py:184-191), fp32 like the reference
__global__ void timestep_embedding_f32_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, k = i % half;
  const float freq = expf(__fmul_rn(-9.210340371976184f, static_cast<float>(k)) / static_cast<float>(half));
  const float arg = __fmul_rn(t[r], freq);
  out[r * 2 * half + k] = cosf(arg);
  out[r * 2 * half + half + k] = sinf(arg);
}

// ----------------------------- out[T, R] = act_in(x[T, K]) . W[R, K]^T + b : fp32 math on bf16-valued weights (the adaLN tables)
// one warp per output row r, all T <= 32 input rows staged in shared memory
__global__ void __launch_bounds__(256) small_gemm_f32_kernel(const float* __restrict__ x, const bf16* __restrict__ w,
                                                             const bf16* __restrict__ b, float* __restrict__ out, int T, int R,
                                                             int K, int silu_in) {
  extern __shared__ float xs[];
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) {
    float v = x[i];
    if (silu_in) v = v / (1.0f + expf(-v));
    xs[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const bf16* wr = w + static_cast<size_t>(r) * K;
  float acc[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) acc[t] = 0.f;
  for (int k = lane * 2; k < K; k += 64) {
    const float2 wv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wr + k));
#pragma unroll
    for (int t = 0; t < 32; ++t)
      if (t < T) acc[t] += wv.x * xs[t * K + k] + wv.y * xs[t * K + k + 1];
  }
  const float bias = b ? __bfloat162float(b[r]) : 0.f;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if (t < T) {
      const float v = warp_sum(acc[t]);
      if (lane == 0) out[static_cast<size_t>(t) * R + r] = v + bias;
    }
  }
}

}  // namespace wf

using namespace wf;

extern "C" int wf_layer_norm(const void* x, int ldx, int x_is_bf16, void* out, int ldo, int out_is_bf16,
                             const float* scale, const float* shift, const float* weight, const float* bias,
                             int rows, int D, float eps, int round_norm_bf16, int rows_per_group, void* stream) {
  WF_REQUIRE(x && out && rows > 0, "wf_layer_norm: bad arguments");
  WF_REQUIRE(D % 4 == 0 && D <= 4 * LN_THREADS * LN_MAXV, "wf_layer_norm: dim must be a multiple of 4 and <= 8192");
  WF_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "wf_layer_norm: leading dimensions must be multiples of 4");
  WF_REQUIRE((scale == nullptr) == (shift == nullptr) && (weight == nullptr) == (bias == nullptr),
             "wf_layer_norm: scale/shift and weight/bias come in pairs");
  WF_REQUIRE(rows_per_group >= 0, "wf_layer_norm: rows_per_group must be >= 0");
  LnArgs a{x, ldx, x_is_bf16, out, ldo, out_is_bf16, scale, shift, weight, bias, D, eps, round_norm_bf16, rows_per_group};
  const auto al16 = [](const void* q) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  WF_REQUIRE(al16(scale) && al16(shift) && al16(weight) && al16(bias), "wf_layer_norm: scale / shift / weight / bias must be 16-byte aligned");
  const int esize = x_is_bf16 ? 2 : 4;
  const int slot = (D * esize + 127) & ~127;
  // many rows: persistent CTAs with a shared-memory row ring (cp.async.bulk needs 16-byte aligned rows)
  if (rows >= 4 * sm_count() && al16(x) && (static_cast<long long>(ldx) * esize) % 16 == 0 && (D * esize) % 16 == 0) {
    static PerDeviceOnce once;
    const int rc1 = once.run([] {
      WF_CUDA_OK(cudaFuncSetAttribute(layer_norm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NORM_RING * 32768));
      return static_cast<int>(WF_OK);
    });
    if (rc1) return rc1;
    const int grid = std::min(rows, 3 * sm_count());
    layer_norm_kernel<true><<<grid, LN_THREADS, NORM_RING * slot, static_cast<cudaStream_t>(stream)>>>(a, rows);
  } else {
    layer_norm_kernel<false><<<rows, LN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, rows);
  }
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_rms_norm_rope(void* x, int ldx, const float* weight, const double* rope, int rows, int D, float eps,
                                void* stream) {
  WF_REQUIRE(x && weight && rows > 0, "wf_rms_norm_rope: bad arguments");
  WF_REQUIRE(D % 8 == 0 && D <= 8 * RMS_THREADS * RMS_MAXV && ldx % 8 == 0, "wf_rms_norm_rope: dim must be a multiple of 8 and <= 8192");
  WF_REQUIRE(rope == nullptr || D % 128 == 0, "wf_rms_norm_rope: RoPE needs head_dim 128");
  RmsArgs a{static_cast<bf16*>(x), ldx, weight, rope, D, eps};
  if (rows >= 4 * sm_count() && reinterpret_cast<uintptr_t>(x) % 16 == 0 && D % 8 == 0) {
    static PerDeviceOnce once;
    const int rc1 = once.run([] {
      WF_CUDA_OK(cudaFuncSetAttribute(rms_norm_rope_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NORM_RING * 16384));
      return static_cast<int>(WF_OK);
    });
    if (rc1) return rc1;
    const int slot = (D * 2 + 127) & ~127;
    const int grid = std::min(rows, 4 * sm_count());
    rms_norm_rope_kernel<true><<<grid, RMS_THREADS, NORM_RING * slot, static_cast<cudaStream_t>(stream)>>>(a, rows);
  } else {
    rms_norm_rope_kernel<false><<<rows, RMS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, rows);
  }
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_qkv_norm_rope_scatter(const void* qkv, int ldx, const float* weight_q, const float* weight_k,
                                        const double* rope, int rows, int D, float eps, void* const* dst_peers, int n_peers,
                                        int ld_dst, int row0, void* stream) {
  WF_REQUIRE(qkv && weight_q && weight_k && rope && dst_peers && rows > 0, "wf_qkv_norm_rope_scatter: bad arguments");
  WF_REQUIRE(D % 128 == 0 && D <= 8 * RMS_THREADS * RMS_MAXV && ldx % 8 == 0 && ldx >= 3 * D,
             "wf_qkv_norm_rope_scatter: dim must be a multiple of 128 and <= 8192, qkv rows 3*D wide");
  WF_REQUIRE(n_peers >= 1 && n_peers <= QS_MAX_PEERS && (D / 128) % n_peers == 0, "wf_qkv_norm_rope_scatter: heads must split over the peers");
  WF_REQUIRE(ld_dst % 8 == 0 && ld_dst >= 3 * D / n_peers && row0 >= 0, "wf_qkv_norm_rope_scatter: bad destination layout");
  QkvScatterArgs a{};
  a.qkv = static_cast<const bf16*>(qkv); a.ldx = ldx; a.weight_q = weight_q; a.weight_k = weight_k; a.rope = rope;
  a.D = D; a.eps = eps; a.n_peers = n_peers; a.ld_dst = ld_dst; a.row0 = row0;
  for (int i = 0; i < n_peers; ++i) {
    WF_REQUIRE(dst_peers[i] != nullptr, "wf_qkv_norm_rope_scatter: null peer pointer");
    a.dst[i] = static_cast<bf16*>(dst_peers[i]);
  }
  qkv_norm_rope_scatter_kernel<<<dim3(rows, 3), RMS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_patchify(const void* hidden, void* cols, int C, int F, int H, int W, void* stream) {
  WF_REQUIRE(hidden && cols && C > 0 && F > 0, "wf_patchify: bad arguments");
  WF_REQUIRE(H % 2 == 0 && W % 2 == 0, "wf_patchify: H and W must be even (patch 1x2x2)");
  const size_t total = static_cast<size_t>(F) * (H / 2) * (W / 2) * C;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(sm_count()) * 16));
  patchify_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(hidden),
                                                                         static_cast<bf16*>(cols), C, F, H, W);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_dit_head(const void* x, int x_is_bf16, int ldx, int L, int D, const float* scale, const float* shift,
                           const float* w, const float* b, int Cout, float* out, int F, int GH, int GW, float eps,
                           int tok_offset, int rows_per_group, int round_bf16, void* stream) {
  WF_REQUIRE(x && scale && shift && w && b && out, "wf_dit_head: null pointer");
  WF_REQUIRE(tok_offset >= 0 && tok_offset + L <= F * GH * GW, "wf_dit_head: token range does not fit the grid");
  WF_REQUIRE(D % 128 == 0 && ldx % 4 == 0, "wf_dit_head: dim must be a multiple of 128");
  HeadArgs a{x, x_is_bf16, ldx, L, D, scale, shift, w, b, 4 * Cout, out, Cout, F, GH, GW, eps, tok_offset, rows_per_group, round_bf16};
  const size_t shmem = static_cast<size_t>(HEAD_TOK) * D * sizeof(float);
  static size_t configured = 0;
  if (shmem > configured) {
    WF_CUDA_OK(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shmem)));
    configured = shmem;
  }
  head_kernel<<<(L + HEAD_TOK - 1) / HEAD_TOK, HEAD_THREADS, shmem, static_cast<cudaStream_t>(stream)>>>(a);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_gemv_f32(const float* w, const float* x, const float* b, float* out, int N, int K, int silu_in,
                           int silu_out, void* stream) {
  WF_REQUIRE(w && x && out && N > 0, "wf_gemv_f32: bad arguments");
  WF_REQUIRE(K % 4 == 0, "wf_gemv_f32: K must be a multiple of 4");
  const int blocks = (N * 32 + 255) / 256;
  gemv_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, x, b, out, N, K, silu_in, silu_out);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_gelu_erf_bf16(void* x, long long n, void* stream) {
  WF_REQUIRE(x && n > 0, "wf_gelu_erf_bf16: bad arguments");
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(sm_count()) * 8));
  gelu_erf_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<bf16*>(x), static_cast<size_t>(n));
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_time_sinusoid(const long long* timestep, float* out, int freq_dim, void* stream) {
  WF_REQUIRE(timestep && out && freq_dim > 0 && freq_dim % 2 == 0, "wf_time_sinusoid: bad arguments");
  const int half = freq_dim / 2;
  time_sinusoid_kernel<<<(half + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(timestep, out, half);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_add_bcast_f32(const float* a, const float* b, float* out, long long rows, int inner, void* stream) {
  WF_REQUIRE(a && b && out && rows > 0 && inner > 0, "wf_add_bcast_f32: bad arguments");
  const long long total = rows * inner;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sm_count()) * 8));
  add_bcast_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, static_cast<size_t>(rows), inner);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_rms_norm_head_rope(void* x, int ldx, const void* gain_bf16, const float* rope, long long rows, int heads,
                                     float eps, void* stream) {
  WF_REQUIRE(x && gain_bf16 && rows > 0 && heads > 0 && ldx % 4 == 0, "wf_rms_norm_head_rope: bad arguments");
  const long long threads = rows * heads * 32;
  rms_head_rope_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(x), ldx, static_cast<const bf16*>(gain_bf16), rope, static_cast<size_t>(rows), heads, eps);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_swiglu_bf16(const void* in, void* out, long long rows, int F, void* stream) {
  WF_REQUIRE(in && out && rows > 0 && F > 0 && F % 8 == 0, "wf_swiglu_bf16: bad arguments");
  const long long total = rows * (F / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sm_count()) * 16));
  swiglu_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(in), static_cast<bf16*>(out),
                                                                           static_cast<size_t>(rows), F);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_timestep_embedding_f32(const float* t, float* out, int n, int dim, void* stream) {
  WF_REQUIRE(t && out && n > 0 && dim > 0 && dim % 2 == 0, "wf_timestep_embedding_f32: bad arguments");
  const int total = n * (dim / 2);
  timestep_embedding_f32_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, out, n, dim / 2);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_small_gemm_f32(const float* x, const void* w_bf16, const void* b_bf16, float* out, int T, int R, int K,
                                 int silu_in, void* stream) {
  WF_REQUIRE(x && w_bf16 && out && T > 0 && T <= 32 && R > 0 && K > 0 && K % 2 == 0, "wf_small_gemm_f32: 1..32 rows, even K");
  const size_t shmem = static_cast<size_t>(T) * K * sizeof(float);
  WF_REQUIRE(shmem <= 200 * 1024, "wf_small_gemm_f32: T*K too large for shared memory");
  static size_t configured = 0;
  if (shmem > configured) {
    WF_CUDA_OK(cudaFuncSetAttribute(small_gemm_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shmem)));
    configured = shmem;
  }
  const int blocks = (R * 32 + 255) / 256;
  small_gemm_f32_kernel<<<blocks, 256, shmem, static_cast<cudaStream_t>(stream)>>>(x, static_cast<const bf16*>(w_bf16),
                                                                                  static_cast<const bf16*>(b_bf16), out, T, R, K, silu_in);
  WF_LAUNCH_OK();
  return WF_OK;
}
