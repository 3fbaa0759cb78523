// worldforge_b200 - convolutions of the Wan 3D-VAE as implicit GEMMs on tcgen05 (kind::tf32).
//
// Every convolution of the VAE (reference wan/modules/vae.py: CausalConv3d :17-36, the Resample 2-D convs
// :76-96, the 1x1 convs of AttentionBlock :234-235 and WanVAE_.conv1/conv2 :505-506) is
//
//   out[t, y, x, n] = bias[n] + sum_{tap} sum_{c} W_tap[n, c] * in[t*st + dt_tap, y + dy_tap, x + dx_tap, c]
//
// over channels-last fp32 activations in[T][H][W][C].  The A operand of tap `tap` for a 128-pixel output tile
// (TH x TW pixels of one frame) is therefore just the input tile shifted by (dt, dy, dx): one 4-D TMA box per
// (tap, 32-channel chunk), with TMA's out-of-bounds zero fill providing the spatial zero padding AND the causal
// zero history in time.  B is the matching 32-channel slice of the tap's [Cout, Cin] weight matrix.  Operands
// stay fp32 in HBM and shared memory and are consumed as tf32 - the precision cuDNN uses for the reference's
// fp32 VAE on a GPU (torch.backends.cudnn.allow_tf32 defaults to True) - with fp32 accumulation in TMEM.
//
// The same kernel covers, by parameters only:
//   * 3x3x3 causal convs (27 taps, dt in {-2,-1,0}), 3x1x1 temporal convs (stride 1 or 2 in time), 1x1x1 convs
//     and plain GEMMs (attention scores / PV of the mid block);
//   * nearest-exact 2x upsample + 3x3 conv as four 2x2-tap convs on the LOW-resolution input, one per output
//     parity, with pre-summed weights (no upsampled tensor is ever materialised; 2.25x fewer FLOPs);
//   * pad(0,1,0,1) + 3x3 stride-2 conv as a 2x2-tap conv on the space-to-depth input;
//   * the epilogue options the network needs: bias, residual add (ResidualBlock's x + h, :220; AttentionBlock's
//     x + identity, :262), the channel->frame de-interleave of upsample3d (:134-137), strided/offset output
//     placement for the parity convs, and the final planar [3,F,H,W] store with clamp(-1,1).
//
// Structure is the GEMM's: persistent CTAs, warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (smem transpose -> 128-byte coalesced row stores), 4-stage mbarrier ring, two TMEM
// accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

constexpr int CV_BM = 128;          // output pixels per tile
constexpr int CV_BK = 32;           // fp32 channels per k-block = one 128-byte swizzle row
constexpr int CV_MAX_BN = 192;
constexpr int CV_STAGES = 4;
constexpr int CV_A_BYTES = CV_BM * CV_BK * 4;           // 16 KB
constexpr int CV_B_BYTES = CV_MAX_BN * CV_BK * 4;       // 24 KB (slot size; BN*128 bytes are used)
constexpr int CV_STAGE_BYTES = CV_A_BYTES + CV_B_BYTES;
constexpr int CV_EPI_LD = 33;
constexpr int CV_EPI_BYTES = 4 * 32 * CV_EPI_LD * 4;
constexpr int CV_SMEM = CV_STAGES * CV_STAGE_BYTES + 1024 + 256 + CV_EPI_BYTES;
constexpr int CV_THREADS = 256;
constexpr int CV_MAX_TAPS = 27;

struct ConvArgs {
  // iteration space of the output tiles
  int T, H, W;              // output frames / rows / columns this launch produces
  int TH, TW;               // tile shape, TH*TW == 128 (TW a power of two)
  int tw_shift;             // log2(TW)
  int a_stages, b_stages, b_slot;   // halo kernel: ring depths and the byte size of one B slot (BN*128 rounded to 1 KB)
  int Cin, Cout, BN;        // Cin % 32 == 0 (padded), BN in {16,32,48,...,192}
  int ntaps;
  int t_stride, t_off;      // input frame = t*t_stride + t_off + dt
  int8_t dt[CV_MAX_TAPS], dy[CV_MAX_TAPS], dx[CV_MAX_TAPS];
  // output placement: element (t, y, x, n) goes to
  //   frame  t*t_mul + n / c_split,  row y*sy + oy,  column x*sx + ox,  channel n % c_split
  float* out; int ldc; int out_H, out_W;
  int t_mul, c_split, sy, sx, oy, ox;
  const float* bias;        // [Cout] or null
  const float* resid;       // same addressing as out, or null
  int planar_clamp;         // 1: out is planar [c_split][frames][out_H][out_W], values clamped to [-1,1]
  long long planar_cstride; // elements between channels in planar mode
  int round_out;            // 1: the stored value is rounded to tf32 (its only consumers are convolutions: see tf32_round)
  // fused RMS_norm (+SiLU) of the result (vae.py:51-54,195-197): with gamma != null (and one N tile covering all Cout
  // channels) every output pixel v[0..Cout) also yields a = silu(v / max(|v|, 1e-12) * sqrt(Cout) * gamma), rounded to tf32,
  // stored to norm_out - or INSTEAD of v when norm_out is null (the raw result has no other reader)
  const float* norm_gamma;
  float* norm_out;
  int norm_silu;
  int vec_bytes;            // shared-memory bytes reserved for bias | gamma (row epilogue), a multiple of 256
  int rows_epi;             // 1: row-per-thread epilogue with TMA stores (plain channels-last output, BN % 32 == 0, TW <= 32)
  long long* prof;          // development: per-role wait / work clocks of CTA 0 (wf_debug_conv_profile), or null
};

// Epilogue of one 128-pixel x BN tile for epilogue warp q (rows 32q..32q+31 of the tile): TMEM -> registers -> shared-memory
// transpose -> 128-byte coalesced row stores (bias, residual, placement, clamp).  The per-row part of the output address
// ((y*sy+oy)*out_W + x*sx+ox) is computed once per tile with shifts (TW is a power of two); per 32-column chunk a row costs
// one multiply-add - the addressing used to be two integer divisions and 64-bit products per row and chunk, which made the
// four epilogue warps the busiest part of the kernel.
__device__ __forceinline__ void conv_epilogue_tile(const ConvArgs& p, float* tile_s, uint32_t t_row, int q, int lane, int t,
                                                   int y0, int x0, int n0) {
  int row_part[32];
  uint32_t row_ok = 0;
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    const int r_in_tile = q * 32 + rr;
    const int y = y0 + (r_in_tile >> p.tw_shift), x = x0 + (r_in_tile & (p.TW - 1));
    row_part[rr] = (y * p.sy + p.oy) * p.out_W + (x * p.sx + p.ox);
    row_ok |= (y < p.H && x < p.W) ? (1u << rr) : 0u;
  }
  const size_t frame_elems = static_cast<size_t>(p.out_H) * p.out_W;
  // ---- fused RMS-norm statistics: TMEM hands each thread one pixel ROW, so the sum of squares over the channels is a
  // per-thread loop (bias and residual included, exactly the value that is stored); the normalised values are produced in the
  // store loop below, where a lane owns a channel, with the row's scale fetched from its owner lane by a shuffle
  float inv = 0.f;
  if (p.norm_gamma) {
    const int r_in_tile = q * 32 + lane;
    const int y = y0 + (r_in_tile >> p.tw_shift), x = x0 + (r_in_tile & (p.TW - 1));
    const bool own_ok = y < p.H && x < p.W;
    const size_t own = (static_cast<size_t>(t) * p.t_mul * frame_elems + static_cast<size_t>((y * p.sy + p.oy) * p.out_W + (x * p.sx + p.ox))) * p.ldc;
    float ss = 0.f;
#pragma unroll 1
    for (int c = 0; c < p.BN; c += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(t_row + c) : "memory");
      tmem_ld_wait();
      float rs[16];
      if (p.resid && own_ok) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 v4 = (c + j < p.Cout) ? *reinterpret_cast<const float4*>(p.resid + own + c + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          rs[j] = v4.x; rs[j + 1] = v4.y; rs[j + 2] = v4.z; rs[j + 3] = v4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) rs[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (c + j < p.Cout) {
          const float v = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + c + j) : 0.f) + rs[j];
          ss = fmaf(v, v, ss);
        }
      }
    }
    inv = sqrtf(static_cast<float>(p.Cout)) / fmaxf(sqrtf(ss), 1e-12f);
  }
#pragma unroll 1
  for (int c = 0; c < p.BN; c += 32) {
    const int n = n0 + c;
    if (n >= p.Cout) break;
    if (p.BN - c >= 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) tile_s[lane * CV_EPI_LD + j] = __uint_as_float(r[j]);
    } else {   // BN = 16 or 48: the last chunk has 16 valid columns; TMEM beyond BN is not ours to read
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(t_row + c) : "memory");
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) tile_s[lane * CV_EPI_LD + j] = __uint_as_float(r[j]);
    }
    __syncwarp();
    const int col = n + lane;
    const bool col_ok = col < p.Cout && (c + lane) < p.BN;
    const float b0 = (col_ok && p.bias) ? p.bias[col] : 0.f;
    const int fo = col_ok ? col / p.c_split : 0, ch = col_ok ? col % p.c_split : 0;
    const size_t frame_off = static_cast<size_t>(t * p.t_mul + fo) * frame_elems;
    // element (row rr, this lane's column): base + row_part[rr] * pitch
    const size_t base = p.planar_clamp ? static_cast<size_t>(ch) * p.planar_cstride + frame_off : frame_off * p.ldc + ch;
    const int pitch = p.planar_clamp ? 1 : p.ldc;
    const uint32_t okm = col_ok ? row_ok : 0u;
    float rv[32];
    if (p.resid) {
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        rv[rr] = ((okm >> rr) & 1u) ? p.resid[base + static_cast<size_t>(row_part[rr]) * pitch] : 0.f;
    }
    if (p.norm_gamma) {
      const float gm = col_ok ? p.norm_gamma[col] : 0.f;
      float* nout = p.norm_out ? p.norm_out : p.out;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        const float inv_r = __shfl_sync(0xffffffffu, inv, rr);         // row rr's scale lives in lane rr
        if ((okm >> rr) & 1u) {
          float v = tile_s[rr * CV_EPI_LD + lane] + b0;
          if (p.resid) v += rv[rr];
          const size_t at = base + static_cast<size_t>(row_part[rr]) * pitch;
          if (p.norm_out) p.out[at] = p.round_out ? tf32_round(v) : v;
          float a = v * inv_r * gm;
          if (p.norm_silu) a = a / (1.0f + expf(-a));
          nout[at] = tf32_round(a);
        }
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        if ((okm >> rr) & 1u) {
          float v = tile_s[rr * CV_EPI_LD + lane] + b0;
          if (p.resid) v += rv[rr];
          if (p.planar_clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
          if (p.round_out) v = tf32_round(v);
          p.out[base + static_cast<size_t>(row_part[rr]) * pitch] = v;
        }
      }
    }
    __syncwarp();
  }
}


// Row-per-thread epilogue for the common placement (plain channels-last output, every channel of the N tile stored): TMEM
// hands each thread one output PIXEL, and a pixel's channels are contiguous in memory - so nothing is transposed.  Per
// 32-channel chunk a thread adds bias (+ residual, read with 16-byte loads from its own pixel row), writes its 128 bytes into a
// SWIZZLE_128B staging tile (16-byte chunk c of row r at c ^ (r & 7): conflict-free) and lane 0 hands the warp's 32 rows
// ({32 channels, TW pixels, 32 / TW rows} of the output tensor) to a TMA store, which also clips partial tiles.  The fused
// RMS-norm is thread-local here (a thread owns all channels of its pixel): pass 1 forms v = acc + bias + residual, sums
// v^2 and writes v back to TMEM; pass 2 re-reads it and stores v (if it has another reader) and silu(v * inv * gamma).
// The transposing epilogue above spent 57 K clocks per 4-tile block on a plain store and 160-213 K with the fused norm
// (1296 MMAs of the block: 98 K); this one is bounded by the ~100 instructions per chunk and thread.
__device__ __forceinline__ void conv_epilogue_rows(const ConvArgs& p, const CUtensorMap* tmOut, const CUtensorMap* tmNorm,
                                                   uint8_t* stage, const float* bias_s, const float* gamma_s, uint32_t t_row,
                                                   int q, int lane, int t, int y0, int x0, int n0) {
  const int nchunks = p.BN >> 5;
  const int yq = y0 + ((q * 32) >> p.tw_shift);             // first output row of this warp's 32 pixels
  uint8_t* srow = stage + lane * 128;
  const int sw = lane & 7;
  auto put = [&](const float (&o)[32], const CUtensorMap* tm, int c, bool round) {
    if (lane == 0) bulk_wait_read_all();                    // the previous store has read the staging tile
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float4 f = make_float4(o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
      if (round) f = make_float4(tf32_round(f.x), tf32_round(f.y), tf32_round(f.z), tf32_round(f.w));
      *reinterpret_cast<float4*>(srow + ((ch ^ sw) << 4)) = f;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) { tma_store_4d(tm, stage, n0 + c * 32, x0, yq, t); bulk_commit_group(); }
  };
  // The residual chunk of the warp's 32 pixels comes in by coalesced 16-byte loads (instruction i covers rows 4i .. 4i+3,
  // lane -> row 4i + lane / 8, piece lane % 8; warp 2 prefetched the lines into L2 during the main loop), issued one chunk
  // AHEAD into registers, and passes through the staging tile so that every thread gets its own row.
  float4 nxt[8];
  auto resid_issue = [&](int c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rt = q * 32 + 4 * i + (lane >> 3);
      const int yy = y0 + (rt >> p.tw_shift), xx = x0 + (rt & (p.TW - 1));
      nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy < p.H && xx < p.W)
        nxt[i] = __ldg(reinterpret_cast<const float4*>(p.resid + ((static_cast<size_t>(t) * p.out_H + yy) * p.out_W + xx) * p.ldc + n0 + c * 32) + (lane & 7));
    }
  };
  // v = accumulator chunk c (+ bias + residual when `add`)
  auto load_v = [&](int c, float (&v)[32], bool add) {
    {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    }
    if (add && p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + c * 32 + j);
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    }
    if (add && p.resid) {
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + (lane >> 3);
        *reinterpret_cast<float4*>(stage + row * 128 + (((lane & 7) ^ (row & 7)) << 4)) = nxt[i];
      }
      __syncwarp();
      if (c + 1 < nchunks) resid_issue(c + 1);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float4 v4 = *reinterpret_cast<const float4*>(srow + ((ch ^ sw) << 4));
        v[4 * ch] += v4.x; v[4 * ch + 1] += v4.y; v[4 * ch + 2] += v4.z; v[4 * ch + 3] += v4.w;
      }
      __syncwarp();                                         // every row read before the tile is rewritten
    }
  };
  auto normalise = [&](float (&v)[32], int c, float inv) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(gamma_s + n0 + c * 32 + j);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a = v[j + u] * inv * gg[u];
        if (p.norm_silu) a = __fdividef(a, 1.0f + __expf(-a));     // result is rounded to tf32: the approximate ex2 / rcp are exact enough
        v[j + u] = a;
      }
    }
  };
  if (p.resid) resid_issue(0);
  if (!p.norm_gamma) {
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      float v[32];
      load_v(c, v, true);
      put(v, tmOut, c, p.round_out != 0);
    }
    return;
  }
  {
    float ss = 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      float v[32];
      load_v(c, v, true);
#pragma unroll
      for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
      uint32_t w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v[j]);
      tmem_st_32x32b_x32(t_row + c * 32, w);
    }
    tmem_st_wait();
    const float inv = sqrtf(static_cast<float>(p.Cout)) / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      float v[32];
      load_v(c, v, false);
      if (p.norm_out) put(v, tmOut, c, p.round_out != 0);
      normalise(v, c, inv);
      put(v, p.norm_out ? tmNorm : tmOut, c, true);
    }
  }
}

__global__ void __launch_bounds__(CV_THREADS, 1)
conv_tf32_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CV_STAGES * CV_STAGE_BYTES);
  uint64_t* empty = full + CV_STAGES;
  uint64_t* acc_full = empty + CV_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tile = reinterpret_cast<float*>(smem + CV_STAGES * CV_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int tiles_x = (p.W + p.TW - 1) / p.TW, tiles_y = (p.H + p.TH - 1) / p.TH;
  const int tiles_n = (p.Cout + p.BN - 1) / p.BN;
  const int tiles_pix = tiles_x * tiles_y * p.T;
  const int num_tiles = tiles_pix * tiles_n;
  const int kchunks = p.Cin / CV_BK;
  const int num_kb = p.ntaps * kchunks;
  const uint32_t stage_tx = CV_A_BYTES + static_cast<uint32_t>(p.BN) * CV_BK * 4;

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < CV_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (n-tile fastest: the n-tiles of one pixel tile run back to back and share the A traffic in L2)
  auto decode = [&](int tile, int& t, int& y0, int& x0, int& n0) {
    const int nt = tile % tiles_n; int pt = tile / tiles_n;
    const int xt = pt % tiles_x; pt /= tiles_x;
    const int yt = pt % tiles_y; t = pt / tiles_y;
    y0 = yt * p.TH; x0 = xt * p.TW; n0 = nt * p.BN;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int t, y0, x0, n0; decode(tile, t, y0, x0, n0);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int ti = t * p.t_stride + p.t_off + p.dt[tap], yi = y0 + p.dy[tap], xi = x0 + p.dx[tap];
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* a_dst = smem + stage * CV_STAGE_BYTES;
            mbar_arrive_expect_tx(&full[stage], stage_tx);
            tma_load_4d(a_dst, &tmA, &full[stage], kc * CV_BK, xi, yi, ti);
            tma_load_2d(a_dst + CV_A_BYTES, &tmB, &full[stage], kc * CV_BK, tap * p.Cout + n0);
          }
          __syncwarp();
          if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer.  The whole warp runs the (uniform) loop
    // and one elected lane issues: with the loop inside an "if (lane == 0)" the compiler cannot prove the descriptors
    // warp-uniform and wraps EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~19 instructions,
    // ~90 clocks per MMA - the whole kernel was issue-bound at 123 clocks per 128x96x8 MMA against a floor of 56)
    const uint32_t idesc = umma_idesc(2, CV_BM, static_cast<uint32_t>(p.BN), 0, 0);   // tf32 x tf32 -> fp32
    const uint32_t a_addr0 = smem_u32(smem);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = a_addr0 + stage * CV_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < CV_BK / 8; ++k)
            umma_tf32_ss(d_tmem, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(a_addr + CV_A_BYTES + k * 32, 16, 1024),
                         idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3, lane = lane_id();
    float* tile_s = epi_tile + q * (32 * CV_EPI_LD);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int t, y0, x0, n0; decode(tile, t, y0, x0, n0);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      conv_epilogue_tile(p, tile_s, t_row, q, lane, t, y0, x0, n0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// 3x3x3 causal convolution: one halo tile per (frame, 32-channel chunk), several pixel tiles per CTA.
//
// Two measurements shape this kernel (tools/umma_offset_probe.cu, tools/umma_rate_probe.cu; profiles/):
//  * tcgen05.mma instructions that accumulate into the SAME TMEM tile issue ~157 clocks apart whatever their width (N = 48
//    .. 256, tf32 or bf16): a single accumulator chain only reaches the tensor-pipe floor (N/2 clocks) at N = 256.  The VAE's
//    convolutions are N = 96 / 192 wide, so one chain runs at 31 % / 61 % - which is where the per-tap kernel above sat
//    (39 %), and neither cheaper addressing, nor a leaner issue loop, nor deeper TMA rings moved it.  Independent chains DO
//    overlap: the CTA therefore owns MT pixel tiles (MT x BN = 384 accumulator columns) and issues their MMAs round-robin.
//  * a SWIZZLE_128B K-major operand may start at any 128-byte row and use any row-group pitch (the swizzle follows absolute
//    address bits).  The MT tiles of 16 x 8 pixels form one block (MT tiles side by side, 16 rows tall) whose halo is ONE 4-D TMA box per
//    (dt, chunk) (out-of-bounds zero fill = padding and causal history); the A operand of tile (ty, tx) and tap (dy, dx) is
//    that box read through a descriptor starting at halo row (16 ty + dy + 1) * PW + 8 tx + dx + 1 with group pitch PW rows.
// One B tile (a tap's [BN, 32] weight slice) feeds all MT tiles: B traffic per MMA drops MT-fold, A traffic ~6-fold.
// K loop order (dt, chunk, in-plane tap); A and B in separate rings fed by two producer warps; one lane issues every MMA.
// The accumulators fill TMEM, so the epilogue of a block is not overlapped with the next block's main loop (it is ~10 %
// of a block now that its addressing is cheap).
// =====================================================================================================================
constexpr int CH_TH = 16, CH_TW = 8;
constexpr int CH_MAX_A_STAGES = 2, CH_MAX_B_STAGES = 16;
constexpr int CH_BAR_BYTES = 512;
constexpr int CH_SMEM_MAX = 232448;
constexpr int CH_THREADS = 384;                     // 4 role warps + 8 epilogue warps (two per scheduler: one hides the other's latencies)
// 384 threads x 168 registers: no setmaxnreg split - capping the role warps made ptxas spill their code, and the row epilogue fits
constexpr int CH_STAGE_BYTES = 8 * 4096;            // row epilogue: one 32-row x 128-byte SWIZZLE_128B staging tile per warp
constexpr int CH_VEC_BYTES = 3072;                  // bias[Cout] and gamma[Cout] in shared memory (Cout <= 384); the host sizes it per launch
constexpr int CH_RING_BUDGET = CH_SMEM_MAX - 1024 - CH_BAR_BYTES - CH_STAGE_BYTES;   // minus the launch's bias / gamma bytes

template <int MT> struct ChGeom {
  // tiles per block along x / y: the MT tiles of 16 x 8 pixels sit side by side (a block is 16 rows tall), so that the row
  // slabs of the sharded VAE (60 rows + halo at 8 ranks) quantise to 16 rows instead of 32; the halo box has the same size
  static constexpr int TX = MT, TY = 1;
  static constexpr int BW = TX * CH_TW, BH = TY * CH_TH;                   // block of output pixels
  static constexpr int PW = BW + 2, PH = BH + 2;                           // halo box
  static constexpr int A_BYTES = PW * PH * CV_BK * 4;
  static constexpr int A_SLOT = (A_BYTES + 1023) / 1024 * 1024;
};

template <int MT, bool PROF>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv333_halo_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmNorm,
                     const __grid_constant__ CUtensorMap tmRes, ConvArgs p) {
  using G = ChGeom<MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int off_b = p.a_stages * G::A_SLOT;
  const int off_epi = off_b + p.b_stages * p.b_slot;          // 1024-byte aligned: staging tiles, then bias / gamma
  const int off_bar = off_epi + CH_STAGE_BYTES + p.vec_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + off_bar);
  uint64_t* a_empty = a_full + CH_MAX_A_STAGES;
  uint64_t* b_full = a_empty + CH_MAX_A_STAGES;
  uint64_t* b_empty = b_full + CH_MAX_B_STAGES;
  uint64_t* acc_full = b_empty + CH_MAX_B_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* epi_tile = reinterpret_cast<float*>(smem + off_epi);
  float* vec_s = reinterpret_cast<float*>(smem + off_epi + CH_STAGE_BYTES);   // bias[Cout] | gamma[Cout]

  const int warp = threadIdx.x >> 5;
  const bool prof = PROF && p.prof != nullptr && blockIdx.x == 0;
  long long pw[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
  const int blocks_x = (p.W + G::BW - 1) / G::BW, blocks_y = (p.H + G::BH - 1) / G::BH;
  const int tiles_n = (p.Cout + p.BN - 1) / p.BN;
  const int num_blocks = blocks_x * blocks_y * p.T * tiles_n;
  const int kchunks = p.Cin / CV_BK;
  const int num_ab = 3 * kchunks;                 // (dt, chunk) pairs = halo boxes per block
  const uint32_t b_tx = static_cast<uint32_t>(p.BN) * CV_BK * 4;
  const int acc_cols = (p.BN + 31) / 32 * 32;     // TMEM columns per tile accumulator

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmNorm); tma_prefetch_desc(&tmRes); }
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < p.Cout; i += CH_THREADS - 128) {
      vec_s[i] = p.bias ? p.bias[i] : 0.f;
      vec_s[p.Cout + i] = p.norm_gamma ? p.norm_gamma[i] : 0.f;
    }
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // block -> (n-tile fastest, then the FRAME, then x, y): the CTAs running at one time work on the same pixel block of
  // consecutive frames, so the three input frames a temporal tap set touches are shared through L2 instead of being read from
  // DRAM once per tap (frame-major order: 5.25 GB of DRAM reads for a 1.38 GB input)
  auto decode = [&](int blk, int& t, int& y0, int& x0, int& n0) {
    const int nt = blk % tiles_n; int pb = blk / tiles_n;
    t = pb % p.T; pb /= p.T;
    const int xb = pb % blocks_x;
    const int yb = pb / blocks_x;
    y0 = yb * G::BH; x0 = xb * G::BW; n0 = nt * p.BN;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ A producer: one halo box per (dt, chunk)
    int stage = 0; uint32_t phase = 0;
    for (int blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
      int t, y0, x0, n0; decode(blk, t, y0, x0, n0);
      for (int dt = 0; dt < 3; ++dt)
        for (int kc = 0; kc < kchunks; ++kc) {
          const long long w0 = prof ? clock64() : 0;
          mbar_wait_trap(&a_empty[stage], phase ^ 1);
          if (prof) pw[0] += clock64() - w0;
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[stage], G::A_BYTES);
            tma_load_4d(smem + stage * G::A_SLOT, &tmA, &a_full[stage], kc * CV_BK, x0 - 1, y0 - 1, t + dt - 2);
          }
          __syncwarp();
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ B producer: the nine in-plane taps of (dt, chunk)
    int stage = 0; uint32_t phase = 0;
    for (int blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
      int t, y0, x0, n0; decode(blk, t, y0, x0, n0);
      for (int dt = 0; dt < 3; ++dt)
        for (int kc = 0; kc < kchunks; ++kc)
          for (int s = 0; s < 9; ++s) {
            const long long w0 = prof ? clock64() : 0;
            mbar_wait_trap(&b_empty[stage], phase ^ 1);
            if (prof) pw[1] += clock64() - w0;
            if (elect_one()) {
              mbar_arrive_expect_tx(&b_full[stage], b_tx);
              tma_load_2d(smem + off_b + stage * p.b_slot, &tmB, &b_full[stage], kc * CV_BK, (dt * 9 + s) * p.Cout + n0);
            }
            __syncwarp();
            if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 2 && p.resid) {
    // ------------------------------------------------------------ residual prefetch: the block's residual boxes into L2 while
    // its main loop runs (the epilogue reads them ~100 K clocks later at L2 instead of DRAM latency)
    uint32_t it = 0;
    for (int blk = blockIdx.x; blk < num_blocks; blk += gridDim.x, ++it) {
      int t, y0, x0, n0; decode(blk, t, y0, x0, n0);
      if (it > 0) mbar_wait_trap(acc_full, (it - 1) & 1);      // one block ahead of the epilogue, not the whole launch
      if (elect_one()) {
        for (int m = 0; m < MT; ++m)
          for (int qq = 0; qq < 4; ++qq)
            for (int c = 0; c < (p.BN >> 5); ++c)
              tma_prefetch_4d(&tmRes, n0 + c * 32, x0 + (m % G::TX) * CH_TW, y0 + (m / G::TX) * CH_TH + qq * (32 / CH_TW), t);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer: MT accumulator chains round-robin; the warp
    // runs the uniform loop, one elected lane issues (see conv_tf32_tcgen05)
    const uint32_t idesc = umma_idesc(2, CV_BM, static_cast<uint32_t>(p.BN), 0, 0);
    const uint32_t a_addr0 = smem_u32(smem), b_addr0 = smem_u32(smem + off_b);
    int as = 0; uint32_t aph = 0;
    int bs = 0; uint32_t bph = 0;
    uint32_t acc_phase = 0;
    for (int blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
      long long w0 = prof ? clock64() : 0;
      mbar_wait_trap(acc_empty, acc_phase ^ 1);
      if (prof) pw[2] += clock64() - w0;
      tc_fence_after();
      for (int ab = 0; ab < num_ab; ++ab) {
        w0 = prof ? clock64() : 0;
        mbar_wait_trap(&a_full[as], aph);
        if (prof) pw[3] += clock64() - w0;
        const uint32_t a_addr = a_addr0 + as * G::A_SLOT;
#pragma unroll
        for (int s = 0; s < 9; ++s) {
          w0 = prof ? clock64() : 0;
          mbar_wait_trap(&b_full[bs], bph);
          if (prof) pw[4] += clock64() - w0;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_addr = b_addr0 + bs * p.b_slot;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < CV_BK / 8; ++k) {
                // tile m = (ty, tx) of the block; tap s = (dy + 1, dx + 1): halo row (16 ty + dy + 1) * PW + 8 tx + dx + 1
                const int ty = m / G::TX, tx = m % G::TX;
                const int row = (ty * CH_TH + s / 3) * G::PW + tx * CH_TW + s % 3;
                umma_tf32_ss(tmem_base + m * acc_cols, umma_desc_sw128(a_addr + row * 128 + k * 32, 16, G::PW * 128),
                             umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc, (ab | s | k) != 0);
              }
            }
            umma_commit(&b_empty[bs]);
            if (s == 8) {
              umma_commit(&a_empty[as]);
              if (ab == num_ab - 1) umma_commit(acc_full);
            }
          }
          __syncwarp();
          if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
        }
        if (++as == p.a_stages) { as = 0; aph ^= 1; }
      }
      acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: the MT tiles one after the other
    // warps 4-7 and 8-11 cover the four TMEM lane quarters twice and take alternate tiles: two warps per scheduler hide each
    // other's TMEM / global / TMA-store latencies.  (Placements the row epilogue does not cover run the per-tap kernel.)
    const int q = warp & 3, lane = lane_id(), wg = (warp - 4) >> 2;
    uint32_t acc_phase = 0;
    for (int blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
      int t, y0, x0, n0; decode(blk, t, y0, x0, n0);
      const long long w0 = prof ? clock64() : 0;
      mbar_wait_trap(acc_full, acc_phase);
      const long long w1 = prof ? clock64() : 0;
      if (prof) pw[5] += w1 - w0;
      tc_fence_after();
#pragma unroll 1
      for (int m = wg; m < MT; m += 2) {
        const int ty = m / G::TX, tx = m % G::TX;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + m * acc_cols;
        if (y0 + ty * CH_TH < p.H && x0 + tx * CH_TW < p.W)
          conv_epilogue_rows(p, &tmOut, &tmNorm, reinterpret_cast<uint8_t*>(epi_tile) + (warp - 4) * 4096, vec_s, vec_s + p.Cout, t_row,
                             q, lane, t, y0 + ty * CH_TH, x0 + tx * CH_TW, n0);
      }
      tc_fence_before();
      __syncwarp();
      if (prof) pw[6] += clock64() - w1;
      if (lane == 0) mbar_arrive(acc_empty);
      acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_all();      // shared memory must outlive the last TMA store's read
  }
  if (prof && lane_id() == 0 && (warp == 0 || warp == 1 || warp == 3 || warp == 4)) {   // (warp 4's epilogue clocks: tiles 0, 2)
    for (int i = 0; i < 7; ++i) if (pw[i]) p.prof[i] = pw[i];
    if (warp == 1) { p.prof[7] = clock64() - t_begin; p.prof[8] = (num_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MT>
static int launch_conv333(const float* in, int in_T, int in_H, int in_W, int Cin, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const CUtensorMap& tmNorm, const CUtensorMap& tmRes, ConvArgs a, void* stream, int grid_cap) {
  using G = ChGeom<MT>;
  CUtensorMap tmH;
  uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(in_W), static_cast<uint64_t>(in_H), static_cast<uint64_t>(in_T)};
  uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 4, static_cast<uint64_t>(in_W) * Cin * 4, static_cast<uint64_t>(in_H) * in_W * Cin * 4};
  uint32_t box[4] = {CV_BK, G::PW, G::PH, 1};
  int rc = make_tmap(&tmH, in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  a.TW = CH_TW; a.TH = CH_TH; a.tw_shift = 3;
  a.b_slot = (a.BN * CV_BK * 4 + 1023) / 1024 * 1024;
  a.a_stages = CH_MAX_A_STAGES;                     // an A slot feeds nine B slots
  a.vec_bytes = (2 * a.Cout * 4 + 255) / 256 * 256;
  a.b_stages = std::min(CH_MAX_B_STAGES, (CH_RING_BUDGET - a.vec_bytes - a.a_stages * G::A_SLOT) / a.b_slot);
  if (a.b_stages < 2) return fail(WF_EINVAL, "wf_conv_tf32: no room for the B ring");
  const int smem = a.a_stages * G::A_SLOT + a.b_stages * a.b_slot + CH_BAR_BYTES + CH_STAGE_BYTES + a.vec_bytes + 1024;
  static PerDeviceOnce once;
  rc = once.run([] {
    WF_CUDA_OK((cudaFuncSetAttribute(conv333_halo_tcgen05<MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX)));
    WF_CUDA_OK((cudaFuncSetAttribute(conv333_halo_tcgen05<MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX)));
    return static_cast<int>(WF_OK);
  });
  if (rc) return rc;
  const long long blocks = static_cast<long long>((a.W + G::BW - 1) / G::BW) * ((a.H + G::BH - 1) / G::BH) * a.T * ((a.Cout + a.BN - 1) / a.BN);
  const int grid = static_cast<int>(std::min<long long>(blocks, grid_cap));
  if (a.prof) conv333_halo_tcgen05<MT, true><<<grid, CH_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmH, tmB, tmOut, tmNorm, tmRes, a);
  else conv333_halo_tcgen05<MT, false><<<grid, CH_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmH, tmB, tmOut, tmNorm, tmRes, a);
  WF_LAUNCH_OK();
  return WF_OK;
}

}  // namespace wf

using namespace wf;

static long long* g_conv_prof = nullptr;
// development hook: device buffer of >= 16 int64 that CTA 0 of the next 3x3x3 launches fills with its per-role clocks
extern "C" int wf_debug_conv_profile(long long* device_buf) { g_conv_prof = device_buf; return WF_OK; }

// in: channels-last fp32 [in_T][in_H][in_W][Cin]; weights fp32 [ntaps*Cout][Cin]; taps: int8 triples (dt,dy,dx).
extern "C" int wf_conv_tf32(const float* in, int in_T, int in_H, int in_W, int Cin, const float* weights, const float* bias,
                            int Cout, int ntaps, const signed char* taps, int T, int H, int W, int t_stride, int t_off,
                            float* out, int ldc, int out_H, int out_W, int t_mul, int c_split, int sy, int sx, int oy,
                            int ox, const float* resid, int planar_clamp, long long planar_cstride, int tile_w,
                            int round_out_tf32, const float* norm_gamma, float* norm_out, int norm_silu, void* stream) {
  WF_REQUIRE(in && weights && out && taps, "wf_conv_tf32: null pointer");
  WF_REQUIRE(Cin > 0 && Cin % 4 == 0, "wf_conv_tf32: Cin must be a positive multiple of 4 (16-byte pixel rows)");
  WF_REQUIRE(Cout > 0 && ntaps > 0 && ntaps <= CV_MAX_TAPS, "wf_conv_tf32: bad Cout / tap count");
  WF_REQUIRE(tile_w == 8 || tile_w == 16 || tile_w == 32 || tile_w == 64 || tile_w == 128, "wf_conv_tf32: tile_w must be 8..128, power of two");
  WF_REQUIRE(T > 0 && H > 0 && W > 0 && c_split > 0 && t_mul > 0, "wf_conv_tf32: empty output");
  WF_REQUIRE(reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(weights) % 16 == 0, "wf_conv_tf32: 16-byte alignment");
  ConvArgs a{};
  a.T = T; a.H = H; a.W = W; a.TW = tile_w; a.TH = CV_BM / tile_w;
  a.tw_shift = 0;
  while ((1 << a.tw_shift) < tile_w) ++a.tw_shift;
  a.Cin = (Cin + CV_BK - 1) / CV_BK * CV_BK;   // the K loop runs over whole 32-channel chunks; TMA zero-fills the tail
  a.Cout = Cout;
  // widest N tile that divides the work evenly enough: 192 for the wide layers, Cout rounded up to 16 otherwise
  int bn = Cout <= CV_MAX_BN ? (Cout + 15) / 16 * 16 : CV_MAX_BN;
  if (Cout > CV_MAX_BN && Cout % 192 != 0 && Cout % 128 == 0) bn = 128;
  a.BN = bn;
  a.ntaps = ntaps; a.t_stride = t_stride; a.t_off = t_off;
  for (int i = 0; i < ntaps; ++i) { a.dt[i] = taps[3 * i]; a.dy[i] = taps[3 * i + 1]; a.dx[i] = taps[3 * i + 2]; }
  a.out = out; a.ldc = ldc; a.out_H = out_H; a.out_W = out_W; a.t_mul = t_mul; a.c_split = c_split;
  a.sy = sy; a.sx = sx; a.oy = oy; a.ox = ox; a.bias = bias; a.resid = resid;
  a.planar_clamp = planar_clamp; a.planar_cstride = planar_cstride;
  a.round_out = round_out_tf32 != 0;
  a.norm_gamma = norm_gamma; a.norm_out = norm_out; a.norm_silu = norm_silu;
  a.prof = g_conv_prof;
  if (norm_gamma) {
    WF_REQUIRE(Cout <= CV_MAX_BN && Cout % 4 == 0, "wf_conv_tf32: the fused RMS-norm needs all channels of a pixel in one N tile (Cout <= 192)");
    WF_REQUIRE(!planar_clamp && c_split == Cout && t_mul == 1 && ldc % 4 == 0, "wf_conv_tf32: the fused RMS-norm needs a plain channels-last output");
    WF_REQUIRE(!resid || reinterpret_cast<uintptr_t>(resid) % 16 == 0, "wf_conv_tf32: fused RMS-norm: residual must be 16-byte aligned");
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(in_W), static_cast<uint64_t>(in_H), static_cast<uint64_t>(in_T)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 4, static_cast<uint64_t>(in_W) * Cin * 4, static_cast<uint64_t>(in_H) * in_W * Cin * 4};
    uint32_t box[4] = {CV_BK, static_cast<uint32_t>(a.TW), static_cast<uint32_t>(a.TH), 1};
    int rc = make_tmap(&tmA, in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(ntaps) * Cout};
    uint64_t strides[1] = {static_cast<uint64_t>(Cin) * 4};
    uint32_t box[2] = {CV_BK, static_cast<uint32_t>(a.BN)};
    int rc = make_tmap(&tmB, weights, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static PerDeviceOnce once;
  {
    const int rc1 = once.run([] {
      WF_CUDA_OK(cudaFuncSetAttribute(conv_tf32_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM));
      return static_cast<int>(WF_OK);
    });
    if (rc1) return rc1;
  }
  static bool attr_set = false;
  static int use_halo = 1, grid_cap = 0;
  if (!attr_set) {
    const char* e = getenv("WF_CONV_HALO");            // 0: always the per-tap kernel (A/B measurements)
    use_halo = e ? atoi(e) : 1;
    const char* g = getenv("WF_CONV_GRID");            // cap on the number of CTAs (development)
    grid_cap = g ? atoi(g) : 0;
    if (grid_cap <= 0) grid_cap = sm_count();
    attr_set = true;
  }
  // the plain 3x3x3 causal convolution (taps in (dt, dy, dx) order, unit strides) runs the halo / multi-tile kernel:
  // MT tiles x BN accumulator columns = 384 independent columns in flight
  bool is333 = use_halo && ntaps == 27 && t_stride == 1 && t_off == 0 && in_H == H && in_W == W;
  for (int i = 0; is333 && i < 27; ++i)
    is333 = taps[3 * i] == i / 9 - 2 && taps[3 * i + 1] == (i / 3) % 3 - 1 && taps[3 * i + 2] == i % 3 - 1;
  if (is333) {
    // the halo kernel has the row-per-thread epilogue with TMA stores only: plain dense channels-last placement, whole 32-channel chunks, 16 x 8 tiles
    CUtensorMap tmOut = tmB, tmNorm = tmB, tmRes = tmB;           // placeholders when the transposing epilogue runs
    static const int rows_epi_on = [] { const char* e = getenv("WF_CONV_ROWS_EPI"); return e ? atoi(e) : 1; }();
    a.rows_epi = rows_epi_on && !planar_clamp && c_split == Cout && t_mul == 1 && sy == 1 && sx == 1 && oy == 0 && ox == 0 && ldc == Cout &&
                 Cout % a.BN == 0 && a.BN % 32 == 0 && 2 * Cout * 4 <= CH_VEC_BYTES && out_H == H && out_W == W &&
                 reinterpret_cast<uintptr_t>(out) % 16 == 0 && (!resid || reinterpret_cast<uintptr_t>(resid) % 16 == 0) &&
                 (!norm_out || reinterpret_cast<uintptr_t>(norm_out) % 16 == 0);
    if (a.rows_epi) {
      uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(out_W), static_cast<uint64_t>(out_H), static_cast<uint64_t>(T)};
      uint64_t strides[3] = {static_cast<uint64_t>(Cout) * 4, static_cast<uint64_t>(out_W) * Cout * 4, static_cast<uint64_t>(out_H) * out_W * Cout * 4};
      uint32_t box[4] = {32, CH_TW, 32 / CH_TW, 1};
      int rc = make_tmap(&tmOut, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      if (norm_out) {
        rc = make_tmap(&tmNorm, norm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
      }
      if (resid) {
        rc = make_tmap(&tmRes, resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
      }
    }
    if (a.rows_epi)
      return a.BN <= 96 ? launch_conv333<4>(in, in_T, in_H, in_W, Cin, tmB, tmOut, tmNorm, tmRes, a, stream, grid_cap)
                        : launch_conv333<2>(in, in_T, in_H, in_W, Cin, tmB, tmOut, tmNorm, tmRes, a, stream, grid_cap);
  }
  const long long tiles = static_cast<long long>((W + a.TW - 1) / a.TW) * ((H + a.TH - 1) / a.TH) * T * ((Cout + a.BN - 1) / a.BN);
  const int grid = static_cast<int>(std::min<long long>(tiles, sm_count()));
  conv_tf32_tcgen05<<<grid, CV_THREADS, CV_SMEM, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, a);
  WF_LAUNCH_OK();
  return WF_OK;
}
