// worldforge_b200 - bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) with a fused epilogue.
//
// This is every nn.Linear of the Wan DiT block (reference wan/modules/model.py:123-126,
// 197-198, 271-273) under autocast(bf16): bf16 operands, fp32 accumulation in TMEM, one
// rounding to bf16, then the epilogue the block applies to that value:
//   EPI_BF16       y = bf16(acc + b)                                   (q,k,v, cross q, text/img embeds)
//   EPI_GELU_BF16  y = bf16(gelu_tanh(bf16(acc + b)))                  (ffn.0 + nn.GELU('tanh'), :272)
//   EPI_RESID_F32  x[m,n] += float(bf16(acc + b)) * gate[n]            (x + y*e, :306,:313; gate==null -> x + y, :310)
//   EPI_F32_OF_BF16 y = float(bf16(acc + b))                           (patch embedding, :534)
//   EPI_RESID_BF16 x[m,n] = bf16(float(x) + gate * float(bf16(acc + b)))  (LongCat's bf16 residual stream,
//                  longcat_video_dit.py:101-103,109,114-116; gate may be a per-frame table)
//
// Layout: A row-major [M,K] (K contiguous) and W row-major [N,K] - nn.Linear's native
// layout - are both "K-major" UMMA operands, so no transposes exist anywhere.
// Persistent, warp-specialised CTA (one per SM): warp 0 TMA producer, warp 1 MMA issuer
// (one elected lane), warp 2 TMEM allocator, warps 4-7 epilogue.  128x256 output tile,
// 64-wide K blocks in 128-byte-swizzled smem, 4-stage mbarrier ring, two 256-column TMEM
// accumulators so tile i's epilogue overlaps tile i+1's main loop.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;   // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int EPI_LD = 65;                              // padded row of the per-warp transpose tile (floats)
constexpr int GEMM_EPI_BYTES = 4 * 32 * EPI_LD * 4;     // one 32 x 64 fp32 tile per epilogue warp
constexpr int GEMM_SMEM = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + GEMM_EPI_BYTES;
constexpr int GEMM_THREADS = 256;

enum { EPI_BF16 = 0, EPI_GELU_BF16 = 1, EPI_RESID_F32 = 2, EPI_F32_OF_BF16 = 3, EPI_RESID_BF16 = 4 };

struct GemmArgs {
  int M, N, K;
  const bf16* bias;    // [N] or null
  void* out;           // bf16 [M,ldo] or float [M,ldo]
  int ldo;
  const float* gate;   // [N] or null (EPI_RESID_F32)
  int group_n;         // n-tiles per raster group (L2 working-set control)
  int gate_rows;       // > 0: gate is a [groups, N] table, row m uses group m / gate_rows (per-frame adaLN gates)
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}

// Epilogue of one accumulator for one epilogue warp: rows m0..m0+31 (TMEM lanes of this warp's quarter), columns n0..n0+255.
// TMEM gives each thread one accumulator ROW; global memory wants a warp on one row.  Each warp transposes 32x32 (fp32 out) /
// 32x64 (bf16 out) chunks through a private padded smem tile so that every global load / store instruction covers 128
// contiguous bytes of one output row.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_warp(const GemmArgs& p, float* tile, uint32_t t_row, int m0, int n0, int lane) {
  constexpr bool kOutBf16 = (EPI == EPI_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_RESID_BF16);
  constexpr int CH = kOutBf16 ? 64 : 32;
#pragma unroll 1
  for (int c = 0; c < GEMM_BN; c += CH) {
    const int n = n0 + c;
    if (n >= p.N) break;                 // warp-uniform
    {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) tile[lane * EPI_LD + j] = __uint_as_float(r[j]);
      if (CH == 64) {
        tmem_ld_32x32b_x32(t_row + c + 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[lane * EPI_LD + 32 + j] = __uint_as_float(r[j]);
      }
    }
    __syncwarp();
    if (kOutBf16) {
      const int col = n + 2 * lane;
      const bool ok = col < p.N;
      const float b0 = (ok && p.bias) ? __bfloat162float(p.bias[col]) : 0.f;
      const float b1 = (ok && p.bias) ? __bfloat162float(p.bias[col + 1]) : 0.f;
      bf16* dst = static_cast<bf16*>(p.out) + static_cast<size_t>(m0) * p.ldo + col;
      if (EPI == EPI_RESID_BF16) {
        const int rows_ok = min(32, p.M - m0);
        uint32_t xv[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr)
          xv[rr] = (ok && rr < rows_ok) ? *reinterpret_cast<const uint32_t*>(dst + static_cast<size_t>(rr) * p.ldo) : 0u;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          if (ok && rr < rows_ok) {
            float g0 = 1.f, g1 = 1.f;
            if (p.gate) {
              const float* gp = p.gate + (p.gate_rows > 0 ? static_cast<size_t>((m0 + rr) / p.gate_rows) * p.N : 0) + col;
              g0 = gp[0]; g1 = gp[1];
            }
            const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv[rr]));
            const float y0 = bf16_round(tile[rr * EPI_LD + 2 * lane] + b0);
            const float y1 = bf16_round(tile[rr * EPI_LD + 2 * lane + 1] + b1);
            *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(rr) * p.ldo) =
                pack_bf16x2(__fadd_rn(x.x, __fmul_rn(g0, y0)), __fadd_rn(x.y, __fmul_rn(g1, y1)));
          }
        }
      } else {
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        if (m0 + rr < p.M && ok) {
          float y0 = bf16_round(tile[rr * EPI_LD + 2 * lane] + b0);
          float y1 = bf16_round(tile[rr * EPI_LD + 2 * lane + 1] + b1);
          if (EPI == EPI_GELU_BF16) { y0 = gelu_tanh_f(y0); y1 = gelu_tanh_f(y1); }
          *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(rr) * p.ldo) = pack_bf16x2(y0, y1);
        }
      }
      }
    } else {
      const int col = n + lane;
      const float b0 = p.bias ? __bfloat162float(p.bias[col]) : 0.f;
      const float g0 = (EPI == EPI_RESID_F32 && p.gate && p.gate_rows == 0) ? p.gate[col] : 1.f;
      float* dst = static_cast<float*>(p.out) + static_cast<size_t>(m0) * p.ldo + col;
      const int rows_ok = min(32, p.M - m0);          // warp-uniform
      if (EPI == EPI_RESID_F32) {
        // all 32 residual loads are issued before the first store: one round trip, not 32
        float xv[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) xv[rr] = (rr < rows_ok) ? dst[static_cast<size_t>(rr) * p.ldo] : 0.f;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          if (rr < rows_ok) {
            const float y = bf16_round(tile[rr * EPI_LD + lane] + b0);
            const float g = (p.gate && p.gate_rows > 0) ? p.gate[static_cast<size_t>((m0 + rr) / p.gate_rows) * p.N + col] : g0;
            dst[static_cast<size_t>(rr) * p.ldo] = __fadd_rn(xv[rr], __fmul_rn(y, g));   // x + y*e: two roundings, like torch
          }
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr)
          if (rr < rows_ok) dst[static_cast<size_t>(rr) * p.ldo] = bf16_round(tile[rr * EPI_LD + lane] + b0);
      }
    }
    __syncwarp();
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
  uint64_t* empty = full + GEMM_STAGES;
  uint64_t* acc_full = empty + GEMM_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tile = reinterpret_cast<float*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  // Raster order: groups of `group_n` n-tiles are swept over ALL m-tiles before the next group starts, n fastest inside
  // a group.  The CTAs in flight then share a few A row-blocks and at most group_n weight column-blocks - a working
  // set sized to stay in the 126 MB L2 - instead of streaming the whole weight matrix from HBM once per wave.
  auto tile_origin = [&](int tile, int& m0, int& n0) {
    const int per_group = p.group_n * tiles_m;
    const int g = tile / per_group, r = tile - g * per_group;
    const int gw = min(p.group_n, tiles_n - g * p.group_n);     // the last group may be narrower
    m0 = (r / gw) * GEMM_BM;
    n0 = (g * p.group_n + r % gw) * GEMM_BN;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m0, n0; tile_origin(tile, m0, n0);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + stage * GEMM_STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], GEMM_STAGE_BYTES);
          tma_load_2d(a_dst, &tmA, &full[stage], kb * GEMM_BK, m0);
          tma_load_2d(a_dst + GEMM_A_BYTES, &tmB, &full[stage], kb * GEMM_BK, n0);
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc(1, GEMM_BM, GEMM_BN, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * GEMM_STAGE_BYTES);
          const uint32_t b_addr = a_addr + GEMM_A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
            uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);                       // smem slot free once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&acc_full[acc]); // accumulator complete
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    // TMEM gives each thread one accumulator ROW; global memory wants a warp on one row.  Each warp
    // transposes 32x32 (fp32 out) / 32x64 (bf16 out) chunks through a private padded smem tile so
    // that every global load / store instruction covers 128 contiguous bytes of one output row.
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int lane = lane_id();
    float* tile = epi_tile + q * (32 * EPI_LD);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_idx = blockIdx.x; tile_idx < num_tiles; tile_idx += gridDim.x) {
      int m0, n0; tile_origin(tile_idx, m0, n0); m0 += q * 32;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * GEMM_BN;
      gemm_epilogue_warp<EPI>(p, tile, t_row, m0, n0, lane);
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

template <int EPI>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, const GemmArgs& args, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BN};
    int rc = make_tmap(&tmB, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  // the attribute is per device and the call is cheap: set it every time (a process may drive several GPUs)
  WF_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  const int tiles = ((args.M + GEMM_BM - 1) / GEMM_BM) * ((args.N + GEMM_BN - 1) / GEMM_BN);
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  gemm_bf16_tcgen05<EPI><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(tmA, tmB, args);
  WF_LAUNCH_OK();
  return WF_OK;
}


// =====================================================================================================================
// CTA-pair form (cta_group::2): one cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  CTA r of the pair stages
// rows [m0 + 128 r, +128) of A and rows [n0 + 128 r, +128) of W per 64-wide K block - 32 KB per stage instead of 48 KB, so
// the ring is six stages deep - and ONE tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), issued by the leader, reads both
// halves of B from the two shared memories: per SM the tensor core fetches 8 KB per MMA instead of 12 KB, which is what holds
// the single-CTA kernel at ~190 clocks per 128-clock MMA (shared-memory operand bandwidth, and power).
//   full[s]      leader's barrier: the leader's producer arms it for the bytes of BOTH CTAs; the peer's TMA signals it remotely
//   empty[s]     one per CTA, released in both by the leader's multicast tcgen05.commit
//   acc_full[a]  one per CTA, multicast commit after the last K block
//   acc_empty[a] leader's barrier, 8 arrivals: the four epilogue warps of each CTA (the peer's arrive through the cluster)
// The epilogue is the single-CTA one on each CTA's 128 rows.
// =====================================================================================================================
constexpr int GEMM2_STAGES = 6;
constexpr int GEMM2_B_BYTES = 128 * GEMM_BK * 2;                       // this CTA's half of the W tile: 16 KB
constexpr int GEMM2_STAGE_BYTES = GEMM_A_BYTES + GEMM2_B_BYTES;        // 32 KB
constexpr int GEMM2_SMEM = GEMM2_STAGES * GEMM2_STAGE_BYTES + 1024 + 256 + GEMM_EPI_BYTES;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GEMM2_STAGES * GEMM2_STAGE_BYTES);
  uint64_t* empty = full + GEMM2_STAGES;
  uint64_t* acc_full = empty + GEMM2_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tile = reinterpret_cast<float*>(smem + GEMM2_STAGES * GEMM2_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int tiles_n = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  auto tile_origin = [&](int tile, int& m0, int& n0) {       // the L2-aware grouped raster of the single-CTA kernel, 256-row tiles
    const int per_group = p.group_n * tiles_m;
    const int g = tile / per_group, r = tile - g * per_group;
    const int gw = min(p.group_n, tiles_n - g * p.group_n);
    m0 = (r / gw) * (2 * GEMM_BM);
    n0 = (g * p.group_n + r % gw) * GEMM_BN;
  };

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < GEMM2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                       // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; bytes counted on the leader's barrier)
    int stage = 0; uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m0, n0; tile_origin(tile, m0, n0);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + stage * GEMM2_STAGE_BYTES;
          const uint32_t bar = mapa_shared(smem_u32(&full[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * GEMM2_STAGE_BYTES);
          tma_load_2d_pair(a_dst, &tmA, bar, kb * GEMM_BK, m0 + static_cast<int>(rank) * GEMM_BM);
          tma_load_2d_pair(a_dst + GEMM_A_BYTES, &tmB, bar, kb * GEMM_BK, n0 + static_cast<int>(rank) * 128);
        }
        __syncwarp();
        if (++stage == GEMM2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc(1, 2 * GEMM_BM, GEMM_BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * GEMM2_STAGE_BYTES);
            const uint32_t b_addr = a_addr + GEMM_A_BYTES;
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
              uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
              umma_f16_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0);
            }
            umma_commit_pair(&empty[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&acc_full[acc]);
          }
          __syncwarp();
          if (++stage == GEMM2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: this CTA's 128 rows of the pair's tile
    const int q = warp & 3;
    const int lane = lane_id();
    float* tile = epi_tile + q * (32 * EPI_LD);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_idx = pair; tile_idx < num_tiles; tile_idx += num_pairs) {
      int m0, n0; tile_origin(tile_idx, m0, n0); m0 += static_cast<int>(rank) * GEMM_BM + q * 32;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * GEMM_BN;
      if (m0 < p.M) gemm_epilogue_warp<EPI>(p, tile, t_row, m0, n0, lane);      // warp-uniform
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();                       // nobody leaves while the peer may still signal into this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int EPI>
static int launch_gemm_pair(const void* A, int lda, const void* W, int ldw, const GemmArgs& args, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {GEMM_BK, 128};
    int rc = make_tmap(&tmB, W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  WF_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tcgen05_pair<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM));
  // persistent pairs: as many clusters as the device can keep resident at once (a GPC with an odd number of free SMs
  // cannot host a pair on its last SM - a grid of sm_count/2 pairs could then need a second wave)
  static int max_pairs = 0;
  if (max_pairs == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() / 2 * 2); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = GEMM2_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    WF_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, gemm_bf16_tcgen05_pair<EPI>, &cfg));
    max_pairs = std::max(1, std::min(n, sm_count() / 2));
  }
  const int tiles = ((args.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * ((args.N + GEMM_BN - 1) / GEMM_BN);
  int grid = std::min(max_pairs, tiles) * 2;
  gemm_bf16_tcgen05_pair<EPI><<<grid, GEMM_THREADS, GEMM2_SMEM, stream>>>(tmA, tmB, args);
  WF_LAUNCH_OK();
  return WF_OK;
}

}  // namespace wf

extern "C" int wf_gemm_bf16(const void* a, int lda, const void* w, int ldw, const void* bias, void* out, int ldo,
                            const float* gate, int gate_rows, int M, int N, int K, int epilogue, void* stream) {
  using namespace wf;
  WF_REQUIRE(a && w && out, "wf_gemm_bf16: null pointer");
  WF_REQUIRE(M > 0 && N > 0 && K > 0, "wf_gemm_bf16: empty problem");
  WF_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "wf_gemm_bf16: K, lda, ldw must be multiples of 8 (16-byte TMA rows)");
  WF_REQUIRE(N % 32 == 0, "wf_gemm_bf16: N must be a multiple of 32");
  WF_REQUIRE(ldo % 8 == 0, "wf_gemm_bf16: ldo must be a multiple of 8");
  WF_REQUIRE((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
             "wf_gemm_bf16: pointers must be 16-byte aligned");
  // n-tiles per raster group: keep group_n weight column-blocks (256 x K bf16 each) within ~40 MB of L2
  const long long col_block_bytes = 2ll * GEMM_BN * K;
  int group_n = static_cast<int>(std::max<long long>(1, (40ll << 20) / col_block_bytes));
  WF_REQUIRE(gate_rows >= 0, "wf_gemm_bf16: gate_rows must be >= 0");
  GemmArgs args{M, N, K, static_cast<const bf16*>(bias), out, ldo, gate, group_n, gate_rows};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // CTA pairs (256 x 256 tiles) for every problem with at least one full pair tile; WF_GEMM_PAIR=0 keeps the single-CTA kernel
  static const int use_pair = [] { const char* e = getenv("WF_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (use_pair && M >= 2 * GEMM_BM && N >= 128) {
    switch (epilogue) {
      case EPI_BF16: return launch_gemm_pair<EPI_BF16>(a, lda, w, ldw, args, s);
      case EPI_GELU_BF16: return launch_gemm_pair<EPI_GELU_BF16>(a, lda, w, ldw, args, s);
      case EPI_RESID_F32: return launch_gemm_pair<EPI_RESID_F32>(a, lda, w, ldw, args, s);
      case EPI_F32_OF_BF16: return launch_gemm_pair<EPI_F32_OF_BF16>(a, lda, w, ldw, args, s);
      case EPI_RESID_BF16: return launch_gemm_pair<EPI_RESID_BF16>(a, lda, w, ldw, args, s);
    }
    return fail(WF_EINVAL, "wf_gemm_bf16: unknown epilogue");
  }
  switch (epilogue) {
    case EPI_BF16: return launch_gemm<EPI_BF16>(a, lda, w, ldw, args, s);
    case EPI_GELU_BF16: return launch_gemm<EPI_GELU_BF16>(a, lda, w, ldw, args, s);
    case EPI_RESID_F32: return launch_gemm<EPI_RESID_F32>(a, lda, w, ldw, args, s);
    case EPI_F32_OF_BF16: return launch_gemm<EPI_F32_OF_BF16>(a, lda, w, ldw, args, s);
    case EPI_RESID_BF16: return launch_gemm<EPI_RESID_BF16>(a, lda, w, ldw, args, s);
  }
  return fail(WF_EINVAL, "wf_gemm_bf16: unknown epilogue");
}
