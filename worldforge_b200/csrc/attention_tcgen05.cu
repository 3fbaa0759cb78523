// worldforge_b200 - non-causal multi-head attention (head_dim 128) on tcgen05 / TMEM / TMA.
//
//   O[q, h, :] = softmax_k( Q[q,h,:] . K[k,h,:] / sqrt(128) ) V[k,h,:]
//
// This is flash_attention() of the reference (wan/modules/attention.py:24-130) for both call
// sites of the Wan DiT block: the 3-D spatio-temporal self-attention (model.py:149-154;
// Lq = Lk = 32 760 tokens at 480p/81f) and the text / image cross-attention (model.py:220-222;
// Lk = 512 and 257).  bf16 q,k,v in, fp32 scores and softmax statistics, bf16 probabilities
// into the PV product, fp32 accumulation, bf16 out.  ``add_in`` fuses the reference's
// "x = x + img_x" (model.py:227): out = bf16(float(bf16(o)) + float(add_in)).
//
// Layout: q,k,v,out are the [tokens, heads*128] row-major matrices the q/k/v projections
// produce - head h is the 128-column slice at h*128; nothing is transposed or re-packed.
//
// One CTA owns TWO 128-row query tiles of one head and walks the keys in blocks:
//   warp 0   TMA producer   (Q once; K/V blocks through an mbarrier ring)
//   warp 1   MMA issuer     S_t = Q_t K^T  and  O_t += P_t V
//   warp 2   TMEM allocator
//   warps 4-7 / 8-11  softmax for tile 0 / tile 1: one thread per query row
// The issuer alternates the two tiles so that while tile 0's rows are in softmax the tensor
// core runs tile 1's MMAs (and vice versa).  V is consumed as an MN-major UMMA operand
// straight from its TMA box.  O stays in TMEM for the whole key loop; it is rescaled there
// only when a row maximum grows by more than 2^8 (lazy rescaling), so the common path never
// touches O.  Two kernels: v2 (64-key blocks; the short ragged key sets of the cross-attention)
// and v3 (128-key blocks; the default).  (Round 1's first kernel, single S buffer with P
// through shared memory, is gone: every variant below supersedes it.)
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "attn_math.cuh"
#include "host_util.h"

namespace wf {

constexpr int AT_BM = 128, AT_BN = 64, AT_D = 128, AT_STAGES = 4;
constexpr int AT_Q_BYTES = AT_BM * AT_D * 2;    // 32 KB per query tile
constexpr int AT_K_BYTES = AT_BN * AT_D * 2;    // 16 KB
constexpr int AT_KV_BYTES = 2 * AT_K_BYTES;     // K + V of one stage
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;   // 16 KB per query tile
constexpr int AT_OFF_KV = 2 * AT_Q_BYTES;
constexpr int AT_OFF_P = AT_OFF_KV + AT_STAGES * AT_KV_BYTES;
constexpr int AT_OFF_BAR = AT_OFF_P + 2 * AT_P_BYTES;
constexpr int AT_SMEM = AT_OFF_BAR + 256 + 1024;
constexpr int AT_THREADS = 384;
constexpr uint32_t AT_TMEM_S = 0, AT_TMEM_O = 128;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;

constexpr int AT_MAX_PEERS = 8;

struct AttnArgs {
  int Lq, Lk;
  bf16* out; int ldo;
  // one process per GPU: query row q is stored to out_peer[q / rows_per_peer] + (q % rows_per_peer) * ldo (n_peers > 0)
  bf16* out_peer[AT_MAX_PEERS]; int n_peers; int rows_per_peer;
  const bf16* add_in; int ld_add;
  float scale_log2;     // softmax scale * log2(e)
};

// =====================================================================================================================
// v2: 64-key blocks with (a) S double-buffered in TMEM so the tensor core computes S(j+1) for both tiles while the
// softmax warps still work on block j - the softmax never waits for the MMA and vice versa except through P;
// (b) packed fp32x2 arithmetic (FFMA2 / FADD2), 3-input max (FMNMX3) and one fused multiply-add per score for
// "scale, subtract the running maximum"; (c) optionally, every fourth pair of exponentials evaluated on the FMA pipe
// (Cody-Waite range reduction + a degree-3 polynomial, relative error 7.5e-5, far below the bf16 rounding P gets
// anyway), because at 64-key blocks the MUFU pipe (16 ex2 / clk / SM) needs exactly as many cycles as the MMAs.
// TMEM columns: S_t[buf] at (2t+buf)*64, O_t at 256 + 128 t.
// =====================================================================================================================
constexpr uint32_t AT2_TMEM_O = 256;

// 2^x for a pair, x <= ~8, on the FMA / ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5], 2^f by a cubic
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& e0, float& e1) {
  float x0, x1; unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t magic = pack2(12582912.0f, 12582912.0f), nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t t2 = fadd2(x2, magic);                       // low mantissa bits = round(x)
  const uint64_t n2 = fadd2(t2, nmagic);
  const uint64_t f2 = ffma2(n2, pack2(-1.0f, -1.0f), x2);
  uint64_t p2 = ffma2(pack2(0.0551716685f, 0.0551716685f), f2, pack2(0.2426111251f, 0.2426111251f));
  p2 = ffma2(p2, f2, pack2(0.6932609677f, 0.6932609677f));
  p2 = ffma2(p2, f2, pack2(0.9999280572f, 0.9999280572f));
  float p0, p1, t0, t1; unpack2(p2, p0, p1); unpack2(t2, t0, t1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// PTMEM: P_t(j) is stored back into the first 32 columns of the S buffer it was computed from (64 bf16 = 32 columns)
// and consumed by the PV product as a TMEM operand: no shared-memory round trip for P, which with 64-key blocks is
// what pushes shared-memory reads (Q 32 KB + K 16 KB per S product, P 16 KB + V 16 KB per PV product) past the MMA time.
template <int POLY, bool PTMEM>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tcgen05_v2(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + AT_STAGES;
  uint64_t* s_full = kv_empty + AT_STAGES;   // [tile][buf] -> 4
  uint64_t* p_full = s_full + 4;             // [tile][buf] -> 4: with P in TMEM the softmax of block j+1 does not wait for
                                             // PV(j), so a tile can signal P(j) and P(j+1) before the issuer looks - one
                                             // barrier per S buffer keeps the two completions apart (a single barrier
                                             // would wrap its phase parity and the issuer would wait forever)
  uint64_t* o_done = p_full + 4;             // 2: one completion per PV product; a waiter may be at most ONE phase behind
  uint64_t* o_final = o_done + 2;            // 2: completes once, after the last PV product of the tile.  The epilogue cannot
                                             // use o_done: after its last softmax step only PV(nblk-3) is known complete, and a
                                             // parity wait two phases behind is satisfied by the wrong completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);

  const int warp = threadIdx.x >> 5;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * AT_BM);
  const int nblk = (p.Lk + AT_BN - 1) / AT_BN;
  const int col0 = head * AT_D;

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
    for (int t = 0; t < 2; ++t) { mbar_init(&o_done[t], 1); mbar_init(&o_final[t], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * AT_Q_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int half = 0; half < 2; ++half)
          tma_load_2d(smem + t * AT_Q_BYTES + half * (AT_Q_BYTES / 2), &tmQ, q_full, col0 + half * 64, q0 + t * AT_BM);
    }
    __syncwarp();
    for (int j = 0; j < nblk; ++j) {
      const int stage = j % AT_STAGES;
      mbar_wait(&kv_empty[stage], ((j / AT_STAGES) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* kdst = smem + AT_OFF_KV + stage * AT_KV_BYTES;
        uint8_t* vdst = kdst + AT_K_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], AT_KV_BYTES);
        for (int half = 0; half < 2; ++half) {
          tma_load_2d(kdst + half * (AT_K_BYTES / 2), &tmK, &kv_full[stage], col0 + half * 64, j * AT_BN);
          tma_load_2d(vdst + half * (AT_K_BYTES / 2), &tmV, &kv_full[stage], col0 + half * 64, j * AT_BN);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(1, AT_BM, AT_BN, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc(1, AT_BM, AT_D, 0, 1);
    const uint32_t q_addr = smem_u32(smem);
    const uint32_t p_addr = smem_u32(smem + AT_OFF_P);
    auto issue_s = [&](int t, int stage, int buf) {
      const uint32_t k_addr = smem_u32(smem + AT_OFF_KV + stage * AT_KV_BYTES);
#pragma unroll
      for (int ks = 0; ks < AT_D / 16; ++ks) {
        uint64_t da = umma_desc_sw128(q_addr + t * AT_Q_BYTES + (ks >> 2) * (AT_Q_BYTES / 2) + (ks & 3) * 32, 16, 1024);
        uint64_t db = umma_desc_sw128(k_addr + (ks >> 2) * (AT_K_BYTES / 2) + (ks & 3) * 32, 16, 1024);
        umma_f16_ss(tmem_base + (2 * t + buf) * AT_BN, da, db, idesc_s, ks != 0);
      }
      umma_commit(&s_full[2 * t + buf]);
    };
    auto issue_pv = [&](int t, int stage, int j) {
      const uint32_t v_addr = smem_u32(smem + AT_OFF_KV + stage * AT_KV_BYTES + AT_K_BYTES);
#pragma unroll
      for (int ks = 0; ks < AT_BN / 16; ++ks) {
        uint64_t db = umma_desc_sw128(v_addr + ks * 2048, AT_K_BYTES / 2, 1024);
        if (PTMEM) {
          umma_f16_ts(tmem_base + AT2_TMEM_O + t * AT_D, tmem_base + (2 * t + (j & 1)) * AT_BN + ks * 8, db, idesc_o, (j | ks) != 0);
        } else {
          uint64_t da = umma_desc_sw128(p_addr + t * AT_P_BYTES + ks * 32, 16, 1024);
          umma_f16_ss(tmem_base + AT2_TMEM_O + t * AT_D, da, db, idesc_o, (j | ks) != 0);
        }
      }
      umma_commit(&o_done[t]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (elect_one()) { issue_s(0, 0, 0); issue_s(1, 0, 0); }
    __syncwarp();
    for (int j = 0; j < nblk; ++j) {
      const int stage = j % AT_STAGES;
      if (j + 1 < nblk) {
        // S(j+1) goes to the other S buffer: its last reader, softmax(j-1), signalled p_full(j-1), which was waited below
        const int nstage = (j + 1) % AT_STAGES;
        mbar_wait(&kv_full[nstage], ((j + 1) / AT_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) { issue_s(0, nstage, (j + 1) & 1); issue_s(1, nstage, (j + 1) & 1); }
        __syncwarp();
      }
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) { issue_pv(0, stage, j); if (j + 1 == nblk) umma_commit(&o_final[0]); }
      __syncwarp();
      mbar_wait(&p_full[2 + (j & 1)], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) { issue_pv(1, stage, j); umma_commit(&kv_empty[stage]); if (j + 1 == nblk) umma_commit(&o_final[1]); }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane_id();
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_addr + AT2_TMEM_O + t * AT_D;
    uint8_t* p_row = smem + AT_OFF_P + t * AT_P_BYTES + row * 128;
    const float c = p.scale_log2;
    const uint64_t c2 = pack2(c, c);
    float m_ref = 0.f, l = 0.f;
    // MASKED = the last, partial key block.  Two instantiations behind a warp-uniform branch: as a predicate inside one
    // body the mask costs an ISETP + SEL per score in EVERY block
    auto softmax_block = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int buf = j & 1;
      mbar_wait(&s_full[2 * t + buf], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[64];
      {
        uint32_t (&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
        uint32_t (&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
        const uint32_t s_tmem = tmem_base + lane_addr + (2 * t + buf) * AT_BN;
        tmem_ld_32x32b_x32(s_tmem, r0);
        tmem_ld_32x32b_x32(s_tmem + 32, r1);
        tmem_ld_wait();
      }
      if (MASKED) {
        const int valid = p.Lk - j * AT_BN;
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i >= valid) r[i] = 0xff800000u;   // -inf
      }
      float mx = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
#pragma unroll
      for (int i = 3; i < 63; i += 2) mx = fmax3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      mx = fmaxf(mx, __uint_as_float(r[63]));
      const float m_blk = mx * c;
      float alpha = 1.0f;
      bool grow = false;
      if (j == 0) {
        m_ref = m_blk;
      } else if (m_blk - m_ref > AT_RESCALE_THRESHOLD) {
        alpha = ex2(m_ref - m_blk);
        m_ref = m_blk;
        grow = true;
      }
      const uint64_t nm2 = pack2(-m_ref, -m_ref);
      uint64_t sum2 = pack2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint64_t x2 = ffma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nm2);
        float e0, e1;
        if (((POLY == 1 && (i & 3) == 3) || (POLY == 2 && (i & 1))) && !MASKED) {
          exp2_poly2(x2, e0, e1);
        } else {
          float x0, x1; unpack2(x2, x0, x1);
          e0 = ex2(x0); e1 = ex2(x1);
        }
        sum2 = fadd2(sum2, pack2(e0, e1));
        pk[i] = pack_bf16x2(e0, e1);
      }
      float s_lo, s_hi; unpack2(sum2, s_lo, s_hi);
      l = l * alpha + (s_lo + s_hi);
      const bool any_grow = __any_sync(0xffffffffu, grow);
      if (j > 0 && (!PTMEM || any_grow)) {
        // S(j) complete implies PV(j-2) complete (in-order MMA pipe), so this parity wait can only mean PV(j-1)
        mbar_wait(&o_done[t], (j - 1) & 1);
        tc_fence_after();
      }
      if (j > 0 && any_grow) {
#pragma unroll 1
        for (int cc = 0; cc < AT_D; cc += 32) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_tmem + cc, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(o_tmem + cc, o);
        }
        tmem_st_wait();
      }
      if (PTMEM) {
        tmem_st_32x32b_x32(tmem_base + lane_addr + (2 * t + buf) * AT_BN, pk);
        tmem_st_wait();
      } else {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 v = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
          *reinterpret_cast<uint4*>(p_row + ((ch ^ (row & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&p_full[2 * t + buf]);
    };
    const int nfull = p.Lk / AT_BN;
#pragma unroll 1
    for (int j = 0; j < nfull; ++j) softmax_block(j, std::false_type{});
    if (nfull < nblk) softmax_block(nfull, std::true_type{});
    mbar_wait(&o_final[t], 0);
    tc_fence_after();
    const int q = q0 + t * AT_BM + row;
    const float inv_l = 1.0f / l;
#pragma unroll 1
    for (int cc = 0; cc < AT_D; cc += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_tmem + cc, o);
      tmem_ld_wait();
      if (q < p.Lq) {
        bf16* dst = (p.n_peers > 0 ? p.out_peer[q / p.rows_per_peer] + static_cast<size_t>(q % p.rows_per_peer) * p.ldo
                                   : p.out + static_cast<size_t>(q) * p.ldo) + col0 + cc;
        const bf16* add = p.add_in ? p.add_in + static_cast<size_t>(q) * p.ld_add + col0 + cc : nullptr;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = __uint_as_float(o[i + u]) * inv_l;
          if (add) {
            uint4 a = *reinterpret_cast<const uint4*>(add + i);
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float2 f = __bfloat1622float2(a2[u]);
              w[2 * u] = __fadd_rn(bf16_round(w[2 * u]), f.x);
              w[2 * u + 1] = __fadd_rn(bf16_round(w[2 * u + 1]), f.y);
            }
          }
          uint4 v = make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]),
                               pack_bf16x2(w[6], w[7]));
          *reinterpret_cast<uint4*>(dst + i) = v;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// v3: 128-key blocks.  With 64-key blocks every S product re-reads the 32 KB query tile from shared memory for 64 keys'
// worth of math - the UMMA operand fetch (Q 32 KB + K 16 KB per 256 MMA clocks = 192 B/clk against the 128 B/clk shared
// memory delivers) is what holds v2's tensor pipe at 58 %.  At 128 keys the S product reads 64 KB per 512 MMA clocks and
// the PV product (P from TMEM) 32 KB per 512: the tensor core is fed.  TMEM cannot hold two S buffers per tile at that
// width (2 tiles x 2 x 128 + 2 x 128 columns of O = 768 > 512), so the two query tiles ping-pong instead: S_t lives in
// 128 columns, P_t (bf16) overwrites its first 64, and the issuer's order  PV_0(j) S_0(j+1) PV_1(j) S_1(j+1)  leaves each
// tile's softmax the time of the other tile's two products.  tcgen05.mma retires in issue order, so S_t(j+1) cannot
// overwrite P_t(j) before PV_t(j) has read it, and "S_t(j+1) complete" tells the softmax that O_t is quiescent.
// K and V travel separately through one 5-slot ring of 32 KB boxes in consumption order K0 V0 K1 V1 ...
// The softmax warpgroups hold a whole 128-score row in registers (setmaxnreg moves the producer warps' registers to
// them), and POLY of every 8 exponential pairs are evaluated on the FMA pipe (Cody-Waite + cubic) because 128x128
// MUFU.EX2 per tile-block cost exactly the 1024 clocks the two products take.
// TMEM columns: S_t / P_t at 128 t, O_t at 256 + 128 t.
// =====================================================================================================================
constexpr int A3_BN = 128, A3_SLOTS = 5;
constexpr int A3_SLOT_BYTES = A3_BN * AT_D * 2;          // 32 KB: one K or one V block
constexpr int A3_OFF_RING = 2 * AT_Q_BYTES;
constexpr int A3_OFF_BAR = A3_OFF_RING + A3_SLOTS * A3_SLOT_BYTES;
constexpr int A3_SMEM = A3_OFF_BAR + 256 + 1024;
constexpr int A3_REGS_PRODUCER = 40, A3_REGS_SOFTMAX = 232;   // (2*232 + 40) * 128 = 168 * 384


template <uint32_t POLY_MASK>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tcgen05_v3(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A3_OFF_BAR);
  uint64_t* q_full = bars;                       // 1
  uint64_t* ring_full = bars + 1;                // A3_SLOTS
  uint64_t* ring_empty = ring_full + A3_SLOTS;   // A3_SLOTS
  uint64_t* s_full = ring_empty + A3_SLOTS;      // 2
  uint64_t* p_full = s_full + 2;                 // [tile][half] -> 4: P is handed over in two 64-key halves so that the first
                                                 // half of the PV product runs while the second half is still being exponentiated
  uint64_t* o_done = p_full + 4;                 // 2 (final PV of each tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * AT_BM);
  const int nblk = (p.Lk + A3_BN - 1) / A3_BN;
  const int col0 = head * AT_D;

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int s = 0; s < A3_SLOTS; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&o_done[t], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<A3_REGS_PRODUCER>();
    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer: Q, then K0 V0 K1 V1 ...
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * AT_Q_BYTES);
        for (int t = 0; t < 2; ++t)
          for (int half = 0; half < 2; ++half)
            tma_load_2d(smem + t * AT_Q_BYTES + half * (AT_Q_BYTES / 2), &tmQ, q_full, col0 + half * 64, q0 + t * AT_BM);
      }
      __syncwarp();
      int slot = 0; uint32_t phase = 0;
      for (int i = 0; i < 2 * nblk; ++i) {
        mbar_wait_trap(&ring_empty[slot], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = smem + A3_OFF_RING + slot * A3_SLOT_BYTES;
          const CUtensorMap* tm = (i & 1) ? &tmV : &tmK;
          mbar_arrive_expect_tx(&ring_full[slot], A3_SLOT_BYTES);
          tma_load_2d(dst, tm, &ring_full[slot], col0, (i >> 1) * A3_BN);
          tma_load_2d(dst + A3_SLOT_BYTES / 2, tm, &ring_full[slot], col0 + 64, (i >> 1) * A3_BN);
        }
        __syncwarp();
        if (++slot == A3_SLOTS) { slot = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc_s = umma_idesc(1, AT_BM, A3_BN, 0, 0);   // S = Q K^T, both K-major
      constexpr uint32_t idesc_o = umma_idesc(1, AT_BM, AT_D, 0, 1);    // O += P V, V MN-major
      const uint32_t q_addr = smem_u32(smem);
      const uint32_t ring_addr = smem_u32(smem + A3_OFF_RING);
      auto issue_s = [&](int t, int slot) {
        const uint32_t k_addr = ring_addr + slot * A3_SLOT_BYTES;
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks) {
          uint64_t da = umma_desc_sw128(q_addr + t * AT_Q_BYTES + (ks >> 2) * (AT_Q_BYTES / 2) + (ks & 3) * 32, 16, 1024);
          uint64_t db = umma_desc_sw128(k_addr + (ks >> 2) * (A3_SLOT_BYTES / 2) + (ks & 3) * 32, 16, 1024);
          umma_f16_ss(tmem_base + t * A3_BN, da, db, idesc_s, ks != 0);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int slot, int j, int half) {
        const uint32_t v_addr = ring_addr + slot * A3_SLOT_BYTES;
#pragma unroll
        for (int ks = 4 * half; ks < 4 * half + 4; ++ks) {
          uint64_t db = umma_desc_sw128(v_addr + ks * 2048, A3_SLOT_BYTES / 2, 1024);
          umma_f16_ts(tmem_base + AT2_TMEM_O + t * AT_D, tmem_base + t * A3_BN + ks * 8, db, idesc_o, (j | ks) != 0);
        }
      };
      int slot = 0; uint32_t phase = 0;          // ring position of the next block to consume
      auto advance = [&]() { if (++slot == A3_SLOTS) { slot = 0; phase ^= 1; } };
      mbar_wait_trap(q_full, 0);
      mbar_wait_trap(&ring_full[slot], phase);        // K0
      tc_fence_after();
      if (elect_one()) { issue_s(0, slot); issue_s(1, slot); umma_commit(&ring_empty[slot]); }
      __syncwarp();
      advance();
      for (int j = 0; j < nblk; ++j) {
        const bool more = (j + 1 < nblk);
        const int v_slot = slot; const uint32_t v_phase = phase;
        advance();
        const int k_slot = slot; const uint32_t k_phase = phase;
        if (more) advance();
        mbar_wait_trap(&ring_full[v_slot], v_phase);                 // V_j
        mbar_wait_trap(&p_full[0], j & 1);
        tc_fence_after();
        if (elect_one()) issue_pv(0, v_slot, j, 0);
        __syncwarp();
        mbar_wait_trap(&p_full[1], j & 1);
        tc_fence_after();
        if (elect_one()) issue_pv(0, v_slot, j, 1);
        __syncwarp();
        if (more) {
          mbar_wait_trap(&ring_full[k_slot], k_phase);               // K_{j+1}
          tc_fence_after();
          if (elect_one()) issue_s(0, k_slot);
          __syncwarp();
        } else if (elect_one()) {
          umma_commit(&o_done[0]);
        }
        mbar_wait_trap(&p_full[2], j & 1);
        tc_fence_after();
        if (elect_one()) issue_pv(1, v_slot, j, 0);
        __syncwarp();
        mbar_wait_trap(&p_full[3], j & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(1, v_slot, j, 1);
          umma_commit(&ring_empty[v_slot]);
          if (more) { issue_s(1, k_slot); umma_commit(&ring_empty[k_slot]); }
          else umma_commit(&o_done[1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------------- softmax + epilogue
    setmaxnreg_inc<A3_REGS_SOFTMAX>();
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane_id();
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_addr + t * A3_BN;
    const uint32_t o_tmem = tmem_base + lane_addr + AT2_TMEM_O + t * AT_D;
    const float c = p.scale_log2;
    const uint64_t c2 = pack2(c, c);
    float m_ref = 0.f, l = 0.f;
    // one key block; MASKED = the last, partial block (scores of keys >= Lk become -inf).  Two instantiations behind a
    // warp-uniform branch: as a predicate inside one body the mask costs an ISETP + SEL per score in EVERY block
    auto softmax_block = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      mbar_wait_trap(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t r[128];
      {
        uint32_t (&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
        uint32_t (&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
        uint32_t (&r2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[64]);
        uint32_t (&r3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[96]);
        tmem_ld_32x32b_x32(s_tmem, r0);
        tmem_ld_32x32b_x32(s_tmem + 32, r1);
        tmem_ld_32x32b_x32(s_tmem + 64, r2);
        tmem_ld_32x32b_x32(s_tmem + 96, r3);
        tmem_ld_wait();
      }
      if (MASKED) {
        const int valid = p.Lk - j * A3_BN;
#pragma unroll
        for (int i = 0; i < A3_BN; ++i) if (i >= valid) r[i] = 0xff800000u;   // -inf
      }
      // four independent max chains (3-input max), then combine
      float mx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        mx[u] = fmax3(__uint_as_float(r[32 * u]), __uint_as_float(r[32 * u + 1]), __uint_as_float(r[32 * u + 2]));
#pragma unroll
        for (int i = 3; i < 31; i += 2) mx[u] = fmax3(mx[u], __uint_as_float(r[32 * u + i]), __uint_as_float(r[32 * u + i + 1]));
        mx[u] = fmaxf(mx[u], __uint_as_float(r[32 * u + 31]));
      }
      const float m_blk = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]) * c;
      float alpha = 1.0f;
      bool grow = false;
      if (j == 0) {
        m_ref = m_blk;
      } else if (m_blk - m_ref > AT_RESCALE_THRESHOLD) {
        alpha = ex2(m_ref - m_blk);
        m_ref = m_blk;
        grow = true;
      }
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // O_t is quiescent: S_t(j) was issued after PV_t(j-1) and the MMA pipe retires in order
#pragma unroll 1
        for (int cc = 0; cc < AT_D; cc += 32) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_tmem + cc, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(o_tmem + cc, o);
        }
        tmem_st_wait();
      }
      const uint64_t nm2 = pack2(-m_ref, -m_ref);
      uint64_t sum2 = pack2(0.f, 0.f), sum2b = pack2(0.f, 0.f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int e = 64 * h + 2 * i;
          const uint64_t x2 = ffma2(pack2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, nm2);
          float e0, e1;
          if (((POLY_MASK >> (i & 7)) & 1u) && !MASKED) {
            exp2_poly2(x2, e0, e1);
          } else {
            float x0, x1; unpack2(x2, x0, x1);
            e0 = ex2(x0); e1 = ex2(x1);
          }
          if (i & 1) sum2b = fadd2(sum2b, pack2(e0, e1)); else sum2 = fadd2(sum2, pack2(e0, e1));
          pk[i] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x32(s_tmem + 32 * h, pk);     // P_t: bf16 pairs over the first 64 columns of S_t
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&p_full[2 * t + h]);
      }
      float s_lo, s_hi; unpack2(fadd2(sum2, sum2b), s_lo, s_hi);
      l = l * alpha + (s_lo + s_hi);
    };
    const int nfull = p.Lk / A3_BN;               // blocks whose 128 keys all exist
#pragma unroll 1
    for (int j = 0; j < nfull; ++j) softmax_block(j, std::false_type{});
    if (nfull < nblk) softmax_block(nfull, std::true_type{});
    mbar_wait_trap(&o_done[t], 0);
    tc_fence_after();
    const int q = q0 + t * AT_BM + row;
    const float inv_l = 1.0f / l;
#pragma unroll 1
    for (int cc = 0; cc < AT_D; cc += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_tmem + cc, o);
      tmem_ld_wait();
      if (q < p.Lq) {
        bf16* dst = (p.n_peers > 0 ? p.out_peer[q / p.rows_per_peer] + static_cast<size_t>(q % p.rows_per_peer) * p.ldo
                                   : p.out + static_cast<size_t>(q) * p.ldo) + col0 + cc;
        const bf16* add = p.add_in ? p.add_in + static_cast<size_t>(q) * p.ld_add + col0 + cc : nullptr;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = __uint_as_float(o[i + u]) * inv_l;
          if (add) {
            uint4 a = *reinterpret_cast<const uint4*>(add + i);
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float2 f = __bfloat1622float2(a2[u]);
              w[2 * u] = __fadd_rn(bf16_round(w[2 * u]), f.x);
              w[2 * u + 1] = __fadd_rn(bf16_round(w[2 * u + 1]), f.y);
            }
          }
          uint4 v = make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]),
                               pack_bf16x2(w[6], w[7]));
          *reinterpret_cast<uint4*>(dst + i) = v;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wf

static int attention_launch(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                            const void* add_in, int ld_add, int Lq, int Lk, int heads, float softmax_scale,
                            void* const* out_peers, int n_peers, int rows_per_peer, void* stream) {
  using namespace wf;
  WF_REQUIRE(q && k && v && (out || n_peers > 0), "wf_attention_bf16: null pointer");
  WF_REQUIRE(n_peers >= 0 && n_peers <= AT_MAX_PEERS, "wf_attention_bf16_peers: at most 8 peers");
  WF_REQUIRE(Lq > 0 && Lk > 0 && heads > 0, "wf_attention_bf16: empty problem");
  WF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ld_add % 8 == 0,
             "wf_attention_bf16: leading dimensions must be multiples of 8");
  WF_REQUIRE(ldq >= heads * AT_D && ldk >= heads * AT_D && ldv >= heads * AT_D && ldo >= heads * AT_D,
             "wf_attention_bf16: leading dimension smaller than heads*128");
  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* m, const void* base, int ld, int rows, uint32_t box_rows) {
    uint64_t dims[2] = {static_cast<uint64_t>(heads) * AT_D, static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
    uint32_t box[2] = {64, box_rows};
    return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  if ((rc = mk(&tmQ, q, ldq, Lq, AT_BM))) return rc;
  if ((rc = mk(&tmK, k, ldk, Lk, AT_BN))) return rc;
  if ((rc = mk(&tmV, v, ldv, Lk, AT_BN))) return rc;
  // WF_ATTN=2 (1 is an alias): v2 (double-buffered S, packed fp32x2 softmax, P through smem);
  // 3: v2 + polynomial ex2 offload; 4: v2 with P kept in TMEM as the A operand of the PV product; 5/6: 4 + polynomial ex2;
  // 7/8/9: v3 (128-key blocks, P handed over in halves) with 0 / 2 / 3 of every 8 exponential pairs on the FMA pipe.
  // Default 9: 26.0 M SM cycles per launch at the 480p shape against 31.9 M for variant 4 (tensor pipe 71 % vs 58 % active).
  static int variant = 0;
  static bool forced = false;        // WF_ATTN set: every call runs that variant (development / A-B measurements)
  if (variant == 0) {
    const char* e = getenv("WF_ATTN");
    forced = e != nullptr;
    int vv = e ? atoi(e) : 9;
    variant = (vv < 1 || vv > 9) ? 9 : vv;
  }
  static PerDeviceOnce once;         // the shared-memory opt-in is a per-device function attribute
  rc = once.run([] {
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v3<0x00u>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v3<0x88u>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v3<0xa4u>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v2<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v2<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v2<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v2<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_tcgen05_v2<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)));
    return static_cast<int>(WF_OK);
  });
  if (rc) return rc;
  AttnArgs args{};
  args.Lq = Lq; args.Lk = Lk; args.out = static_cast<bf16*>(out); args.ldo = ldo;
  args.add_in = static_cast<const bf16*>(add_in); args.ld_add = ld_add; args.scale_log2 = softmax_scale * 1.4426950408889634f;
  args.n_peers = n_peers; args.rows_per_peer = rows_per_peer > 0 ? rows_per_peer : 1;
  for (int i = 0; i < n_peers; ++i) args.out_peer[i] = static_cast<bf16*>(out_peers[i]);
  dim3 grid((Lq + 2 * AT_BM - 1) / (2 * AT_BM), heads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // short key sequences whose tail would leave a 128-key block nearly empty (the 257 image tokens of the cross-attention)
  // run the 64-key kernel
  const bool short_tail = Lk < 2048 && (Lk % A3_BN) != 0 && (Lk % A3_BN) <= AT_BN;
  if (variant >= 7 && short_tail && !forced) {
    attention_tcgen05_v2<0, true><<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, args);
    WF_LAUNCH_OK();
    return WF_OK;
  }
  if (variant >= 7) {
    // 7..9: 128-key blocks (v3); K and V boxes are 128 rows.  8: 2 of 8 exponential pairs on the FMA pipe, 9: 3 of 8
    if ((rc = mk(&tmK, k, ldk, Lk, A3_BN))) return rc;
    if ((rc = mk(&tmV, v, ldv, Lk, A3_BN))) return rc;
    if (variant == 7) attention_tcgen05_v3<0x00u><<<grid, AT_THREADS, A3_SMEM, st>>>(tmQ, tmK, tmV, args);
    else if (variant == 8) attention_tcgen05_v3<0x88u><<<grid, AT_THREADS, A3_SMEM, st>>>(tmQ, tmK, tmV, args);
    else attention_tcgen05_v3<0xa4u><<<grid, AT_THREADS, A3_SMEM, st>>>(tmQ, tmK, tmV, args);
    WF_LAUNCH_OK();
    return WF_OK;
  }
  if (variant <= 2) attention_tcgen05_v2<0, false><<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, args);
  else if (variant == 3) attention_tcgen05_v2<1, false><<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, args);
  else if (variant == 5) attention_tcgen05_v2<1, true><<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, args);   // 4 + 25% polynomial ex2
  else if (variant == 6) attention_tcgen05_v2<2, true><<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, args);   // 4 + 50% polynomial ex2
  else attention_tcgen05_v2<0, true><<<grid, AT_THREADS, AT_SMEM, st>>>(tmQ, tmK, tmV, args);   // 4: P through TMEM
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_attention_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                                 int ldo, const void* add_in, int ld_add, int Lq, int Lk, int heads,
                                 float softmax_scale, void* stream) {
  return attention_launch(q, ldq, k, ldk, v, ldv, out, ldo, add_in, ld_add, Lq, Lk, heads, softmax_scale, nullptr, 0, 0, stream);
}

extern "C" int wf_attention_bf16_peers(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                                       void* const* out_peers, int n_peers, int rows_per_peer, int ldo, int Lq, int Lk,
                                       int heads, float softmax_scale, void* stream) {
  WF_REQUIRE(out_peers && n_peers >= 1 && rows_per_peer > 0, "wf_attention_bf16_peers: bad peer table");
  WF_REQUIRE(static_cast<long long>(n_peers) * rows_per_peer >= Lq, "wf_attention_bf16_peers: the peers do not cover every query row");
  for (int i = 0; i < n_peers; ++i) WF_REQUIRE(out_peers[i] != nullptr, "wf_attention_bf16_peers: null peer pointer");
  return attention_launch(q, ldq, k, ldk, v, ldv, nullptr, ldo, nullptr, 0, Lq, Lk, heads, softmax_scale, out_peers, n_peers,
                          rows_per_peer, stream);
}
