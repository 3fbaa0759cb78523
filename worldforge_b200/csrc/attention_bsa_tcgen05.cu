// worldforge_b200 - LongCat-Video's block-sparse attention (the 720p refine pass) on tcgen05 / TMEM / TMA.
//
// Reference: longcat_video/block_sparse_attention/bsa_interface.py:612-659 (flash_attn_bsa_3d: re-order the tokens
// chunk-major, gate, sparse attention, re-order back) and flash_attn_bsa_varlen_mask.py:174-285 (the Triton forward
// kernel: per 64- or 128-token query chunk, online softmax over the SELECTED key chunks only).
//
//   O[q, h, :] = softmax_{k in selected chunks of chunk(q)}( Q[q,h,:] . K[k,h,:] / sqrt(128) ) V[k,h,:]
//
// Nothing is re-ordered in memory.  q, k, v, out stay the [T*H*W, heads*128] matrices of the DiT in (t,h,w) token
// order; a chunk (ct x ch x cw tokens) is one 4-D TMA box of the (d, w, h, t) view, which lands in shared memory in
// exactly the chunk-major order the reference builds with two permute+contiguous passes per tensor.
//
// CTA = two 128-row query tiles of one head (ping-pong between softmax and MMA as in attention_tcgen05_v2), keys walked
// in steps of 64 through per-tile K and V rings:
//   chunk = 128 tokens: a tile is one query chunk; a step is half of a selected key chunk.
//   chunk =  64 tokens: a tile is TWO query chunks (rows 0-63 / 64-127) with different selections; a step stacks 32
//                       keys of the upper chunk's selection on 32 keys of the lower chunk's.  S is 128x64 with the
//                       useful quadrants on the diagonal; each row's softmax reads only its own 32 columns and writes
//                       zeros into the other 32 of P, so the PV product adds nothing from the foreign quadrant.  The
//                       tensor pipe runs M=128 instructions either way (an M=64 MMA costs the same issue time).
//   warp 0 K producer | warp 1 MMA issuer | warp 2 TMEM allocator | warp 3 V producer | warps 4-7 / 8-11 softmax
// P stays in TMEM (A operand of the PV product); O is rescaled lazily (row max grown by more than 2^8).
#include "attn_math.cuh"
#include "common.cuh"
#include "host_util.h"

namespace wf {

constexpr int BS_BM = 128, BS_BN = 64, BS_D = 128, BS_STAGES = 5;
constexpr int BS_Q_BYTES = BS_BM * BS_D * 2;      // 32 KB per query tile
constexpr int BS_T_BYTES = BS_BN * BS_D * 2;      // 16 KB: one K (or V) step
constexpr int BS_OFF_K = 2 * BS_Q_BYTES;
constexpr int BS_OFF_V = BS_OFF_K + BS_STAGES * BS_T_BYTES;
constexpr int BS_OFF_BAR = BS_OFF_V + BS_STAGES * BS_T_BYTES;
constexpr int BS_SMEM = BS_OFF_BAR + 512 + 1024;
constexpr int BS_THREADS = 384;
constexpr uint32_t BS_TMEM_O = 256;
constexpr float BS_RESCALE_THRESHOLD = 8.0f;

struct BsaArgs {
  const int32_t* idx;    // [heads][nq_chunks][max_sel]
  const int32_t* lens;   // [heads][nq_chunks] or null (every list has max_sel entries)
  int max_sel;
  int nq_chunks;
  int Hq, Wq;            // query grid is [Tq][Hq][Wq] tokens; chunks tile it
  int nhq, nwq;          // chunks per axis of the query grid
  int nhk, nwk;          // ... of the key grid
  int ct, ch, cw;        // chunk shape
  bf16* out; int ldo;
  float scale_log2;
};

struct ChunkList {
  const int32_t* p;
  int len;
};

// one lane-parallel window of 16 list entries for the two sources (a, b) of a tile; entry e of source s sits in lane s*16 + (e & 15)
__device__ __forceinline__ int load_window(const ChunkList& a, const ChunkList& b, int e0) {
  const int l = lane_id();
  const ChunkList& s = (l < 16) ? a : b;
  int e = e0 + (l & 15);
  if (s.len == 0) return 0;
  if (e >= s.len) e = s.len - 1;      // exhausted list: re-read a valid chunk, its probabilities are masked to zero
  return __ldg(s.p + e);
}

template <int CHUNK>
__global__ void __launch_bounds__(BS_THREADS, 1)
attention_bsa_tcgen05(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, BsaArgs p) {
  constexpr int SRC = (CHUNK == 64) ? 2 : 1;            // query chunks (= key selections) per 128-row tile
  constexpr int HALF_ROWS = BS_BN / SRC;                // key rows one source contributes to a step
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BS_OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + BS_STAGES;
  uint64_t* v_full = k_empty + BS_STAGES;
  uint64_t* v_empty = v_full + BS_STAGES;
  uint64_t* s_full = v_empty + BS_STAGES;    // [tile][buf] -> 4
  uint64_t* p_full = s_full + 4;             // [tile][buf] -> 4 (a tile may signal P(j) and P(j+1) before the issuer looks:
                                             // one barrier per S buffer, or the phase parity would wrap)
  uint64_t* o_done = p_full + 4;             // 2: one completion per PV product; a waiter may be at most ONE phase behind
  uint64_t* o_final = o_done + 2;            // 2: completes once, after the tile's last PV product (the epilogue's wait: after
                                             // the last softmax step only PV(nsteps-3) is known complete, and a parity wait two
                                             // phases behind is satisfied by the wrong completion - with 4-6 steps that is most of O)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);

  const int warp = threadIdx.x >> 5;
  const int head = blockIdx.y;
  const int col0 = head * BS_D;
  // query chunks of this CTA: tile t covers chunks (2*blockIdx.x + t)*SRC + {0 .. SRC-1}
  ChunkList lists[2][2];
  int nsteps = 0;
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int c = (2 * blockIdx.x + t) * SRC + (SRC == 2 ? s : 0);
      ChunkList& L = lists[t][s];
      if (c < p.nq_chunks) {
        const size_t row = static_cast<size_t>(head) * p.nq_chunks + c;
        L.p = p.idx + row * p.max_sel;
        L.len = p.lens ? min(__ldg(p.lens + row), p.max_sel) : p.max_sel;
      } else {
        L.p = p.idx; L.len = 0;
      }
      nsteps = max(nsteps, 2 * L.len);
    }

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int s = 0; s < BS_STAGES; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
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

  if ((warp == 0 || warp == 3) && nsteps > 0) {
    // ------------------------------------------------------------ K producer (warp 0) / V producer (warp 3)
    const bool is_k = warp == 0;
    if (is_k && elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * BS_Q_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int s = 0; s < SRC; ++s) {
          const int c = (2 * blockIdx.x + t) * SRC + s;       // chunks past the end read out of bounds: zero fill
          const int cw_i = c % p.nwq, ch_i = (c / p.nwq) % p.nhq, ct_i = c / (p.nwq * p.nhq);
          for (int half = 0; half < 2; ++half)
            tma_load_4d(smem + t * BS_Q_BYTES + half * (BS_Q_BYTES / 2) + s * (CHUNK * 128), &tmQ, q_full, col0 + half * 64,
                        cw_i * p.cw, ch_i * p.ch, ct_i * p.ct);
        }
    }
    __syncwarp();
    const CUtensorMap* tm = is_k ? &tmK : &tmV;
    uint64_t* full = is_k ? k_full : v_full;
    uint64_t* empty = is_k ? k_empty : v_empty;
    uint8_t* base = smem + (is_k ? BS_OFF_K : BS_OFF_V);
    int win[2] = {0, 0};
    for (int j = 0; j < nsteps; ++j) {
      const int e = j >> 1;
      if ((j & 31) == 0) {
        win[0] = load_window(lists[0][0], lists[0][1], e);
        win[1] = load_window(lists[1][0], lists[1][1], e);
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int seq = 2 * j + t;
        const int stage = seq % BS_STAGES;
        const int ca = __shfl_sync(0xffffffffu, win[t], e & 15);
        const int cb = __shfl_sync(0xffffffffu, win[t], 16 + (e & 15));
        mbar_wait(&empty[stage], ((seq / BS_STAGES) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* dst = base + stage * BS_T_BYTES;
          mbar_arrive_expect_tx(&full[stage], BS_T_BYTES);
#pragma unroll
          for (int s = 0; s < SRC; ++s) {
            const int c = s ? cb : ca;
            const int cw_i = c % p.nwk, ch_i = (c / p.nwk) % p.nhk, ct_i = c / (p.nwk * p.nhk);
            const int t0 = ct_i * p.ct + (j & 1) * (p.ct / 2);
            for (int half = 0; half < 2; ++half)
              tma_load_4d(dst + half * (BS_T_BYTES / 2) + s * (HALF_ROWS * 128), tm, &full[stage], col0 + half * 64, cw_i * p.cw,
                          ch_i * p.ch, t0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && nsteps > 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc(1, BS_BM, BS_BN, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc(1, BS_BM, BS_D, 0, 1);
    const uint32_t q_addr = smem_u32(smem);
    auto issue_s = [&](int t, int j) {
      const int seq = 2 * j + t, stage = seq % BS_STAGES, buf = j & 1;
      const uint32_t k_addr = smem_u32(smem + BS_OFF_K + stage * BS_T_BYTES);
#pragma unroll
      for (int ks = 0; ks < BS_D / 16; ++ks) {
        uint64_t da = umma_desc_sw128(q_addr + t * BS_Q_BYTES + (ks >> 2) * (BS_Q_BYTES / 2) + (ks & 3) * 32, 16, 1024);
        uint64_t db = umma_desc_sw128(k_addr + (ks >> 2) * (BS_T_BYTES / 2) + (ks & 3) * 32, 16, 1024);
        umma_f16_ss(tmem_base + (2 * t + buf) * BS_BN, da, db, idesc_s, ks != 0);
      }
      umma_commit(&s_full[2 * t + buf]);
      umma_commit(&k_empty[stage]);
    };
    auto issue_pv = [&](int t, int j) {
      const int seq = 2 * j + t, stage = seq % BS_STAGES;
      const uint32_t v_addr = smem_u32(smem + BS_OFF_V + stage * BS_T_BYTES);
#pragma unroll
      for (int ks = 0; ks < BS_BN / 16; ++ks) {
        uint64_t db = umma_desc_sw128(v_addr + ks * 2048, BS_T_BYTES / 2, 1024);
        umma_f16_ts(tmem_base + BS_TMEM_O + t * BS_D, tmem_base + (2 * t + (j & 1)) * BS_BN + ks * 8, db, idesc_o, (j | ks) != 0);
      }
      umma_commit(&o_done[t]);
      umma_commit(&v_empty[stage]);
    };
    auto wait_k = [&](int t, int j) {
      const int seq = 2 * j + t;
      mbar_wait(&k_full[seq % BS_STAGES], (seq / BS_STAGES) & 1);
    };
    mbar_wait(q_full, 0);
    for (int t = 0; t < 2; ++t) {
      wait_k(t, 0);
      tc_fence_after();
      if (elect_one()) issue_s(t, 0);
      __syncwarp();
    }
    for (int j = 0; j < nsteps; ++j) {
      if (j + 1 < nsteps) {
        // S(j+1) goes to the other S buffer; its previous content P(j-1) was consumed by PV(j-1), issued earlier
        for (int t = 0; t < 2; ++t) {
          wait_k(t, j + 1);
          tc_fence_after();
          if (elect_one()) issue_s(t, j + 1);
          __syncwarp();
        }
      }
      for (int t = 0; t < 2; ++t) {
        const int seq = 2 * j + t;
        mbar_wait(&v_full[seq % BS_STAGES], (seq / BS_STAGES) & 1);
        mbar_wait(&p_full[2 * t + (j & 1)], (j >> 1) & 1);
        tc_fence_after();
        if (elect_one()) { issue_pv(t, j); if (j + 1 == nsteps) umma_commit(&o_final[t]); }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax: one thread per query row
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane_id();
    const int src = (SRC == 2) ? (qd >> 1) : 0;                 // which query chunk of the tile this row belongs to
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_addr + BS_TMEM_O + t * BS_D;
    const int my_steps = 2 * lists[t][src].len;
    const float c = p.scale_log2;
    const uint64_t c2 = pack2(c, c);
    float m_ref = 0.f, l = 0.f;
    bool first = true;
    for (int j = 0; j < nsteps; ++j) {
      const int buf = j & 1;
      mbar_wait(&s_full[2 * t + buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_tmem = tmem_base + lane_addr + (2 * t + buf) * BS_BN;
      const bool live = j < my_steps;                            // warp-uniform
      constexpr int NV = BS_BN / SRC;                            // score columns this row owns in a step
      uint32_t r[NV];
      {
        uint32_t (&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
        tmem_ld_32x32b_x32(s_tmem + (SRC == 2 ? src * 32 : 0), r0);
        if (SRC == 1) {
          uint32_t (&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[NV - 32]);
          tmem_ld_32x32b_x32(s_tmem + 32, r1);
        }
        tmem_ld_wait();
      }
      uint32_t pk[32];
      float alpha = 1.0f;
      bool grow = false;
      if (live) {
        float mx = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
#pragma unroll
        for (int i = 3; i < NV - 1; i += 2) mx = fmax3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        mx = fmaxf(mx, __uint_as_float(r[NV - 1]));
        const float m_blk = mx * c;
        if (first) {
          m_ref = m_blk;
        } else if (m_blk - m_ref > BS_RESCALE_THRESHOLD) {
          alpha = ex2(m_ref - m_blk);
          m_ref = m_blk;
          grow = true;
        }
        const uint64_t nm2 = pack2(-m_ref, -m_ref);
        uint64_t sum2 = pack2(0.f, 0.f);
        uint32_t e[NV / 2];
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
          const uint64_t x2 = ffma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nm2);
          float x0, x1; unpack2(x2, x0, x1);
          const float e0 = ex2(x0), e1 = ex2(x1);
          sum2 = fadd2(sum2, pack2(e0, e1));
          e[i] = pack_bf16x2(e0, e1);
        }
        float s_lo, s_hi; unpack2(sum2, s_lo, s_hi);
        l = l * alpha + (s_lo + s_hi);
        if (SRC == 2) {
          if (src == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { pk[i] = e[i % (NV / 2)]; pk[16 + i] = 0u; }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) { pk[i] = 0u; pk[16 + i] = e[i % (NV / 2)]; }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) pk[i] = e[i % (NV / 2)];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) pk[i] = 0u;
      }
      const bool any_grow = __any_sync(0xffffffffu, grow);
      if (!first && any_grow) {
        // S(j) complete implies PV(j-2) complete (in-order MMA pipe), so this parity wait can only mean PV(j-1)
        mbar_wait(&o_done[t], (j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < BS_D; cc += 32) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_tmem + cc, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(o_tmem + cc, o);
        }
        tmem_st_wait();
      }
      if (live) first = false;
      tmem_st_32x32b_x32(s_tmem, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&p_full[2 * t + buf]);
    }
    if (nsteps > 0) {
      mbar_wait(&o_final[t], 0);
      tc_fence_after();
    }
    // this row's token: chunk c, position i inside the chunk in (t,h,w) order
    const int c_idx = (2 * blockIdx.x + t) * SRC + src;
    const int i_in = (SRC == 2) ? (row & 63) : row;
    const bool valid = c_idx < p.nq_chunks;
    const int cw_i = c_idx % p.nwq, ch_i = (c_idx / p.nwq) % p.nhq, ct_i = c_idx / (p.nwq * p.nhq);
    const int tw = i_in % p.cw, th = (i_in / p.cw) % p.ch, tt = i_in / (p.cw * p.ch);
    const size_t tok = (static_cast<size_t>(ct_i * p.ct + tt) * p.Hq + (ch_i * p.ch + th)) * p.Wq + (cw_i * p.cw + tw);
    const float inv_l = (my_steps > 0) ? 1.0f / l : 0.f;
#pragma unroll 1
    for (int cc = 0; cc < BS_D; cc += 32) {
      uint32_t o[32];
      if (my_steps > 0) {
        tmem_ld_32x32b_x32(o_tmem + cc, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;     // an empty selection gives zeros (the reference's acc = 0, l = 1)
      }
      if (valid) {
        bf16* dst = p.out + tok * p.ldo + col0 + cc;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v = make_uint4(pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l),
                               pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l),
                               pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l),
                               pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l));
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

// --------------------------------------------------------------------------------------------- gating
// mean over the tokens of each chunk, fp32 accumulation, rounded to bf16 (bsa_interface.py:176-186 on bf16 tensors)
__global__ void bsa_mean_pool_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ out, int H, int W, int nh, int nw,
                                     int ct, int ch, int cw, int nchunks) {
  const int c = blockIdx.x, head = blockIdx.y;
  const int cw_i = c % nw, ch_i = (c / nw) % nh, ct_i = c / (nw * nh);
  const int ntok = ct * ch * cw;
  // 128 threads: thread d owns channel d; tokens walked in the chunk's (t,h,w) order (the reference's summation order
  // is torch's reduction tree; fp32 sums of <= 128 bf16 values differ from it by fp32 ulps only)
  const int d = threadIdx.x;
  float acc = 0.f;
  for (int i = 0; i < ntok; ++i) {
    const int tw = i % cw, th = (i / cw) % ch, tt = i / (cw * ch);
    const size_t tok = (static_cast<size_t>(ct_i * ct + tt) * H + (ch_i * ch + th)) * W + (cw_i * cw + tw);
    acc += __bfloat162float(x[tok * ldx + head * BS_D + d]);
  }
  out[(static_cast<size_t>(head) * nchunks + c) * BS_D + d] = __float2bfloat16_rn(acc / static_cast<float>(ntok));
}

// order-preserving map of a bf16 bit pattern to uint16 (larger value -> larger key)
__device__ __forceinline__ uint32_t bf16_key(bf16 v) {
  const uint32_t b = __bfloat16_as_ushort(v);
  return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u);
}

constexpr int SEL_ROWS = 8, SEL_THREADS = 256;

// scores = bf16(q_cmp . k_cmp) (bsa_interface.py:188-192), then the n_sel largest per row (:221-232), written in
// ascending chunk order; ties at the threshold go to the lower chunk index (torch.topk leaves that unspecified).
// One CTA = 8 query chunks of one head (k_cmp rows are read once for the 8); one warp selects one row by a two-pass
// radix select over the 16-bit keys.
__global__ void __launch_bounds__(SEL_THREADS)
bsa_select_topk_kernel(const bf16* __restrict__ q_cmp, const bf16* __restrict__ k_cmp, int32_t* __restrict__ idx, int Nq, int Nk,
                       int n_sel) {
  extern __shared__ uint8_t sel_smem[];
  float* qs = reinterpret_cast<float*>(sel_smem);                              // [SEL_ROWS][128]
  uint32_t* hist = reinterpret_cast<uint32_t*>(qs + SEL_ROWS * BS_D);          // [SEL_ROWS][256]
  uint16_t* keys = reinterpret_cast<uint16_t*>(hist + SEL_ROWS * 256);         // [SEL_ROWS][Nk]
  const int head = blockIdx.y, r0 = blockIdx.x * SEL_ROWS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < SEL_ROWS * BS_D; i += SEL_THREADS) {
    const int r = r0 + i / BS_D;
    qs[i] = r < Nq ? __bfloat162float(q_cmp[(static_cast<size_t>(head) * Nq + r) * BS_D + (i % BS_D)]) : 0.f;
  }
  for (int i = tid; i < SEL_ROWS * 256; i += SEL_THREADS) hist[i] = 0;
  __syncthreads();
  for (int n = tid; n < Nk; n += SEL_THREADS) {
    const uint4* kr = reinterpret_cast<const uint4*>(k_cmp + (static_cast<size_t>(head) * Nk + n) * BS_D);
    float acc[SEL_ROWS];
#pragma unroll
    for (int r = 0; r < SEL_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int v = 0; v < BS_D / 8; ++v) {
      const uint4 kk = __ldg(kr + v);
      const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __bfloat1622float2(k2[u]);
#pragma unroll
        for (int r = 0; r < SEL_ROWS; ++r) {
          acc[r] = fmaf(qs[r * BS_D + v * 8 + 2 * u], f.x, acc[r]);
          acc[r] = fmaf(qs[r * BS_D + v * 8 + 2 * u + 1], f.y, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SEL_ROWS; ++r) keys[r * Nk + n] = static_cast<uint16_t>(bf16_key(__float2bfloat16_rn(acc[r])));
  }
  __syncthreads();
  const int row = r0 + warp;
  if (row >= Nq) return;
  const uint16_t* kr = keys + warp * Nk;
  uint32_t* h = hist + warp * 256;
  // pass 1: high byte
  for (int n = lane; n < Nk; n += 32) atomicAdd(&h[kr[n] >> 8], 1u);
  __syncwarp();
  auto find_bin = [&](int need, int& bin, int& above) {
    // largest bin b such that count(bins > b) < need <= count(bins >= b); lane owns bins 8*lane .. 8*lane+7
    uint32_t cnt[8]; uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { cnt[i] = h[lane * 8 + i]; mine += cnt[i]; }
    uint32_t suffix = mine;                       // inclusive suffix sum over lanes (higher lanes = larger keys)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_down_sync(0xffffffffu, suffix, o);
      if (lane + o < 32) suffix += v;
    }
    uint32_t run = suffix - mine;                 // count in lanes above this one
    int b = -1, ab = 0;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (b < 0 && run + cnt[i] >= static_cast<uint32_t>(need)) { b = lane * 8 + i; ab = run; }
      run += cnt[i];
    }
    const uint32_t has = __ballot_sync(0xffffffffu, b >= 0);
    const int srcl = 31 - __clz(has);             // the highest lane that found it holds the largest such bin
    bin = __shfl_sync(0xffffffffu, b, srcl);
    above = __shfl_sync(0xffffffffu, ab, srcl);
  };
  int hi_bin, above_hi;
  find_bin(n_sel, hi_bin, above_hi);
  __syncwarp();
  for (int i = lane; i < 256; i += 32) h[i] = 0;
  __syncwarp();
  for (int n = lane; n < Nk; n += 32) if ((kr[n] >> 8) == hi_bin) atomicAdd(&h[kr[n] & 0xff], 1u);
  __syncwarp();
  int lo_bin, above_lo;
  find_bin(n_sel - above_hi, lo_bin, above_lo);
  const uint32_t thr = (static_cast<uint32_t>(hi_bin) << 8) | static_cast<uint32_t>(lo_bin);
  int ties_left = n_sel - above_hi - above_lo;    // how many keys equal to thr are taken (lowest chunk index first)
  int32_t* dst = idx + (static_cast<size_t>(head) * Nq + row) * n_sel;
  int written = 0;
  for (int n0 = 0; n0 < Nk; n0 += 32) {
    const int n = n0 + lane;
    const uint32_t key = n < Nk ? kr[n] : 0u;
    const bool gt = n < Nk && key > thr;
    const bool eq = n < Nk && key == thr;
    const uint32_t eq_mask = __ballot_sync(0xffffffffu, eq);
    const bool take = gt || (eq && __popc(eq_mask & ((1u << lane) - 1)) < ties_left);
    const uint32_t take_mask = __ballot_sync(0xffffffffu, take);
    if (take) dst[written + __popc(take_mask & ((1u << lane) - 1))] = n;
    written += __popc(take_mask);
    ties_left -= min(ties_left, __popc(eq_mask));
  }
}


// get_select_indices_cdf / _cdf_topk (bsa_interface.py:234-275): per query chunk the key chunks sorted by softmax(score /
// sqrt(128)) in descending order and the number of them whose cumulative mass stays <= cdf_threshold (torch.searchsorted,
// right=True), optionally floored by the top-k count.  Scores are the same bf16 values as above, so sorting the 16-bit
// keys sorts the weights; ties go to the lower chunk index.  One CTA = 8 query chunks of one head, one warp per row: a
// bitonic sort of (key << 16 | 0xffff - index) in shared memory, then a running sum over the sorted weights.
__device__ __forceinline__ float bf16_from_key(uint32_t key) {
  const uint32_t b = (key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu);
  return __uint_as_float(b << 16);
}

__global__ void __launch_bounds__(SEL_THREADS)
bsa_select_cdf_kernel(const bf16* __restrict__ q_cmp, const bf16* __restrict__ k_cmp, int32_t* __restrict__ idx,
                      int32_t* __restrict__ lens, int Nq, int Nk, int Npad, float cdf_threshold, int n_floor, float scale_log2) {
  extern __shared__ uint8_t sel_smem[];
  float* qs = reinterpret_cast<float*>(sel_smem);                              // [SEL_ROWS][128]
  uint32_t* comp = reinterpret_cast<uint32_t*>(qs + SEL_ROWS * BS_D);          // [SEL_ROWS][Npad]
  const int head = blockIdx.y, r0 = blockIdx.x * SEL_ROWS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < SEL_ROWS * BS_D; i += SEL_THREADS) {
    const int r = r0 + i / BS_D;
    qs[i] = r < Nq ? __bfloat162float(q_cmp[(static_cast<size_t>(head) * Nq + r) * BS_D + (i % BS_D)]) : 0.f;
  }
  __syncthreads();
  for (int n = tid; n < Npad; n += SEL_THREADS) {
    if (n >= Nk) {
#pragma unroll
      for (int r = 0; r < SEL_ROWS; ++r) comp[r * Npad + n] = 0u;              // padding sorts to the end
      continue;
    }
    const uint4* kr = reinterpret_cast<const uint4*>(k_cmp + (static_cast<size_t>(head) * Nk + n) * BS_D);
    float acc[SEL_ROWS];
#pragma unroll
    for (int r = 0; r < SEL_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int v = 0; v < BS_D / 8; ++v) {
      const uint4 kk = __ldg(kr + v);
      const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __bfloat1622float2(k2[u]);
#pragma unroll
        for (int r = 0; r < SEL_ROWS; ++r) {
          acc[r] = fmaf(qs[r * BS_D + v * 8 + 2 * u], f.x, acc[r]);
          acc[r] = fmaf(qs[r * BS_D + v * 8 + 2 * u + 1], f.y, acc[r]);
        }
      }
    }
    // keys are >= 1 for every real score (bf16_key never returns 0 for a finite value: -max maps to 0x0080), padding is 0
#pragma unroll
    for (int r = 0; r < SEL_ROWS; ++r)
      comp[r * Npad + n] = (bf16_key(__float2bfloat16_rn(acc[r])) << 16) | (0xffffu - static_cast<uint32_t>(n));
  }
  __syncthreads();
  const int row = r0 + warp;
  if (row >= Nq) return;
  uint32_t* c = comp + warp * Npad;
  // bitonic sort, descending
  for (int k = 2; k <= Npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < Npad; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t a = c[i], b = c[l];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { c[i] = b; c[l] = a; }
        }
      }
      __syncwarp();
    }
  }
  // softmax statistics over the row (the maximum is the first sorted element)
  const float smax = bf16_from_key(c[0] >> 16);
  float sum = 0.f;
  for (int n = lane; n < Nk; n += 32) sum += exp2f((bf16_from_key(c[n] >> 16) - smax) * scale_log2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  int32_t* dst = idx + (static_cast<size_t>(head) * Nq + row) * Nk;
  float carry = 0.f;
  int count = 0;
  bool open = true;                               // the cumulative mass has not passed the threshold yet
  for (int n0 = 0; n0 < Nk; n0 += 32) {
    const int n = n0 + lane;
    const uint32_t v = n < Nk ? c[n] : 0u;
    if (n < Nk) dst[n] = static_cast<int32_t>(0xffffu - (v & 0xffffu));
    if (open) {
      float w = n < Nk ? exp2f((bf16_from_key(v >> 16) - smax) * scale_log2) * inv : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      const float cdf = carry + w;
      const uint32_t in = __ballot_sync(0xffffffffu, n < Nk && cdf <= cdf_threshold);
      count += __popc(in);
      if (in != 0xffffffffu) open = false;
      carry = __shfl_sync(0xffffffffu, cdf, 31);
    }
  }
  if (lane == 0) lens[static_cast<size_t>(head) * Nq + row] = min(max(count, n_floor), Nk);
}

}  // namespace wf

extern "C" int wf_bsa_mean_pool(const void* x, int ldx, void* out, int T, int H, int W, int ct, int ch, int cw, int heads,
                                void* stream) {
  using namespace wf;
  WF_REQUIRE(x && out, "wf_bsa_mean_pool: null pointer");
  WF_REQUIRE(T > 0 && H > 0 && W > 0 && heads > 0 && ct > 0 && ch > 0 && cw > 0, "wf_bsa_mean_pool: empty problem");
  WF_REQUIRE(T % ct == 0 && H % ch == 0 && W % cw == 0, "wf_bsa_mean_pool: the grid must be a whole number of chunks");
  WF_REQUIRE(ldx >= heads * BS_D, "wf_bsa_mean_pool: leading dimension smaller than heads*128");
  const int nchunks = (T / ct) * (H / ch) * (W / cw);
  bsa_mean_pool_kernel<<<dim3(nchunks, heads), BS_D, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), ldx, static_cast<bf16*>(out), H, W, H / ch, W / cw, ct, ch, cw, nchunks);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_bsa_select_topk(const void* q_cmp, const void* k_cmp, int32_t* idx, int Nq, int Nk, int heads, int n_sel,
                                  void* stream) {
  using namespace wf;
  WF_REQUIRE(q_cmp && k_cmp && idx, "wf_bsa_select_topk: null pointer");
  WF_REQUIRE(Nq > 0 && Nk > 0 && heads > 0, "wf_bsa_select_topk: empty problem");
  WF_REQUIRE(n_sel >= 1 && n_sel <= Nk, "wf_bsa_select_topk: need 1 <= n_sel <= Nk");
  const size_t smem = SEL_ROWS * BS_D * 4 + SEL_ROWS * 256 * 4 + static_cast<size_t>(SEL_ROWS) * Nk * 2;
  WF_REQUIRE(smem <= 200 * 1024, "wf_bsa_select_topk: too many key chunks for the shared-memory score table");
  WF_CUDA_OK(cudaFuncSetAttribute(bsa_select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  bsa_select_topk_kernel<<<dim3((Nq + SEL_ROWS - 1) / SEL_ROWS, heads), SEL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(q_cmp), static_cast<const bf16*>(k_cmp), idx, Nq, Nk, n_sel);
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_bsa_select_cdf(const void* q_cmp, const void* k_cmp, int32_t* idx, int32_t* lens, int Nq, int Nk, int heads,
                                 float cdf_threshold, int n_floor, void* stream) {
  using namespace wf;
  WF_REQUIRE(q_cmp && k_cmp && idx && lens, "wf_bsa_select_cdf: null pointer");
  WF_REQUIRE(Nq > 0 && Nk > 0 && Nk < 65536 && heads > 0, "wf_bsa_select_cdf: empty problem (or more than 65535 key chunks)");
  WF_REQUIRE(n_floor >= 0 && n_floor <= Nk, "wf_bsa_select_cdf: need 0 <= n_floor <= Nk");
  int Npad = 32;
  while (Npad < Nk) Npad <<= 1;
  const size_t smem = SEL_ROWS * BS_D * 4 + static_cast<size_t>(SEL_ROWS) * Npad * 4;
  WF_REQUIRE(smem <= 200 * 1024, "wf_bsa_select_cdf: too many key chunks for the shared-memory sort");
  WF_CUDA_OK(cudaFuncSetAttribute(bsa_select_cdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  bsa_select_cdf_kernel<<<dim3((Nq + SEL_ROWS - 1) / SEL_ROWS, heads), SEL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(q_cmp), static_cast<const bf16*>(k_cmp), idx, lens, Nq, Nk, Npad, cdf_threshold, n_floor,
      1.4426950408889634f / sqrtf(static_cast<float>(BS_D)));
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_attention_bsa_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                                     const int32_t* block_idx, const int32_t* block_lens, int max_sel, int Tq, int Tk, int H, int W,
                                     int ct, int ch, int cw, int heads, float softmax_scale, void* stream) {
  using namespace wf;
  WF_REQUIRE(q && k && v && out && block_idx, "wf_attention_bsa_bf16: null pointer");
  WF_REQUIRE(Tq > 0 && Tk > 0 && H > 0 && W > 0 && heads > 0 && max_sel > 0, "wf_attention_bsa_bf16: empty problem");
  const int chunk = ct * ch * cw;
  WF_REQUIRE(chunk == 64 || chunk == 128, "wf_attention_bsa_bf16: chunks of 64 or 128 tokens (4x4x4 / 4x4x8) only");
  WF_REQUIRE(ct % 2 == 0, "wf_attention_bsa_bf16: the temporal chunk extent must be even");
  WF_REQUIRE(Tq % ct == 0 && Tk % ct == 0 && H % ch == 0 && W % cw == 0,
             "wf_attention_bsa_bf16: the grids must be whole numbers of chunks");
  WF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "wf_attention_bsa_bf16: leading dimensions must be multiples of 8");
  WF_REQUIRE(ldq >= heads * BS_D && ldk >= heads * BS_D && ldv >= heads * BS_D && ldo >= heads * BS_D,
             "wf_attention_bsa_bf16: leading dimension smaller than heads*128");
  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* m, const void* base, int ld, int T, uint32_t box_t) {
    uint64_t dims[4] = {static_cast<uint64_t>(heads) * BS_D, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(T)};
    uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(W) * ld * 2, static_cast<uint64_t>(H) * W * ld * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(cw), static_cast<uint32_t>(ch), box_t};
    return make_tmap(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  if ((rc = mk(&tmQ, q, ldq, Tq, ct))) return rc;
  if ((rc = mk(&tmK, k, ldk, Tk, ct / 2))) return rc;
  if ((rc = mk(&tmV, v, ldv, Tk, ct / 2))) return rc;
  static PerDeviceOnce once;
  rc = once.run([] {
    WF_CUDA_OK((cudaFuncSetAttribute(attention_bsa_tcgen05<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMEM)));
    WF_CUDA_OK((cudaFuncSetAttribute(attention_bsa_tcgen05<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMEM)));
    return static_cast<int>(WF_OK);
  });
  if (rc) return rc;
  BsaArgs a;
  a.idx = block_idx; a.lens = block_lens; a.max_sel = max_sel;
  a.nhq = H / ch; a.nwq = W / cw; a.nhk = a.nhq; a.nwk = a.nwq;
  a.nq_chunks = (Tq / ct) * a.nhq * a.nwq;
  a.Hq = H; a.Wq = W; a.ct = ct; a.ch = ch; a.cw = cw;
  a.out = static_cast<bf16*>(out); a.ldo = ldo;
  a.scale_log2 = softmax_scale * 1.4426950408889634f;
  const int chunks_per_cta = 2 * (BS_BM / chunk);
  dim3 grid((a.nq_chunks + chunks_per_cta - 1) / chunks_per_cta, heads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (chunk == 64) attention_bsa_tcgen05<64><<<grid, BS_THREADS, BS_SMEM, st>>>(tmQ, tmK, tmV, a);
  else attention_bsa_tcgen05<128><<<grid, BS_THREADS, BS_SMEM, st>>>(tmQ, tmK, tmV, a);
  WF_LAUNCH_OK();
  return WF_OK;
}
