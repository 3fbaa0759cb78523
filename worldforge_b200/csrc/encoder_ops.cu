// worldforge_b200 - kernels of the once-per-video encoders: the UMT5-XXL text encoder
// (wan_for_worldforge/wan/modules/t5.py; transformers' UMT5EncoderModel in the diffusers pipeline,
// utils/pipeline_wan_i2v_clean.py:167-211) and the CLIP ViT-H/14 image encoder (wan/modules/clip.py:209-300).
// Their Linears run on the tcgen05 GEMM; what is specific to them is here:
//
//   wf_attention_small_bf16  attention over <= 1024 keys for head_dim 64 (T5: 64 heads) or 80 (CLIP: 16 heads) - shapes the
//                            128-wide tcgen05 attention kernel does not take.  Two rounding modes:
//                              mode 0 "einsum" (t5.py:107-109): scores = bf16(q.k), + bf16 relative-position bias -> bf16,
//                                      masked keys -> finfo.min, softmax in fp32, probabilities -> bf16, PV in fp32 -> bf16;
//                                      no softmax scale (T5 does not use one)
//                              mode 1 "flash" (clip.py:85): fp32 scores * scale, fp32 softmax, bf16 probabilities
//                            The T5 bias is emb[bucket[j - i + Lk - 1]][head]: the bucket of every relative distance is a
//                            (2 Lk - 1)-entry table the host fills with T5RelativeEmbedding._relative_position_bucket.
//   wf_geglu_bf16            T5FeedForward (t5.py:136-137): fc1(x) * GELU_tanh(gate(x)) with every intermediate of the
//                            reference's bf16 tensor expression rounded to bf16 where torch rounds it.
//   wf_gelu_erf / quick      reused from dit_ops (LongCat) where needed.
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace wf {

constexpr int SA_THREADS = 128;      // one query row per thread

template <int D>
__global__ void __launch_bounds__(SA_THREADS)
attention_small_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, int ldk, const bf16* __restrict__ v, int ldv,
                       bf16* __restrict__ out, int ldo, const bf16* __restrict__ emb, const int* __restrict__ bucket, int heads,
                       int Lq, int Lk, int Lk_valid, float scale, int mode) {
  extern __shared__ __align__(16) uint8_t sa_smem[];
  bf16* ks = reinterpret_cast<bf16*>(sa_smem);                 // [Lk][D]
  bf16* vs = ks + static_cast<size_t>(Lk) * D;                 // [Lk][D]
  float* bias_s = reinterpret_cast<float*>(vs + static_cast<size_t>(Lk) * D);   // [2 Lk - 1] bias by relative distance (mode 0)
  const int head = blockIdx.y;
  for (int i = threadIdx.x; i < Lk * (D / 8); i += SA_THREADS) {
    const int r = i / (D / 8), c = (i % (D / 8)) * 8;
    *reinterpret_cast<uint4*>(ks + r * D + c) = *reinterpret_cast<const uint4*>(k + static_cast<size_t>(r) * ldk + head * D + c);
    *reinterpret_cast<uint4*>(vs + r * D + c) = *reinterpret_cast<const uint4*>(v + static_cast<size_t>(r) * ldv + head * D + c);
  }
  if (emb)
    for (int i = threadIdx.x; i < 2 * Lk - 1; i += SA_THREADS) bias_s[i] = __bfloat162float(emb[bucket[i] * heads + head]);
  __syncthreads();
  const int row = blockIdx.x * SA_THREADS + threadIdx.x;
  if (row >= Lq) return;
  float qr[D];
  {
    const bf16* qp = q + static_cast<size_t>(row) * ldq + head * D;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(qp + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) { const float2 f = __bfloat1622float2(h[t]); qr[c + 2 * t] = f.x; qr[c + 2 * t + 1] = f.y; }
    }
  }
  const float NEG = -3.3895313892515355e38f;                    // torch.finfo(torch.bfloat16).min
  auto score = [&](int j) -> float {
    const bf16* kr = ks + j * D;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < D; c += 2) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kr + c));
      acc = fmaf(qr[c], f.x, acc);
      acc = fmaf(qr[c + 1], f.y, acc);
    }
    if (mode == 0) {
      float s = bf16_round(acc);
      if (emb) s = bf16_round(s + bias_s[j - row + Lk - 1]);
      return j < Lk_valid ? s : NEG;
    }
    return j < Lk_valid ? acc * scale : -INFINITY;
  };
  // pass 1: row maximum and the fp32 normaliser
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < Lk; ++j) {
    const float s = score(j);
    if (s > m) { l *= __expf(m - s); m = s; }
    l += __expf(s - m);
  }
  const float inv = 1.0f / l;
  // pass 2: probabilities rounded to bf16 (softmax(...).type_as(attn)), PV accumulated in fp32
  float o[D];
#pragma unroll
  for (int c = 0; c < D; ++c) o[c] = 0.f;
  for (int j = 0; j < Lk; ++j) {
    const float p = bf16_round(__expf(score(j) - m) * inv);
    const bf16* vr = vs + j * D;
#pragma unroll
    for (int c = 0; c < D; c += 2) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vr + c));
      o[c] = fmaf(p, f.x, o[c]);
      o[c + 1] = fmaf(p, f.y, o[c + 1]);
    }
  }
  bf16* op = out + static_cast<size_t>(row) * ldo + head * D;
#pragma unroll
  for (int c = 0; c < D; c += 2) *reinterpret_cast<uint32_t*>(op + c) = pack_bf16x2(o[c], o[c + 1]);
}

// h = [gate | fc1] (one fused GEMM), out = fc1 * gelu(gate) with the bf16 roundings of
//   0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))        (t5.py:48-50)
// evaluated on a bf16 tensor: every tensor-valued sub-expression is a bf16 tensor.
__global__ void geglu_bf16_kernel(const bf16* __restrict__ h, bf16* __restrict__ out, size_t rows, int F) {
  const size_t n = rows * static_cast<size_t>(F / 2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / (F / 2);
    const int c = static_cast<int>(i % (F / 2)) * 2;
    const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h + r * 2 * F + c));
    const float2 u = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h + r * 2 * F + F + c));
    auto gelu = [](float x) {
      const float t1 = bf16_round(x * x * x);                     // torch.pow(x, 3.0)
      const float t2 = bf16_round(0.044715f * t1);
      const float t3 = bf16_round(x + t2);
      const float t4 = bf16_round(0.7978845608028654f * t3);
      const float t5 = bf16_round(tanhf(t4));
      const float t6 = bf16_round(1.0f + t5);
      const float t7 = bf16_round(0.5f * x);
      return bf16_round(t7 * t6);
    };
    *reinterpret_cast<uint32_t*>(out + r * F + c) = pack_bf16x2(u.x * gelu(g.x), u.y * gelu(g.y));
  }
}

}  // namespace wf

using namespace wf;

extern "C" int wf_attention_small_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                                       const void* bias_emb, const int* bias_bucket, int Lq, int Lk, int Lk_valid, int heads,
                                       int head_dim, float scale, int mode, void* stream) {
  WF_REQUIRE(q && k && v && out, "wf_attention_small_bf16: null pointer");
  WF_REQUIRE(Lq > 0 && Lk > 0 && Lk <= 1024 && heads > 0 && Lk_valid >= 1 && Lk_valid <= Lk, "wf_attention_small_bf16: bad sizes (1 <= Lk <= 1024)");
  WF_REQUIRE(head_dim == 64 || head_dim == 80, "wf_attention_small_bf16: head_dim 64 or 80");
  WF_REQUIRE((bias_emb == nullptr) == (bias_bucket == nullptr), "wf_attention_small_bf16: bias table and bucket map come together");
  WF_REQUIRE(!bias_emb || (mode == 0 && Lq == Lk), "wf_attention_small_bf16: the relative-position bias is for self-attention in einsum mode");
  WF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "wf_attention_small_bf16: leading dimensions must be multiples of 8");
  const size_t smem = static_cast<size_t>(Lk) * head_dim * 4 + static_cast<size_t>(2 * Lk) * 4;
  WF_REQUIRE(smem <= 200 * 1024, "wf_attention_small_bf16: keys and values do not fit in shared memory");
  dim3 grid((Lq + SA_THREADS - 1) / SA_THREADS, heads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (head_dim == 64) {
    WF_CUDA_OK(cudaFuncSetAttribute(attention_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attention_small_kernel<64><<<grid, SA_THREADS, smem, st>>>(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk,
        static_cast<const bf16*>(v), ldv, static_cast<bf16*>(out), ldo, static_cast<const bf16*>(bias_emb), bias_bucket, heads, Lq, Lk,
        Lk_valid, scale, mode);
  } else {
    WF_CUDA_OK(cudaFuncSetAttribute(attention_small_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attention_small_kernel<80><<<grid, SA_THREADS, smem, st>>>(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk,
        static_cast<const bf16*>(v), ldv, static_cast<bf16*>(out), ldo, static_cast<const bf16*>(bias_emb), bias_bucket, heads, Lq, Lk,
        Lk_valid, scale, mode);
  }
  WF_LAUNCH_OK();
  return WF_OK;
}

extern "C" int wf_geglu_bf16(const void* h, void* out, long long rows, int F, void* stream) {
  WF_REQUIRE(h && out && rows > 0 && F > 0 && F % 2 == 0, "wf_geglu_bf16: bad arguments");
  const long long n = rows * (F / 2);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16));
  geglu_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(h), static_cast<bf16*>(out),
                                                                         static_cast<size_t>(rows), F);
  WF_LAUNCH_OK();
  return WF_OK;
}
