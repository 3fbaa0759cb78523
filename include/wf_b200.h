/* worldforge_b200 - C ABI of libwf_b200.so (sm_100a kernels for the WorldForge sampling hot path).
 *
 * The reference (Westlake-AGI-Lab/WorldForge) is 100% Python and has no FFI: its hot path sits
 * behind duck-typed objects (pipe.transformer / pipe.vae / pipe.scheduler).  This
 * header is therefore the boundary a maintainer would bind with ctypes from those objects; every
 * entry point names the reference expression it replaces (paths relative to
 * wan_for_worldforge/).  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain C, raw device pointers (tensor.data_ptr()), sizes in elements unless a name
 * says bytes; every function returns 0 (WF_OK) or a negative WF_E* code and never throws; the
 * message is available from wf_last_error() (thread-local).  All device work is asynchronous on
 * the cudaStream_t passed as `stream` (0 = legacy default stream).  The library owns no memory:
 * callers allocate inputs, outputs and the workspaces whose sizes the *_workspace_bytes functions
 * report.  `*_bf16` / `is_bf16` flags say whether a buffer holds bf16 (1) or fp32 (0) elements.
 */
#ifndef WF_B200_H
#define WF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define WF_ABI_VERSION 1

#define WF_OK 0
#define WF_EINVAL (-1) /* bad argument (shape, alignment, null pointer) */
#define WF_ECUDA (-2)  /* CUDA runtime / driver error */

const char* wf_last_error(void);
int wf_abi_version(void);
int wf_sm_count(void);

/* ---- DiT tensor-core kernels (tcgen05 / TMEM / TMA) ------------------------------------------- */

/* epilogues of wf_gemm_bf16 */
#define WF_EPI_BF16 0        /* out bf16 [M,ldo]   = bf16(acc + bias)                                   */
#define WF_EPI_GELU_BF16 1   /* out bf16           = bf16(gelu_tanh(bf16(acc + bias)))                  */
#define WF_EPI_RESID_F32 2   /* out fp32 (in place) += float(bf16(acc + bias)) * gate[n]  (gate NULL: 1) */
#define WF_EPI_F32_OF_BF16 3 /* out fp32           = float(bf16(acc + bias))                            */
#define WF_EPI_RESID_BF16 4  /* out bf16 (in place) = bf16(float(out) + gate * float(bf16(acc + bias)))   */

/* nn.Linear under autocast(bf16): out = a[M,K] . w[N,K]^T + bias, fused epilogue.
 * Replaces every F.linear of WanAttentionBlock / text_embedding / img_emb / patch_embedding
 * (wan/modules/model.py:123-126,197-198,271-273,456-460) together with the elementwise op that
 * follows it (GELU :272; gated residual :306,:313; residual :310).
 * a, w: bf16 row-major, K contiguous (lda, ldw in elements, multiples of 8); bias: bf16 [N] or NULL.
 * gate: fp32 [N] (gate_rows = 0) or a [groups, N] table where row m uses group m / gate_rows - the per-frame adaLN
 * gates of LongCat (longcat_video/modules/longcat_video_dit.py:83-85,101-103). */
int wf_gemm_bf16(const void* a, int lda, const void* w, int ldw, const void* bias, void* out, int ldo,
                 const float* gate, int gate_rows, int M, int N, int K, int epilogue, void* stream);

/* flash_attention(q, k, v) for head_dim 128, non-causal (wan/modules/attention.py:24-130; call sites
 * model.py:149-154 self-attention, :220-222 cross-attention).  q [Lq, heads*128], k/v [Lk, heads*128],
 * out [Lq, heads*128], all bf16 row-major with the given leading dimensions (multiples of 8).
 * add_in (bf16, same shape as out, or NULL): out = bf16(float(bf16(attn)) + add_in), the
 * "x = x + img_x" of WanI2VCrossAttention (model.py:227). */
int wf_attention_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                      const void* add_in, int ld_add, int Lq, int Lk, int heads, float softmax_scale,
                      void* stream);

/* ---- LongCat block-sparse attention (the 720p refine pass) ---------------------------------------- */

/* mean_pooling_compression (longcat_video/block_sparse_attention/bsa_interface.py:176-186) of a [T*H*W, heads*128]
 * bf16 matrix in (t,h,w) token order over ct x ch x cw chunks, WITHOUT the chunk-major re-ordering of
 * rearrange_THW_to_3d_block (:598-603): out [heads][chunks][128] bf16, chunks in (Nt,Nh,Nw) order. */
int wf_bsa_mean_pool(const void* x, int ldx, void* out, int T, int H, int W, int ct, int ch, int cw, int heads,
                     void* stream);

/* get_select_indices_topk (bsa_interface.py:214-232): scores = bf16(q_cmp k_cmp^T) per head, the n_sel best key
 * chunks of every query chunk.  q_cmp [heads][Nq][128], k_cmp [heads][Nk][128] bf16; idx [heads][Nq][n_sel] int32 in
 * ascending chunk order (the order topk_sort gives, :530-534); ties at the threshold take the lower chunk index. */
int wf_bsa_select_topk(const void* q_cmp, const void* k_cmp, int32_t* idx, int Nq, int Nk, int heads, int n_sel,
                       void* stream);
/* get_select_indices_cdf / get_select_indices_cdf_topk (bsa_interface.py:234-275): idx [heads][Nq][Nk] = the key chunks of
 * every query chunk sorted by softmax(score / sqrt(128)) descending (ties: lower chunk first), lens [heads][Nq] = how many
 * of them have a cumulative softmax mass <= cdf_threshold (searchsorted right=True; may be 0), raised to n_floor (the
 * top-k count int((1 - sparsity) * Nk), or 0 for the plain cdf rule). */
int wf_bsa_select_cdf(const void* q_cmp, const void* k_cmp, int32_t* idx, int32_t* lens, int Nq, int Nk, int heads,
                      float cdf_threshold, int n_floor, void* stream);

/* flash_attn_bsa_3d minus the gating (bsa_interface.py:612-659 -> _attn_fwd_bsa_varlen_align,
 * flash_attn_bsa_varlen_mask.py:174-285): every query chunk attends to the key chunks block_idx[h][chunk][0 ..
 * len) only (len = block_lens[h][chunk], or max_sel when block_lens is NULL).  q [Tq*H*W, heads*128], k/v
 * [Tk*H*W, heads*128], out like q: bf16 row-major in (t,h,w) token order - the kernel gathers chunks with 4-D TMA
 * boxes, no re-ordered copy exists.  ct*ch*cw must be 64 or 128 (the checkpoint's 4x4x4, the code default 4x4x8). */
int wf_attention_bsa_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                          const int32_t* block_idx, const int32_t* block_lens, int max_sel, int Tq, int Tk, int H,
                          int W, int ct, int ch, int cw, int heads, float softmax_scale, void* stream);

/* ---- DiT HBM-bound kernels ------------------------------------------------------------------------ */

/* WanLayerNorm (+ modulation or affine), model.py:97-102 with :303/:311 (scale,shift = e[1],e[0] /
 * e[4],e[3]) or norm3's weight,bias (:262-264).  round_norm_bf16 = 1 reproduces .type_as(x) for the
 * bf16 token stream of block 0. */
int wf_layer_norm(const void* x, int ldx, int x_is_bf16, void* out, int ldo, int out_is_bf16, const float* scale,
                  const float* shift, const float* weight, const float* bias, int rows, int D, float eps,
                  int round_norm_bf16, int rows_per_group, void* stream);   /* rows_per_group > 0: scale/shift are [groups, D] */

/* ---- one process per GPU: heads <-> tokens exchange of Ulysses attention through NVLink peer memory ------------------
 * (reference pattern: longcat_video/context_parallel/ulysses_wrapper.py:87-105 = four NCCL all-to-alls + staging copies
 * per attention; wan/distributed/xdit_context_parallel.py:160-176 for the token split.)  Here the kernels that PRODUCE
 * the data store it straight into the consuming rank's buffer: wf_qkv_norm_rope_scatter (RMSNorm + RoPE of q and k, and
 * v) writes [all tokens, this peer's heads], wf_attention_bf16_peers writes each output row to the rank that owns the
 * token.  Buffers come from wf_peer_alloc (cudaMalloc + CUDA IPC handle); peers map them with wf_peer_open.  The caller
 * orders producers and consumers with a barrier across the ranks (any stream-ordered collective). */
int wf_peer_alloc(long long bytes, void** ptr, void* handle64);       /* handle64: 64 bytes to send to the peers */
int wf_peer_open(const void* handle64, void** ptr);
int wf_peer_close(void* ptr);
int wf_peer_free(void* ptr);
/* qkv: bf16 [rows, 3*D] (q|k|v of this rank's tokens, all heads).  For which in {q,k}: WanRMSNorm over D with weight_q /
 * weight_k, then RoPE (same arithmetic as wf_rms_norm_rope); v is copied.  Head h of row r goes to peer h / (heads/n_peers):
 * dst_peers[peer][(row0 + r) * ld_dst + (which*(heads/n_peers) + h % (heads/n_peers))*128 ...]. */
int wf_qkv_norm_rope_scatter(const void* qkv, int ldx, const float* weight_q, const float* weight_k, const double* rope,
                             int rows, int D, float eps, void* const* dst_peers, int n_peers, int ld_dst, int row0,
                             void* stream);
/* wf_attention_bf16 whose output row q goes to out_peers[q / rows_per_peer] + (q % rows_per_peer)*ldo + head*128
 * (each pointer already includes the column offset of this rank's head block). */
int wf_attention_bf16_peers(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* const* out_peers,
                            int n_peers, int rows_per_peer, int ldo, int Lq, int Lk, int heads, float softmax_scale,
                            void* stream);

/* WanRMSNorm over the full model dim (model.py:81-89) followed by rope_apply (:43-70), in place on a
 * bf16 [rows, D] slice (leading dimension ldx).  rope: fp64 [rows, 64, 2] (cos, sin) per token and
 * complex pair, or NULL for the cross-attention q/k, which carry no RoPE. */
int wf_rms_norm_rope(void* x, int ldx, const float* weight, const double* rope, int rows, int D, float eps,
                     void* stream);

/* im2col of patch_embedding = Conv3d(k = s = (1,2,2)) (model.py:456-457,534-537):
 * hidden bf16 [C,F,H,W] -> cols bf16 [F*(H/2)*(W/2), 4*C]. */
int wf_patchify(const void* hidden, void* cols, int C, int F, int H, int W, void* stream);

/* Head.forward + unpatchify (model.py:337-347, 584-607): fp32 LN, modulation, fp32 Linear(D -> 4*Cout),
 * scatter to out fp32 [Cout, F, 2*GH, 2*GW].  x holds the L tokens starting at global token tok_offset
 * (0 and L = F*GH*GW on one GPU; a contiguous shard under sequence parallelism). */
int wf_dit_head(const void* x, int x_is_bf16, int ldx, int L, int D, const float* scale, const float* shift, const float* w,
                const float* b, int Cout, float* out, int F, int GH, int GW, float eps, int tok_offset, int rows_per_group,
                int round_bf16, void* stream);
/* rows_per_group / round_bf16: LongCat's FinalLayer_FP32 (longcat_video/modules/blocks.py:147-156): per-frame shift / scale
 * tables and the modulated row rounded to bf16 before the fp32 projection. */

/* ---- LongCat-Video DiT extras (longcat_for_worldforge/longcat_video/modules) ---------------------------------------- */
/* per-head RMSNorm with a bf16 gain (blocks.py:46-52) + fp32 rotate-half RoPE (rope_3d.py:99-119), in place on a bf16
 * [rows, heads*128] slice; rope: fp32 [rows, 64, 2] (cos, sin) or NULL */
int wf_rms_norm_head_rope(void* x, int ldx, const void* gain_bf16, const float* rope, long long rows, int heads, float eps,
                          void* stream);
/* FeedForwardSwiGLU gate (blocks.py:39): in bf16 [rows, 2F] = [w1 x | w3 x] -> out bf16 [rows, F] = silu(w1 x) * (w3 x) */
int wf_swiglu_bf16(const void* in, void* out, long long rows, int F, void* stream);
/* TimestepEmbedder.timestep_embedding (blocks.py:184-191) in fp32: t [n] -> out [n, dim] = [cos | sin] */
int wf_timestep_embedding_f32(const float* t, float* out, int n, int dim, void* stream);
/* out[T,R] = act(x[T,K]) . W[R,K]^T + b in fp32 on bf16 weights, T <= 32: the t-embedder MLP and every block's
 * adaLN_modulation inside the reference's autocast(float32) islands (longcat_video_dit.py:82-85,312-313) */
int wf_small_gemm_f32(const float* x, const void* w_bf16, const void* b_bf16, float* out, int T, int R, int K, int silu_in,
                      void* stream);

/* fp32 matrix-vector product with optional SiLU on the input and/or output: the time_embedding /
 * time_projection MLPs at batch 1 (model.py:546-550). */
int wf_gemv_f32(const float* w, const float* x, const float* b, float* out, int N, int K, int silu_in,
                int silu_out, void* stream);

/* sinusoidal_embedding_1d (model.py:18-28) of the int64 timestep held on the device, float64 math: out fp32 [freq_dim] */
int wf_time_sinusoid(const long long* timestep, float* out, int freq_dim, void* stream);

/* out[r,:] = a[r,:] + b[:] - "self.modulation + e" for all blocks at once (model.py:298, 345) */
int wf_add_bcast_f32(const float* a, const float* b, float* out, long long rows, int inner, void* stream);

/* nn.GELU() (erf form) in place on bf16 - the MLPProj activation (model.py:355-358). */
int wf_gelu_erf_bf16(void* x, long long n, void* stream);

/* ---- WorldForge sampler ops (utils/pipeline_wan_i2v_clean.py, utils/scheduling_unipc_multistep_clean.py) */

/* noise_pred + s*(noise_pred - noise_uncond), pipeline :611 */
int wf_cfg_combine(const void* cond, const void* uncond, void* out, int is_bf16, float scale, long long n,
                   void* stream);
/* x0 = sample - sigma*model_output, scheduler :952-958; out dtype = bf16 iff both inputs are bf16 */
int wf_x0_convert(const void* sample, int sample_bf16, const void* v, int v_bf16, void* out, float sigma,
                  long long n, void* stream);
/* multistep_uni_p_bh_update, scheduler :1084-1099, with the scalar coefficients evaluated by the host:
 * c_x = sigma_t/sigma_s0, c_m0 = alpha_t*expm1(-h), c_res = alpha_t*B_h; out has x's dtype.
 * rk: r_1 itself (rk_is_reciprocal = 0: (m1-m0)/r_1) or 1/r_1 (rk_is_reciprocal = 1: (m1-m0)*(1/r_1), which is
 * how torch's CUDA division kernel treats a host-side scalar divisor). */
int wf_unip_update(const void* x, int x_bf16, const void* m0, int m0_bf16, const void* m1, int m1_bf16, void* out,
                   int order, float c_x, float c_m0, float rk, int rk_is_reciprocal, float c_res, long long n,
                   void* stream);
/* add_noise, scheduler :1584: out fp32 = (1-sigma)*x0 + sigma*noise, sigma held in x0's dtype */
int wf_renoise(const void* x0, int x0_bf16, const float* noise, float* out, float one_minus_sigma, float sigma,
               long long n, void* stream);
/* DSG, pipeline :664-681: two passes (three reductions, then the guided combination).
 * stats (device float[3], may be NULL) receives cos, sin, magnitude ratio. */
long long wf_dsg_workspace_bytes(void);
int wf_dsg(const void* g, const void* w, void* out, int is_bf16, float omega, long long n, void* workspace,
           float* stats, void* stream);
/* FLF pixel blend, scheduler :1375-1381: out = (2*ref-1)*mask + decoded*(1-mask); fp32;
 * decoded/ref/out [channels, plane], mask [plane] broadcast over channels. */
int wf_flf_blend(const float* decoded, const float* ref, const float* mask, float* out, int channels,
                 long long plane, void* stream);
/* Input preparation of the entry script (wan_for_worldforge/infer_worldforge.py).
 * wf_soften_mask replaces soften_mask (:105-150) for one clip: mask / out fp32 [frames][H][W]; a set pixel (!= 0) whose exact
 * Euclidean distance d to the nearest unset pixel of its frame satisfies d*d <= max_d2 becomes lut[d*d]; every other pixel is
 * copied.  radius = floor(transition_distance) <= 31, max_d2 = floor(transition_distance^2), lut: device float[max_d2 + 1]
 * holding smooth(sqrt(d2) / transition_distance) as the reference's float64 numpy expression rounds it to fp32.
 * wf_clip_from_u8 replaces the frame stacking of :232-238: uint8 [pixels][3] -> planar fp32 [3][pixels] / 255. */
int wf_soften_mask(const float* mask, float* out, int frames, int H, int W, int radius, int max_d2, const float* lut,
                   void* stream);
int wf_clip_from_u8(const unsigned char* frames, float* out, long long pixels, void* stream);
/* Encoders (once per video).
 * wf_attention_small_bf16: attention over <= 1024 keys, head_dim 64 (UMT5-XXL, wan/modules/t5.py:86-116) or 80 (CLIP ViT-H/14,
 * wan/modules/clip.py:74-86); q/k/v/out bf16 [tokens][heads*head_dim] views.  mode 0 = the rounding points of T5Attention's
 * bf16 einsum form (scores and score + bias rounded to bf16, keys >= Lk_valid at finfo.min, fp32 softmax, bf16 probabilities,
 * no scale) with bias_emb bf16 [buckets][heads] and bias_bucket int32 [2*Lk-1] (bucket of relative distance d = j - i at
 * index d + Lk - 1; both NULL for no bias); mode 1 = flash-attention rounding (fp32 scores * scale).
 * wf_geglu_bf16: h bf16 [rows][2F] = [gate | fc1] -> out [rows][F] = fc1 * GELU_tanh(gate), bf16 intermediates (t5.py:48-50,136). */
int wf_attention_small_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                            const void* bias_emb, const int* bias_bucket, int Lq, int Lk, int Lk_valid, int heads, int head_dim,
                            float scale, int mode, void* stream);
int wf_geglu_bf16(const void* h, void* out, long long rows, int F, void* stream);
/* LongCat generate_refine step 5 (longcat_video/pipeline_longcat_video.py:1393-1419): stage-1 video uint8 [F,H0,W0,3]
 * -> bilinear (align_corners) to [H,W] -> /255 -> trilinear to F2 frames -> *2-1, bf16 intermediates as in the reference,
 * first / last frame repeated pad_front / pad_back times.  out: planar fp32 [3][pad_front+F2+pad_back][H][W]. */
int wf_refine_upsample(const unsigned char* video, int F, int H0, int W0, float* out, int F2, int H, int W,
                       int pad_front, int pad_back, void* stream);
/* scheduler :1281-1282: out fp32 = x0/inv_std + mean evaluated in x0's dtype ([channels, per_channel]) */
int wf_latent_denorm(const void* x0, int is_bf16, float* out, const float* mean, const float* inv_std, int channels,
                     long long per_channel, void* stream);
/* scheduler :1385 + :1410-1417: out (x0's dtype) = replace_mask bit c ? x0[c] : (enc[c]-mean[c])*inv_std[c] */
int wf_latent_norm_replace(const float* enc, const void* x0, int is_bf16, void* out, const float* mean,
                           const float* inv_std, unsigned replace_mask, int channels, long long per_channel,
                           void* stream);
/* FLF scoring front end: min-max normalise the n elements and quantise to uint8.
 * mode 0 - Wan selector (scheduler :376-378 + :175-176): n*255, truncated, over the whole 16-channel tensor;
 * mode 1 - LongCat selector (scheduling_flow_match_euler_discrete.py:329-330 + :147-151), called per channel:
 *          ((n+1)*127.5).clip(0,255), truncated. */
long long wf_quantise_workspace_bytes(void);
int wf_quantise_u8(const void* x, int is_bf16, unsigned char* out, long long n, int mode, void* workspace, void* stream);
/* FLF channel scoring on the device (reference: cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0) per
 * consecutive frame pair, scheduling_unipc_multistep_clean.py:220-224, and the flow metrics :541-604).  clips_u8: uint8
 * [clips][T][H][W]; flow: fp32 [clips][T-1][H][W][2] (dx, dy).  Only frames whose Farneback pyramid has ONE level
 * (min(H, W) < 64: the 60 x 104 latent frames of the 480p configuration); restated in oracle/farneback.py, pinned to OpenCV. */
long long wf_farneback_workspace_bytes(int clips, int T, int H, int W);
int wf_farneback_u8(const unsigned char* clips_u8, int clips, int T, int H, int W, int winsize, int iterations, float* flow,
                    void* workspace, void* stream);
/* per channel: mean end-point error, mean outlier indicator (epe > 3 AND epe > 0.05 |ref|; with outlier_or = 1 the two tests are
 * OR-ed - LongCat's variant, longcat_video/modules/scheduling_flow_match_euler_discrete.py:226), mean angular error in degrees
 * between two flow fields [channels][per_channel][2]; out3: fp32 [channels][3]. */
int wf_flow_metrics(const float* flow_ref, const float* flow_cand, float* out3, int channels, long long per_channel,
                    int outlier_or, void* stream);
/* LongCat CFG-zero + sign flip (pipeline_longcat_video.py:374-383, 875-888), fp32:
 * st = <cond,uncond>/(|uncond|^2 + 1e-8); out = -(uncond*st + scale*(cond - uncond*st)).  workspace: wf_dsg_workspace_bytes().
 * stats (device float[1] or NULL) receives st. */
int wf_cfg_zero(const float* cond, const float* uncond, float* out, float scale, long long n, void* workspace, float* stats,
                void* stream);

/* ---- Wan 3D-VAE (wan/modules/vae.py), channels-last fp32 activations [T][H][W][C] ------------------------ */

/* Implicit-GEMM convolution on the tensor cores (tf32 operands, fp32 accumulation):
 *   out(t,y,x,n) = bias[n] + sum_tap sum_c W[tap*Cout + n, c] * in(t*t_stride + t_off + dt_tap, y + dy_tap, x + dx_tap, c)
 * with zero fill outside the input.  taps: ntaps int8 triples (dt, dy, dx).  The output element goes to frame
 * t*t_mul + n/c_split, row y*sy+oy, column x*sx+ox, channel n%c_split of a channels-last tensor with out_H x out_W
 * pixels per frame and ldc floats per pixel (planar_clamp = 1: planar [c][frame][out_H][out_W] with channel stride
 * planar_cstride, values clamped to [-1,1]).  resid (same addressing as out) is added when not NULL.
 * Replaces CausalConv3d.forward (vae.py:28-36), the Resample convs (:76-96,:128-140,:156-157), the 1x1 convs of
 * AttentionBlock (:246,:260) and, as a plain GEMM, its q.k^T and p.v products (:252-256).
 * Operand precision: kind::tf32 reads fp32 operands with the 13 low mantissa bits IGNORED (truncation) where cuDNN - the
 * reference's fp32 VAE on a GPU - rounds to nearest; callers therefore hand this function operands already rounded to
 * tf32 (weights at load; activations by their producer: round_tf32 of wf_rms_norm_cl / wf_planar_to_cl / wf_round_tf32,
 * or round_out_tf32 = 1 here when every consumer of `out` is a convolution).
 * Fused RMS_norm (+SiLU) (vae.py:51-54, :195-197 - the norm that FOLLOWS this convolution in ResidualBlock): with
 * norm_gamma != NULL (needs Cout <= 192 so that one accumulator tile holds all channels of a pixel, plain channels-last
 * output) a = silu?(v / max(|v|_2, 1e-12) * sqrt(Cout) * gamma), rounded to tf32, is stored to norm_out, or replaces v in
 * `out` when norm_out is NULL. */
int wf_conv_tf32(const float* in, int in_T, int in_H, int in_W, int Cin, const float* weights, const float* bias,
                 int Cout, int ntaps, const signed char* taps, int T, int H, int W, int t_stride, int t_off, float* out,
                 int ldc, int out_H, int out_W, int t_mul, int c_split, int sy, int sx, int oy, int ox,
                 const float* resid, int planar_clamp, long long planar_cstride, int tile_w, int round_out_tf32,
                 const float* norm_gamma, float* norm_out, int norm_silu, void* stream);
/* Development hook (tools/conv_probe.py): a device buffer of >= 16 int64 that CTA 0 of every following 3x3x3 launch fills
   with its per-role wait / work clocks; NULL switches it off.  Not part of the reference-facing surface. */
int wf_debug_conv_profile(long long* device_buf);
/* RMS_norm (vae.py:51-54) over the channels of every pixel, optionally followed by SiLU (:195-197); round_tf32 = 1 stores
   the result rounded to tf32 (it feeds a convolution) */
int wf_rms_norm_cl(const float* x, int ldx, float* out, int ldo, const float* gamma, long long pixels, int C, int silu,
                   int round_tf32, void* stream);
/* planar [C][n] -> channels-last [n][Cp] (channels C..Cp-1 zero; optionally rounded to tf32) and back (first C of ld channels) */
int wf_planar_to_cl(const float* src, float* dst, long long n, int C, int Cp, int round_tf32, void* stream);
/* dst = src rounded to the nearest tf32 (ties away from zero), n a multiple of 4 */
int wf_round_tf32(const float* src, float* dst, long long n, void* stream);
int wf_cl_to_planar(const float* src, float* dst, long long n, int C, int ld, void* stream);
/* [T][H][W][C] -> [T][H/2][W/2][4C]: turns ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) (vae.py:87-96) into a 2x2-tap conv */
int wf_space_to_depth(const float* src, float* dst, int T, int H, int W, int C, void* stream);
/* in-place softmax(scale*x) over the rows of x [rows][ld] (the mid-block attention, vae.py:252-256) */
int wf_softmax_rows(float* x, int rows, int cols, int ld, float scale, void* stream);
/* dst[c][r] = src[r][c] */
int wf_transpose_f32(const float* src, float* dst, int R, int C, int lds, int ldd, void* stream);
/* hi = x truncated to tf32 (13 low mantissa bits cleared), lo = x - hi: operands of the three-product fp32-accurate form of
   the VAE mid-block attention matmuls (vae.py:252-256 - F.scaled_dot_product_attention on fp32 tensors; matmul TF32 is off
   by default in torch, unlike cuDNN convolutions) */
int wf_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WF_B200_H */
