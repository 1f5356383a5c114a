/* rg_b200.h -- C ABI of the B200-native guided-DDIM + exemplar-retrieval hot path of RAG-Gesture.
 *
 * The reference (m-hamza-mughal/RAG-Gesture) is pure Python and has no FFI: its seam is the mmcv
 * registry plus Python call signatures.  This library sits UNDER re-registered
 * Python classes of the same names (rag_gesture_b200/mogen_api.py); every entry point below
 * names the reference function(s) it replaces.  Conventions:
 *   - plain C: raw pointers + sizes; all tensors are dense row-major fp32 unless stated otherwise;
 *   - every `dev` pointer is device memory on the CURRENT cuda device, borrowed for the call;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls enqueue
 *     work and return without synchronising unless stated otherwise;
 *   - return value 0 = ok, otherwise rg_last_error() describes the failure (thread-local);
 *   - a handle is re-entrant from one thread at a time; distinct handles are independent.
 * There is no CPU fallback anywhere: without a CUDA device every compute entry point fails.
 */
#ifndef RG_B200_H
#define RG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define RG_ABI_VERSION 1

typedef struct rg_model* rg_handle;

/* Hyper-parameters: configs/raggesture_beatx/basegesture_len150_beat.py:32-42.
 * latent_dim must be 512 and num_heads 16 (head dim 32 == warp width is baked into the kernels). */
typedef struct rg_config {
    int32_t latent_dim;      /* 512  */
    int32_t num_heads;       /* 16   */
    int32_t ffn_dim;         /* 1024 */
    int32_t time_embed_dim;  /* 2048 */
    int32_t num_layers;      /* 8    */
    int32_t n_tokens;        /* 43 = 4 * n_chunks + 3 */
    int32_t n_chunks;        /* 10   */
    int32_t text_dim;        /* 768  (text_pre_proj / audio_pre_proj input width) */
    int32_t num_speakers;    /* 25   */
    int32_t precision;       /* RG_PREC_*: arithmetic of the dense contractions */
} rg_config;

enum { RG_PREC_FP32 = 0,      /* fp32 FMA GEMMs: the exact tier (rel-L2 <= 1e-3 vs reference)      */
       RG_PREC_BF16 = 1,      /* tcgen05 bf16 x bf16 -> fp32 TMEM accumulate (rel-L2 <= 2e-2)       */
       RG_PREC_BF16X3 = 2 };  /* tcgen05, operands split hi+lo bf16, 3 products: fp32-class accuracy */

const char* rg_last_error(void);
int rg_abi_version(void);
/* number of CUDA kernels launched by this library in the calling process so far */
int64_t rg_launch_count(void);

/* Build a denoiser from the reference state dict of ReGestureTransformer (minus the VAEs):
 * names[i] is the state-dict key without the leading "model." (SURVEY.md 8b lists them),
 * ptrs[i] points to numels[i] fp32 values on the HOST.  The library keeps its own packed device
 * copies (fused QKV, LayerNorm affine folded into the consuming Linear, ...).
 * Replaces: DiffusionTransformer.__init__ + load_checkpoint (diffusion_transformer.py:335-425). */
int rg_create(const rg_config* cfg, int n_tensors, const char* const* names,
              const float* const* ptrs, const int64_t* numels, rg_handle* out);
int rg_destroy(rg_handle h);

/* Install the respaced sampling schedule and build the timestep table (K7):
 * timestep_map[s] = original timestep of step s (SpacedDiffusion.timestep_map,
 * gaussian_diffusion.py:1723-1738); coef[s*8 + i], fp32 values the reference uses at step s:
 *   0 sqrt_recip_alphas_cumprod  1 sqrt_recipm1_alphas_cumprod     (:437-438, cast at :1623)
 *   2 sqrt(alphas_cumprod_prev)  3 sqrt(1 - alphas_cumprod_prev)   (:994-995)
 *   4 sqrt(alphas_cumprod_next)  5 sqrt(1 - alphas_cumprod_next)   (:1035-1036)
 *   6 sqrt_alphas_cumprod        7 sqrt_one_minus_alphas_cumprod   (:473-476)
 * The table holds, for every step, every layer and each of its 5 StylizationBlocks, the
 * (scale | shift) vector emb_layers(SiLU(time_embed(timestep_embedding(t)))) -- computed once
 * instead of per step per clip (diffusion_transformer.py:27-46,404-408,643; stylization_block.py:35-37).
 * Synchronises the stream. */
int rg_set_schedule(rg_handle h, int n_steps, const int32_t* timestep_map, const float* coef,
                    void* stream);

/* xf_text = text_pre_proj(word), xf_audio = audio_pre_proj(audio), xf_spk = speaker_embedding[ids]
 * Replaces: encode_text / encode_audio / encode_spks (diffusion_transformer.py:544-606) as called
 * from ReGestureTransformer.get_precompute_condition (raggesture.py:978-987).
 * word [B,n_text,text_dim], audio [B,n_audio,text_dim], spk_ids int64 [B,n_spk] (all dev);
 * outputs [B,n_*,512] (dev). */
int rg_encode_conditions(rg_handle h, const float* word, const float* audio,
                         const int64_t* spk_ids, int B, int n_text, int n_audio, int n_spk,
                         float* xf_text, float* xf_audio, float* xf_spk, void* stream);

/* Cross-attention key/value state of B clips for all layers and the 3 conditions (K6):
 * state[b][layer][cond][head][32][32] = softmax_tokens(key(text_norm(xf)))^T value(text_norm(xf)).
 * The reference recomputes this on every denoiser call (efficient_attention.py:74,78-89); it
 * depends on neither x nor t.  rg_state_floats_per_clip() gives the per-clip size. */
int64_t rg_state_floats_per_clip(rg_handle h);
int rg_precompute_clip_state(rg_handle h, const float* xf_text, const float* xf_audio,
                             const float* xf_spk, int n_text, int n_audio, int n_spk, int B,
                             float* state, void* stream);

/* One denoiser evaluation for B clips that share one timestep:  x0 = model(x, t).
 * Replaces: DiffusionTransformer.forward + ReGestureTransformer.forward_test, single conditional
 * branch (diffusion_transformer.py:620-668, raggesture.py:1041-1086) incl. all 8 DecoderLayers.
 * step_idx >= 0 selects a row of the schedule's timestep table; step_idx < 0 evaluates at the
 * original-scale timestep `tau` (table row computed on the fly).
 * x,x0_out [B,T,512]; src_mask [B,T] (motion_mask); query_mask [3,B,T] or NULL; state from
 * rg_precompute_clip_state.  x0_out may alias x. */
int rg_denoise(rg_handle h, const float* x, int B, int step_idx, int tau, const float* src_mask,
               const float* query_mask, const float* state, float* x0_out, void* stream);

/* The same evaluation for a batch made of `n_groups` (<= 8) consecutive clip ranges, each at its OWN schedule
 * level: clips [sum(group_clips[:g]), +group_clips[g]) are evaluated at step_idx group_step_idx[g] (host arrays).
 * The reference never mixes timesteps in a batch (gaussian_diffusion.py:1283 fills `t` with one value); no op of
 * the denoiser couples clips, so every clip's output equals what rg_denoise gives it in a single-level batch.
 * Used to run the guided sampling of one batch and the DDIM inversion of the next batch's exemplars as ONE
 * kernel chain (the per-kernel fixed cost is paid once for both). */
int rg_denoise_groups(rg_handle h, const float* x, int B, int n_groups, const int32_t* group_clips,
                      const int32_t* group_step_idx, const float* src_mask, const float* query_mask,
                      const float* state, float* x0_out, void* stream);

/* S consecutive levels of the samplers in ONE call (the host thread is then free for the next batch's
 * retrieval; per level this enqueues exactly what the single-step entry points would):
 *   clips     xj[0:B]   : guided / plain sampling, levels S-1 .. 0 (ddim_guided_sample_loop / ddim_sample_loop,
 *                         gaussian_diffusion.py:1233-1395, 1042-1135): in_seq = in_seq0 at level S-1 and, when
 *                         inv_list [S,B,T,512] is given, inv_list[i] at every level i below it (else in_seq0 at
 *                         every level); a non-NULL in_seq is blended with noise[j] (j = S-1-i, [S,B,T,512]);
 *                         guidance_iters[i] dead gradient steps are executed only if run_dead_guidance;
 *   exemplars xj[B:B+E] : DDIM inversion, levels 0 .. S-1 (ddim_reverse_sample_loop, :1137-1230); the latent
 *                         after level j is written to samples_out[j] ([S,E,T,512]).
 * Pass j evaluates both ranges in one kernel chain (rg_denoise_groups).  B or E may be 0.  src_mask [B+E,T],
 * query_mask [3,B+E,T] or NULL, state for B+E clips, x0_scratch [B+E,T,512].  xj holds the results. */
int rg_run_levels(rg_handle h, int S, float* xj, int B, int E, const float* in_seq0, const float* inv_list,
                  const float* noise, const int32_t* guidance_iters, float guidance_lr, int run_dead_guidance,
                  const float* src_mask, const float* query_mask, const float* state, float* samples_out,
                  float* x0_scratch, void* stream);

/* How many concurrent kernel chains ("lanes", contiguous clip ranges on separate streams, joined back
 * onto the caller's stream before the call returns; inside the CUDA graph they are parallel branches) one
 * rg_denoise / rg_denoise_groups evaluation uses: 0 = automatic (2 lanes from 64 clips on -- each lane's kernels
 * run while the other lane sits in a launch's fixed latency -- else 1), 1..4 fixed.  Results do not depend on the
 * setting: clips never interact inside a step. */
int rg_set_lanes(rg_handle h, int lanes);
/* CUDA graphs of the evaluation chain (default on; RG_GRAPHS=0 in the environment also disables them).  The
 * ~100 kernel launches of one rg_denoise / rg_denoise_groups evaluation depend only on buffer addresses and B:
 * the second call with the same addresses captures them (programmatic-dependent-launch edges included) and
 * later calls replay the graph with one cudaGraphLaunch.  Results are bit-identical either way. */
int rg_set_graphs(rg_handle h, int on);

/* DDIM update (eta = 0), bit-exact fp32 op order of the reference:
 *   eps = (c0*x - x0)/c1 ; out = x0*ca + cb*eps ; direction -1: (ca,cb) = coef 2,3 (ddim_sample,
 *   gaussian_diffusion.py:983-1001); direction +1: coef 4,5 (ddim_reverse_sample :1032-1038).
 * n = number of floats (multiple of 4).  out may alias x or x0. */
int rg_ddim_update(rg_handle h, const float* x, const float* x0, int step_idx, int direction,
                   float* out, int64_t n, void* stream);

/* in_seq outpainting blend of ddim_sample (gaussian_diffusion.py:934-947): rows of in_seq with any
 * non-zero entry replace the same rows of x by q_sample(in_seq, t, noise).  rows = B*T rows of 512.
 * out may alias x. */
int rg_blend_in_seq(rg_handle h, const float* x, const float* in_seq, const float* noise,
                    int step_idx, float* out, int64_t rows, void* stream);

/* 2-branch mixing of ReGestureTransformer.forward_test with scale_func_cfg (raggesture.py:1087-1111): out2 [2B,T,512]
 * holds the text-branch rows first, then the "none"-branch rows (same clips evaluated against the unconditional
 * cross-attention state); coef [B,4] (device) = (both, text, retr, none) per clip, joint_scale [T] (device) = the
 * per-body-part scale of each token row.  out [B,T,512] = x_text*both*js + x_text*text*js + x_none*retr/js +
 * x_none*none/js, every product and sum rounded to fp32 in that order. */
int rg_mix_branches(rg_handle h, const float* out2, int B, const float* coef, const float* joint_scale, float* out,
                    void* stream);
/* `iters` closed-form gradient steps of the insertion guidance (gaussian_diffusion.py:1351-1378):
 * x <- x - lr * grad_x mse(x * m, in_seq), m = any(in_seq != 0, -1); numel = B*T*512 of the batch the
 * reference would have run (mse_loss 'mean').  In place. */
int rg_guidance_steps(rg_handle h, float* x, const float* in_seq, int64_t rows, int iters,
                      float lr, int64_t numel, void* stream);

/* ---- op-level entry points (used by the per-module Python mirrors and the unit parity tests) -- */

/* y = epilogue(x W^T + b).  x [M,K] (row stride ldx), W [N,K], b [N] or NULL, residual [M,N] or
 * NULL (epilogue RG_OP_RESIDUAL), out [M,N].  Replaces nn.Linear on the path. */
enum { RG_OP_NONE = 0, RG_OP_RESIDUAL = 1, RG_OP_GELU = 2, RG_OP_SILU = 4 };
int rg_op_linear(const float* x, int ldx, const float* W, const float* b, const float* residual,
                 float* out, int M, int N, int K, int epilogue, void* stream);
/* The same contraction on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulator, TMA-fed):
 * operands are converted to bf16 on the fly; split=0: bf16 x bf16 -> fp32 (RG_PREC_BF16);
 * split=1: hi+lo operand split, 3 products (RG_PREC_BF16X3, fp32-class accuracy).
 * N % 128 == 0, K % 64 == 0.  out (fp32 [M,N]) and/or out_bf16 (bf16 [M,N], or [M,2N] = hi|lo planes
 * when split) may be NULL.  Used by the unit parity tests and bench.py's roofline probe. */
int rg_op_linear_tc(const float* x, const float* W, const float* b, const float* residual, float* out,
                    void* out_bf16, int M, int N, int K, int epilogue, int split, void* stream);
/* The codec's projections (gesture_vae.py / detr_utils.py nn.Linear and MultiheadAttention in/out projections, the
 * callers either side of the path: SURVEY 8f.1) keep their weights converted: rg_op_split_bf16 writes W [rows, cols]
 * as bf16 planes ([rows, cols], or [rows, 2*cols] = hi | lo when split) into out16 once, rg_op_linear_tc_w16 is
 * rg_op_linear_tc with that buffer in place of W. */
int rg_op_split_bf16(const float* w, void* out16, int rows, int cols, int split, void* stream);
int rg_op_linear_tc_w16(const float* x, const void* w16, const float* b, const float* residual, float* out,
                        void* out_bf16, int M, int N, int K, int epilogue, int split, void* stream);
/* Softmax multi-head attention of those VAE blocks (torch nn.MultiheadAttention core between the projections):
 * q [N, Sq, .] with rows ldq floats apart and head h in columns [h*dh, (h+1)*dh), k and v [N, Sk, .] likewise,
 * keep [N, Sk] bytes (non-zero = the key is attended; key_padding_mask inverted) or NULL, out [N, Sq, H*dh].
 * softmax(q k^T / sqrt(dh)) v in fp32; dh in {16, 32, 64, 128}; pointers 16-byte aligned, strides % 4 == 0. */
int rg_op_mha(const float* q, const float* k, const float* v, const unsigned char* keep, float* out, int N, int Sq,
              int Sk, int H, int dh, int64_t ldq, int64_t ldk, int64_t ldv, void* stream);
/* Measurement probe for bench.py's roofline: converts operands once, then times `reps` launches of the
 * tcgen05 GEMM alone with CUDA events on `stream` (flush_buf, if given, is overwritten before every
 * launch to evict L2); *median_ms receives the median launch time.  Synchronises. */
int rg_probe_gemm_tc(const float* x, const float* W, const float* b, float* out, int M, int N, int K,
                     int split, int reps, void* flush_buf, int64_t flush_bytes, float* median_ms, void* stream);
/* Measurement probe for bench.py's roofline: while on, rg_denoise / rg_denoise_groups of this handle (tensor-core
 * tiers) launch ONLY their dense contractions -- the same programmatic-dependent-launch chain, weights, shapes and
 * epilogues as in the step, without the attention and row kernels between them (outputs are then meaningless).
 * Time a few evaluations with CUDA events, then switch it off. */
int rg_probe_gemm_only(rg_handle h, int on);
/* Measurement probe for bench.py's second GEMM roofline: read bandwidth L2 -> SMs.  Every SM streams `buf`
 * (bytes <= ~64 MB so that it stays L2-resident) `passes` times with 128-bit loads; *gb_per_s receives
 * bytes * passes / CUDA-event time.  Synchronises. */
int rg_probe_l2_read(const void* buf, int64_t bytes, int passes, float* gb_per_s, void* stream);
/* Which tcgen05 GEMM kernel the tensor-core tiers launch (process-wide; results are bit-identical: all kernels
 * accumulate the K blocks in the same order in an fp32 TMEM accumulator and share the epilogue arithmetic):
 *   mode 0 (default): automatic -- by row count: the pair128 kernel from `pair128_min_rows` rows on, the 2-CTA 256x256
 *          kernel from `min_rows` rows on (N a multiple of 256), the 128x128 one-tile-per-CTA kernel otherwise;
 *   mode 1: always the 128x128 kernel (tcgen05.mma.cta_group::1, 64 flop per operand byte);
 *   mode 2: the 2-CTA 256x256 kernel whenever the shape allows (cta_group::2, 128 flop/B, TMA-store epilogue; one tile
 *          per pair with two CTAs per SM below `persist_tiles` pair tiles, persistent with two TMEM accumulator stages
 *          from there on);
 *   mode 3: the pair128 kernel whenever the shape allows (cta_group::2, a CTA pair shares each 128-row weight tile:
 *          85 flop/B, the 128x128 kernel's epilogue and CTA count).
 * Arguments <= 0 keep the current thresholds.  Used by the parity tests and bench.py. */
int rg_set_gemm_kernel(int mode, int min_rows, int persist_tiles, int pair128_min_rows);
/* Diagnostics: one traced launch of the tcgen05 GEMM on zero operands (after 3 untraced ones).  trace_host
 * receives 10 int64 per CTA (grid order x-fastest): clock64 at [0] entry, [1] prologue done, [2] producer
 * past griddepcontrol.wait, [3] first operand stage landed, [4] last MMA committed, [5] accumulator visible
 * to the epilogue, [6] TMEM drained to smem, [7] all stores issued; [8] globaltimer ns at entry, [9] SM id.
 * epilogue 0: bias -> fp32; 1: bias + fp32 residual -> fp32; 2: bias + GELU -> bf16 planes. */
int rg_probe_gemm_trace(int M, int N, int K, int split, int epilogue, int64_t* trace_host,
                        int64_t n_trace, void* stream);
/* LayerNorm over 512-wide rows, eps 1e-5; gamma/beta may be NULL (no affine). */
int rg_op_layernorm(const float* x, const float* gamma, const float* beta, float* out, int M,
                    void* stream);
int rg_op_silu(const float* x, float* out, int64_t n, void* stream);
/* StylizationBlock row part (stylization_block.py:38-39 up to the SiLU):
 * out = silu(LN(y)*(1+scale)+shift); ss [n_clips or 1, 1024] = emb_layers output. */
int rg_op_stylization_rows(const float* y, const float* gamma, const float* beta, const float* ss,
                           int ss_per_clip, int rows_per_clip, float* out, int M, void* stream);
/* EfficientSelfAttention core (efficient_attention.py:30-39) from fused qkv [B*T,1536]:
 * with_styl=1: out = silu(LN(Y)*(1+scale)+shift) (input of proj_out.out_layers);
 * with_styl=0: out = x_res + Y (time_embed_dim=None variant, :41-42). */
int rg_op_self_attention(const float* qkv, const float* src_mask, const float* gamma,
                         const float* beta, const float* ss, int ss_per_clip, const float* x_res,
                         float* out, int B, int T, int with_styl, void* stream);
/* EfficientCrossAttention core for ONE condition (efficient_attention.py:72-101 minus the
 * projections): q [B*T,512] pre-softmax queries, state [B,16,32,32], query_mask [B,T] or NULL. */
int rg_op_cross_attention(const float* q, const float* state, const float* query_mask,
                          const float* gamma, const float* beta, const float* ss, int ss_per_clip,
                          float* out, int B, int T, void* stream);
/* The attention cores of the fused denoiser without the Stylization prologue: Y before norm/scale/shift.
 * mode 0: fp32 FMA; 1: TF32 mma.sync (the RG_PREC_BF16 tier); 2: 3xTF32 hi/lo split (RG_PREC_BF16X3).
 * self: qkv [B*T,1536], src_mask [B,T] -> y [B*T,512].
 * cross: q3 [B*T,1536] pre-softmax queries of the three conditions, state [B,3,16,32,32],
 *        query_mask [3,B,T] or NULL -> y [B*T,1536]. */
int rg_op_self_attention_core(const float* qkv, const float* src_mask, float* y, int B, int T, int mode,
                              void* stream);
int rg_op_cross_attention_core(const float* q3, const float* state, const float* query_mask, float* y,
                               int B, int T, int mode, void* stream);
/* The same cores fused with the Stylization prologue of proj_out (LayerNorm -> *(1+scale)+shift -> SiLU),
 * as the tensor-core tiers run them: split 0 = TF32, 1 = 3xTF32.  ss = [scale(512) | shift(512)] rows:
 * self: ss [1024] or [B,1024] (ss_per_clip); cross: gamma3/beta3 [3,512], ss3 [3,1024] or [B,3,1024].
 * out: self [B*T,512], cross [B*T,1536] (fp32). */
int rg_op_self_attention_tc(const float* qkv, const float* src_mask, const float* gamma, const float* beta,
                            const float* ss, int ss_per_clip, float* out, int B, int T, int split, void* stream);
int rg_op_cross_attention_tc(const float* q3, const float* state, const float* query_mask, const float* gamma3,
                             const float* beta3, const float* ss3, int ss_per_clip, float* out, int B, int T,
                             int split, void* stream);
/* K/V -> state for ONE condition/layer: kv [B*N,1024] = [key | value] projections. */
int rg_op_kv_state(const float* kv, int n_tokens, int B, float* state, void* stream);

/* ---- exemplar retrieval (K10/K11) -------------------------------------------------------- */

/* Token-aligned text-similarity scores of one query against a padded exemplar database:
 *   score[i] = (1/m) * sum_{t<m} <query[t,:], db[i,t,:]>,  m = min(tq, db_len[i])
 * == mean(diag(Q D_i^T)) of sort_sidx_by_textsimilarity (rag/utils.py:86-132), un-normalised.
 * db [n, max_len, dim] zero-padded, db_len int32 [n]; subset int32 [n_sub] (indices to score) or
 * NULL for all; scores fp32 [n_sub or n]. */
int rg_text_similarity(const float* db, const int32_t* db_len, int64_t n, int max_len, int dim,
                       const float* query, int tq, const int32_t* subset, int64_t n_sub,
                       float* scores, void* stream);

/* Exact fp32 top-k by dot product of q queries against a database shard (flat embeddings):
 * out_idx int64 [q,k] (global index = local + idx_base), out_score fp32 [q,k], ordered by
 * (score desc, index asc) -- the order of Python's stable sorted(..., reverse=True). */
int rg_knn_topk(const float* db, int64_t n, int dim, const float* queries, int q, int k,
                int64_t idx_base, int64_t* out_idx, float* out_score, void* stream);
/* Measurement probe for bench.py's roofline: like rg_knn_topk for q <= 8 queries, but times ONLY the scan
 * kernel (`reps` launches, CUDA events on `stream`, flush_buf overwritten before each) -> *median_ms. */
int rg_probe_knn_scan(const float* db, int64_t n, int dim, const float* queries, int q, int k, int reps,
                      void* flush_buf, int64_t flush_bytes, float* median_ms, void* stream);
/* Merge `parts` per-shard candidate lists [parts,q,k] (as all-gathered) into the global top-k. */
int rg_knn_merge(const int64_t* idx_parts, const float* score_parts, int parts, int q, int k,
                 int64_t* out_idx, float* out_score, void* stream);
/* The same merge over the packed buffer that parallel.sharded_knn all-gathers in ONE collective: per rank one
 * block of part_stride_bytes (>= 12*q*k, multiple of 8) holding [idx int64 q*k | score fp32 q*k]. */
int rg_knn_merge_packed(const void* packed, int64_t part_stride_bytes, int parts, int q, int k,
                        int64_t* out_idx, float* out_score, void* stream);

/* ---- large query batches: tensor-core similarity + certified over-selection (knn_tc.cu) ----
 * Same contract and bit-identical results as rg_knn_topk, for the Q = 4096 sweep of configs[3] where one exact
 * pass per 8 queries would stream the shard 512 times.  The reference has no batched path that is used
 * (sort_sidx_by_textsimilarity_batched, rag/utils.py:135-168, is dead code with a different ranking); the
 * operation replaced is the per-candidate loop of rag/utils.py:107-118 applied to many queries at once.
 *
 * rg_knn_index_create: one-off bf16 copy of the shard [n, dim] (dim % 64 == 0, n < 2^31-256) and its largest
 * row norm; owns 2*n*dim bytes of device memory until rg_knn_index_destroy.  `db` is only read during the call.
 * rg_knn_topk_tc: S = Q*D^T on tcgen05 (bf16 operands, fp32 TMEM accumulators) with the top-16 of every
 * (query, chunk) selected in the epilogue, exact fp32 re-score of the best 64 candidates per query in the
 * summation order of rg_knn_topk, and a per-query error-bound certificate; queries without one are re-run
 * through rg_knn_topk and counted in *n_uncertified (may be NULL).  `db` must be the fp32 shard the index
 * was built from.  Synchronises `stream` once (reads the uncertified count). */
int rg_knn_index_create(const float* db, int64_t n, int dim, void** index, void* stream);
int rg_knn_index_destroy(void* index);
int rg_knn_topk_tc(void* index, const float* db, const float* queries, int q, int k, int64_t idx_base,
                   int64_t* out_idx, float* out_score, int32_t* n_uncertified, void* stream);
/* Measurement probe: times ONLY knn_tc_kernel (`reps` launches, CUDA events on `stream`, flush_buf overwritten
 * before each) -> *median_ms; algorithmic work of one launch = 2*q*n*dim flop. */
int rg_probe_knn_tc(void* index, const float* queries, int q, int reps, void* flush_buf, int64_t flush_bytes,
                    float* median_ms, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RG_B200_H */
