// Shared declarations for the sm_100a kernels of the guided-DDIM hot path.
// Fixed by the reference config (basegesture_len150_beat.py:32-42): latent_dim 512, 16 heads of
// 32 -- one head == one warp lane per feature -- so these are compile-time; T (tokens), F, E, L
// and the condition lengths are runtime.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define RG_D 512        // latent_dim
#define RG_HD 32        // head dim == warp size
#define RG_H 16         // heads
#define RG_MAX_T 64     // smem rows reserved per clip (T = 43 in the shipped config)
#define RG_NEG_MASK (-1000000.0f)

enum RgEpilogue : int {
    RG_EPI_BIAS = 0,        // C = A W^T + b
    RG_EPI_BIAS_RESIDUAL,   // C = A W^T + b + R
    RG_EPI_BIAS_GELU,       // C = gelu_erf(A W^T + b)
    RG_EPI_BIAS_POS,        // C = A W^T + b + pos[row % pos_T]      (joint_embed)
    RG_EPI_BIAS_SILU,       // C = silu(A W^T + b)                   (time_embed.0)
};

// C[M,N] = A[M,K] * W[N,K]^T (+ epilogue); both operands K-major (nn.Linear layout).
// blockIdx.z selects a group: every pointer advances by its *_g element stride.
struct RgGemm {
    const float* A; const float* W; const float* bias; float* C; const float* R; const float* pos;
    int M, N, K;
    int lda, ldw, ldc, ldr;
    long long a_g, w_g, b_g, c_g, r_g;
    int groups;
    int pos_T;
    int epi;
};

struct RgStylParams {           // one StylizationBlock's row-wise part
    const float* gamma;         // norm.weight [512]
    const float* beta;          // norm.bias   [512]
    const float* ss;            // [scale(512) | shift(512)] per clip
    long long ss_clip_stride;   // 0: the same row for every clip (one timestep for the batch)
};

// Destination of a row-wise kernel: fp32 rows, or bf16 rows (hi plane, plus the lo = bf16(x - hi)
// plane lo_off columns further when lo_off != 0) for the tensor-core GEMM that consumes them.
struct RgRowOut {
    float* f32;
    __nv_bfloat16* b16;
    int ld;
    int lo_off;
};
static inline RgRowOut rg_out_f32(float* p, int ld) { RgRowOut o = {p, nullptr, ld, 0}; return o; }
static inline RgRowOut rg_out_b16(void* p, int ld, int lo_off) {
    RgRowOut o = {nullptr, reinterpret_cast<__nv_bfloat16*>(p), ld, lo_off};
    return o;
}

cudaError_t rg_launch_gemm_f32(const RgGemm& g, cudaStream_t st);

// rowops.cu
cudaError_t rg_launch_ln_rows(const float* x, int ldx, const float* gamma, const float* beta,
                              RgRowOut out, int M, cudaStream_t st);
cudaError_t rg_launch_styl_rows(const float* y, int ldy, RgStylParams sp, int rows_per_clip,
                                RgRowOut out, int M, cudaStream_t st);
cudaError_t rg_launch_silu(const float* x, float* out, long long n, cudaStream_t st);
cudaError_t rg_launch_gather_rows(const float* table, const long long* idx, float* out,
                                  long long n_rows, int n_table, cudaStream_t st);
cudaError_t rg_launch_fold_ln(const float* W, const float* b, const float* gamma,
                              const float* beta, float* Wf, float* bf, int N, int K,
                              cudaStream_t st);
cudaError_t rg_launch_ddim_update(const float* x, const float* x0, float* out, long long n,
                                  float c_recip, float c_recipm1, float c_a, float c_b,
                                  cudaStream_t st);
cudaError_t rg_launch_blend(const float* x, const float* in_seq, const float* noise, float* out,
                            long long rows, float s_ab, float s_1mab, cudaStream_t st);
cudaError_t rg_launch_mix_branches(const float* out2, const float* coef, const float* joint_scale, float* out,
                                   long long rows, int T, cudaStream_t st);
cudaError_t rg_launch_guidance(float* x, const float* in_seq, long long rows, int iters,
                               float lr_2_over_n, cudaStream_t st);
cudaError_t rg_launch_transpose_sq(const float* in, float* out, int n, cudaStream_t st);
cudaError_t rg_launch_sum3_blocks(const float* a, int lda, float* out, int ldo, int n, int rows, cudaStream_t st);
cudaError_t rg_launch_pos_table(const float* seq_pe, const float* glob_pe, float* pos, int T,
                                int n_chunks, cudaStream_t st);

// attention.cu
cudaError_t rg_launch_sa_attention(const float* qkv, const float* src_mask, RgStylParams sp,
                                   const float* x_res, RgRowOut out, int B, int T, int with_styl,
                                   cudaStream_t st);
cudaError_t rg_launch_ca_attention(const float* q3, int ldq, const float* state,
                                   long long state_clip_stride, long long state_cond_stride,
                                   const float* qmask, long long qmask_cond_stride,
                                   const RgStylParams* sp3, RgRowOut out, int B, int T,
                                   int n_cond, cudaStream_t st);
cudaError_t rg_launch_sa_core(const float* qkv, const float* src_mask, float* Y, int B, int T, int mode, cudaStream_t st);
cudaError_t rg_launch_ca_core(const float* q3, int ldq, const float* state, long long state_clip_stride,
                              long long state_cond_stride, const float* qmask, long long qmask_cond_stride,
                              float* Y, int ldy, int B, int T, int mode, cudaStream_t st);
cudaError_t rg_launch_sa_styl(const float* qkv, const float* src_mask, RgStylParams sp, RgRowOut out, int B, int T,
                              int split, cudaStream_t st);
cudaError_t rg_launch_ca_styl(const float* q3, int ldq, const float* state, long long state_clip_stride,
                              long long state_cond_stride, const float* qmask, long long qmask_cond_stride,
                              const RgStylParams* sp3, RgRowOut out, int B, int T, int split, cudaStream_t st);
cudaError_t rg_launch_styl_rows3(const float* y, int ldy, const RgStylParams* sp3, int rows_per_clip,
                                 RgRowOut out, int M, cudaStream_t st);
cudaError_t rg_launch_kv_state(const float* kv, int ldkv, int k_off, int v_off, int n_tokens,
                               float* state, long long state_clip_stride, int B, int n_sets,
                               int kv_set_stride, long long state_set_stride, cudaStream_t st);

// mha.cu
cudaError_t rg_launch_mha(const float* q, const float* k, const float* v, const unsigned char* keep, float* out,
                          int N, int Sq, int Sk, int H, int dh, long long ldq, long long ldk, long long ldv,
                          float scale, cudaStream_t st);

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// Every kernel of the per-step chain is launched with programmaticStreamSerialization and starts with
// rg_pdl_launch() (lets the NEXT kernel's CTAs be scheduled early) and rg_pdl_wait() (blocks until the
// PREVIOUS kernel has completed and its writes are visible) before touching dependent data: the launch
// latency and prologue of kernel N+1 overlap the tail of kernel N.  Without the launch attribute both
// instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void rg_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// RULE for every kernel launched with PDL: pointers to data the PREVIOUS kernel may have produced must not
// be `const T* __restrict__`.  Such loads compile to ld.global.nc, which nvcc/ptxas treat as invariant and
// hoist ABOVE griddepcontrol.wait -- reading the predecessor's output before it is written (seen on
// sa_core_mma_kernel).  Plain loads are ordered by the "memory" clobber below.  tools/pdl_lint.py checks the
// SASS of every kernel for global loads ahead of the wait and runs in the CPU test suite (tests/test_abi.py).
__device__ __forceinline__ void rg_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t rg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float rg_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float rg_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// exp via ex2.approx (<= 2 ulp + |x|*6e-8 relative) and approximate division: these sit in the inner
// loops of every row/attention kernel; their error is far below every parity tier (tests).
__device__ __forceinline__ float rg_exp(float v) { return __expf(v); }
__device__ __forceinline__ float rg_silu(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// GELU(erf) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 absolute, ~15 instructions against
// ~40 for erff): used in the tensor-core tiers' GEMM epilogue, where it was half of the epilogue's time.
__device__ __forceinline__ float rg_gelu_fast(float v) {
    const float x = v * 0.70710678118654752440f, ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = fmaf(-p * t, __expf(-ax * ax), 1.0f);
    return 0.5f * v * (1.0f + copysignf(e, x));
}
__device__ __forceinline__ float rg_gelu_erf(float v) {
    return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}
