// Fused "efficient" (linear) attention kernels (efficient_attention.py:23-102).
//   softmax over the 32 head features of Q, softmax over tokens of K, A = K^T V (32x32 per head),
//   Y = Q A, then the StylizationBlock prologue (LN, modulation, SiLU) on the assembled rows.
// One head == one warp (head dim 32 == warp width): feature softmaxes and the 32x32 contractions
// are warp shuffles, A lives in 32 registers per lane.  One CTA (16 warps) per clip so that the
// full 512-wide rows needed by the LayerNorm are assembled in shared memory (T*2 KB) and never
// round-trip through HBM.  fp32 throughout: the -1e6 additive masks and LN statistics of the
// reference only make sense in fp32.
#include <math.h>
#include "rg_common.cuh"
#include "rg_rows.cuh"

namespace {

struct RgStyl3 { RgStylParams p[3]; };

// rows of Ysm -> LN/modulate/SiLU (or + residual) -> out; 16 warps stride over the T rows
__device__ __forceinline__ void finish_rows(const float* Ysm, int T, int clip, const RgStylParams& sp,
                                            int with_styl, const float* x_res, int ldr, const RgRowOut& out,
                                            int col0, int warp, int lane) {
    for (int n = warp; n < T; n += RG_H) {
        float4 v[4];
        load_row(Ysm + n * RG_D, lane, v);
        const long long row = (long long)clip * T + n;
        if (with_styl) {
            rg_styl_row(v, sp, clip, lane);
        } else {
            float4 r[4];
            load_row(x_res + row * ldr, lane, r);
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[j].x += r[j].x; v[j].y += r[j].y; v[j].z += r[j].z; v[j].w += r[j].w; }
        }
        rg_store_row_out(out, row, col0, lane, v);
    }
}

// Y[n, lane] = sum_d softmax_d(q[n, :])[d] * A[d]   for one head
__device__ __forceinline__ float q_dot_A(float q, const float A[RG_HD]) {
    const float qmax = rg_warp_max(q);
    const float e = expf(q - qmax);
    const float qs = e / rg_warp_sum(e);
    float y = 0.f;
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) y = fmaf(__shfl_sync(0xffffffffu, qs, d), A[d], y);
    return y;
}

__global__ void __launch_bounds__(512) sa_attn_kernel(const float* __restrict__ qkv,
                                                     const float* __restrict__ src_mask,
                                                     RgStylParams sp, const float* __restrict__ x_res,
                                                     RgRowOut out, int T, int with_styl) {
    extern __shared__ __align__(16) float Ysm[];   // [T][512]
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* base = qkv + (long long)b * T * (3 * RG_D) + warp * RG_HD + lane;
    const float* mrow = src_mask + (long long)b * T;

    // token softmax of K (column `lane` of this head): max, then normaliser
    float kmax = -INFINITY;
    for (int n = 0; n < T; ++n)
        kmax = fmaxf(kmax, base[(long long)n * (3 * RG_D) + RG_D] + (1.0f - mrow[n]) * RG_NEG_MASK);
    float ksum = 0.f;
    for (int n = 0; n < T; ++n)
        ksum += expf(base[(long long)n * (3 * RG_D) + RG_D] + (1.0f - mrow[n]) * RG_NEG_MASK - kmax);

    // A[d][lane] = sum_n softmaxK[n][d] * (V[n][lane] * mask[n])
    float A[RG_HD];
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) A[d] = 0.f;
    for (int n = 0; n < T; ++n) {
        const float m = mrow[n];
        const float kk = base[(long long)n * (3 * RG_D) + RG_D] + (1.0f - m) * RG_NEG_MASK;
        const float ks = expf(kk - kmax) / ksum;
        const float vv = base[(long long)n * (3 * RG_D) + 2 * RG_D] * m;
#pragma unroll
        for (int d = 0; d < RG_HD; ++d) A[d] = fmaf(__shfl_sync(0xffffffffu, ks, d), vv, A[d]);
    }
    for (int n = 0; n < T; ++n)
        Ysm[n * RG_D + warp * RG_HD + lane] = q_dot_A(base[(long long)n * (3 * RG_D)], A);
    __syncthreads();
    finish_rows(Ysm, T, b, sp, with_styl, x_res, RG_D, out, 0, warp, lane);
}

__global__ void __launch_bounds__(512) ca_attn_kernel(const float* __restrict__ q3, int ldq,
                                                     const float* __restrict__ state,
                                                     long long state_clip_stride,
                                                     long long state_cond_stride,
                                                     const float* __restrict__ qmask,
                                                     long long qmask_cond_stride, RgStyl3 sp3,
                                                     RgRowOut out, int T) {
    extern __shared__ __align__(16) float Ysm[];
    const int b = blockIdx.x, c = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* Ap = state + (long long)b * state_clip_stride + (long long)c * state_cond_stride +
                      (long long)warp * RG_HD * RG_HD + lane;
    float A[RG_HD];
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) A[d] = __ldg(Ap + d * RG_HD);
    const float* qb = q3 + (long long)b * T * ldq + c * RG_D + warp * RG_HD + lane;
    const float* qm = qmask ? qmask + (long long)c * qmask_cond_stride + (long long)b * T : nullptr;
    for (int n = 0; n < T; ++n) {
        float y = q_dot_A(qb[(long long)n * ldq], A);
        if (qm) y = y + (1.0f - qm[n]) * RG_NEG_MASK;     // fp32 add: y - 1e6 rounds to a 1/16 grid
        Ysm[n * RG_D + warp * RG_HD + lane] = y;
    }
    __syncthreads();
    finish_rows(Ysm, T, b, sp3.p[c], 1, nullptr, 0, out, c * RG_D, warp, lane);
}

// state[b][set][h][d][l] = sum_n softmax_n(K[b,n,h,d]) * V[b,n,h,l]; warp = d, lane = l
__global__ void __launch_bounds__(1024) kv_state_kernel(const float* __restrict__ kv, int ldkv,
                                                       int k_off, int v_off, int N,
                                                       float* __restrict__ state,
                                                       long long state_clip_stride,
                                                       int kv_set_stride, long long state_set_stride) {
    const int b = blockIdx.x, h = blockIdx.y, set = blockIdx.z;
    const int d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* rows = kv + (long long)b * N * ldkv + (long long)set * kv_set_stride;
    const float* kcol = rows + k_off + h * RG_HD + d;
    const float* vcol = rows + v_off + h * RG_HD + lane;
    float m = -INFINITY;
    for (int n = lane; n < N; n += 32) m = fmaxf(m, kcol[(long long)n * ldkv]);
    m = rg_warp_max(m);
    float s = 0.f;
    for (int n = lane; n < N; n += 32) s += expf(kcol[(long long)n * ldkv] - m);
    s = rg_warp_sum(s);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
        acc = fmaf(expf(kcol[(long long)n * ldkv] - m), vcol[(long long)n * ldkv], acc);
    state[(long long)b * state_clip_stride + (long long)set * state_set_stride +
          ((long long)h * RG_HD + d) * RG_HD + lane] = acc / s;
}

}  // namespace

cudaError_t rg_launch_sa_attention(const float* qkv, const float* src_mask, RgStylParams sp,
                                   const float* x_res, RgRowOut out, int B, int T, int with_styl,
                                   cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > RG_MAX_T) return cudaErrorInvalidValue;
    const size_t smem = (size_t)T * RG_D * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(sa_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         RG_MAX_T * RG_D * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    sa_attn_kernel<<<B, 512, smem, st>>>(qkv, src_mask, sp, x_res, out, T, with_styl);
    return cudaGetLastError();
}

cudaError_t rg_launch_ca_attention(const float* q3, int ldq, const float* state,
                                   long long state_clip_stride, long long state_cond_stride,
                                   const float* qmask, long long qmask_cond_stride,
                                   const RgStylParams* sp3, RgRowOut out, int B, int T,
                                   int n_cond, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > RG_MAX_T || n_cond < 1 || n_cond > 3) return cudaErrorInvalidValue;
    RgStyl3 s3;
    for (int c = 0; c < 3; ++c) s3.p[c] = sp3[c < n_cond ? c : 0];
    const size_t smem = (size_t)T * RG_D * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(ca_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         RG_MAX_T * RG_D * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    ca_attn_kernel<<<dim3(B, n_cond), 512, smem, st>>>(q3, ldq, state, state_clip_stride,
                                                       state_cond_stride, qmask, qmask_cond_stride,
                                                       s3, out, T);
    return cudaGetLastError();
}

cudaError_t rg_launch_kv_state(const float* kv, int ldkv, int k_off, int v_off, int n_tokens,
                               float* state, long long state_clip_stride, int B, int n_sets,
                               int kv_set_stride, long long state_set_stride, cudaStream_t st) {
    if (B <= 0 || n_tokens <= 0) return cudaSuccess;
    kv_state_kernel<<<dim3(B, RG_H, n_sets), 1024, 0, st>>>(kv, ldkv, k_off, v_off, n_tokens, state,
                                                            state_clip_stride, kv_set_stride,
                                                            state_set_stride);
    return cudaGetLastError();
}
