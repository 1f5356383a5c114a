// Fused "efficient" (linear) attention kernels (efficient_attention.py:23-102).
//   softmax over the 32 head features of Q, softmax over tokens of K, A = K^T V (32x32 per head),
//   Y = Q A, then the StylizationBlock prologue (LN, modulation, SiLU) on the assembled rows.
// One head == one warp (head dim 32 == warp width): feature softmaxes and the 32x32 contractions
// are warp shuffles, A lives in 32 registers per lane.  One CTA (16 warps) per clip so that the
// full 512-wide rows needed by the LayerNorm are assembled in shared memory (T*2 KB) and never
// round-trip through HBM.  fp32 throughout: the -1e6 additive masks and LN statistics of the
// reference only make sense in fp32.
#include <math.h>
#include <stdlib.h>
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

// Per head (= per warp) the kernels keep A[32][32] as 32 registers per lane (lane = output feature l)
// and use this warp's [T][32] slice of the shared Y buffer three times over: first for exp(K - max),
// then for softmax(Q), finally for Y itself.  Rows are read back as broadcast 128-bit LDS (all lanes
// read the same 4 values), so the 32x32 contractions run on the FMA pipe instead of 32 shuffles per
// token (shuffle issue rate, not arithmetic, bounded the first version: 56 us -> see profiles/).
__device__ __forceinline__ void row_times_A(const float* row, const float A[RG_HD], float& y) {
    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;       // four chains: the 32 FMAs are not serialised
#pragma unroll
    for (int d4 = 0; d4 < RG_HD / 4; ++d4) {
        const float4 q = *reinterpret_cast<const float4*>(row + d4 * 4);
        y0 = fmaf(q.x, A[d4 * 4 + 0], y0); y1 = fmaf(q.y, A[d4 * 4 + 1], y1);
        y2 = fmaf(q.z, A[d4 * 4 + 2], y2); y3 = fmaf(q.w, A[d4 * 4 + 3], y3);
    }
    y = (y0 + y1) + (y2 + y3);
}

// softmax over the 32 head features of Q for tokens [0,T) -> slice[n][lane]; loads batched CH-wide
__device__ __forceinline__ void q_softmax_to_smem(const float* qcol, long long stride, int T, float* slice, int lane,
                                                  int pitch = RG_D) {
    constexpr int CH = 8;
    for (int n0 = 0; n0 < T; n0 += CH) {
        float qq[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) qq[i] = (n0 + i < T) ? qcol[(n0 + i) * stride] : 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (n0 + i < T) {
                const float e = rg_exp(qq[i] - rg_warp_max(qq[i]));
                slice[(n0 + i) * pitch + lane] = __fdividef(e, rg_warp_sum(e));
            }
        }
    }
}

// One head of EfficientSelfAttention on one warp: on return slice[n*pitch + lane] = Y[n][head*32 + lane].
__device__ __forceinline__ void sa_head(const float* base, const float* mrow, int T, float* slice, int pitch, int lane) {
    constexpr int CH = 8;
    const long long RS = 3 * RG_D;
    // sweep 1: column max of K (+ -1e6 on masked tokens, efficient_attention.py:32); lane = feature d
    float kmax = -INFINITY;
    for (int n0 = 0; n0 < T; n0 += CH) {
        float kk[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int n = n0 + i;
            kk[i] = n < T ? base[n * RS + RG_D] + (1.0f - mrow[n]) * RG_NEG_MASK : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) kmax = fmaxf(kmax, kk[i]);
    }
    // sweep 2a: E[n][d] = exp(K - max) -> smem, column sums in registers
    float ksum = 0.f;
    for (int n0 = 0; n0 < T; n0 += CH) {
        float kk[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int n = n0 + i;
            kk[i] = n < T ? base[n * RS + RG_D] + (1.0f - mrow[n]) * RG_NEG_MASK : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (n0 + i < T) {
                const float e = rg_exp(kk[i] - kmax);
                ksum += e;
                slice[(n0 + i) * pitch + lane] = e;
            }
        }
    }
    __syncwarp();
    // sweep 2b: A[d][lane] = sum_n E[n][d] * (V[n][lane] * mask[n]); lane = feature l
    float A[RG_HD];
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) A[d] = 0.f;
    for (int n0 = 0; n0 < T; n0 += CH) {
        float vv[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) vv[i] = (n0 + i < T) ? base[(n0 + i) * RS + 2 * RG_D] * mrow[n0 + i] : 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (n0 + i < T) {
                const float* row = slice + (n0 + i) * pitch;
#pragma unroll
                for (int d4 = 0; d4 < RG_HD / 4; ++d4) {
                    const float4 e = *reinterpret_cast<const float4*>(row + d4 * 4);
                    A[d4 * 4 + 0] = fmaf(e.x, vv[i], A[d4 * 4 + 0]); A[d4 * 4 + 1] = fmaf(e.y, vv[i], A[d4 * 4 + 1]);
                    A[d4 * 4 + 2] = fmaf(e.z, vv[i], A[d4 * 4 + 2]); A[d4 * 4 + 3] = fmaf(e.w, vv[i], A[d4 * 4 + 3]);
                }
            }
        }
    }
    __syncwarp();
    slice[lane] = ksum;                              // row 0 is free again: publish the normalisers
    __syncwarp();
#pragma unroll
    for (int d4 = 0; d4 < RG_HD / 4; ++d4) {
        const float4 s4 = *reinterpret_cast<const float4*>(slice + d4 * 4);
        A[d4 * 4 + 0] = __fdividef(A[d4 * 4 + 0], s4.x); A[d4 * 4 + 1] = __fdividef(A[d4 * 4 + 1], s4.y);
        A[d4 * 4 + 2] = __fdividef(A[d4 * 4 + 2], s4.z); A[d4 * 4 + 3] = __fdividef(A[d4 * 4 + 3], s4.w);
    }
    __syncwarp();
    // sweep 3: softmax_features(Q) -> smem, then Y = Q A written in place
    q_softmax_to_smem(base, RS, T, slice, lane, pitch);
    __syncwarp();
    for (int n = 0; n < T; ++n) {
        float y;
        row_times_A(slice + n * pitch, A, y);
        __syncwarp();
        slice[n * pitch + lane] = y;
    }
}

__global__ void __launch_bounds__(512, 2) sa_attn_kernel(const float* __restrict__ qkv,
                                                        const float* __restrict__ src_mask,
                                                        RgStylParams sp, const float* __restrict__ x_res,
                                                        RgRowOut out, int T, int with_styl) {
    extern __shared__ __align__(16) float Ysm[];   // [T][512]
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sa_head(qkv + (long long)b * T * (3 * RG_D) + warp * RG_HD + lane, src_mask + (long long)b * T, T,
            Ysm + warp * RG_HD, RG_D, lane);
    __syncthreads();
    finish_rows(Ysm, T, b, sp, with_styl, x_res, RG_D, out, 0, warp, lane);
}

// Finer-grained variants for the fused denoiser: ONE head per CTA, its tokens split over 4 warps
// (warp w takes tokens n = w, w+4, ...), Y goes to global fp32 and the Stylization prologue runs as a
// row kernel.  At 64-160 clips the one-CTA-per-clip kernels left most SMs idle and each warp walked a
// ~6k-instruction dependent chain (33 us measured); here there are 16x/48x more CTAs, every chain is
// 4x shorter and the per-warp partials of max / sum / A are combined through shared memory.
constexpr int TW = 4;                                   // warps (token groups) per head
struct HeadSmem {
    float row[RG_MAX_T][RG_HD];                         // E, then softmax(Q), per token
    float Apart[TW][RG_HD][RG_HD];                      // per-warp partial K^T V
    float kmaxp[TW][RG_HD], ksump[TW][RG_HD];
};

template <int MAXN>
__global__ void __launch_bounds__(TW * 32, 8) sa_core_kernel(const float* qkv,
                                                         const float* src_mask,
                                                         float* __restrict__ Y, int T) {
    __shared__ __align__(16) HeadSmem sm;
    rg_pdl_launch();
    rg_pdl_wait();
    const int b = blockIdx.x, head = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* base = qkv + (long long)b * T * (3 * RG_D) + head * RG_HD + lane;
    const float* mrow = src_mask + (long long)b * T;
    const long long RS = 3 * RG_D;
    // own tokens' keys (+ -1e6 on masked tokens) and values in registers: all loads in flight at once
    float kk[MAXN], vv[MAXN];
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        const float m = n < T ? mrow[n] : 0.f;
        kk[i] = n < T ? base[n * RS + RG_D] + (1.0f - m) * RG_NEG_MASK : -INFINITY;
        vv[i] = n < T ? base[n * RS + 2 * RG_D] * m : 0.f;
    }
    float kmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXN; ++i) kmax = fmaxf(kmax, kk[i]);
    sm.kmaxp[w][lane] = kmax;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TW; ++j) kmax = fmaxf(kmax, sm.kmaxp[j][lane]);
    float ksum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        if (n < T) {
            const float e = rg_exp(kk[i] - kmax);
            ksum += e;
            sm.row[n][lane] = e;
        }
    }
    sm.ksump[w][lane] = ksum;
    __syncwarp();
    // partial A over own tokens: A[d][lane] += E[n][d] * V[n][lane]
    float A[RG_HD];
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) A[d] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        if (n < T) {
#pragma unroll
            for (int d4 = 0; d4 < RG_HD / 4; ++d4) {
                const float4 e = *reinterpret_cast<const float4*>(&sm.row[n][d4 * 4]);
                A[d4 * 4 + 0] = fmaf(e.x, vv[i], A[d4 * 4 + 0]); A[d4 * 4 + 1] = fmaf(e.y, vv[i], A[d4 * 4 + 1]);
                A[d4 * 4 + 2] = fmaf(e.z, vv[i], A[d4 * 4 + 2]); A[d4 * 4 + 3] = fmaf(e.w, vv[i], A[d4 * 4 + 3]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) sm.Apart[w][d][lane] = A[d];
    // queries of own tokens (loads overlap the barrier)
    float qq[MAXN];
#pragma unroll
    for (int i = 0; i < MAXN; ++i) qq[i] = (w + i * TW < T) ? base[(w + i * TW) * RS] : 0.f;
    __syncthreads();
    // full normalised A = sum of the partials / column sums
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) {
        float a = 0.f, s = 0.f;
#pragma unroll
        for (int j = 0; j < TW; ++j) { a += sm.Apart[j][d][lane]; s += sm.ksump[j][d]; }
        A[d] = __fdividef(a, s);
    }
    // softmax_features(Q) then Y = Q A for own tokens
    float* o = Y + (long long)b * T * RG_D + head * RG_HD + lane;
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        if (n < T) {
            const float e = rg_exp(qq[i] - rg_warp_max(qq[i]));
            sm.row[n][lane] = __fdividef(e, rg_warp_sum(e));
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        if (n < T) {
            float y;
            row_times_A(&sm.row[n][0], A, y);
            o[(long long)n * RG_D] = y;
        }
    }
}

template <int MAXN>
__global__ void __launch_bounds__(TW * 32, 8) ca_core_kernel(const float* q3, int ldq,
                                                         const float* __restrict__ state,
                                                         long long state_clip_stride, long long state_cond_stride,
                                                         const float* qmask,
                                                         long long qmask_cond_stride, float* __restrict__ Y,
                                                         int ldy, int T) {
    __shared__ __align__(16) float rows[RG_MAX_T][RG_HD];
    rg_pdl_launch();
    rg_pdl_wait();
    const int b = blockIdx.x, c = blockIdx.y, head = blockIdx.z, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* qb = q3 + (long long)b * T * ldq + c * RG_D + head * RG_HD + lane;
    const float* qm = qmask ? qmask + (long long)c * qmask_cond_stride + (long long)b * T : nullptr;
    float qq[MAXN], mk[MAXN];
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        qq[i] = n < T ? qb[(long long)n * ldq] : 0.f;
        mk[i] = (qm && n < T) ? qm[n] : 1.f;
    }
    const float* Ap = state + (long long)b * state_clip_stride + (long long)c * state_cond_stride +
                      (long long)head * RG_HD * RG_HD + lane;
    float A[RG_HD];
#pragma unroll
    for (int d = 0; d < RG_HD; ++d) A[d] = __ldg(Ap + d * RG_HD);
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        if (n < T) {
            const float e = rg_exp(qq[i] - rg_warp_max(qq[i]));
            rows[n][lane] = __fdividef(e, rg_warp_sum(e));
        }
    }
    __syncwarp();
    float* o = Y + (long long)b * T * ldy + c * RG_D + head * RG_HD + lane;
#pragma unroll
    for (int i = 0; i < MAXN; ++i) {
        const int n = w + i * TW;
        if (n < T) {
            float y;
            row_times_A(&rows[n][0], A, y);
            if (qm) y = y + (1.0f - mk[i]) * RG_NEG_MASK;   // fp32 add: y - 1e6 rounds to a 1/16 grid
            o[(long long)n * ldy] = y;
        }
    }
}

// ---- tensor-core cores (mma.sync m16n8k8, TF32 operands, fp32 accumulate) --------------------------------
// The per-head products are 43x32x32: far below a tcgen05 tile (M >= 64, one issuing thread per CTA), but
// the warp-level mma.sync fits them exactly.  ONE WARP owns one (clip, head): everything stays in
// registers, there is no shared memory and no barrier.  Fragment ownership (g = lane/4, t = lane%4):
//   A (16x8, row): a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4);  B (8x8, col): b0=(k=t,n=g) b1=(k=t+4,n=g)
//   C (16x8): c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)
// Reduction indices may be permuted freely as long as A and B agree, so thread t takes
//   * feature reductions (32 = 4 k-steps): logical (ks, t) -> d = 8ks+2t, (ks, t+4) -> d = 8ks+2t+1
//     (one float2 load per row and k-step, and exactly the columns of a C fragment), and
//   * token reductions (16*MT tokens = 2*MT k-steps): thread group t owns tokens [t*4MT, (t+1)*4MT).
// With A^T = (V*m)^T softmax_n(K) computed as the first product, its C fragments ARE the B fragments of
// Y = softmax_d(Q) A: no transpose, no exchange.  Softmaxes reduce over 8 (features) or 4MT (tokens)
// in-thread values plus two xor-shuffles over t.
// SPLIT (bf16x3 tier): operands are split hi + lo in TF32 and three products accumulate (lo*hi, hi*lo,
// hi*hi): fp32-class accuracy.  Unsplit (bf16 tier): operands rounded to TF32 (rel. 2^-11).
__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float c[4], const uint32_t a[4], const uint32_t b[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <bool SPLIT>
__device__ __forceinline__ void mma_f32(float c[4], const float a[4], const float b[2]) {
    uint32_t ah[4], bh[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) ah[i] = f2tf32(a[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) bh[i] = f2tf32(b[i]);
    if (SPLIT) {
        uint32_t al[4], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) al[i] = f2tf32(a[i] - __uint_as_float(ah[i]));
#pragma unroll
        for (int i = 0; i < 2; ++i) bl[i] = f2tf32(b[i] - __uint_as_float(bh[i]));
        mma_tf32(c, al, bh);
        mma_tf32(c, ah, bl);
    }
    mma_tf32(c, ah, bh);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// softmax over the 32 features of rows r0 / r1 held as 4 float2 per row across the 4 threads of a quad
__device__ __forceinline__ void quad_row_softmax(float2 q[4]) {
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) mx = fmaxf(mx, fmaxf(q[k].x, q[k].y));
    mx = quad_max(mx);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { q[k].x = rg_exp(q[k].x - mx); q[k].y = rg_exp(q[k].y - mx); s += q[k].x + q[k].y; }
    s = quad_sum(s);
#pragma unroll
    for (int k = 0; k < 4; ++k) { q[k].x = __fdividef(q[k].x, s); q[k].y = __fdividef(q[k].y, s); }
}

// y[mt][nt][*] = C fragments of softmax_d(Q) (softmax_n(K + mask)^T (V m)) for one (clip, head):
// rows 16mt+g (+8), columns 8nt+2t (+1)   (efficient_attention.py:146-160)
template <int MT, bool SPLIT>
__device__ __forceinline__ void sa_head_frags(const float* base, const float* mrow, int T, int g, int t,
                                              float (&y)[MT][4][4]) {
    constexpr int KS = 2 * MT, TPT = 2 * KS;             // token k-steps; tokens owned by one thread group
    const long long RS = 3 * RG_D;
    const int n0 = t * TPT;
    float mk[TPT], kk[TPT][4], vv[TPT][4];               // K[n][8j+g] (+ -1e6 mask), V[n][8j+g] * m
#pragma unroll
    for (int i = 0; i < TPT; ++i) mk[i] = (n0 + i < T) ? mrow[n0 + i] : 0.f;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
        const float* row = base + (long long)(n0 + i) * RS + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            kk[i][j] = (n0 + i < T) ? row[RG_D + 8 * j] : 0.f;
            vv[i][j] = (n0 + i < T) ? row[2 * RG_D + 8 * j] : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < TPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            kk[i][j] = (n0 + i < T) ? kk[i][j] + (1.0f - mk[i]) * RG_NEG_MASK : -INFINITY;
            vv[i][j] *= mk[i];
        }
    // softmax over the tokens, per feature column 8j+g, normalised here (a/s == sum((e/s) v) up to rounding)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < TPT; ++i) mx = fmaxf(mx, kk[i][j]);
        mx = quad_max(mx);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < TPT; ++i) { kk[i][j] = rg_exp(kk[i][j] - mx); s += kk[i][j]; }
        const float r = __fdividef(1.0f, quad_sum(s));
#pragma unroll
        for (int i = 0; i < TPT; ++i) kk[i][j] *= r;
    }
    // A^T[l][d] = sum_n (V m)[n][l] E[n][d]:  M = l (2 tiles), N = d (4 tiles), K = tokens
    float at[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) at[mt][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int i0 = 2 * ks, i1 = 2 * ks + 1;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float a[4] = {vv[i0][2 * mt], vv[i0][2 * mt + 1], vv[i1][2 * mt], vv[i1][2 * mt + 1]};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float bb[2] = {kk[i0][nt], kk[i1][nt]};
                mma_f32<SPLIT>(at[mt][nt], a, bb);
            }
        }
    }
    // Y = softmax_d(Q) A, 16 query rows per tile; B fragments are the C fragments above
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int r0 = 16 * mt + g, r1 = r0 + 8;
        float2 q0[4], q1[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            q0[ks] = r0 < T ? *reinterpret_cast<const float2*>(base + (long long)r0 * RS + 8 * ks + 2 * t) : make_float2(0.f, 0.f);
            q1[ks] = r1 < T ? *reinterpret_cast<const float2*>(base + (long long)r1 * RS + 8 * ks + 2 * t) : make_float2(0.f, 0.f);
        }
        quad_row_softmax(q0);
        quad_row_softmax(q1);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) y[mt][nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float a[4] = {q0[ks].x, q1[ks].x, q0[ks].y, q1[ks].y};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float bb[2] = {at[nt >> 1][ks][2 * (nt & 1)], at[nt >> 1][ks][2 * (nt & 1) + 1]};
                mma_f32<SPLIT>(y[mt][nt], a, bb);
            }
        }
    }
}

// y = C fragments of softmax_d(Q_c) state[b, c, head]  (+ -1e6 on rows with query_mask 0) for one (clip, cond, head)
template <int MT, bool SPLIT>
__device__ __forceinline__ void ca_head_frags(const float* qb, int ldq, const float* Ap, const float* qm, int T,
                                              int g, int t, float (&y)[MT][4][4]) {
    float bf[4][4][2];                                   // [ks][nt][j] = A[d = 8ks+2t+j][l = 8nt+g]
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[ks][nt][j] = __ldg(Ap + (8 * ks + 2 * t + j) * RG_HD + 8 * nt + g);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int r0 = 16 * mt + g, r1 = r0 + 8;
        float2 q0[4], q1[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            q0[ks] = r0 < T ? *reinterpret_cast<const float2*>(qb + (long long)r0 * ldq + 8 * ks) : make_float2(0.f, 0.f);
            q1[ks] = r1 < T ? *reinterpret_cast<const float2*>(qb + (long long)r1 * ldq + 8 * ks) : make_float2(0.f, 0.f);
        }
        const float add0 = (qm && r0 < T) ? (1.0f - qm[r0]) * RG_NEG_MASK : 0.f;
        const float add1 = (qm && r1 < T) ? (1.0f - qm[r1]) * RG_NEG_MASK : 0.f;
        quad_row_softmax(q0);
        quad_row_softmax(q1);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) y[mt][nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float a[4] = {q0[ks].x, q1[ks].x, q0[ks].y, q1[ks].y};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_f32<SPLIT>(y[mt][nt], a, bf[ks][nt]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {      // fp32 add: y - 1e6 rounds to a 1/16 grid, as in the reference
            y[mt][nt][0] += add0; y[mt][nt][1] += add0; y[mt][nt][2] += add1; y[mt][nt][3] += add1;
        }
    }
}

template <int MT>
__device__ __forceinline__ void store_frags(const float (&y)[MT][4][4], float* o, int ld, int T, int g) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int r0 = 16 * mt + g, r1 = r0 + 8;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            if (r0 < T) *reinterpret_cast<float2*>(o + (long long)r0 * ld + 8 * nt) = make_float2(y[mt][nt][0], y[mt][nt][1]);
            if (r1 < T) *reinterpret_cast<float2*>(o + (long long)r1 * ld + 8 * nt) = make_float2(y[mt][nt][2], y[mt][nt][3]);
        }
    }
}

// The cores alone (Y before the Stylization prologue): 4 heads per CTA, one per warp.
template <int MT, bool SPLIT>
__global__ void __launch_bounds__(128) sa_core_mma_kernel(const float* qkv, const float* src_mask,
                                                         float* __restrict__ Y, int T) {
    rg_pdl_launch();
    rg_pdl_wait();
    const int b = blockIdx.x, head = blockIdx.y * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float y[MT][4][4];
    sa_head_frags<MT, SPLIT>(qkv + (long long)b * T * (3 * RG_D) + head * RG_HD, src_mask + (long long)b * T, T, g, t, y);
    store_frags<MT>(y, Y + (long long)b * T * RG_D + head * RG_HD + 2 * t, RG_D, T, g);
}
template <int MT, bool SPLIT>
__global__ void __launch_bounds__(128) ca_core_mma_kernel(const float* q3, int ldq,
                                                         const float* __restrict__ state,
                                                         long long state_clip_stride, long long state_cond_stride,
                                                         const float* qmask, long long qmask_cond_stride,
                                                         float* __restrict__ Y, int ldy, int T) {
    rg_pdl_launch();
    rg_pdl_wait();
    const int b = blockIdx.x, c = blockIdx.y, head = blockIdx.z * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float y[MT][4][4];
    ca_head_frags<MT, SPLIT>(q3 + (long long)b * T * ldq + c * RG_D + head * RG_HD + 2 * t, ldq,
                             state + (long long)b * state_clip_stride + (long long)c * state_cond_stride + (long long)head * RG_HD * RG_HD,
                             qmask ? qmask + (long long)c * qmask_cond_stride + (long long)b * T : nullptr, T, g, t, y);
    store_frags<MT>(y, Y + (long long)b * T * ldy + c * RG_D + head * RG_HD + 2 * t, ldy, T, g);
}

// ---- cores fused with the Stylization prologue of proj_out ---------------------------------------------------
// One CTA = one clip (x one condition), 16 warps = the 16 heads, so a token's whole 512-wide output row lives
// in this CTA's registers: LayerNorm statistics are combined across the warps through shared memory (per-warp
// mean / centred sum of squares over its 32 columns, merged with Chan's formula), then norm affine,
// *(1+scale)+shift and SiLU are applied to the fragments and the rows go straight out as the bf16 operand
// planes of the projection GEMM.  Saves the fp32 Y round trip and one kernel per attention block.
template <int MT>
__device__ __forceinline__ void styl_frags_store(float (&y)[MT][4][4], int T, int head, int g, int t,
                                                 float2 (*part)[RG_H], float2* stat, const RgStylParams& sp,
                                                 int clip, const RgRowOut& out, long long row_base, int col0) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float s = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) s += y[mt][nt][2 * h] + y[mt][nt][2 * h + 1];
            const float mw = quad_sum(s) * (1.0f / RG_HD);
            float q = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float d0 = y[mt][nt][2 * h] - mw, d1 = y[mt][nt][2 * h + 1] - mw;
                q += d0 * d0 + d1 * d1;
            }
            q = quad_sum(q);
            const int r = 16 * mt + g + 8 * h;
            if (t == 0 && r < T) part[r][head] = make_float2(mw, q);
        }
    __syncthreads();
    if ((int)threadIdx.x < T) {
        float mean = 0.f;
#pragma unroll
        for (int w = 0; w < RG_H; ++w) mean += part[threadIdx.x][w].x;
        mean *= (1.0f / RG_H);
        float m2 = 0.f;
#pragma unroll
        for (int w = 0; w < RG_H; ++w) {
            const float2 pw = part[threadIdx.x][w];
            const float d = pw.x - mean;
            m2 += pw.y + (float)RG_HD * d * d;
        }
        stat[threadIdx.x] = make_float2(mean, 1.0f / sqrtf(m2 * (1.0f / RG_D) + 1e-5f));
    }
    __syncthreads();
    const float* ss = sp.ss + (long long)clip * sp.ss_clip_stride;
    float2 ga[4], be[4], sc[4], sh[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int col = head * RG_HD + 8 * nt + 2 * t;
        ga[nt] = __ldg(reinterpret_cast<const float2*>(sp.gamma + col));
        be[nt] = __ldg(reinterpret_cast<const float2*>(sp.beta + col));
        sc[nt] = __ldg(reinterpret_cast<const float2*>(ss + col));
        sh[nt] = __ldg(reinterpret_cast<const float2*>(ss + RG_D + col));
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 16 * mt + g + 8 * h;
            if (r >= T) continue;
            const float2 st = stat[r];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float v0 = (y[mt][nt][2 * h] - st.x) * st.y, v1 = (y[mt][nt][2 * h + 1] - st.x) * st.y;
                v0 = rg_silu((v0 * ga[nt].x + be[nt].x) * (1.0f + sc[nt].x) + sh[nt].x);
                v1 = rg_silu((v1 * ga[nt].y + be[nt].y) * (1.0f + sc[nt].y) + sh[nt].y);
                const long long off = (row_base + r) * out.ld + col0 + head * RG_HD + 8 * nt + 2 * t;
                if (out.f32) {
                    *reinterpret_cast<float2*>(out.f32 + off) = make_float2(v0, v1);
                } else {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                    *reinterpret_cast<__nv_bfloat162*>(out.b16 + off) = hh;
                    if (out.lo_off)
                        *reinterpret_cast<__nv_bfloat162*>(out.b16 + off + out.lo_off) =
                            __floats2bfloat162_rn(v0 - __low2float(hh), v1 - __high2float(hh));
                }
            }
        }
}

template <int MT, bool SPLIT>
__global__ void __launch_bounds__(512) sa_styl_mma_kernel(const float* qkv, const float* src_mask, RgStylParams sp,
                                                         RgRowOut out, int T) {
    __shared__ float2 part[RG_MAX_T][RG_H];
    __shared__ float2 stat[RG_MAX_T];
    rg_pdl_launch();
    rg_pdl_wait();
    const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float y[MT][4][4];
    sa_head_frags<MT, SPLIT>(qkv + (long long)b * T * (3 * RG_D) + head * RG_HD, src_mask + (long long)b * T, T, g, t, y);
    styl_frags_store<MT>(y, T, head, g, t, part, stat, sp, b, out, (long long)b * T, 0);
}
template <int MT, bool SPLIT>
__global__ void __launch_bounds__(512) ca_styl_mma_kernel(const float* q3, int ldq, const float* __restrict__ state,
                                                         long long state_clip_stride, long long state_cond_stride,
                                                         const float* qmask, long long qmask_cond_stride, RgStyl3 sp3,
                                                         RgRowOut out, int T) {
    __shared__ float2 part[RG_MAX_T][RG_H];
    __shared__ float2 stat[RG_MAX_T];
    rg_pdl_launch();
    rg_pdl_wait();
    const int b = blockIdx.x, c = blockIdx.y, head = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float y[MT][4][4];
    ca_head_frags<MT, SPLIT>(q3 + (long long)b * T * ldq + c * RG_D + head * RG_HD + 2 * t, ldq,
                             state + (long long)b * state_clip_stride + (long long)c * state_cond_stride + (long long)head * RG_HD * RG_HD,
                             qmask ? qmask + (long long)c * qmask_cond_stride + (long long)b * T : nullptr, T, g, t, y);
    styl_frags_store<MT>(y, T, head, g, t, part, stat, sp3.p[c], b, out, (long long)b * T, c * RG_D);
}

// Stylization prologue of the three cross-attention blocks over Y[M,1536]: blockIdx.y = condition;
// 4 rows per warp with the parameters held in registers (see styl_rows_kernel)
__global__ void __launch_bounds__(128) styl_rows3_kernel(const float* y, int ldy, RgStyl3 sp3,
                                                        int rows_per_clip, RgRowOut out, int M) {
    rg_pdl_launch();
    rg_pdl_wait();
    constexpr int RPW = 4;
    const int row0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * RPW, c = blockIdx.y, lane = threadIdx.x & 31;
    if (row0 >= M) return;
    float4 v[RPW][4];
#pragma unroll
    for (int i = 0; i < RPW; ++i)
        if (row0 + i < M) load_row(y + (long long)(row0 + i) * ldy + c * RG_D, lane, v[i]);
    const RgStylParams& sp = sp3.p[c];
    RgStylRegs r;
    int clip = row0 / rows_per_clip;
    rg_styl_load(r, sp, clip, lane);
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int row = row0 + i;
        if (row >= M) break;
        const int cl = row / rows_per_clip;
        if (cl != clip && sp.ss_clip_stride != 0) { clip = cl; rg_styl_load(r, sp, clip, lane); }
        rg_styl_apply(v[i], r);
        rg_store_row_out(out, row, c * RG_D, lane, v[i]);
    }
}

__global__ void __launch_bounds__(512, 2) ca_attn_kernel(const float* __restrict__ q3, int ldq,
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
    float* slice = Ysm + warp * RG_HD;
    q_softmax_to_smem(qb, ldq, T, slice, lane);
    __syncwarp();
    for (int n = 0; n < T; ++n) {
        float y;
        row_times_A(slice + n * RG_D, A, y);
        if (qm) y = y + (1.0f - qm[n]) * RG_NEG_MASK;       // fp32 add: y - 1e6 rounds to a 1/16 grid
        __syncwarp();
        slice[n * RG_D + lane] = y;
    }
    __syncthreads();
    finish_rows(Ysm, T, b, sp3.p[c], 1, nullptr, 0, out, c * RG_D, warp, lane);
}

// state[b][set][h][d][l] = sum_n softmax_n(K[b,n,h,d]) * V[b,n,h,l]   (K6; efficient_attention.py:78-89)
// One CTA (256 threads) per (clip, head, set), ONE pass over K and V (they are read from HBM exactly once: the
// [rows, 8192] fp32 buffer of a chunk is ~1 GB, far beyond L2): tiles of 64 tokens staged in shared memory with
// coalesced 128-bit loads, column softmax kept online -- running maximum m[d] and sum s[d] per feature column; when a
// tile raises the maximum, the accumulator row A[d][:] and s[d] are rescaled by exp(m_old - m_new) -- then
// E = exp(K - m) once per element and E^T V (32x32 outputs, 4 per thread) accumulates from shared memory.
constexpr int KV_TILE = 64;
constexpr int KV_EP = RG_HD + 4;        // row pitch of the E tile: 16-byte aligned rows for the 128-bit reads of the product
__global__ void __launch_bounds__(256) kv_state_kernel(const float* __restrict__ kv, int ldkv,
                                                      int k_off, int v_off, int N,
                                                      float* __restrict__ state,
                                                      long long state_clip_stride,
                                                      int kv_set_stride, long long state_set_stride) {
    __shared__ __align__(16) float Es[KV_TILE][KV_EP];
    __shared__ __align__(16) float Vs[KV_TILE][RG_HD];
    __shared__ __align__(16) float part[4][RG_HD][RG_HD];      // per token-quarter partial products, summed at the end
    __shared__ float red[8][RG_HD];
    __shared__ float m_s[RG_HD], scale_s[RG_HD], sum_s[RG_HD];
    const int h = blockIdx.x, set = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, wg = tid >> 5;
    const float* rows = kv + (long long)b * N * ldkv + (long long)set * kv_set_stride;
    const float* kbase = rows + k_off + h * RG_HD;
    const float* vbase = rows + v_off + h * RG_HD;
    if (tid < RG_HD) { m_s[tid] = -INFINITY; sum_s[tid] = 0.f; }
    // The product E^T V is issue-bound when every thread owns 4 outputs (1 + 1 shared loads per 4 FMAs): here a
    // thread owns a 4 x 4 block A[d4..d4+3][l4..l4+3] (2 x 128-bit shared loads per 16 FMAs) for a QUARTER of the
    // tile's tokens; the four quarters are summed through shared memory once, at the end.
    const int grp = tid >> 6, t64 = tid & 63;
    const int d4 = (t64 >> 3) * 4, l4 = (t64 & 7) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int ld_row = tid >> 3, ld_c4 = (tid & 7) * 4;    // tile loader: 32 rows x 8 float4 per pass
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += KV_TILE) {
        // raw K and V of the tile -> shared memory
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = ld_row + half * 32, n = n0 + r;
            float4 kq = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), vq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
                kq = *reinterpret_cast<const float4*>(kbase + (long long)n * ldkv + ld_c4);
                vq = *reinterpret_cast<const float4*>(vbase + (long long)n * ldkv + ld_c4);
            }
            *reinterpret_cast<float4*>(&Es[r][ld_c4]) = kq;
            *reinterpret_cast<float4*>(&Vs[r][ld_c4]) = vq;
        }
        __syncthreads();
        // column maximum of the tile (lane = column, 8 warps stride over the tile's rows)
        float m = -INFINITY;
#pragma unroll
        for (int r = wg; r < KV_TILE; r += 8) m = fmaxf(m, Es[r][lane]);
        red[wg][lane] = m;
        __syncthreads();
        if (wg == 0) {
#pragma unroll
            for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i][lane]);
            const float mo = m_s[lane], mn = fmaxf(mo, m);
            scale_s[lane] = (mo == -INFINITY) ? 0.f : expf(mo - mn);       // first tile: nothing accumulated yet
            m_s[lane] = mn;
        }
        __syncthreads();
        // rescale what was accumulated under the old maximum, then exponentiate the tile in place
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sc = scale_s[d4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] *= sc;
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = ld_row + half * 32;
            float4 e = *reinterpret_cast<const float4*>(&Es[r][ld_c4]);
            e.x = expf(e.x - m_s[ld_c4 + 0]); e.y = expf(e.y - m_s[ld_c4 + 1]);       // exp(-inf) = 0 pads
            e.z = expf(e.z - m_s[ld_c4 + 2]); e.w = expf(e.w - m_s[ld_c4 + 3]);
            *reinterpret_cast<float4*>(&Es[r][ld_c4]) = e;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < KV_TILE / 4; ++rr) {
            const int r = grp * (KV_TILE / 4) + rr;
            const float4 e = *reinterpret_cast<const float4*>(&Es[r][d4]);
            const float4 v = *reinterpret_cast<const float4*>(&Vs[r][l4]);
            const float ee[4] = {e.x, e.y, e.z, e.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ee[i], vv[j], acc[i][j]);
        }
        if (tid < RG_HD) {
            float colsum = 0.f;
#pragma unroll 8
            for (int r = 0; r < KV_TILE; ++r) colsum += Es[r][tid];
            sum_s[tid] = sum_s[tid] * scale_s[tid] + colsum;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(&part[grp][d4 + i][l4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    const int d = tid >> 3, lo4 = (tid & 7) * 4;            // final pass: A[d][lo4..lo4+3], quarters in fixed order
    float4 a = *reinterpret_cast<const float4*>(&part[0][d][lo4]);
#pragma unroll
    for (int g = 1; g < 4; ++g) {
        const float4 p = *reinterpret_cast<const float4*>(&part[g][d][lo4]);
        a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    const float s = sum_s[d];
    float* o = state + (long long)b * state_clip_stride + (long long)set * state_set_stride +
               ((long long)h * RG_HD + d) * RG_HD + lo4;
    *reinterpret_cast<float4*>(o) = make_float4(a.x / s, a.y / s, a.z / s, a.w / s);
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

template <bool SPLIT>
static cudaError_t launch_sa_mma(const float* qkv, const float* src_mask, float* Y, int B, int T, cudaStream_t st) {
    const dim3 grid(B, RG_H / 4), block(128);
    switch ((T + 15) / 16) {
        case 1: return rg_launch_pdl(sa_core_mma_kernel<1, SPLIT>, grid, block, 0, st, qkv, src_mask, Y, T);
        case 2: return rg_launch_pdl(sa_core_mma_kernel<2, SPLIT>, grid, block, 0, st, qkv, src_mask, Y, T);
        case 3: return rg_launch_pdl(sa_core_mma_kernel<3, SPLIT>, grid, block, 0, st, qkv, src_mask, Y, T);
        default: return rg_launch_pdl(sa_core_mma_kernel<4, SPLIT>, grid, block, 0, st, qkv, src_mask, Y, T);
    }
}
template <bool SPLIT>
static cudaError_t launch_ca_mma(const float* q3, int ldq, const float* state, long long scs, long long sds,
                                 const float* qmask, long long qms, float* Y, int ldy, int B, int T, cudaStream_t st) {
    const dim3 grid(B, 3, RG_H / 4), block(128);
    switch ((T + 15) / 16) {
        case 1: return rg_launch_pdl(ca_core_mma_kernel<1, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, Y, ldy, T);
        case 2: return rg_launch_pdl(ca_core_mma_kernel<2, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, Y, ldy, T);
        case 3: return rg_launch_pdl(ca_core_mma_kernel<3, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, Y, ldy, T);
        default: return rg_launch_pdl(ca_core_mma_kernel<4, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, Y, ldy, T);
    }
}

template <bool SPLIT>
static cudaError_t launch_sa_styl(const float* qkv, const float* src_mask, RgStylParams sp, RgRowOut out, int B, int T,
                                  cudaStream_t st) {
    const dim3 grid(B), block(512);
    switch ((T + 15) / 16) {
        case 1: return rg_launch_pdl(sa_styl_mma_kernel<1, SPLIT>, grid, block, 0, st, qkv, src_mask, sp, out, T);
        case 2: return rg_launch_pdl(sa_styl_mma_kernel<2, SPLIT>, grid, block, 0, st, qkv, src_mask, sp, out, T);
        case 3: return rg_launch_pdl(sa_styl_mma_kernel<3, SPLIT>, grid, block, 0, st, qkv, src_mask, sp, out, T);
        default: return rg_launch_pdl(sa_styl_mma_kernel<4, SPLIT>, grid, block, 0, st, qkv, src_mask, sp, out, T);
    }
}
template <bool SPLIT>
static cudaError_t launch_ca_styl(const float* q3, int ldq, const float* state, long long scs, long long sds,
                                  const float* qmask, long long qms, RgStyl3 s3, RgRowOut out, int B, int T, cudaStream_t st) {
    const dim3 grid(B, 3), block(512);
    switch ((T + 15) / 16) {
        case 1: return rg_launch_pdl(ca_styl_mma_kernel<1, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, s3, out, T);
        case 2: return rg_launch_pdl(ca_styl_mma_kernel<2, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, s3, out, T);
        case 3: return rg_launch_pdl(ca_styl_mma_kernel<3, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, s3, out, T);
        default: return rg_launch_pdl(ca_styl_mma_kernel<4, SPLIT>, grid, block, 0, st, q3, ldq, state, scs, sds, qmask, qms, s3, out, T);
    }
}
// attention core + Stylization prologue in one kernel (tensor-core tiers): split = 3xTF32
cudaError_t rg_launch_sa_styl(const float* qkv, const float* src_mask, RgStylParams sp, RgRowOut out, int B, int T,
                              int split, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > RG_MAX_T || (out.ld % 2) || (out.lo_off % 2)) return cudaErrorInvalidValue;
    return split ? launch_sa_styl<true>(qkv, src_mask, sp, out, B, T, st) : launch_sa_styl<false>(qkv, src_mask, sp, out, B, T, st);
}
cudaError_t rg_launch_ca_styl(const float* q3, int ldq, const float* state, long long state_clip_stride,
                              long long state_cond_stride, const float* qmask, long long qmask_cond_stride,
                              const RgStylParams* sp3, RgRowOut out, int B, int T, int split, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > RG_MAX_T || (ldq % 2) || (out.ld % 2) || (out.lo_off % 2)) return cudaErrorInvalidValue;
    RgStyl3 s3;
    for (int c = 0; c < 3; ++c) s3.p[c] = sp3[c];
    return split ? launch_ca_styl<true>(q3, ldq, state, state_clip_stride, state_cond_stride, qmask, qmask_cond_stride, s3, out, B, T, st)
                 : launch_ca_styl<false>(q3, ldq, state, state_clip_stride, state_cond_stride, qmask, qmask_cond_stride, s3, out, B, T, st);
}

// mode 0: fp32 SIMT cores; 1: TF32 mma.sync; 2: 3xTF32 (hi/lo split) mma.sync
cudaError_t rg_launch_sa_core(const float* qkv, const float* src_mask, float* Y, int B, int T, int mode, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > RG_MAX_T) return cudaErrorInvalidValue;
    if (mode == 1) return launch_sa_mma<false>(qkv, src_mask, Y, B, T, st);
    if (mode == 2) return launch_sa_mma<true>(qkv, src_mask, Y, B, T, st);
    if (T <= 11 * TW) return rg_launch_pdl(sa_core_kernel<11>, dim3(B, RG_H), dim3(TW * 32), 0, st, qkv, src_mask, Y, T);
    return rg_launch_pdl(sa_core_kernel<16>, dim3(B, RG_H), dim3(TW * 32), 0, st, qkv, src_mask, Y, T);
}

cudaError_t rg_launch_ca_core(const float* q3, int ldq, const float* state, long long state_clip_stride,
                              long long state_cond_stride, const float* qmask, long long qmask_cond_stride,
                              float* Y, int ldy, int B, int T, int mode, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > RG_MAX_T || (ldq % 2) || (ldy % 2)) return cudaErrorInvalidValue;
    if (mode == 1)
        return launch_ca_mma<false>(q3, ldq, state, state_clip_stride, state_cond_stride, qmask, qmask_cond_stride, Y, ldy, B, T, st);
    if (mode == 2)
        return launch_ca_mma<true>(q3, ldq, state, state_clip_stride, state_cond_stride, qmask, qmask_cond_stride, Y, ldy, B, T, st);
    if (T <= 11 * TW)
        return rg_launch_pdl(ca_core_kernel<11>, dim3(B, 3, RG_H), dim3(TW * 32), 0, st, q3, ldq, state,
                             state_clip_stride, state_cond_stride, qmask, qmask_cond_stride, Y, ldy, T);
    return rg_launch_pdl(ca_core_kernel<16>, dim3(B, 3, RG_H), dim3(TW * 32), 0, st, q3, ldq, state, state_clip_stride,
                         state_cond_stride, qmask, qmask_cond_stride, Y, ldy, T);
}

cudaError_t rg_launch_styl_rows3(const float* y, int ldy, const RgStylParams* sp3, int rows_per_clip,
                                 RgRowOut out, int M, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    RgStyl3 s3;
    for (int c = 0; c < 3; ++c) s3.p[c] = sp3[c];
    return rg_launch_pdl(styl_rows3_kernel, dim3((M + 15) / 16, 3), dim3(128), 0, st, y, ldy, s3, rows_per_clip, out, M);
}

cudaError_t rg_launch_kv_state(const float* kv, int ldkv, int k_off, int v_off, int n_tokens,
                               float* state, long long state_clip_stride, int B, int n_sets,
                               int kv_set_stride, long long state_set_stride, cudaStream_t st) {
    if (B <= 0 || n_tokens <= 0) return cudaSuccess;
    // grid order: (head, set) fastest, clip slowest -- CTAs that are resident together then read neighbouring 128-byte
    // slices of the SAME rows of the [rows, sets * 1024] buffer (whole DRAM pages), not the same slice of different rows
    kv_state_kernel<<<dim3(RG_H, n_sets, B), 256, 0, st>>>(kv, ldkv, k_off, v_off, n_tokens, state,
                                                            state_clip_stride, kv_set_stride,
                                                            state_set_stride);
    return cudaGetLastError();
}
