// Row-wise and elementwise kernels of the path: LayerNorm, the StylizationBlock prologue
// (LN * (1+scale) + shift -> SiLU, stylization_block.py:38-39), the DDIM / reverse-DDIM update
// (gaussian_diffusion.py:693-697, 983-1001, 1032-1038), the in_seq blend (:934-947) and the
// closed form of the insertion-guidance gradient step (:1351-1378).
// All are HBM-bound: one warp per 512-wide row, 128-bit coalesced accesses, no shared memory.
// The sampler arithmetic uses explicit round-to-nearest intrinsics so that nvcc cannot contract
// a*b+c into an FMA: those kernels are bit-exact against the reference's fp32 op sequence.
#include "rg_common.cuh"
#include "rg_rows.cuh"

namespace {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps

__global__ void __launch_bounds__(256) ln_rows_kernel(const float* x, int ldx,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta,
                                                     RgRowOut out, int M) {
    rg_pdl_launch();
    rg_pdl_wait();
    const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float4 v[4];
    load_row(x + (long long)row * ldx, lane, v);
    rg_ln_normalize(v);
    if (gamma) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
            v[j] = affine4(v[j], g, b);
        }
    }
    rg_store_row_out(out, row, 0, lane, v);
}

// Each warp takes RPW consecutive rows and keeps the block's 4 x 2 KB of parameters (gamma, beta, scale,
// shift) in registers across them: re-reading them per row made this kernel move 5x its data.
constexpr int RPW = 4;
__global__ void __launch_bounds__(128) styl_rows_kernel(const float* y, int ldy,
                                                       RgStylParams sp, int rows_per_clip,
                                                       RgRowOut out, int M) {
    rg_pdl_launch();
    rg_pdl_wait();
    const int row0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * RPW;
    const int lane = threadIdx.x & 31;
    if (row0 >= M) return;
    float4 v[RPW][4];
#pragma unroll
    for (int i = 0; i < RPW; ++i)
        if (row0 + i < M) load_row(y + (long long)(row0 + i) * ldy, lane, v[i]);
    RgStylRegs r;
    int clip = row0 / rows_per_clip;
    rg_styl_load(r, sp, clip, lane);
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int row = row0 + i;
        if (row >= M) break;
        const int c = row / rows_per_clip;
        if (c != clip && sp.ss_clip_stride != 0) { clip = c; rg_styl_load(r, sp, clip, lane); }
        rg_styl_apply(v[i], r);
        rg_store_row_out(out, row, 0, lane, v[i]);
    }
}

__global__ void silu_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = rg_silu(x[i]);
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table,
                                                         const long long* __restrict__ idx,
                                                         float* __restrict__ out, long long n_rows,
                                                         int n_table) {
    const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    long long id = idx[row];
    id = id < 0 ? 0 : (id >= n_table ? n_table - 1 : id);
    float4 v[4];
    load_row(table + id * RG_D, lane, v);
    store_row(out + row * RG_D, lane, v);
}

// Wf[n,k] = W[n,k]*gamma[k];  bf[n] = b[n] + sum_k W[n,k]*beta[k]   (LayerNorm affine folded
// into the Linear that consumes it; one warp per output feature)
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W,
                                                     const float* __restrict__ b,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta,
                                                     float* __restrict__ Wf, float* __restrict__ bf,
                                                     int N, int K) {
    const int n = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = W[(long long)n * K + k];
        Wf[(long long)n * K + k] = w * gamma[k];
        acc = fmaf(w, beta[k], acc);
    }
    acc = rg_warp_sum(acc);
    if (lane == 0) bf[n] = (b ? b[n] : 0.f) + acc;
}

// x' = x0*c_a + c_b*eps,  eps = (c_recip*x - x0)/c_recipm1      -- no FMA contraction
__global__ void __launch_bounds__(256) ddim_update_kernel(const float4* x, const float4* x0,
                                                         float4* out, long long n4,
                                                         float c_recip, float c_recipm1, float c_a,
                                                         float c_b) {
    rg_pdl_launch();
    rg_pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        const float4 a = x[i], p = x0[i];
        float4 o;
#define RG_UPD(f)                                                                           \
    {                                                                                       \
        const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c_recip, a.f), p.f), c_recipm1);     \
        o.f = __fadd_rn(__fmul_rn(p.f, c_a), __fmul_rn(c_b, eps));                           \
    }
        RG_UPD(x) RG_UPD(y) RG_UPD(z) RG_UPD(w)
#undef RG_UPD
        out[i] = o;
    }
}

// rows where in_seq has any non-zero entry are replaced by q_sample(in_seq, t, noise)
__global__ void __launch_bounds__(256) blend_kernel(const float* x,
                                                   const float* in_seq,
                                                   const float* noise,
                                                   float* out, long long rows,
                                                   float s_ab, float s_1mab) {
    rg_pdl_launch();
    rg_pdl_wait();
    const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 s[4];
    load_row(in_seq + row * RG_D, lane, s);
    bool nz = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) nz |= (s[j].x != 0.f) | (s[j].y != 0.f) | (s[j].z != 0.f) | (s[j].w != 0.f);
    float4 v[4];
    if (__any_sync(0xffffffffu, nz)) {
        load_row(noise + row * RG_D, lane, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j].x = __fadd_rn(__fmul_rn(s_ab, s[j].x), __fmul_rn(s_1mab, v[j].x));
            v[j].y = __fadd_rn(__fmul_rn(s_ab, s[j].y), __fmul_rn(s_1mab, v[j].y));
            v[j].z = __fadd_rn(__fmul_rn(s_ab, s[j].z), __fmul_rn(s_1mab, v[j].z));
            v[j].w = __fadd_rn(__fmul_rn(s_ab, s[j].w), __fmul_rn(s_1mab, v[j].w));
        }
    } else {
        load_row(x + row * RG_D, lane, v);
    }
    store_row(out + row * RG_D, lane, v);
}

// `iters` steps of x <- x - lr * d/dx mse(x*m, in_seq) = x - (2 lr / numel) * m * (x*m - in_seq)
__global__ void __launch_bounds__(256) guidance_kernel(float* __restrict__ x,
                                                      const float* in_seq,
                                                      long long rows, int iters, float c) {
    const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 s[4];
    load_row(in_seq + row * RG_D, lane, s);
    bool nz = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) nz |= (s[j].x != 0.f) | (s[j].y != 0.f) | (s[j].z != 0.f) | (s[j].w != 0.f);
    if (!__any_sync(0xffffffffu, nz)) return;
    float4 v[4];
    load_row(x + row * RG_D, lane, v);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j].x -= c * (v[j].x - s[j].x); v[j].y -= c * (v[j].y - s[j].y);
            v[j].z -= c * (v[j].z - s[j].z); v[j].w -= c * (v[j].w - s[j].w);
        }
    }
    store_row(x + row * RG_D, lane, v);
}

// pos[t] = learned_global[t] + (sine[i] for the i-th chunk of each body part, 0 on separators)
__global__ void pos_table_kernel(const float* __restrict__ seq_pe, const float* __restrict__ glob_pe,
                                 float* __restrict__ pos, int T, int n_chunks) {
    const int t = blockIdx.x;
    const int i = t % (n_chunks + 1);
    for (int c = threadIdx.x; c < RG_D; c += blockDim.x) {
        const float s = (i < n_chunks) ? seq_pe[i * RG_D + c] : 0.f;
        pos[t * RG_D + c] = s + glob_pe[t * RG_D + c];
    }
}

// out[c][r] = in[r][c] for an n x n fp32 matrix (weight folding at model creation)
__global__ void transpose_sq_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
    __shared__ float t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) t[i][threadIdx.x] = in[(long long)(by + i) * n + bx + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) out[(long long)(bx + i) * n + by + threadIdx.x] = t[threadIdx.x][i];
}

// out[r][c] = a[r][c] + a[r][c + n] + a[r][c + 2n]   (sum of the three column blocks of ca_mix.weight)
__global__ void sum3_blocks_kernel(const float* __restrict__ a, int lda, float* __restrict__ out, int ldo, int n) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < n; c += blockDim.x)
        out[(long long)r * ldo + c] = a[(long long)r * lda + c] + a[(long long)r * lda + n + c] + a[(long long)r * lda + 2 * n + c];
}

inline int row_blocks(long long rows) { return (int)((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK); }

}  // namespace

cudaError_t rg_launch_ln_rows(const float* x, int ldx, const float* gamma, const float* beta,
                              RgRowOut out, int M, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    return rg_launch_pdl(ln_rows_kernel, dim3(row_blocks(M)), dim3(256), 0, st, x, ldx, gamma, beta, out, M);
}
cudaError_t rg_launch_styl_rows(const float* y, int ldy, RgStylParams sp, int rows_per_clip,
                                RgRowOut out, int M, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    return rg_launch_pdl(styl_rows_kernel, dim3((M + 4 * RPW - 1) / (4 * RPW)), dim3(128), 0, st, y, ldy, sp, rows_per_clip, out, M);
}
cudaError_t rg_launch_silu(const float* x, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    silu_kernel<<<blocks, 256, 0, st>>>(x, out, n);
    return cudaGetLastError();
}
cudaError_t rg_launch_gather_rows(const float* table, const long long* idx, float* out,
                                  long long n_rows, int n_table, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    gather_rows_kernel<<<row_blocks(n_rows), 256, 0, st>>>(table, idx, out, n_rows, n_table);
    return cudaGetLastError();
}
cudaError_t rg_launch_fold_ln(const float* W, const float* b, const float* gamma, const float* beta,
                              float* Wf, float* bf, int N, int K, cudaStream_t st) {
    fold_ln_kernel<<<row_blocks(N), 256, 0, st>>>(W, b, gamma, beta, Wf, bf, N, K);
    return cudaGetLastError();
}
cudaError_t rg_launch_ddim_update(const float* x, const float* x0, float* out, long long n,
                                  float c_recip, float c_recipm1, float c_a, float c_b,
                                  cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (n % 4) return cudaErrorInvalidValue;
    const long long n4 = n / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    return rg_launch_pdl(ddim_update_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(x),
                         reinterpret_cast<const float4*>(x0), reinterpret_cast<float4*>(out), n4, c_recip,
                         c_recipm1, c_a, c_b);
}
cudaError_t rg_launch_blend(const float* x, const float* in_seq, const float* noise, float* out,
                            long long rows, float s_ab, float s_1mab, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    return rg_launch_pdl(blend_kernel, dim3(row_blocks(rows)), dim3(256), 0, st, x, in_seq, noise, out, rows, s_ab, s_1mab);
}
// 2-branch mixing of forward_test (raggesture.py:1087-1111): rows [0,B*T) of `out2` = text branch, rows [B*T,2*B*T) =
// "none" branch; per clip four coefficients (both, text, retr, none), per token row the joint scale.  Products and sums
// in the reference's order, each rounded to fp32 (no FMA contraction).
__global__ void __launch_bounds__(256) mix_branches_kernel(const float* out2, const float* __restrict__ coef,
                                                          const float* __restrict__ joint_scale, float* out,
                                                          long long rows, int T) {
    const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long clip = row / T;
    const float js = joint_scale[row - clip * T], ijs = __fdiv_rn(1.0f, js);
    const float both = coef[clip * 4], text = coef[clip * 4 + 1], retr = coef[clip * 4 + 2], none = coef[clip * 4 + 3];
    float4 a[4], b[4];
    load_row(out2 + row * RG_D, lane, a);
    load_row(out2 + (rows + row) * RG_D, lane, b);
    auto mix = [&](float xt, float xn) {
        float r = __fmul_rn(__fmul_rn(xt, both), js);
        r = __fadd_rn(r, __fmul_rn(__fmul_rn(xt, text), js));
        r = __fadd_rn(r, __fmul_rn(__fmul_rn(xn, retr), ijs));
        return __fadd_rn(r, __fmul_rn(__fmul_rn(xn, none), ijs));
    };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j].x = mix(a[j].x, b[j].x); a[j].y = mix(a[j].y, b[j].y);
        a[j].z = mix(a[j].z, b[j].z); a[j].w = mix(a[j].w, b[j].w);
    }
    store_row(out + row * RG_D, lane, a);
}

cudaError_t rg_launch_mix_branches(const float* out2, const float* coef, const float* joint_scale, float* out,
                                   long long rows, int T, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    mix_branches_kernel<<<row_blocks(rows), 256, 0, st>>>(out2, coef, joint_scale, out, rows, T);
    return cudaGetLastError();
}
cudaError_t rg_launch_guidance(float* x, const float* in_seq, long long rows, int iters,
                               float lr_2_over_n, cudaStream_t st) {
    if (rows <= 0 || iters <= 0) return cudaSuccess;
    guidance_kernel<<<row_blocks(rows), 256, 0, st>>>(x, in_seq, rows, iters, lr_2_over_n);
    return cudaGetLastError();
}
cudaError_t rg_launch_pos_table(const float* seq_pe, const float* glob_pe, float* pos, int T,
                                int n_chunks, cudaStream_t st) {
    pos_table_kernel<<<T, 128, 0, st>>>(seq_pe, glob_pe, pos, T, n_chunks);
    return cudaGetLastError();
}

cudaError_t rg_launch_transpose_sq(const float* in, float* out, int n, cudaStream_t st) {
    if (n % 32) return cudaErrorInvalidValue;
    transpose_sq_kernel<<<dim3(n / 32, n / 32), dim3(32, 8), 0, st>>>(in, out, n);
    return cudaGetLastError();
}
cudaError_t rg_launch_sum3_blocks(const float* a, int lda, float* out, int ldo, int n, int rows, cudaStream_t st) {
    sum3_blocks_kernel<<<rows, 256, 0, st>>>(a, lda, out, ldo, n);
    return cudaGetLastError();
}
