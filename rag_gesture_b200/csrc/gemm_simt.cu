// fp32 SIMT GEMM  C[M,N] = A[M,K] * W[N,K]^T  with fused epilogues.
// The exact-arithmetic tier of the path (rel-L2 <= 1e-3 mode) and every one-off contraction
// (K7 timestep table, K6 state projections, condition pre-projection).  128x128x16 tiles,
// 256 threads, 8x8 outputs per thread held in registers, register-staged double buffering.
// Roofline: fp32 FMA pipe (not tensor cores); the tcgen05 kernel in gemm_tc.cu is the fast path.
#include "rg_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <int EPI>
__global__ void __launch_bounds__(256, 2) gemm_tn_f32_kernel(RgGemm p) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Ws[2][BK][BN + PAD];

    const int g = blockIdx.z;
    const float* __restrict__ A = p.A + (long long)g * p.a_g;
    const float* __restrict__ W = p.W + (long long)g * p.w_g;
    const float* __restrict__ bias = p.bias ? p.bias + (long long)g * p.b_g : nullptr;
    float* C = p.C + (long long)g * p.c_g;      // may alias R (in-place residual add)
    const float* R = p.R ? p.R + (long long)g * p.r_g : nullptr;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    // global -> register staging: each thread moves two float4 of A and two of W per k-tile
    const int lrow0 = tid >> 2, lkq = tid & 3;       // rows lrow0 and lrow0 + 64
    float4 ra[2], rw[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = lrow0 + i * 64;
            const int am = m0 + row, wn = n0 + row;
            ra[i] = (am < p.M) ? __ldg(reinterpret_cast<const float4*>(A + (long long)am * p.lda + k0 + lkq * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            rw[i] = (wn < p.N) ? __ldg(reinterpret_cast<const float4*>(W + (long long)wn * p.ldw + k0 + lkq * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = lrow0 + i * 64;
            As[buf][lkq * 4 + 0][row] = ra[i].x; As[buf][lkq * 4 + 1][row] = ra[i].y;
            As[buf][lkq * 4 + 2][row] = ra[i].z; As[buf][lkq * 4 + 3][row] = ra[i].w;
            Ws[buf][lkq * 4 + 0][row] = rw[i].x; Ws[buf][lkq * 4 + 1][row] = rw[i].y;
            Ws[buf][lkq * 4 + 2][row] = rw[i].z; Ws[buf][lkq * 4 + 3][row] = rw[i].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = p.K / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= p.M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = n0 + h * 64 + tx * 4;
            if (col >= p.N) continue;
            float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            if (bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (EPI == RG_EPI_BIAS_RESIDUAL) {
                const float4 r = *reinterpret_cast<const float4*>(R + (long long)row * p.ldr + col);
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            } else if (EPI == RG_EPI_BIAS_GELU) {
                v.x = rg_gelu_erf(v.x); v.y = rg_gelu_erf(v.y); v.z = rg_gelu_erf(v.z); v.w = rg_gelu_erf(v.w);
            } else if (EPI == RG_EPI_BIAS_SILU) {
                v.x = rg_silu(v.x); v.y = rg_silu(v.y); v.z = rg_silu(v.z); v.w = rg_silu(v.w);
            } else if (EPI == RG_EPI_BIAS_POS) {
                const float4 pp = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(row % p.pos_T) * p.N + col));
                v.x += pp.x; v.y += pp.y; v.z += pp.z; v.w += pp.w;
            }
            *reinterpret_cast<float4*>(C + (long long)row * p.ldc + col) = v;
        }
    }
}

}  // namespace

cudaError_t rg_launch_gemm_f32(const RgGemm& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return cudaSuccess;
    if (g.K % BK != 0 || g.N % 4 != 0 || g.lda % 4 != 0 || g.ldw % 4 != 0 || g.ldc % 4 != 0 ||
        (g.R && g.ldr % 4 != 0))
        return cudaErrorInvalidValue;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.groups > 0 ? g.groups : 1);
    switch (g.epi) {
        case RG_EPI_BIAS: gemm_tn_f32_kernel<RG_EPI_BIAS><<<grid, 256, 0, st>>>(g); break;
        case RG_EPI_BIAS_RESIDUAL: gemm_tn_f32_kernel<RG_EPI_BIAS_RESIDUAL><<<grid, 256, 0, st>>>(g); break;
        case RG_EPI_BIAS_GELU: gemm_tn_f32_kernel<RG_EPI_BIAS_GELU><<<grid, 256, 0, st>>>(g); break;
        case RG_EPI_BIAS_POS: gemm_tn_f32_kernel<RG_EPI_BIAS_POS><<<grid, 256, 0, st>>>(g); break;
        case RG_EPI_BIAS_SILU: gemm_tn_f32_kernel<RG_EPI_BIAS_SILU><<<grid, 256, 0, st>>>(g); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
