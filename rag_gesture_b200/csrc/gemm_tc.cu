// tcgen05 GEMM  C[M,N] = A[M,K] * W[N,K]^T  for sm_100a: bf16 operands staged by TMA
// (cp.async.bulk.tensor, SWIZZLE_128B) into a 3-stage shared-memory ring, one elected thread issuing
// tcgen05.mma.cta_group::1.kind::f16 (UMMA 128xBNx16) into a TMEM accumulator (fp32, 128 lanes x BN
// columns), epilogue warps reading it back with tcgen05.ld and fusing bias / residual / GELU / SiLU /
// positional add, writing fp32 and/or bf16 (optionally hi+lo split) outputs.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (TMEM lane group = warp % 4, two warps per group split the columns).  Two CTAs are resident per SM
// (96 KB smem, 128 TMEM columns each) so one tile's epilogue overlaps another tile's main loop.
//
// "split" mode (RG_PREC_BF16X3): operands are stored as [hi | lo] bf16 planes (x = hi + lo + O(2^-17 x));
// the K loop runs three passes  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  into the same fp32 accumulator:
// fp32-class accuracy (~2^-16 per product) on the bf16 tensor pipe, without a second kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "rg_common.cuh"
#include "rg_gemm_tc.h"
#include "rg_tcgen05.cuh"

namespace {

using namespace rg_tc;

constexpr int BM = 128, BK = 64;

// STAGES = 3: two CTAs per SM (grids above one wave); STAGES = 6: one CTA per SM with the whole K = 512
// reduction in flight (small grids, where a tile's latency, not throughput, is what is measured).
template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(320, STAGES <= 3 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, RgGemmTc p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    rg_pdl_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* tr = p.trace ? p.trace + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 10 : nullptr;
#define RG_STAMP(slot) do { if (tr) tr[slot] = clock64(); } while (0)
    if (tr && threadIdx.x == 0) {
        unsigned long long gt; unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        tr[0] = clock64(); tr[8] = (long long)gt; tr[9] = sm;
    }
    const int g = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = p.K / BK;
    const int total_kb = p.split ? 3 * nkb : nkb;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
            mbar_init(smem_u32(&tmem_full_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) RG_STAMP(1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            const int a_k0 = g * p.a_goff, w_n0 = g * p.w_goff + n0;
            auto coords = [&](int j, int& ka, int& kw) {
                const int pass = j / nkb, kb = j - pass * nkb;          // pass 0: hi*hi, 1: lo*hi, 2: hi*lo
                ka = a_k0 + (pass == 1 ? p.a_lo_off : 0) + kb * BK;
                kw = (pass == 2 ? p.w_lo_off : 0) + kb * BK;
            };
            // The weight operand does not depend on the previous kernel: its first STAGES tiles are
            // requested BEFORE griddepcontrol.wait, so under PDL they stream in while the producer of
            // our A operand is still finishing.  The activation tiles follow once it has completed.
            const int pre = total_kb < STAGES ? total_kb : STAGES;
            for (int j = 0; j < pre; ++j) {
                int ka, kw;
                coords(j, ka, kw);
                mbar_expect_tx(smem_u32(&full_bar[j]), STAGE_BYTES);
                tma_load_2d(smem_u32(smem + j * STAGE_BYTES) + A_BYTES, &tmW, smem_u32(&full_bar[j]), kw, w_n0);
            }
            rg_pdl_wait();
            RG_STAMP(2);
            for (int j = 0; j < pre; ++j) {
                int ka, kw;
                coords(j, ka, kw);
                tma_load_2d(smem_u32(smem + j * STAGE_BYTES), &tmA, smem_u32(&full_bar[j]), ka, m0);
            }
            for (int j = pre; j < total_kb; ++j) {
                const int s = j % STAGES, ph = (j / STAGES) & 1;
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                int ka, kw;
                coords(j, ka, kw);
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
                mbar_expect_tx(smem_u32(&full_bar[s]), STAGE_BYTES);
                tma_load_2d(sa, &tmA, smem_u32(&full_bar[s]), ka, m0);
                tma_load_2d(sb, &tmW, smem_u32(&full_bar[s]), kw, w_n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(BM, BN);
        for (int j = 0; j < total_kb; ++j) {
            const int s = j % STAGES, ph = (j / STAGES) & 1;
            mbar_wait(smem_u32(&full_bar[s]), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                if (j == 0) RG_STAMP(3);
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
                const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)      // +32 B per UMMA_K inside the 128 B swizzle row
                    umma_bf16(tmem_base, da + (k * UMMA_K * 2 >> 4), db + (k * UMMA_K * 2 >> 4), idesc, (j | k) != 0);
                umma_commit(smem_u32(&empty_bar[s]));                   // frees the smem slot when the MMAs retire
                if (j == total_kb - 1) { umma_commit(smem_u32(&tmem_full_bar)); RG_STAMP(4); }
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue (8 warps): TMEM -> registers -> smem (transpose) -> coalesced global =====
        // tcgen05.ld gives thread `lane` one accumulator ROW (32 columns per load); storing that
        // straight to global would touch 32 different cache lines per instruction.  The operand
        // ring is idle once tmem_full fires, so every warp parks its 32-row x BN/2-column block there
        // (row pitch BN/2+4 floats: conflict-free both ways) and then walks it row by row with lanes
        // along N: bias / residual / pos loads and all stores are contiguous, 8 rows in flight.
        // Two warps share each TMEM lane group (a warp may only read lanes 32*(warp%4)..+31) and split
        // the columns, which halves the serial latency of this phase.
        rg_pdl_wait();          // residual reads and all stores below touch buffers of the previous kernel
        mbar_wait(smem_u32(&tmem_full_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (threadIdx.x == 64) RG_STAMP(5);
        const int lg = warp & 3;                                        // TMEM lane group of this warp
        const int half = (warp - 2) >> 2;                               // which half of the tile's columns
        constexpr int HC = BN / 2, PITCH = HC + 4;
        float* stage = reinterpret_cast<float*>(smem) + ((warp - 2) * 32) * PITCH;
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + half * HC + c * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* dst = stage + lane * PITCH + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                  __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
        __syncwarp();
        if (threadIdx.x == 64) RG_STAMP(6);
        // Row pass: 16 lanes x float4 cover the warp's 64 columns, so one instruction handles TWO rows and
        // every global access is a 256-byte (fp32) / 128-byte (bf16) contiguous run.  The epilogue kind is a
        // template parameter: this loop is ALU-bound (8 warps finish a 128x128 tile), branches and 64-bit
        // address arithmetic per element were most of its instructions.
        static_assert(HC == 64, "row pass assumes 64 columns per epilogue warp");
        const int sub = lane >> 4, cc = (lane & 15) * 4;
        const int tcol = half * HC;                                     // first column of this warp inside the tile
        const int cbase = g * p.c_goff + n0 + tcol + cc;                // ... in C
        const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + g * p.b_goff + n0 + tcol + cc))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        const int row_first = m0 + lg * 32 + sub;
        const float* rsrc = nullptr;                                    // residual / positional rows
        long long rstep = 0;
        if (EPI == RG_EPI_BIAS_RESIDUAL) {
            rsrc = p.R + (long long)row_first * p.ldr + (p.r_grouped ? g * p.c_goff : 0) + n0 + tcol + cc;
            rstep = 2ll * p.ldr;
        }
        float* c32 = p.C32 ? p.C32 + (long long)row_first * p.ldc32 + cbase : nullptr;
        __nv_bfloat16* c16 = p.C16_ ? reinterpret_cast<__nv_bfloat16*>(p.C16_) + (long long)row_first * p.ldc16 + cbase : nullptr;
        const float* srow = stage + sub * PITCH + cc;
        constexpr int RB = 4;                       // row pairs per batch: all loads of a batch are issued first
#pragma unroll 1
        for (int s0 = 0; s0 < 16; s0 += RB) {
            float4 f[RB], rr[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const int row = row_first + 2 * (s0 + i);
                f[i] = *reinterpret_cast<const float4*>(srow + 2 * (s0 + i) * PITCH);
                rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < p.M) {
                    if (EPI == RG_EPI_BIAS_RESIDUAL)
                        rr[i] = *reinterpret_cast<const float4*>(rsrc + (s0 + i) * rstep);
                    else if (EPI == RG_EPI_BIAS_POS)
                        rr[i] = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(row % p.pos_T) * p.N + n0 + tcol + cc));
                }
            }
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const int row = row_first + 2 * (s0 + i);
                if (row >= p.M) break;
                float4 v = f[i];
                v.x += bv.x + rr[i].x; v.y += bv.y + rr[i].y; v.z += bv.z + rr[i].z; v.w += bv.w + rr[i].w;
                if (EPI == RG_EPI_BIAS_GELU) {
                    v.x = rg_gelu_fast(v.x); v.y = rg_gelu_fast(v.y); v.z = rg_gelu_fast(v.z); v.w = rg_gelu_fast(v.w);
                } else if (EPI == RG_EPI_BIAS_SILU) {
                    v.x = rg_silu(v.x); v.y = rg_silu(v.y); v.z = rg_silu(v.z); v.w = rg_silu(v.w);
                }
                if (c32) *reinterpret_cast<float4*>(c32 + 2ll * (s0 + i) * p.ldc32) = v;
                if (c16) {
                    __nv_bfloat16* o = c16 + 2ll * (s0 + i) * p.ldc16;
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&h01); pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                    *reinterpret_cast<uint2*>(o) = pk;
                    if (p.c16_lo_off) {
                        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
                        const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
                        pk.x = *reinterpret_cast<const uint32_t*>(&l01); pk.y = *reinterpret_cast<const uint32_t*>(&l23);
                        *reinterpret_cast<uint2*>(o + p.c16_lo_off) = pk;
                    }
                }
            }
        }
        if (threadIdx.x == 64) RG_STAMP(7);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN));
    }
#undef RG_STAMP
}

// fp32 -> bf16 hi (and lo = bf16(x - hi)) planes; row-major, lo plane at column offset lo_off (0 = none)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* x, int ldx,
                                                        __nv_bfloat16* __restrict__ out, int ldo, int lo_off,
                                                        long long rows, int cols) {
    rg_pdl_launch();
    rg_pdl_wait();
    const long long n4 = rows * (cols / 4);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        const long long r = i / (cols / 4);
        const int c = (int)(i - r * (cols / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { h[k] = __float2bfloat16_rn(f[k]); l[k] = __float2bfloat16_rn(f[k] - __bfloat162float(h[k])); }
        *reinterpret_cast<uint2*>(out + r * ldo + c) = *reinterpret_cast<uint2*>(h);
        if (lo_off) *reinterpret_cast<uint2*>(out + r * ldo + lo_off + c) = *reinterpret_cast<uint2*>(l);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

// row-major bf16 [rows, cols] with pitch ld (elements): box = 64 (K) x box_rows, 128-byte swizzle
cudaError_t rg_make_tensor_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld,
                               int box_rows) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int STAGES, int EPI>
static cudaError_t launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, dim3 grid, cudaStream_t st) {
    constexpr int BN = 128;
    constexpr size_t smem = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (p.no_pdl) {     // weight tiles are prefetched before griddepcontrol.wait: only valid for constant W
        gemm_tc_kernel<BN, STAGES, EPI><<<grid, 320, smem, st>>>(tmA, tmW, p);
        return cudaGetLastError();
    }
    return rg_launch_pdl(gemm_tc_kernel<BN, STAGES, EPI>, grid, dim3(320), smem, st, tmA, tmW, p);
}
template <int EPI>
static cudaError_t launch_tc_epi(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, dim3 grid, cudaStream_t st) {
    // one CTA per SM with the whole K = 512 reduction in flight for grids below one wave, else two per SM
    if ((long long)grid.x * grid.y * grid.z <= 148) return launch_tc<6, EPI>(tmA, tmW, p, grid, st);
    return launch_tc<3, EPI>(tmA, tmW, p, grid, st);
}

int rg_gemm_kernel_mode = 0;
int rg_gemm2_min_rows = 16384;     // measured on B200: below ~11k rows the 128x128 kernel's shorter epilogue tail wins
int rg_gemm2_persist_tiles = 296;

int rg_pair128_min_rows = 1 << 30;   // set from measurements (capi.cu / rg_set_gemm_kernel)

cudaError_t rg_launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, cudaStream_t st) {
    if (rg_gemm_pair128_eligible(p) && (rg_gemm_kernel_mode == 3 || (rg_gemm_kernel_mode == 0 && p.M >= rg_pair128_min_rows)))
        return rg_launch_gemm_pair128(tmA, p, st);
    if (rg_gemm_kernel_mode != 1 && rg_gemm_kernel_mode != 3 && rg_gemm2_eligible(p) && (rg_gemm_kernel_mode == 2 || (p.M >= rg_gemm2_min_rows && !p.trace)))
        return rg_launch_gemm2_tc(tmA, tmW, p, st);
    return rg_launch_gemm1_tc(tmA, tmW, p, st);
}

cudaError_t rg_launch_gemm1_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return cudaSuccess;
    constexpr int BN = 128;
    if (p.K % BK || p.N % BN || (p.C32 && p.ldc32 % 4) || (p.C16_ && p.ldc16 % 8) || (p.R && p.ldr % 4) ||
        (p.C16_ && p.c16_lo_off % 4))
        return cudaErrorInvalidValue;
    if (p.epi == RG_EPI_BIAS_RESIDUAL && !p.R) return cudaErrorInvalidValue;
    if (p.epi == RG_EPI_BIAS_POS && (!p.pos || p.pos_T <= 0)) return cudaErrorInvalidValue;
    const dim3 grid(p.N / BN, (p.M + BM - 1) / BM, p.groups > 0 ? p.groups : 1);
    switch (p.epi) {
        case RG_EPI_BIAS: return launch_tc_epi<RG_EPI_BIAS>(tmA, tmW, p, grid, st);
        case RG_EPI_BIAS_RESIDUAL: return launch_tc_epi<RG_EPI_BIAS_RESIDUAL>(tmA, tmW, p, grid, st);
        case RG_EPI_BIAS_GELU: return launch_tc_epi<RG_EPI_BIAS_GELU>(tmA, tmW, p, grid, st);
        case RG_EPI_BIAS_POS: return launch_tc_epi<RG_EPI_BIAS_POS>(tmA, tmW, p, grid, st);
        case RG_EPI_BIAS_SILU: return launch_tc_epi<RG_EPI_BIAS_SILU>(tmA, tmW, p, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t rg_launch_split_bf16(const float* x, int ldx, void* out, int ldo, int lo_off, long long rows, int cols,
                                 cudaStream_t st, bool pdl) {
    if (rows <= 0) return cudaSuccess;
    if (cols % 4 || ldx % 4 || ldo % 4 || lo_off % 4) return cudaErrorInvalidValue;
    const long long n4 = rows * (cols / 4);
    const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    if (!pdl) {
        split_bf16_kernel<<<blocks, 256, 0, st>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(out), ldo, lo_off, rows, cols);
        return cudaGetLastError();
    }
    return rg_launch_pdl(split_bf16_kernel, dim3(blocks), dim3(256), 0, st, x, ldx, reinterpret_cast<__nv_bfloat16*>(out),
                         ldo, lo_off, rows, cols);
}
