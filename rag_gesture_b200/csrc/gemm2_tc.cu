// Persistent 2-CTA tcgen05 GEMM  C[M,N] = A[M,K] * W[N,K]^T (+ epilogue)  for the denoiser's large-M launches
// (the fused guided+inversion level: M = (B + E) * 43 = 6880 rows).
//
// Why a second kernel: with 128x128 tiles (gemm_tc.cu) every tile pulls 64 flop per operand byte out of L2, and at
// ~42 B/clk/SM of L2->SM bandwidth (6.3 KB/clk chip-wide) the K = 512 GEMMs of a denoiser layer are bound by that,
// not by the tensor pipe.  Here a CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2) owns a 256 x 256
// output tile: each CTA stages its own 128 rows of A and HALF of the 256 weight rows per K-block, the pair's
// UMMA 256x256x16 reads both halves -> 128 flop per operand byte (half the L2 traffic per flop), and the weight
// tile is fetched once per 256 rows instead of once per 128.
//
//   Two instantiations of one kernel (template ACCS = accumulator stages):
//   * ACCS = 2, persistent: grid = 2 * min(tiles, resident pairs), one CTA per SM; each pair walks tiles
//     t = pair, pair + n_pairs, ...; all 512 TMEM columns = 2 accumulator stages of 256 fp32 columns, so the
//     epilogue of tile i overlaps the main loop of tile i + 1 (tmem_full / tmem_empty mbarriers; the peer CTA's
//     epilogue warps arrive remotely on the leader's tmem_empty barrier); dedicated epilogue staging.  For
//     launches with many tiles per pair.
//   * ACCS = 1, one tile per pair: grid = 2 * tiles, 256 TMEM columns and a 96 KB ring that the epilogue reuses
//     as its staging area, so TWO CTAs fit on an SM: under programmatic dependent launch the next GEMM's CTAs are
//     already resident (barriers initialised, TMEM allocated, weight tiles in flight) while this one drains --
//     the denoiser's launches have 1-3 tiles per pair, where that overlap is worth more than double buffering
//     (measured: DESIGN 6);
//   * operands: cp.async.bulk.tensor.2d.cta_group::2 (TMA, 128-byte swizzle) into a 4-stage (3 in split mode)
//     ring of 32 KB per CTA; both CTAs' loads complete on the LEADER's full barrier (2 x 32 KB expected), the
//     leader's single MMA thread issues for the pair and tcgen05.commit multicasts slot-free / accumulator-ready
//     to both CTAs;
//   * epilogue (8 warps per CTA; two warps share a TMEM lane group and split the columns): tcgen05.ld 64 columns ->
//     bias / residual / positional / GELU in registers -> 128B-swizzled staging tiles in shared memory ->
//     TMA STORE (cp.async.bulk.tensor.2d.global.shared::cta, SASS UTMASTG) of fp32 and/or bf16 (hi | lo) tiles:
//     no per-thread global stores, row clipping by the tensor map;
//   * under PDL the weight tiles of the first stages are requested before griddepcontrol.wait (as gemm_tc.cu).
//
// "split" (RG_PREC_BF16X3): three passes A_hi*W_hi + A_lo*W_hi + A_hi*W_lo into the same accumulator.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "rg_common.cuh"
#include "rg_gemm_tc.h"
#include "rg_tcgen05.cuh"

namespace {

using namespace rg_tc;

constexpr int BM = 128;             // rows per CTA (256 per pair)
constexpr int BN = 256;             // columns per pair tile; each CTA stages BN/2 weight rows
constexpr int BK = 64;
// epilogue warps per CTA: 8 (two per TMEM lane group, splitting the columns) in the persistent variant, 4 in the
// one-tile-per-pair variant, where two CTAs share an SM and 170 registers per thread avoid spilling the 64-column
// accumulator + residual registers
__host__ __device__ constexpr int epi_warps(int accs) { return accs == 1 ? 4 : 8; }
__host__ __device__ constexpr int threads_of(int accs) { return 64 + 32 * epi_warps(accs); }
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait: a protocol error traps instead of hanging the device
__device__ __forceinline__ void mbar_wait_g(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();
    } while (!ok);
}
// TMA load issued by either CTA of the pair; completes on the mbarrier at cluster address `bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int EPI_BYTES = 12288;    // per epilogue warp: two fp32 tiles + one bf16 tile (hi, then lo) of 4 KB

__device__ __forceinline__ long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (long long)t;
}

template <int SPLIT, int EPI, int ACCS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads_of(ACCS), ACCS == 1 ? 2 : 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmC16, RgGemmTc p) {
    constexpr int STAGES = ACCS == 1 ? 3 : 4;
    constexpr int TMEM_COLS = ACCS * BN;
    constexpr int EPI_WARPS = epi_warps(ACCS);
    constexpr int WCOLS = BN / (EPI_WARPS / 4);          // columns per epilogue warp
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[ACCS], tmem_empty_bar[ACCS];
    __shared__ uint32_t tmem_base_smem;

    rg_pdl_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    long long* tr = p.trace ? p.trace + (long long)blockIdx.x * 16 : nullptr;     // diagnostics (rg_probe_gemm_trace)
#define RG_STAMP(slot) do { if (tr) tr[slot] = gtime(); } while (0)
    if (tr && threadIdx.x == 0) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        tr[0] = gtime(); tr[10] = sm;
    }
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_n = p.N / BN, tiles_m = (p.M + 2 * BM - 1) / (2 * BM), total = tiles_m * tiles_n;
    const int nkb = p.K / BK;
    const int total_kb = SPLIT ? 3 * nkb : nkb;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC32)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC16)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
            for (int a = 0; a < ACCS; ++a) { mbar_init(smem_u32(&tmem_full_bar[a]), 1); mbar_init(smem_u32(&tmem_empty_bar[a]), 2 * EPI_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                 // barrier inits + TMEM visible to both CTAs of the pair
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) RG_STAMP(1);

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 rows of A, own half of the 256 weight rows =====
        if (elect_one()) {
            auto coords = [&](int j, int& ka, int& kw) {
                const int pass = SPLIT ? j / nkb : 0, kb = j - pass * nkb;      // pass 0: hi*hi, 1: lo*hi, 2: hi*lo
                ka = (pass == 1 ? p.a_lo_off : 0) + kb * BK;
                kw = (pass == 2 ? p.w_lo_off : 0) + kb * BK;
            };
            uint32_t full_leader[STAGES];
#pragma unroll
            for (int s = 0; s < STAGES; ++s) full_leader[s] = mapa_u32(smem_u32(&full_bar[s]), 0);
            int it = 0, j0 = 0;
            if (pair < total) {
                // weights do not depend on the previous kernel: request the first tiles before the PDL wait
                const int w_row = (pair % tiles_n) * BN + (int)rank * (BN / 2), a_row = (pair / tiles_n) * 2 * BM + (int)rank * BM;
                const int pre = total_kb < STAGES ? total_kb : STAGES;
                for (int j = 0; j < pre; ++j) {
                    int ka, kw;
                    coords(j, ka, kw);
                    if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[j]), 2 * STAGE_BYTES);
                    tma_load_2d_pair(smem_u32(smem + j * STAGE_BYTES) + A_BYTES, &tmW, full_leader[j], kw, w_row);
                }
                rg_pdl_wait();
                for (int j = 0; j < pre; ++j) {
                    int ka, kw;
                    coords(j, ka, kw);
                    tma_load_2d_pair(smem_u32(smem + j * STAGE_BYTES), &tmA, full_leader[j], ka, a_row);
                }
                it = pre; j0 = pre;
            } else {
                rg_pdl_wait();
            }
            for (int t = pair; t < total; t += n_pairs) {
                const int w_row = (t % tiles_n) * BN + (int)rank * (BN / 2), a_row = (t / tiles_n) * 2 * BM + (int)rank * BM;
                for (int j = j0; j < total_kb; ++j, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait_g(smem_u32(&empty_bar[s]), ph ^ 1);
                    int ka, kw;
                    coords(j, ka, kw);
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * STAGE_BYTES);
                    tma_load_2d_pair(sa, &tmA, full_leader[s], ka, a_row);
                    tma_load_2d_pair(sa + A_BYTES, &tmW, full_leader[s], kw, w_row);
                }
                j0 = 0;
            }
            RG_STAMP(3);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA issues for the pair =====
        if (rank == 0) {
            const uint32_t idesc = make_idesc(2 * BM, BN);
            int it = 0, tl = 0;
            for (int t = pair; t < total; t += n_pairs, ++tl) {
                const int acc = tl % ACCS, aph = (tl / ACCS) & 1;
                mbar_wait_g(smem_u32(&tmem_empty_bar[acc]), aph ^ 1);      // both CTAs' epilogues drained this stage
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int j = 0; j < total_kb; ++j, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait_g(smem_u32(&full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        if (it == 0) RG_STAMP(4);
                        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
                        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16_pair(d_tmem, da + (k * UMMA_K * 2 >> 4), db + (k * UMMA_K * 2 >> 4), idesc, (j | k) != 0);
                        umma_commit_pair(smem_u32(&empty_bar[s]));
                        if (j == total_kb - 1) { umma_commit_pair(smem_u32(&tmem_full_bar[acc])); RG_STAMP(5); }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue (EPI_WARPS warps per CTA) =====
        rg_pdl_wait();                                  // residual reads / stores touch the previous kernel's buffers
        const int ew = warp - 2, lg = warp & 3, half = ew >> 2;
        // ACCS == 1: one tile per pair, so once tmem_full fires the operand ring is idle in both CTAs: stage there
        const uint32_t stg = smem_u32(smem + (ACCS == 1 ? 0 : STAGES * STAGE_BYTES) + ew * EPI_BYTES);
        const uint32_t s_f0 = stg, s_f1 = stg + 4096, s_h = stg + 8192;
        const uint32_t row_off = lane * 128, sw = lane & 7;
        const uint32_t tmem_empty_leader[2] = {mapa_u32(smem_u32(&tmem_empty_bar[0]), 0), mapa_u32(smem_u32(&tmem_empty_bar[1]), 0)};
        const bool has32 = p.C32 != nullptr, has16 = p.C16_ != nullptr, lo16 = SPLIT && p.c16_lo_off != 0;
        int tl = 0;
        for (int t = pair; t < total; t += n_pairs, ++tl) {
            const int acc = tl % ACCS, aph = (tl / ACCS) & 1;
            const int m0 = (t / tiles_n) * 2 * BM + (int)rank * BM, n0 = (t % tiles_n) * BN;
            const int row = m0 + lg * 32 + lane;
            const bool row_ok = row < p.M;
            mbar_wait_g(smem_u32(&tmem_full_bar[acc]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 64 && tl == 0) RG_STAMP(6);
#pragma unroll 1
            for (int itc = 0; itc < WCOLS / 64; ++itc) {
                const int ncol = n0 + half * WCOLS + itc * 64;
                const uint32_t ta = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + acc * BN + half * WCOLS + itc * 64;
                uint32_t v[64];
                tmem_ld32(ta, v);
                tmem_ld32(ta + 32, v + 32);
                float4 rr[16];
                if (EPI == RG_EPI_BIAS_RESIDUAL || EPI == RG_EPI_BIAS_POS) {
                    const float* src = EPI == RG_EPI_BIAS_RESIDUAL ? p.R + (long long)row * p.ldr + ncol
                                                                   : p.pos + (long long)(row % p.pos_T) * p.N + ncol;
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        rr[q] = row_ok ? *reinterpret_cast<const float4*>(src + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (itc == WCOLS / 64 - 1) {            // accumulator stage fully read by this warp: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0) mbar_arrive_remote(tmem_empty_leader[acc]);
                }
                // the TMA stores of the previous iteration must have finished READING the staging tiles
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {        // 8 columns per step = one 16-byte bf16 chunk, two fp32 chunks
                    float f[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * qq + h;
                        const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ncol + 4 * q))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 ad = bv;                 // acc + (bias + residual): the op order of gemm_tc.cu (bit-identical)
                        if (EPI == RG_EPI_BIAS_RESIDUAL || EPI == RG_EPI_BIAS_POS) {
                            ad.x += rr[q].x; ad.y += rr[q].y; ad.z += rr[q].z; ad.w += rr[q].w;
                        }
                        float4 x = make_float4(__uint_as_float(v[4 * q]) + ad.x, __uint_as_float(v[4 * q + 1]) + ad.y,
                                               __uint_as_float(v[4 * q + 2]) + ad.z, __uint_as_float(v[4 * q + 3]) + ad.w);
                        if (EPI == RG_EPI_BIAS_GELU) {
                            x.x = rg_gelu_fast(x.x); x.y = rg_gelu_fast(x.y); x.z = rg_gelu_fast(x.z); x.w = rg_gelu_fast(x.w);
                        } else if (EPI == RG_EPI_BIAS_SILU) {
                            x.x = rg_silu(x.x); x.y = rg_silu(x.y); x.z = rg_silu(x.z); x.w = rg_silu(x.w);
                        }
                        f[4 * h] = x.x; f[4 * h + 1] = x.y; f[4 * h + 2] = x.z; f[4 * h + 3] = x.w;
                        if (has32) {                    // fp32 tile: 32 columns per 128-byte row, chunk = 4 floats
                            const uint32_t dst = (q < 8 ? s_f0 : s_f1) + row_off + ((static_cast<uint32_t>(q & 7) ^ sw) << 4);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
                        }
                    }
                    if (has16) {                        // bf16 tile: 64 columns per 128-byte row, chunk = 8 bf16
                        const uint32_t h0 = pack_bf16(f[0], f[1]), h1 = pack_bf16(f[2], f[3]), h2 = pack_bf16(f[4], f[5]), h3 = pack_bf16(f[6], f[7]);
                        const uint32_t off = row_off + ((static_cast<uint32_t>(qq) ^ sw) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_h + off), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
                        if (lo16) {                     // lo = bf16(x - hi): kept in the accumulator registers for the second store
                            const uint32_t hh[4] = {h0, h1, h2, h3};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&hh[i]);
                                v[4 * qq + i] = pack_bf16(f[2 * i] - __low2float(hb), f[2 * i + 1] - __high2float(hb));
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to TMA
                __syncwarp();
                const int r0 = m0 + lg * 32;
                if (lane == 0) {
                    if (has32) {
                        tma_store_2d(&tmC32, s_f0, p.c32_col0 + ncol, r0);
                        tma_store_2d(&tmC32, s_f1, p.c32_col0 + ncol + 32, r0);
                    }
                    if (has16) tma_store_2d(&tmC16, s_h, p.c16_col0 + ncol, r0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (has16 && lo16) {                    // the lo plane goes through the same 4 KB tile once the hi store has read it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        const uint32_t off = row_off + ((static_cast<uint32_t>(qq) ^ sw) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_h + off), "r"(v[4 * qq]), "r"(v[4 * qq + 1]),
                                     "r"(v[4 * qq + 2]), "r"(v[4 * qq + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC16, s_h, p.c16_col0 + p.c16_lo_off + ncol, r0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
        if (threadIdx.x == 64) RG_STAMP(7);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
        if (threadIdx.x == 64) RG_STAMP(8);
    }
    __syncwarp();                                       // reconverge the role branches before the aligned barrier
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                 // nobody exits while the peer may still signal / read us
    if (threadIdx.x == 0) RG_STAMP(9);
#undef RG_STAMP
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ---- 2-CTA kernel with a SHARED WEIGHT TILE and the 128x128 kernel's epilogue ("pair128") ------------------------
// A CTA pair owns 256 rows x 128 columns: each CTA keeps a 128 x 128 fp32 accumulator (128 TMEM columns, the same
// epilogue work and the same CTA count as gemm_tc_kernel), stages its own 128 rows of A but only 64 of the tile's
// 128 weight rows per K block -- tcgen05.mma.cta_group::2 (UMMA 256x128x16) reads both halves.  Operand bytes per
// CTA and K block drop from 32 KB to 24 KB (64 -> 85 flop per operand byte) and every weight tile is fetched once
// per 256 rows: the chain is L2-bound, so that is where the time is.  4-stage ring of 24 KB, two CTAs per
// SM, weight tiles requested before griddepcontrol.wait; epilogue: TMEM -> registers -> transpose through the idle
// ring -> coalesced 256-byte row segments (bias / residual / pos loads and all stores contiguous), identical
// arithmetic and operation order to gemm_tc_kernel, hence bit-identical results.
constexpr int P_BN = 128, P_STAGES = 4;
constexpr int P_A_BYTES = BM * BK * 2, P_B_BYTES = (P_BN / 2) * BK * 2, P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;

template <int SPLIT, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 2)
gemm_pair128_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, RgGemmTc p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[P_STAGES], empty_bar[P_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    rg_pdl_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int tiles_n = p.N / P_BN;
    const int m_pair = pair / tiles_n, n_tile = pair - m_pair * tiles_n;
    const int m0 = m_pair * 2 * BM + (int)rank * BM, n0 = n_tile * P_BN;
    const int nkb = p.K / BK;
    const int total_kb = SPLIT ? 3 * nkb : nkb;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < P_STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
            mbar_init(smem_u32(&tmem_full_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(P_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (elect_one()) {
            auto coords = [&](int j, int& ka, int& kw) {
                const int pass = SPLIT ? j / nkb : 0, kb = j - pass * nkb;
                ka = (pass == 1 ? p.a_lo_off : 0) + kb * BK;
                kw = (pass == 2 ? p.w_lo_off : 0) + kb * BK;
            };
            auto full_leader = [&](int s) { return mapa_u32(smem_u32(&full_bar[s]), 0); };
            const int w_row = n0 + (int)rank * (P_BN / 2);
            const int pre = total_kb < P_STAGES ? total_kb : P_STAGES;
            for (int j = 0; j < pre; ++j) {
                int ka, kw;
                coords(j, ka, kw);
                if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[j]), 2 * P_STAGE_BYTES);
                tma_load_2d_pair(smem_u32(smem + j * P_STAGE_BYTES) + P_A_BYTES, &tmW, full_leader(j), kw, w_row);
            }
            rg_pdl_wait();
            for (int j = 0; j < pre; ++j) {
                int ka, kw;
                coords(j, ka, kw);
                tma_load_2d_pair(smem_u32(smem + j * P_STAGE_BYTES), &tmA, full_leader(j), ka, m0);
            }
            for (int j = pre; j < total_kb; ++j) {
                const int s = j % P_STAGES, ph = (j / P_STAGES) & 1;
                mbar_wait_g(smem_u32(&empty_bar[s]), ph ^ 1);
                int ka, kw;
                coords(j, ka, kw);
                const uint32_t sa = smem_u32(smem + s * P_STAGE_BYTES);
                if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * P_STAGE_BYTES);
                tma_load_2d_pair(sa, &tmA, full_leader(s), ka, m0);
                tma_load_2d_pair(sa + P_A_BYTES, &tmW, full_leader(s), kw, w_row);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA) =====
        if (rank == 0) {
            const uint32_t idesc = make_idesc(2 * BM, P_BN);
            for (int j = 0; j < total_kb; ++j) {
                const int s = j % P_STAGES, ph = (j / P_STAGES) & 1;
                mbar_wait_g(smem_u32(&full_bar[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + s * P_STAGE_BYTES), sb = sa + P_A_BYTES;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16_pair(tmem_base, da + (k * UMMA_K * 2 >> 4), db + (k * UMMA_K * 2 >> 4), idesc, (j | k) != 0);
                    umma_commit_pair(smem_u32(&empty_bar[s]));
                    if (j == total_kb - 1) umma_commit_pair(smem_u32(&tmem_full_bar));
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue (8 warps): the row pass of gemm_tc_kernel =====
        rg_pdl_wait();
        mbar_wait_g(smem_u32(&tmem_full_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3, half = (warp - 2) >> 2;
        constexpr int HC = P_BN / 2, PITCH = HC + 4;
        float* stage = reinterpret_cast<float*>(smem) + ((warp - 2) * 32) * PITCH;     // the ring is idle in both CTAs now
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
        const int sub = lane >> 4, cc = (lane & 15) * 4;
        const int tcol = half * HC;
        const int cbase = n0 + tcol + cc;
        const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + tcol + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int row_first = m0 + lg * 32 + sub;
        const float* rsrc = nullptr;
        long long rstep = 0;
        if (EPI == RG_EPI_BIAS_RESIDUAL) {
            rsrc = p.R + (long long)row_first * p.ldr + n0 + tcol + cc;
            rstep = 2ll * p.ldr;
        }
        float* c32 = p.C32 ? p.C32 + (long long)row_first * p.ldc32 + cbase : nullptr;
        __nv_bfloat16* c16 = p.C16_ ? reinterpret_cast<__nv_bfloat16*>(p.C16_) + (long long)row_first * p.ldc16 + cbase : nullptr;
        const float* srow = stage + sub * PITCH + cc;
        constexpr int RB = 4;
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
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P_BN));
    }
}

template <int SPLIT, int EPI>
cudaError_t launch_pair128(const CUtensorMap& tmA, const CUtensorMap& tmW64, const RgGemmTc& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)P_STAGES * P_STAGE_BYTES + 1024;
    static_assert(P_STAGES * P_STAGE_BYTES >= 8 * 32 * (P_BN / 2 + 4) * 4, "the ring must hold the epilogue staging");
    auto kern = gemm_pair128_kernel<SPLIT, EPI>;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int pairs = (p.N / P_BN) * ((p.M + 2 * BM - 1) / (2 * BM));
    const dim3 grid(2 * pairs);
    if (p.no_pdl) {
        kern<<<grid, 320, smem, st>>>(tmA, tmW64, p);
        return cudaGetLastError();
    }
    return rg_launch_pdl(kern, grid, dim3(320), smem, st, tmA, tmW64, p);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode2 = nullptr;

template <int SPLIT, int EPI, int ACCS>
cudaError_t launch2(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC32, const CUtensorMap& tmC16,
                    const RgGemmTc& p, cudaStream_t st) {
    constexpr int STAGES = ACCS == 1 ? 3 : 4;
    constexpr int EPI_WARPS = epi_warps(ACCS), THREADS = threads_of(ACCS);
    constexpr size_t smem = (size_t)STAGES * STAGE_BYTES + (ACCS == 1 ? 0 : (size_t)EPI_WARPS * EPI_BYTES) + 1024;
    static_assert(ACCS == 2 || STAGES * STAGE_BYTES >= EPI_WARPS * EPI_BYTES, "the ring must hold the epilogue staging");
    auto kern = gemm2_tc_kernel<SPLIT, EPI, ACCS>;
    static int max_pairs = 0;                   // resident CTA pairs of the persistent variant (one CTA per SM)
    static int attr_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (attr_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (ACCS == 2) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(2 * 74); q.blockDim = dim3(THREADS); q.dynamicSmemBytes = smem;
            int n = 0;
            e = cudaOccupancyMaxActiveClusters(&n, kern, &q);
            if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = 74; }     // any grid is correct; this only sizes it
            max_pairs = n > 74 ? 74 : n;
        }
        attr_dev = dev;
    }
    const int tiles = (p.N / BN) * ((p.M + 2 * BM - 1) / (2 * BM));
    const int pairs = ACCS == 1 ? tiles : (tiles < max_pairs ? tiles : max_pairs);
    const dim3 grid(2 * pairs);
    if (p.no_pdl) {
        kern<<<grid, THREADS, smem, st>>>(tmA, tmW, tmC32, tmC16, p);
        return cudaGetLastError();
    }
    return rg_launch_pdl(kern, grid, dim3(THREADS), smem, st, tmA, tmW, tmC32, tmC16, p);
}

template <int SPLIT, int ACCS>
cudaError_t launch2_epi(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& c32, const CUtensorMap& c16,
                        const RgGemmTc& p, cudaStream_t st) {
    switch (p.epi) {
        case RG_EPI_BIAS: return launch2<SPLIT, RG_EPI_BIAS, ACCS>(tmA, tmW, c32, c16, p, st);
        case RG_EPI_BIAS_RESIDUAL: return launch2<SPLIT, RG_EPI_BIAS_RESIDUAL, ACCS>(tmA, tmW, c32, c16, p, st);
        case RG_EPI_BIAS_GELU: return launch2<SPLIT, RG_EPI_BIAS_GELU, ACCS>(tmA, tmW, c32, c16, p, st);
        case RG_EPI_BIAS_POS: return launch2<SPLIT, RG_EPI_BIAS_POS, ACCS>(tmA, tmW, c32, c16, p, st);
        case RG_EPI_BIAS_SILU: return launch2<SPLIT, RG_EPI_BIAS_SILU, ACCS>(tmA, tmW, c32, c16, p, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// Row-major [rows, cols] tensor with pitch ld (elements) as a TMA STORE destination of the epilogue: box = one
// 128-byte swizzle row (32 fp32 / 64 bf16 columns) x 32 rows (one epilogue warp's TMEM lane group).
cudaError_t rg_make_store_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int elem_bytes) {
    if (!g_encode2) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        g_encode2 = reinterpret_cast<EncodeTiledFn>(fn);
    }
    if (elem_bytes != 2 && elem_bytes != 4) return cudaErrorInvalidValue;
    if ((ld * elem_bytes) % 16 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return cudaErrorInvalidValue;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode2(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

bool rg_gemm_pair128_eligible(const RgGemmTc& p) {
    return p.tmW64 != nullptr && p.N % P_BN == 0 && p.K % BK == 0 && p.groups <= 1 && !p.trace && (p.C32 || p.C16_) &&
           (!p.C32 || p.ldc32 % 4 == 0) && (!p.C16_ || (p.ldc16 % 8 == 0 && p.c16_lo_off % 4 == 0)) && (!p.R || p.ldr % 4 == 0);
}

// tmA: box 64 x 128 rows over A; p.tmW64: box 64 x 64 rows over W (each CTA stages half of a 128-row weight tile)
cudaError_t rg_launch_gemm_pair128(const CUtensorMap& tmA, const RgGemmTc& p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return cudaSuccess;
    if (!rg_gemm_pair128_eligible(p)) return cudaErrorInvalidValue;
    if (p.epi == RG_EPI_BIAS_RESIDUAL && !p.R) return cudaErrorInvalidValue;
    if (p.epi == RG_EPI_BIAS_POS && (!p.pos || p.pos_T <= 0)) return cudaErrorInvalidValue;
    const CUtensorMap& w = *p.tmW64;
#define RG_P128(SP)                                                                                      \
    switch (p.epi) {                                                                                     \
        case RG_EPI_BIAS: return launch_pair128<SP, RG_EPI_BIAS>(tmA, w, p, st);                         \
        case RG_EPI_BIAS_RESIDUAL: return launch_pair128<SP, RG_EPI_BIAS_RESIDUAL>(tmA, w, p, st);       \
        case RG_EPI_BIAS_GELU: return launch_pair128<SP, RG_EPI_BIAS_GELU>(tmA, w, p, st);               \
        case RG_EPI_BIAS_POS: return launch_pair128<SP, RG_EPI_BIAS_POS>(tmA, w, p, st);                 \
        case RG_EPI_BIAS_SILU: return launch_pair128<SP, RG_EPI_BIAS_SILU>(tmA, w, p, st);               \
        default: return cudaErrorInvalidValue;                                                           \
    }
    if (p.split) { RG_P128(1) }
    RG_P128(0)
#undef RG_P128
}

bool rg_gemm2_eligible(const RgGemmTc& p) {
    return p.N % BN == 0 && p.K % BK == 0 && p.groups <= 1 && (p.C32 || p.C16_) &&
           (!p.C32 || p.tmC32) && (!p.C16_ || p.tmC16) && (!p.R || (p.ldr % 4 == 0)) &&
           (p.epi != RG_EPI_BIAS_POS || (p.N % 4 == 0 && p.pos && p.pos_T > 0));
}

// tmA: box 64 x 128 rows over A; tmW: box 64 x 128 rows over W (each CTA stages half of a 256-row weight tile);
// p.tmC32 / p.tmC16: store maps (rg_make_store_map) of the outputs in use.
cudaError_t rg_launch_gemm2_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return cudaSuccess;
    if (!rg_gemm2_eligible(p)) return cudaErrorInvalidValue;
    if (p.epi == RG_EPI_BIAS_RESIDUAL && !p.R) return cudaErrorInvalidValue;
    const CUtensorMap& c32 = p.tmC32 ? *p.tmC32 : tmA;      // unused maps still need a valid descriptor
    const CUtensorMap& c16 = p.tmC16 ? *p.tmC16 : tmA;
    // many tiles per pair: persistent with two accumulator stages; else one tile per pair, two CTAs per SM
    const int tiles = (p.N / BN) * ((p.M + 2 * BM - 1) / (2 * BM));
    if (tiles >= rg_gemm2_persist_tiles)
        return p.split ? launch2_epi<1, 2>(tmA, tmW, c32, c16, p, st) : launch2_epi<0, 2>(tmA, tmW, c32, c16, p, st);
    return p.split ? launch2_epi<1, 1>(tmA, tmW, c32, c16, p, st) : launch2_epi<0, 1>(tmA, tmW, c32, c16, p, st);
}
