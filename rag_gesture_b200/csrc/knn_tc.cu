// Exemplar kNN for LARGE query batches (configs[3], Q = 4096): tensor-core similarity + certified
// over-selection + exact fp32 re-score.  Results are bit-identical to the exact scan (rg_knn_topk).
//
//  1. rg_knn_index_create: one-off bf16 copy of the shard (TMA operand) and the largest row norm.
//  2. knn_tc_kernel: persistent tcgen05 GEMM  S = Q_bf16 * D_bf16^T  that never writes S.  One CTA per SM
//     walks work items (256-query tile x chunk of database tiles).  Per database tile of 256 rows the MMA
//     warp issues 2 x 48 UMMAs (128x256x16, two 128-lane accumulators = all 512 TMEM columns); TMA streams
//     both operands through a 3 x 64 KB ring.  TMEM lane == query row, so each of the 256 epilogue threads
//     owns one query: it reads its accumulator row with tcgen05.ld, compares every score with its private
//     threshold and keeps the KP best (approximate score, row) of the chunk in a shared-memory list.
//  3. knn_tc_finish_kernel (one block per query): merges the chunk lists by approximate score, re-scores the
//     best RESCORE candidates in fp32 with the SAME summation order as the exact scan, selects the top-k by
//     (score desc, index asc) and issues a CERTIFICATE: every row that was not re-scored has approximate
//     score <= B, and |approx - exact| <= eps = c*|q|*max|d| (bf16 rounding of both operands, unit roundoff
//     2^-8 each, Cauchy-Schwarz), so B + eps < (k-th exact score) proves no such row belongs to the top-k.
//  4. Queries without a certificate (adversarial data, k > KP clusters, NaNs) are re-run through the exact
//     scan by rg_knn_topk_tc; their number is returned.  For unit-norm N(0,1) data about 1 query in 4000 fails at
//     1M rows (its k-th best score sits within eps of a chunk's KP-th best).
//
// Work decomposition: ONE wave of items (q_tiles x chunks ~ 148), so a list sees a long chunk (a running top-KP
// costs ~KP*ln(rows/KP) replacements) and the CTAs running together are the q_tiles tiles of the same few chunks:
// a database tile comes from DRAM once (ncu: 1.544 GB read for the 1.536 GB shard, L2 hit 95 %).
// Roofline: tensor pipe.  Operands are 768 KB from L2 per 100.7 MFLOP tile; the accumulators are single-buffered
// (two of them fill TMEM), so the epilogue serialises with the MMAs and its length decides the rest: with the
// branch-per-4-scores test and the FMNMX-tree list update the kernel runs at 1277 TF/s = 92 % of the measured
// sustained bf16 peak (4096 x 1M x 768 in 4.93 ms).
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

#include "../../include/rg_b200.h"
#include "rg_common.cuh"
#include "rg_gemm_tc.h"
#include "rg_internal.h"
#include "rg_tcgen05.cuh"

namespace {

using namespace rg_tc;

constexpr int TQ = 256;            // queries per work item: two 128-lane accumulators
constexpr int TN = 256;            // database rows per tile (UMMA N)
constexpr int BK = 64;             // one 128-byte swizzle row of bf16
constexpr int STAGES = 3;
constexpr int KP = 16;             // candidates kept per (query, chunk)
constexpr int RESCORE = 64;        // candidates re-scored exactly per query
constexpr int MAX_CHUNKS = 256;    // chunks * KP <= 4096 entries sorted per query
constexpr int STAGE_BYTES = (TQ + TN) * BK * 2;                    // 64 KB
constexpr int LIST_BYTES = TQ * KP * (int)(sizeof(float) + sizeof(int));   // 32 KB
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + LIST_BYTES;
constexpr int N_EPI_WARPS = 8;
constexpr long long IDX_NONE = 0x7fffffffffffffffLL;
// |approx - exact| <= C_EPS * |q| * |d|: 2u + u^2 with u = 2^-8 (two bf16 roundings per product), plus the
// fp32 accumulation error of both sums (<= 2 * 768 * 2^-23 relative to sum|q_i d_i| at dim 768; scaled with
// dim below), plus slack for the fp32 rounding of the norms themselves.
// The slot bits in the stored scores move a score / threshold by at most 2^-19 relative (|score| <= |q||d|).
constexpr float C_EPS_BF16 = 0.0078125f + 0.0000153f + 0.0000077f;

struct KnnTcParams {
    int n_rows, q_total, nkb;
    int q_tiles, chunks, tiles_per_chunk, tiles_total;
    int single_half;               // q_total <= 128: the second accumulator is never issued
    float* cand_score;             // [q_total][chunks][KP]
    int* cand_idx;
};

// Candidate list of one query thread: KP (score, row) pairs, UNSORTED, in shared memory ([KP][TQ]: the
// thread index is the fastest dimension, so every access is conflict-free).  The low 4 mantissa bits of a
// stored score hold its slot number (a 2^-19 relative perturbation, charged to the certificate's eps), so
// the minimum of the list -- the thread's threshold `thr` -- also names the slot a new candidate overwrites,
// and the new minimum is one 4-level FMNMX tree over KP independent loads (no compare/select chain).
// One out-of-line copy: 256 call sites sit in the unrolled epilogue loop.
static_assert(KP == 16, "slot number is kept in 4 mantissa bits");
__device__ __forceinline__ float slot_init(int j) { return __uint_as_float(0xff7ffff0u | (unsigned)j); }   // ~ -FLT_MAX
__device__ __noinline__ float cand_replace(float* ls, int* li, float v, int id, int n_rows, float thr) {
    if (id >= n_rows) return thr;              // TMA zero fill of the last, partial database tile
    const int slot = (int)(__float_as_uint(thr) & 15u);
    ls[slot * TQ] = __uint_as_float((__float_as_uint(v) & ~15u) | (unsigned)slot);
    li[slot * TQ] = id;
    float u[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) u[j] = ls[j * TQ];
#pragma unroll
    for (int w = KP / 2; w >= 1; w >>= 1)
#pragma unroll
        for (int j = 0; j < w; ++j) u[j] = fminf(u[j], u[j + w]);
    return u[0];
}

__global__ void __launch_bounds__(320, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmD, KnnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* list_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);     // [KP][TQ]
    int* list_i = reinterpret_cast<int*>(list_s + KP * TQ);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = p.q_tiles * p.chunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmD)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
            mbar_init(smem_u32(&tmem_full_bar), 1);
            mbar_init(smem_u32(&tmem_empty_bar), N_EPI_WARPS);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer: both operands of every (tile, k-block), in ring order =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int qt = item % p.q_tiles, ch = item / p.q_tiles;
                const int t0 = ch * p.tiles_per_chunk;
                const int t1 = min(t0 + p.tiles_per_chunk, p.tiles_total);
                for (int t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                        const int s = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + TQ * BK * 2;
                        mbar_expect_tx(smem_u32(&full_bar[s]), STAGE_BYTES);
                        tma_load_2d(sa, &tmQ, smem_u32(&full_bar[s]), kb * BK, qt * TQ);
                        tma_load_2d(sb, &tmD, smem_u32(&full_bar[s]), kb * BK, t * TN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: acc0 = queries 0..127 of the tile, acc1 = queries 128..255 =====
        const uint32_t idesc = make_idesc(128, TN);
        uint32_t it = 0, tl = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int ch = item / p.q_tiles;
            const int t0 = ch * p.tiles_per_chunk;
            const int t1 = min(t0 + p.tiles_per_chunk, p.tiles_total);
            for (int t = t0; t < t1; ++t, ++tl) {
                mbar_wait(smem_u32(&tmem_empty_bar), (tl & 1) ^ 1);     // epilogue has drained the previous tile
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + TQ * BK * 2;
                        const uint64_t da0 = make_smem_desc(sa), da1 = make_smem_desc(sa + 128 * BK * 2);
                        const uint64_t db = make_smem_desc(sb);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t ko = (uint64_t)(k * UMMA_K * 2 >> 4);
                            umma_bf16(tmem_base, da0 + ko, db + ko, idesc, (kb | k) != 0);
                            if (!p.single_half) umma_bf16(tmem_base + TN, da1 + ko, db + ko, idesc, (kb | k) != 0);
                        }
                        umma_commit(smem_u32(&empty_bar[s]));
                        if (kb == p.nkb - 1) umma_commit(smem_u32(&tmem_full_bar));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: thread == query row; private top-KP of the chunk =====
        const int lg = warp & 3;                       // TMEM lane group this warp may read
        const int half = (warp - 2) >> 2;              // which accumulator
        const int qrow = half * 128 + lg * 32 + lane;  // query row inside the tile
        float* ls = list_s + qrow;
        int* li = list_i + qrow;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + half * TN;
        uint32_t tl = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int qt = item % p.q_tiles, ch = item / p.q_tiles;
            const int t0 = ch * p.tiles_per_chunk;
            const int t1 = min(t0 + p.tiles_per_chunk, p.tiles_total);
#pragma unroll
            for (int j = 0; j < KP; ++j) { ls[j * TQ] = slot_init(j); li[j * TQ] = -1; }
            float thr = slot_init(KP - 1);          // the most negative of the initial entries
            for (int t = t0; t < t1; ++t, ++tl) {
                mbar_wait(smem_u32(&tmem_full_bar), tl & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int row0 = t * TN;
#pragma unroll 1
                for (int c = 0; c < TN / 64; ++c) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + c * 64, v0);
                    tmem_ld32(taddr + c * 64 + 32, v1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c == TN / 64 - 1) {            // accumulator fully read: hand TMEM back to the MMA warp
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar));
                    }
                    // four scores per branch: the common case (none above the threshold) costs a 3-input max
                    // tree and one compare; only a group with a hit is examined element by element
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t* v = h ? v1 : v0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                            const float s2 = __uint_as_float(v[i + 2]), s3 = __uint_as_float(v[i + 3]);
                            if (fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)) > thr) {
                                const int r = row0 + c * 64 + h * 32 + i;
                                if (s0 > thr) thr = cand_replace(ls, li, s0, r, p.n_rows, thr);
                                if (s1 > thr) thr = cand_replace(ls, li, s1, r + 1, p.n_rows, thr);
                                if (s2 > thr) thr = cand_replace(ls, li, s2, r + 2, p.n_rows, thr);
                                if (s3 > thr) thr = cand_replace(ls, li, s3, r + 3, p.n_rows, thr);
                            }
                        }
                    }
                }
            }
            const int q = qt * TQ + qrow;
            if (q < p.q_total) {
                const size_t o = ((size_t)q * p.chunks + ch) * KP;
#pragma unroll
                for (int j = 0; j < KP; ++j) { p.cand_score[o + j] = ls[j * TQ]; p.cand_idx[o + j] = li[j * TQ]; }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// fp32 -> bf16 copy of the shard (round to nearest even) + largest row norm; one warp per row
__global__ void __launch_bounds__(256) knn_index_build_kernel(const float* __restrict__ db, long long n, int dim,
                                                             __nv_bfloat16* __restrict__ out, float* dmax) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float4* src = reinterpret_cast<const float4*>(db + r * dim);
    uint2* dst = reinterpret_cast<uint2*>(out + r * dim);
    float ss = 0.f;
    for (int e = lane; e < dim / 4; e += 32) {
        const float4 v = __ldg(src + e);
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
        dst[e] = pk;
    }
    ss = rg_warp_sum(ss);
    // norms are >= 0, so the IEEE bit pattern is monotone as an unsigned integer; NaN (0x7fc00000) wins the max
    if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(dmax), __float_as_uint(sqrtf(ss) * 1.0000005f));
}

__global__ void __launch_bounds__(256) knn_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(out)[i] = pk;
}

__device__ __forceinline__ bool better(float s, long long i, float s2, long long i2) {
    return s > s2 || (s == s2 && i < i2);
}
// the warp-distributed sorted list of knn.cu (lane j = j-th best)
__device__ __forceinline__ void list_insert(float& ls, long long& li, float s, long long id, int k, int lane) {
    const unsigned ahead = __ballot_sync(0xffffffffu, lane < k && better(ls, li, s, id));
    const int pos = __popc(ahead);
    if (pos >= k) return;
    const float us = __shfl_up_sync(0xffffffffu, ls, 1);
    const long long ui = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == pos) { ls = s; li = id; }
    else if (lane > pos) { ls = us; li = ui; }
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red, int warp, int lane) {
    v = is_max ? rg_warp_max(v) : rg_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

__global__ void __launch_bounds__(256)
knn_tc_finish_kernel(const float* __restrict__ db, int dim, const float* __restrict__ queries,
                     const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int chunks, int ncp,
                     int k, long long idx_base, const float* __restrict__ dmax, float c_eps,
                     long long* __restrict__ out_idx, float* __restrict__ out_score,
                     int* __restrict__ fail_count, int* __restrict__ fail_list) {
    extern __shared__ __align__(16) uint8_t fsm[];
    float* sa = reinterpret_cast<float*>(fsm);             // [ncp] approximate scores
    int* si = reinterpret_cast<int*>(sa + ncp);            // [ncp] shard-local rows
    __shared__ float red[8];
    __shared__ float ex[RESCORE];
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nc = chunks * KP;
    int n_valid_t = 0;
    for (int e = tid; e < ncp; e += 256) {
        float s = -INFINITY;
        int id = -1;
        if (e < nc) {
            id = cand_idx[(size_t)q * nc + e];
            s = id >= 0 ? cand_score[(size_t)q * nc + e] : -INFINITY;
            n_valid_t += id >= 0;
        }
        sa[e] = s;
        si[e] = id;
    }
    __syncthreads();
    // rows outside the lists: a FULL list (unsorted) bounds them by its minimum
    float mt = -INFINITY;
    for (int c = tid; c < chunks; c += 256) {
        float mn = INFINITY;
        bool full = true;
        for (int j = 0; j < KP; ++j) { full = full && si[c * KP + j] >= 0; mn = fminf(mn, sa[c * KP + j]); }
        if (full) mt = fmaxf(mt, mn);
    }
    const float m_max = block_reduce(mt, true, red, warp, lane);
    const int n_valid = (int)(block_reduce((float)n_valid_t, false, red, warp, lane) + 0.5f);
    const float* qv = queries + (size_t)q * dim;
    float qs = 0.f;
    for (int e = tid; e < dim; e += 256) qs = fmaf(qv[e], qv[e], qs);
    const float qnorm = sqrtf(block_reduce(qs, false, red, warp, lane)) * 1.0000005f;
    __syncthreads();
    // bitonic sort, descending by approximate score (ties: lower row first, for determinism)
    for (int size = 2; size <= ncp; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < ncp / 2; t += 256) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const float a = sa[lo], b = sa[hi];
                const int ia = si[lo], ib = si[hi];
                const bool a_first = a > b || (a == b && (unsigned)ia < (unsigned)ib);
                if (a_first != desc) { sa[lo] = b; sa[hi] = a; si[lo] = ib; si[hi] = ia; }
            }
            __syncthreads();
        }
    }
    const int R = n_valid < RESCORE ? n_valid : RESCORE;
    // exact fp32 re-score in the summation order of the exact scan (knn.cu): lane handles float4 e = lane,
    // lane+32, ...; fma over x,y,z,w; xor-butterfly 16..1
    for (int r = warp; r < R; r += 8) {
        const float4* d4 = reinterpret_cast<const float4*>(db + (size_t)si[r] * dim);
        const float4* q4 = reinterpret_cast<const float4*>(qv);
        float acc = 0.f;
        for (int e = lane; e < dim / 4; e += 32) {
            const float4 a = __ldg(d4 + e), w = __ldg(q4 + e);
            acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc);
            acc = fmaf(a.z, w.z, acc); acc = fmaf(a.w, w.w, acc);
        }
        acc = rg_warp_sum(acc);
        if (lane == 0) ex[r] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        float bs = -INFINITY;
        long long bi = IDX_NONE;
        for (int r = 0; r < R; ++r) list_insert(bs, bi, ex[r], (long long)si[r], k, lane);
        if (lane < k) {
            out_idx[(size_t)q * k + lane] = (bi == IDX_NONE) ? -1 : bi + idx_base;
            out_score[(size_t)q * k + lane] = bs;
        }
        const float tau = __shfl_sync(0xffffffffu, bs, k - 1);
        const long long tau_i = __shfl_sync(0xffffffffu, bi, k - 1);
        if (lane == 0) {
            const float bound = fmaxf(m_max, n_valid > R ? sa[R] : -INFINITY);
            bool ok;
            if (bound == -INFINITY && n_valid <= R) ok = true;           // every row of the shard was re-scored
            else ok = (tau_i != IDX_NONE) && (bound + c_eps * qnorm * dmax[0] < tau);   // false on NaN
            if (!ok) fail_list[atomicAdd(fail_count, 1)] = q;
        }
    }
}

__global__ void knn_gather_queries_kernel(const float* __restrict__ queries, const int* __restrict__ list, int dim,
                                          float* __restrict__ out) {
    const int src = list[blockIdx.x];
    for (int e = threadIdx.x; e < dim; e += blockDim.x) out[(size_t)blockIdx.x * dim + e] = queries[(size_t)src * dim + e];
}
__global__ void knn_scatter_results_kernel(const long long* __restrict__ idx, const float* __restrict__ score,
                                           const int* __restrict__ list, int k, long long* __restrict__ out_idx,
                                           float* __restrict__ out_score) {
    const int dst = list[blockIdx.x];
    if (threadIdx.x < k) {
        out_idx[(size_t)dst * k + threadIdx.x] = idx[(size_t)blockIdx.x * k + threadIdx.x];
        out_score[(size_t)dst * k + threadIdx.x] = score[(size_t)blockIdx.x * k + threadIdx.x];
    }
}

struct KnnIndex {
    uint32_t magic;
    long long n;
    int dim;
    const float* db;        // the fp32 shard the index was built from (the exact re-score reads it)
    __nv_bfloat16* db16;
    float* dmax;
    CUtensorMap tmD;
};
constexpr uint32_t KNN_MAGIC = 0x4b4e4e31u;

struct Plan { int q_tiles, chunks, tiles_per_chunk, tiles_total, grid; };
Plan make_plan(long long n, int q) {
    Plan pl;
    pl.tiles_total = (int)((n + TN - 1) / TN);
    pl.q_tiles = (q + TQ - 1) / TQ;
    // ONE wave of items over the 148 SMs: a list costs ~KP*ln(rows/KP) replacements per chunk, so long chunks
    // keep the epilogue (which the single-buffered accumulator serialises with the MMAs) short.  The CTAs that
    // run concurrently are the q_tiles tiles of the same few chunks: a database tile is fetched from DRAM once.
    int chunks = std::max(1, 148 / pl.q_tiles);
    chunks = std::max(1, std::min(chunks, std::min(MAX_CHUNKS, pl.tiles_total)));
    pl.tiles_per_chunk = (pl.tiles_total + chunks - 1) / chunks;
    pl.chunks = (pl.tiles_total + pl.tiles_per_chunk - 1) / pl.tiles_per_chunk;
    pl.grid = std::min(148, pl.q_tiles * pl.chunks);
    return pl;
}

// bf16 queries + the persistent kernel; cand_* sized [q][chunks][KP]
cudaError_t launch_tc_scan(const KnnIndex* ix, const __nv_bfloat16* q16, int q, const Plan& pl, float* cand_score,
                           int* cand_idx, cudaStream_t st) {
    CUtensorMap tmQ;
    cudaError_t e = rg_make_tensor_map(&tmQ, q16, q, ix->dim, ix->dim, TQ);
    if (e != cudaSuccess) return e;
    static bool attr_done[64] = {};              // function attributes are per device
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        e = cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    KnnTcParams p;
    p.n_rows = (int)ix->n; p.q_total = q; p.nkb = ix->dim / BK;
    p.q_tiles = pl.q_tiles; p.chunks = pl.chunks; p.tiles_per_chunk = pl.tiles_per_chunk; p.tiles_total = pl.tiles_total;
    p.single_half = q <= 128;
    p.cand_score = cand_score; p.cand_idx = cand_idx;
    knn_tc_kernel<<<pl.grid, 320, SMEM_BYTES, st>>>(tmQ, ix->tmD, p);
    return cudaGetLastError();
}

int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

// stream-ordered scratch released on every exit path of the entry points below
struct Scratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    template <typename T> cudaError_t get(T** p, size_t bytes) {
        cudaError_t e = cudaMallocAsync((void**)p, bytes ? bytes : 4, st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

}  // namespace

extern "C" int rg_knn_index_create(const float* db, int64_t n, int dim, void** index, void* stream) {
    if (!db || !index) return rg_fail("rg_knn_index_create: null argument");
    if (n < 1 || n > 0x7fffff00LL) return rg_fail("rg_knn_index_create: 1 <= n < 2^31-256 rows per shard");
    if (dim < BK || dim % BK) return rg_fail("rg_knn_index_create: dim must be a multiple of 64");
    cudaStream_t st = (cudaStream_t)stream;
    KnnIndex* ix = new KnnIndex();
    ix->magic = KNN_MAGIC; ix->n = n; ix->dim = dim; ix->db = db; ix->db16 = nullptr; ix->dmax = nullptr;
    cudaError_t e = cudaMalloc((void**)&ix->db16, (size_t)n * dim * 2);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ix->dmax, sizeof(float));
    if (e == cudaSuccess) e = cudaMemsetAsync(ix->dmax, 0, sizeof(float), st);
    if (e == cudaSuccess) {
        knn_index_build_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(db, n, dim, ix->db16, ix->dmax);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = rg_make_tensor_map(&ix->tmD, ix->db16, n, dim, dim, TN);
    if (e != cudaSuccess) {
        cudaFree(ix->db16); cudaFree(ix->dmax);
        delete ix;
        return rg_fail("rg_knn_index_create: %s", cudaGetErrorString(e));
    }
    rg_count_launch(1);
    *index = ix;
    return 0;
}

extern "C" int rg_knn_index_destroy(void* index) {
    KnnIndex* ix = reinterpret_cast<KnnIndex*>(index);
    if (!ix) return 0;
    if (ix->magic != KNN_MAGIC) return rg_fail("rg_knn_index_destroy: not an index handle");
    ix->magic = 0;
    cudaFree(ix->db16); cudaFree(ix->dmax);
    delete ix;
    return 0;
}

extern "C" int rg_knn_topk_tc(void* index, const float* db, const float* queries, int q, int k, int64_t idx_base,
                              int64_t* out_idx, float* out_score, int32_t* n_uncertified, void* stream) {
    KnnIndex* ix = reinterpret_cast<KnnIndex*>(index);
    if (!ix || ix->magic != KNN_MAGIC) return rg_fail("rg_knn_topk_tc: not an index handle");
    if (!db || !queries || !out_idx || !out_score) return rg_fail("rg_knn_topk_tc: null argument");
    if (db != ix->db) return rg_fail("rg_knn_topk_tc: db is not the shard this index was built from");
    if (k < 1 || k > 32) return rg_fail("rg_knn_topk_tc: k must be in [1,32]");
    if (n_uncertified) *n_uncertified = 0;
    if (q <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    rg_keep_mempool();
    const int dim = ix->dim;
    const Plan pl = make_plan(ix->n, q);
    const int nc = pl.chunks * KP, ncp = next_pow2(nc);
    Scratch scratch(st);
    __nv_bfloat16* q16 = nullptr;
    float* cand_score = nullptr;
    int *cand_idx = nullptr, *fail = nullptr;          // fail[0] = count, fail[1..] = query ids
    RG_CU(scratch.get(&q16, (size_t)q * dim * 2));
    RG_CU(scratch.get(&cand_score, (size_t)q * nc * sizeof(float)));
    RG_CU(scratch.get(&cand_idx, (size_t)q * nc * sizeof(int)));
    RG_CU(scratch.get(&fail, (size_t)(q + 1) * sizeof(int)));
    RG_CU(cudaMemsetAsync(fail, 0, sizeof(int), st));
    const long long n4 = (long long)q * dim / 4;
    knn_to_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(queries, q16, n4);
    RG_CU(cudaGetLastError());
    RG_CU(launch_tc_scan(ix, q16, q, pl, cand_score, cand_idx, st));
    // fp32 accumulation error of both sums grows with the reduction length
    const float c_eps = (C_EPS_BF16 + 2.5e-7f * (float)dim) * 1.01f;
    const size_t fsm = (size_t)ncp * 8;
    // per-device function attribute: set whenever it is needed (cheap, and correct on every device of the process)
    if (fsm > 48 * 1024)
        RG_CU(cudaFuncSetAttribute(knn_tc_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
    knn_tc_finish_kernel<<<q, 256, fsm, st>>>(db, dim, queries, cand_score, cand_idx, pl.chunks, ncp, k, idx_base,
                                               ix->dmax, c_eps, (long long*)out_idx, out_score, fail, fail + 1);
    RG_CU(cudaGetLastError());
    rg_count_launch(3);
    int nf = 0;
    RG_CU(cudaMemcpyAsync(&nf, fail, sizeof(int), cudaMemcpyDeviceToHost, st));
    RG_CU(cudaStreamSynchronize(st));
    if (nf > 0) {
        // exact scan for the queries without a certificate
        float* qbuf = nullptr; long long* tidx = nullptr; float* tsc = nullptr;
        RG_CU(scratch.get(&qbuf, (size_t)nf * dim * sizeof(float)));
        RG_CU(scratch.get(&tidx, (size_t)nf * k * sizeof(long long)));
        RG_CU(scratch.get(&tsc, (size_t)nf * k * sizeof(float)));
        knn_gather_queries_kernel<<<nf, 256, 0, st>>>(queries, fail + 1, dim, qbuf);
        RG_CU(cudaGetLastError());
        if (rg_knn_topk(db, ix->n, dim, qbuf, nf, k, idx_base, (int64_t*)tidx, tsc, stream)) return 1;
        knn_scatter_results_kernel<<<nf, 32, 0, st>>>(tidx, tsc, fail + 1, k, (long long*)out_idx, out_score);
        RG_CU(cudaGetLastError());
        rg_count_launch(2);
    }
    if (n_uncertified) *n_uncertified = nf;
    return 0;
}

extern "C" int rg_probe_knn_tc(void* index, const float* queries, int q, int reps, void* flush_buf,
                               int64_t flush_bytes, float* median_ms, void* stream) {
    KnnIndex* ix = reinterpret_cast<KnnIndex*>(index);
    if (!ix || ix->magic != KNN_MAGIC) return rg_fail("rg_probe_knn_tc: not an index handle");
    if (!queries || !median_ms || q < 1 || reps < 1 || reps > 64) return rg_fail("rg_probe_knn_tc: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const Plan pl = make_plan(ix->n, q);
    const int nc = pl.chunks * KP;
    __nv_bfloat16* q16 = nullptr; float* cand_score = nullptr; int* cand_idx = nullptr;
    RG_CU(cudaMalloc((void**)&q16, (size_t)q * ix->dim * 2));
    RG_CU(cudaMalloc((void**)&cand_score, (size_t)q * nc * sizeof(float)));
    RG_CU(cudaMalloc((void**)&cand_idx, (size_t)q * nc * sizeof(int)));
    const long long n4 = (long long)q * ix->dim / 4;
    knn_to_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(queries, q16, n4);
    RG_CU(cudaGetLastError());
    cudaEvent_t e0, e1;
    RG_CU(cudaEventCreate(&e0));
    RG_CU(cudaEventCreate(&e1));
    std::vector<float> ts;
    for (int i = 0; i < reps + 2; ++i) {
        if (flush_buf) RG_CU(cudaMemsetAsync(flush_buf, i, (size_t)flush_bytes, st));
        RG_CU(cudaEventRecord(e0, st));
        RG_CU(launch_tc_scan(ix, q16, q, pl, cand_score, cand_idx, st));
        rg_count_launch(1);
        RG_CU(cudaEventRecord(e1, st));
        RG_CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        RG_CU(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= 2) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    *median_ms = ts[ts.size() / 2];
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(q16); cudaFree(cand_score); cudaFree(cand_idx);
    return 0;
}
