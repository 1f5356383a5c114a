// Exemplar retrieval kernels (K10/K11).
//  * rg_text_similarity: the reference's ranking score, mean(diag(Q D_i^T)) over min(Tq,Td) aligned
//    tokens of un-normalised BERT features (rag/utils.py:86-132) == a flat dot product / m.
//  * rg_knn_topk: exact fp32 dot-product top-k of flat embeddings.  The database is streamed from
//    HBM exactly once per tile of QT queries (queries staged in shared memory, two database rows
//    per warp held in registers) -> HBM-bound for Q <= QT; scores never touch memory.  Each warp
//    keeps its top-k as one (score, index) pair per lane, sorted by (score desc, index asc) ==
//    Python's stable sorted(..., reverse=True) / torch sort(descending, stable) tie order.
//  * rg_knn_merge: merges per-shard candidate lists after the NCCL all-gather (K11).
#include <math.h>
#include <algorithm>
#include <vector>
#include "../../include/rg_b200.h"
#include "rg_common.cuh"
#include "rg_internal.h"

namespace {

constexpr int KNN_QT = 8;        // queries per pass over the database
constexpr int KNN_WARPS = 8;
constexpr long long IDX_NONE = 0x7fffffffffffffffLL;

__device__ __forceinline__ bool better(float s, long long i, float s2, long long i2) {
    return s > s2 || (s == s2 && i < i2);
}

// insert (s, id) into the warp-distributed sorted list (lane j = j-th best), length k <= 32
__device__ __forceinline__ void list_insert(float& ls, long long& li, float s, long long id, int k, int lane) {
    const unsigned ahead = __ballot_sync(0xffffffffu, lane < k && better(ls, li, s, id));
    const int pos = __popc(ahead);           // entries that stay in front of the new one
    if (pos >= k) return;
    const float us = __shfl_up_sync(0xffffffffu, ls, 1);
    const long long ui = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == pos) { ls = s; li = id; }
    else if (lane > pos) { ls = us; li = ui; }
}

__global__ void __launch_bounds__(256) text_similarity_kernel(const float* __restrict__ db,
                                                             const int* __restrict__ db_len,
                                                             int max_len, int dim,
                                                             const float* __restrict__ query, int tq,
                                                             const int* __restrict__ subset,
                                                             long long n_out, float* __restrict__ scores) {
    const long long j = (long long)blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= n_out) return;
    const long long i = subset ? subset[j] : j;
    int m = db_len[i] < tq ? db_len[i] : tq;
    const float4* d4 = reinterpret_cast<const float4*>(db + i * (long long)max_len * dim);
    const float4* q4 = reinterpret_cast<const float4*>(query);
    const int n4 = m * dim / 4;              // aligned tokens are contiguous in both operands
    float acc = 0.f;
    for (int e = lane; e < n4; e += 32) {
        const float4 a = __ldg(d4 + e), b = __ldg(q4 + e);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    acc = rg_warp_sum(acc);
    if (lane == 0) scores[j] = m > 0 ? acc / (float)m : nanf("");   // mean of an empty diagonal is nan in torch
}

// one pass over rows [row0, row1) of the shard for queries [q0, q0+QT)
template <int QT>
__global__ void __launch_bounds__(256) knn_scan_kernel(const float* __restrict__ db, long long n,
                                                      int dim, const float* __restrict__ queries,
                                                      int q_total, int q0, int k,
                                                      long long rows_per_block,
                                                      long long* __restrict__ part_idx,
                                                      float* __restrict__ part_score) {
    extern __shared__ __align__(16) float smem[];
    float* qs = smem;                                                   // [QT][dim]
    float* ms = smem + QT * dim;                                        // [WARPS][QT][32] scores
    long long* mi = reinterpret_cast<long long*>(ms + KNN_WARPS * QT * 32);   // [WARPS][QT][32] idx
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = (q_total - q0 < QT) ? q_total - q0 : QT;
    for (int e = threadIdx.x; e < QT * dim; e += blockDim.x)
        qs[e] = (e / dim < nq) ? queries[(long long)q0 * dim + e] : 0.f;
    __syncthreads();

    float ls[QT];
    long long li[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) { ls[q] = -INFINITY; li[q] = IDX_NONE; }

    const long long row0 = (long long)blockIdx.x * rows_per_block;
    long long row1 = row0 + rows_per_block;
    if (row1 > n) row1 = n;
    const int d4n = dim / 4;
    // two rows per warp per iteration: the query chunk read from smem is used twice
    for (long long r = row0 + 2 * warp; r < row1; r += 2 * KNN_WARPS) {
        const bool has2 = (r + 1 < row1);
        const float4* a4 = reinterpret_cast<const float4*>(db + r * dim);
        const float4* b4 = reinterpret_cast<const float4*>(db + (has2 ? r + 1 : r) * dim);
        float acc0[QT], acc1[QT];
#pragma unroll
        for (int q = 0; q < QT; ++q) { acc0[q] = 0.f; acc1[q] = 0.f; }
        for (int e = lane; e < d4n; e += 32) {
            const float4 a = __ldg(a4 + e), b = __ldg(b4 + e);
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(qs + q * dim + e * 4);
                acc0[q] = fmaf(a.x, w.x, acc0[q]); acc0[q] = fmaf(a.y, w.y, acc0[q]);
                acc0[q] = fmaf(a.z, w.z, acc0[q]); acc0[q] = fmaf(a.w, w.w, acc0[q]);
                acc1[q] = fmaf(b.x, w.x, acc1[q]); acc1[q] = fmaf(b.y, w.y, acc1[q]);
                acc1[q] = fmaf(b.z, w.z, acc1[q]); acc1[q] = fmaf(b.w, w.w, acc1[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            if (q < nq) {
                const float s0 = rg_warp_sum(acc0[q]);
                const float s1 = rg_warp_sum(acc1[q]);
                list_insert(ls[q], li[q], s0, r, k, lane);
                if (has2) list_insert(ls[q], li[q], s1, r + 1, k, lane);
            }
        }
    }
    // block merge: every warp publishes its lists, warp w merges query q = w, w + WARPS, ...
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        ms[(warp * QT + q) * 32 + lane] = ls[q];
        mi[(warp * QT + q) * 32 + lane] = li[q];
    }
    __syncthreads();
    for (int q = warp; q < nq; q += KNN_WARPS) {
        float bs = -INFINITY;
        long long bi = IDX_NONE;
        for (int w = 0; w < KNN_WARPS; ++w)
            for (int j = 0; j < k; ++j) {
                const long long id = mi[(w * QT + q) * 32 + j];
                if (id != IDX_NONE) list_insert(bs, bi, ms[(w * QT + q) * 32 + j], id, k, lane);
            }
        if (lane < k) {
            const long long o = ((long long)blockIdx.x * q_total + (q0 + q)) * k + lane;
            part_idx[o] = bi;
            part_score[o] = bs;
        }
    }
}

// ---- fast path: dim == 768 (the embedding width of the path) ------------------------------------------
// Budget at HBM rate: 133 SM-cycles per 3 KB row.  Each warp takes R = 4 rows per iteration and issues all
// 24 128-bit loads first (12 KB in flight per warp); the query chunk read from shared memory then serves 4
// rows (LDS traffic /4), the R*QT partial sums are reduced with a transposing butterfly (31 shuffles instead
// of 160), and the owning lane tests against the replicated k-th best before any cooperative insert.
template <int NV>
__device__ __forceinline__ float reduce_transpose(float (&a)[NV], int lane) {
    int n = NV / 2;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n >= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < n) {
                    const float send = upper ? a[i] : a[i + n];
                    const float keep = upper ? a[i + n] : a[i];
                    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            n >>= 1;
        } else {
            a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
        }
    }
    return a[0];
}

template <int QT>
__global__ void __launch_bounds__(256, 1) knn_scan768_kernel(const float* __restrict__ db, long long n,
                                                            const float* __restrict__ queries, int q_total,
                                                            int q0, int k, long long rows_per_block,
                                                            long long* __restrict__ part_idx,
                                                            float* __restrict__ part_score) {
    constexpr int DIM = 768, NCH = DIM / 128, R = 4, NV = R * QT;
    constexpr int SHIFT = (NV == 32) ? 0 : (NV == 16) ? 1 : (NV == 8) ? 2 : (NV == 4) ? 3 : 4;   // pair = lane >> SHIFT
    extern __shared__ __align__(16) float smem[];
    float* qs = smem;                                                   // [QT][768]
    float* ms = smem + QT * DIM;                                        // [WARPS][QT][32]
    long long* mi = reinterpret_cast<long long*>(ms + KNN_WARPS * QT * 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = (q_total - q0 < QT) ? q_total - q0 : QT;
    for (int e = threadIdx.x; e < QT * DIM; e += blockDim.x)
        qs[e] = (e / DIM < nq) ? queries[(long long)q0 * DIM + e] : 0.f;
    __syncthreads();

    float ls[QT], ts[QT];
    long long li[QT], ti[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) { ls[q] = -INFINITY; li[q] = IDX_NONE; ts[q] = -INFINITY; ti[q] = IDX_NONE; }

    const long long row0 = (long long)blockIdx.x * rows_per_block;
    long long row1 = row0 + rows_per_block;
    if (row1 > n) row1 = n;
    const int pair = lane >> SHIFT, pi = pair / QT, pq = pair % QT;
    const bool holder = (lane & ((1 << SHIFT) - 1)) == 0;
    for (long long r = row0 + R * warp; r < row1; r += R * KNN_WARPS) {
        float4 x[R][NCH];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const long long rr = (r + i < row1) ? r + i : r;            // clamp: duplicates are discarded below
            const float4* p4 = reinterpret_cast<const float4*>(db + rr * DIM);
#pragma unroll
            for (int j = 0; j < NCH; ++j) x[i][j] = __ldg(p4 + lane + 32 * j);
        }
        float acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(qs + q * DIM + (lane + 32 * j) * 4);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    float a = acc[i * QT + q];
                    a = fmaf(x[i][j].x, w.x, a); a = fmaf(x[i][j].y, w.y, a);
                    a = fmaf(x[i][j].z, w.z, a); a = fmaf(x[i][j].w, w.w, a);
                    acc[i * QT + q] = a;
                }
            }
        }
        const float v = reduce_transpose<NV>(acc, lane);                // lane holds the score of (row pi, query pq)
        const long long id = r + pi;
        float tq = ts[0];
        long long tiq = ti[0];
#pragma unroll
        for (int q = 1; q < QT; ++q) if (pq == q) { tq = ts[q]; tiq = ti[q]; }
        unsigned cand = __ballot_sync(0xffffffffu, holder && pq < nq && id < row1 && better(v, id, tq, tiq));
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const float s = __shfl_sync(0xffffffffu, v, src);
            const int sp = src >> SHIFT, sq = sp % QT;
            const long long sid = r + sp / QT;
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                if (q == sq) {
                    list_insert(ls[q], li[q], s, sid, k, lane);
                    ts[q] = __shfl_sync(0xffffffffu, ls[q], k - 1);
                    ti[q] = __shfl_sync(0xffffffffu, li[q], k - 1);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        ms[(warp * QT + q) * 32 + lane] = ls[q];
        mi[(warp * QT + q) * 32 + lane] = li[q];
    }
    __syncthreads();
    for (int q = warp; q < nq; q += KNN_WARPS) {
        float bs = -INFINITY;
        long long bi = IDX_NONE;
        for (int w = 0; w < KNN_WARPS; ++w)
            for (int j = 0; j < k; ++j) {
                const long long id = mi[(w * QT + q) * 32 + j];
                if (id != IDX_NONE) list_insert(bs, bi, ms[(w * QT + q) * 32 + j], id, k, lane);
            }
        if (lane < k) {
            const long long o = ((long long)blockIdx.x * q_total + (q0 + q)) * k + lane;
            part_idx[o] = bi;
            part_score[o] = bs;
        }
    }
}

template <int QT>
static cudaError_t launch_scan768(const float* db, long long n, const float* queries, int q, int q0, int k,
                                  int blocks, long long rows_per_block, long long* part_idx, float* part_score,
                                  cudaStream_t st) {
    const size_t smem = (size_t)QT * 768 * sizeof(float) + (size_t)KNN_WARPS * QT * 32 * (sizeof(float) + sizeof(long long));
    cudaError_t e = cudaFuncSetAttribute(knn_scan768_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    knn_scan768_kernel<QT><<<blocks, 256, smem, st>>>(db, n, queries, q, q0, k, rows_per_block, part_idx, part_score);
    return cudaGetLastError();
}

// one warp per query: merge parts*k candidates.  Part p's lists start p * part_stride BYTES after the base pointers
// (two separate [parts,Q,k] arrays, or the packed all-gather buffer of parallel.sharded_knn: per rank one block
// [idx int64 Q*k | score fp32 Q*k]).
__global__ void __launch_bounds__(256) knn_merge_kernel(const char* __restrict__ idx_parts,
                                                       const char* __restrict__ score_parts,
                                                       long long idx_stride, long long score_stride,
                                                       int parts, int q_total, int k, long long idx_base,
                                                       long long* __restrict__ out_idx,
                                                       float* __restrict__ out_score) {
    const int q = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= q_total) return;
    float bs = -INFINITY;
    long long bi = IDX_NONE;
    for (int p = 0; p < parts; ++p) {
        const long long* ip = reinterpret_cast<const long long*>(idx_parts + p * idx_stride) + (long long)q * k;
        const float* sp = reinterpret_cast<const float*>(score_parts + p * score_stride) + (long long)q * k;
        // read k candidates at once, then insert one by one (warp-uniform loop)
        const long long cid = lane < k ? ip[lane] : IDX_NONE;
        const float cs = lane < k ? sp[lane] : -INFINITY;
        for (int j = 0; j < k; ++j) {
            const long long id = __shfl_sync(0xffffffffu, cid, j);
            const float s = __shfl_sync(0xffffffffu, cs, j);
            if (id != IDX_NONE && id >= 0) list_insert(bs, bi, s, id, k, lane);
        }
    }
    if (lane < k) {
        out_idx[(long long)q * k + lane] = (bi == IDX_NONE) ? -1 : bi + idx_base;
        out_score[(long long)q * k + lane] = bs;
    }
}

}  // namespace

extern "C" int rg_text_similarity(const float* db, const int32_t* db_len, int64_t n, int max_len,
                                  int dim, const float* query, int tq, const int32_t* subset,
                                  int64_t n_sub, float* scores, void* stream) {
    if (!db || !db_len || !query || !scores) return rg_fail("rg_text_similarity: null argument");
    if (dim % 4) return rg_fail("rg_text_similarity: dim must be a multiple of 4");
    const long long n_out = subset ? n_sub : n;
    if (n_out <= 0) return 0;
    text_similarity_kernel<<<(unsigned)((n_out + KNN_WARPS - 1) / KNN_WARPS), 256, 0, (cudaStream_t)stream>>>(
        db, db_len, max_len, dim, query, tq, subset, n_out, scores);
    RG_CU(cudaGetLastError());
    rg_count_launch(1);
    return 0;
}

extern "C" int rg_knn_merge(const int64_t* idx_parts, const float* score_parts, int parts, int q,
                            int k, int64_t* out_idx, float* out_score, void* stream) {
    if (k < 1 || k > 32) return rg_fail("rg_knn_merge: k must be in [1,32]");
    if (q <= 0) return 0;
    knn_merge_kernel<<<(q + KNN_WARPS - 1) / KNN_WARPS, 256, 0, (cudaStream_t)stream>>>(
        (const char*)idx_parts, (const char*)score_parts, (long long)q * k * 8, (long long)q * k * 4, parts, q, k, 0,
        (long long*)out_idx, out_score);
    RG_CU(cudaGetLastError());
    rg_count_launch(1);
    return 0;
}

extern "C" int rg_knn_merge_packed(const void* packed, int64_t part_stride_bytes, int parts, int q, int k,
                                   int64_t* out_idx, float* out_score, void* stream) {
    if (k < 1 || k > 32) return rg_fail("rg_knn_merge_packed: k must be in [1,32]");
    if (!packed || !out_idx || !out_score) return rg_fail("rg_knn_merge_packed: null argument");
    if (part_stride_bytes < (int64_t)q * k * 12 || part_stride_bytes % 8)
        return rg_fail("rg_knn_merge_packed: part stride %lld too small or not a multiple of 8", (long long)part_stride_bytes);
    if (q <= 0) return 0;
    const char* base = (const char*)packed;
    knn_merge_kernel<<<(q + KNN_WARPS - 1) / KNN_WARPS, 256, 0, (cudaStream_t)stream>>>(
        base, base + (long long)q * k * 8, part_stride_bytes, part_stride_bytes, parts, q, k, 0, (long long*)out_idx,
        out_score);
    RG_CU(cudaGetLastError());
    rg_count_launch(1);
    return 0;
}

extern "C" int rg_knn_topk(const float* db, int64_t n, int dim, const float* queries, int q, int k,
                           int64_t idx_base, int64_t* out_idx, float* out_score, void* stream) {
    if (!db || !queries || !out_idx || !out_score) return rg_fail("rg_knn_topk: null argument");
    if (k < 1 || k > 32) return rg_fail("rg_knn_topk: k must be in [1,32]");
    if (dim % 4 || dim > 4096) return rg_fail("rg_knn_topk: dim must be a multiple of 4 and <= 4096");
    if (q <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    rg_keep_mempool();
    // grid: a multiple of the SM count; each block scans a contiguous slab of the shard
    const bool fast = (dim == 768);
    int blocks = fast ? 148 : 148 * 4;
    const int rows_per_iter = (fast ? 4 : 2) * KNN_WARPS;
    if (n < (long long)blocks * rows_per_iter) blocks = (int)((n + rows_per_iter - 1) / rows_per_iter);
    if (blocks < 1) blocks = 1;
    const long long rows_per_block = (n + blocks - 1) / blocks;
    struct AsyncScratch {        // stream-ordered scratch, released on every exit path
        cudaStream_t st; void* p = nullptr;
        explicit AsyncScratch(cudaStream_t s) : st(s) {}
        ~AsyncScratch() { if (p) cudaFreeAsync(p, st); }
    } s_idx(st), s_score(st);
    RG_CU(cudaMallocAsync(&s_idx.p, (size_t)blocks * q * k * sizeof(long long), st));
    RG_CU(cudaMallocAsync(&s_score.p, (size_t)blocks * q * k * sizeof(float), st));
    long long* part_idx = static_cast<long long*>(s_idx.p);
    float* part_score = static_cast<float*>(s_score.p);
    const size_t smem = (size_t)KNN_QT * dim * sizeof(float) + (size_t)KNN_WARPS * KNN_QT * 32 * (sizeof(float) + sizeof(long long));
    if (!fast) RG_CU(cudaFuncSetAttribute(knn_scan_kernel<KNN_QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int q0 = 0; q0 < q;) {
        const int left = q - q0;
        if (fast) {
            // widest tile that is not mostly padding: 8 queries per pass, 4 / 2 / 1 for the tail
            if (left >= 5) { RG_CU(launch_scan768<8>(db, n, queries, q, q0, k, blocks, rows_per_block, part_idx, part_score, st)); q0 += 8; }
            else if (left >= 3) { RG_CU(launch_scan768<4>(db, n, queries, q, q0, k, blocks, rows_per_block, part_idx, part_score, st)); q0 += 4; }
            else if (left == 2) { RG_CU(launch_scan768<2>(db, n, queries, q, q0, k, blocks, rows_per_block, part_idx, part_score, st)); q0 += 2; }
            else { RG_CU(launch_scan768<1>(db, n, queries, q, q0, k, blocks, rows_per_block, part_idx, part_score, st)); q0 += 1; }
        } else {
            knn_scan_kernel<KNN_QT><<<blocks, 256, smem, st>>>(db, n, dim, queries, q, q0, k, rows_per_block,
                                                               part_idx, part_score);
            RG_CU(cudaGetLastError());
            q0 += KNN_QT;
        }
        rg_count_launch(1);
    }
    knn_merge_kernel<<<(q + KNN_WARPS - 1) / KNN_WARPS, 256, 0, st>>>(
        (const char*)part_idx, (const char*)part_score, (long long)q * k * 8, (long long)q * k * 4, blocks, q, k, idx_base,
        (long long*)out_idx, out_score);
    RG_CU(cudaGetLastError());
    rg_count_launch(1);
    return 0;
}

extern "C" int rg_probe_knn_scan(const float* db, int64_t n, int dim, const float* queries, int q, int k, int reps,
                                 void* flush_buf, int64_t flush_bytes, float* median_ms, void* stream) {
    if (dim != 768 || q < 1 || q > 8 || k < 1 || k > 32 || reps < 1 || reps > 64)
        return rg_fail("rg_probe_knn_scan: needs dim 768, 1 <= q <= 8, 1 <= k <= 32");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = 148;
    const long long rows_per_block = (n + blocks - 1) / blocks;
    long long* part_idx = nullptr;
    float* part_score = nullptr;
    RG_CU(cudaMalloc((void**)&part_idx, (size_t)blocks * q * k * sizeof(long long)));
    RG_CU(cudaMalloc((void**)&part_score, (size_t)blocks * q * k * sizeof(float)));
    cudaEvent_t e0, e1;
    RG_CU(cudaEventCreate(&e0));
    RG_CU(cudaEventCreate(&e1));
    std::vector<float> ts;
    for (int i = 0; i < reps + 2; ++i) {
        if (flush_buf) RG_CU(cudaMemsetAsync(flush_buf, i, (size_t)flush_bytes, st));
        RG_CU(cudaEventRecord(e0, st));
        if (q >= 5) RG_CU(launch_scan768<8>(db, n, queries, q, 0, k, blocks, rows_per_block, part_idx, part_score, st));
        else if (q >= 3) RG_CU(launch_scan768<4>(db, n, queries, q, 0, k, blocks, rows_per_block, part_idx, part_score, st));
        else if (q == 2) RG_CU(launch_scan768<2>(db, n, queries, q, 0, k, blocks, rows_per_block, part_idx, part_score, st));
        else RG_CU(launch_scan768<1>(db, n, queries, q, 0, k, blocks, rows_per_block, part_idx, part_score, st));
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
    cudaFree(part_idx); cudaFree(part_score);
    return 0;
}
