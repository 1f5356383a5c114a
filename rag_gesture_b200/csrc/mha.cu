// Softmax multi-head attention for the body-part TransformerVAEs of the latent codec (SURVEY 8f.1;
// mogen/models/utils/detr_utils.py MultiheadAttention calls of the encoder / decoder blocks): short sequences
// (17 tokens per encoder chunk, 160 per decoder clip), wide heads (512/4 = 128), fp32, key-padding mask.
// The library attention kernels that PyTorch picks for fp32 here (mem-efficient sm80 tiles of 64x128) spend
// 250-390 us per call on these shapes; this one is sized for them:
//   grid (N, H, ceil(Sq/32)), 8 warps; a CTA holds 32 query rows of one head in shared memory and walks the keys
//   in tiles of 32 (K padded to DH+4 floats per row: conflict-free 128-bit reads with one key per lane);
//   warp w owns rows w, w+8, w+16, w+24: scores with one key per lane (no shuffles in the dot products), online
//   softmax across the lanes, P.V with one output column group per lane and the probabilities broadcast by shuffle.
// fp32 FMA throughout (the tensor-core tiers keep the codec fp32-class); exp through expf.
#include "rg_common.cuh"

namespace {

constexpr int MHA_QT = 32, MHA_KT = 32, MHA_WARPS = 8, MHA_RPW = MHA_QT / MHA_WARPS;

template <int DH>    // head dim: 16, 32, 64 or 128
__global__ void __launch_bounds__(MHA_WARPS * 32)
mha_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
           const unsigned char* __restrict__ keep, float* __restrict__ out, int Sq, int Sk, int H,
           long long ldq, long long ldk, long long ldv, float scale) {
    constexpr int DPL = DH >= 32 ? DH / 32 : 1, KS = DH + 4;
    extern __shared__ __align__(16) float smem[];
    float* sq = smem;                        // [MHA_QT][DH], pre-scaled
    float* sk = sq + MHA_QT * DH;            // [MHA_KT][KS]
    float* sv = sk + MHA_KT * KS;            // [MHA_KT][DH]
    const int n = blockIdx.x, h = blockIdx.y, q0 = blockIdx.z * MHA_QT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* qb = q + ((long long)n * Sq) * ldq + h * DH;
    const float* kb = k + ((long long)n * Sk) * ldk + h * DH;
    const float* vb = v + ((long long)n * Sk) * ldv + h * DH;

    for (int i = tid; i < MHA_QT * (DH / 4); i += MHA_WARPS * 32) {
        const int r = i / (DH / 4), c = i % (DH / 4);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < Sq) x = *reinterpret_cast<const float4*>(qb + (long long)(q0 + r) * ldq + 4 * c);
        x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
        *reinterpret_cast<float4*>(sq + r * DH + 4 * c) = x;
    }
    float m[MHA_RPW], l[MHA_RPW], acc[MHA_RPW][DPL];
#pragma unroll
    for (int r = 0; r < MHA_RPW; ++r) {
        m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
        for (int c = 0; c < DPL; ++c) acc[r][c] = 0.f;
    }
    const int rows_here = min(MHA_QT, Sq - q0);
    for (int kt = 0; kt < Sk; kt += MHA_KT) {
        __syncthreads();                      // previous tile consumed (and sq written, first time round)
        for (int i = tid; i < MHA_KT * (DH / 4); i += MHA_WARPS * 32) {
            const int r = i / (DH / 4), c = i % (DH / 4);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (kt + r < Sk) {
                a = *reinterpret_cast<const float4*>(kb + (long long)(kt + r) * ldk + 4 * c);
                b = *reinterpret_cast<const float4*>(vb + (long long)(kt + r) * ldv + 4 * c);
            }
            *reinterpret_cast<float4*>(sk + r * KS + 4 * c) = a;
            *reinterpret_cast<float4*>(sv + r * DH + 4 * c) = b;
        }
        __syncthreads();
        const int j = kt + lane;
        const bool valid = j < Sk && (keep == nullptr || keep[(long long)n * Sk + j] != 0);
        // scores of this lane's key against the warp's rows
        float s[MHA_RPW];
#pragma unroll
        for (int r = 0; r < MHA_RPW; ++r) s[r] = 0.f;
        const float* kr = sk + lane * KS;
#pragma unroll 4
        for (int d = 0; d < DH; d += 4) {
            const float4 kk = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
            for (int r = 0; r < MHA_RPW; ++r) {
                const float4 qq = *reinterpret_cast<const float4*>(sq + (warp + MHA_WARPS * r) * DH + d);
                s[r] = fmaf(qq.x, kk.x, s[r]); s[r] = fmaf(qq.y, kk.y, s[r]);
                s[r] = fmaf(qq.z, kk.z, s[r]); s[r] = fmaf(qq.w, kk.w, s[r]);
            }
        }
        float p[MHA_RPW];
#pragma unroll
        for (int r = 0; r < MHA_RPW; ++r) {
            const float sc = valid ? s[r] : -INFINITY;
            const float m_new = fmaxf(m[r], rg_warp_max(sc));
            const float corr = (m[r] == -INFINITY) ? 0.f : expf(m[r] - m_new);
            p[r] = (sc == -INFINITY) ? 0.f : expf(sc - m_new);
            l[r] = l[r] * corr + rg_warp_sum(p[r]);
            m[r] = m_new;
#pragma unroll
            for (int c = 0; c < DPL; ++c) acc[r][c] *= corr;
        }
        const int n_keys = min(MHA_KT, Sk - kt);
        if constexpr (DH >= 32) {
            for (int jj = 0; jj < n_keys; ++jj) {
                float vv[DPL];
#pragma unroll
                for (int c = 0; c < DPL; ++c) vv[c] = sv[jj * DH + lane + 32 * c];
#pragma unroll
                for (int r = 0; r < MHA_RPW; ++r) {
                    const float pj = __shfl_sync(0xffffffffu, p[r], jj);
#pragma unroll
                    for (int c = 0; c < DPL; ++c) acc[r][c] = fmaf(pj, vv[c], acc[r][c]);
                }
            }
        } else {
            // 16 columns: the half-warps take the even and the odd keys of the tile, summed at the end
            const int col = lane & 15, par = lane >> 4;
            for (int jj = 0; jj < n_keys; jj += 2) {
                const int key = jj + par;                 // sv rows beyond Sk are zero-filled, p is 0 there
                const float vv = sv[key * DH + col];
#pragma unroll
                for (int r = 0; r < MHA_RPW; ++r)
                    acc[r][0] = fmaf(__shfl_sync(0xffffffffu, p[r], key), vv, acc[r][0]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MHA_RPW; ++r) {
        if constexpr (DH < 32) acc[r][0] += __shfl_xor_sync(0xffffffffu, acc[r][0], 16);
        const int row = warp + MHA_WARPS * r;
        if (row >= rows_here) continue;
        const float inv = 1.0f / l[r];        // a row whose keys are all masked gives NaN, as softmax over -inf does
        float* o = out + ((long long)n * Sq + q0 + row) * ((long long)H * DH) + h * DH;
        if constexpr (DH >= 32) {
#pragma unroll
            for (int c = 0; c < DPL; ++c) o[lane + 32 * c] = acc[r][c] * inv;
        } else if (lane < DH) {
            o[lane] = acc[r][0] * inv;
        }
    }
}

// Narrow heads (dh = 16: the decoder's 32 heads): one query row per LANE instead of per warp.  With 16 columns the
// warp-per-row form above spends its time in shuffles and half-empty reductions (318 us for 64 clips x 32 heads x
// 160^2, ncu); here a thread keeps its row's q, running max / sum and 16 accumulators in registers, keys and values
// are broadcast from shared memory (every lane reads the same key: conflict-free), and the online softmax advances
// 8 keys at a time -- no shuffles at all.  grid (N, H, ceil(Sq/64)), 64 threads.
constexpr int RL_ROWS = 64, RL_KT = 64, RL_CH = 8;

template <int DH>
__global__ void __launch_bounds__(RL_ROWS)
mha_rowlane_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                   const unsigned char* __restrict__ keep, float* __restrict__ out, int Sq, int Sk, int H,
                   long long ldq, long long ldk, long long ldv, float scale) {
    __shared__ __align__(16) float sk[RL_KT][DH];
    __shared__ __align__(16) float sv[RL_KT][DH];
    __shared__ float sbias[RL_KT];            // 0 for an attended key, -inf for a masked or absent one
    const int n = blockIdx.x, h = blockIdx.y, row = blockIdx.z * RL_ROWS + threadIdx.x, t = threadIdx.x;
    const bool active = row < Sq;
    const float* kb = k + ((long long)n * Sk) * ldk + h * DH;
    const float* vb = v + ((long long)n * Sk) * ldv + h * DH;
    float qr[DH], acc[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { qr[d] = 0.f; acc[d] = 0.f; }
    if (active) {
        const float* qp = q + ((long long)n * Sq + row) * ldq + h * DH;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
            const float4 x = *reinterpret_cast<const float4*>(qp + d);
            qr[d] = x.x * scale; qr[d + 1] = x.y * scale; qr[d + 2] = x.z * scale; qr[d + 3] = x.w * scale;
        }
    }
    float m = -INFINITY, l = 0.f;
    for (int kt = 0; kt < Sk; kt += RL_KT) {
        __syncthreads();
        {
            const int j = kt + t;
            const bool have = j < Sk;
#pragma unroll
            for (int d = 0; d < DH; d += 4) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (have) {
                    a = *reinterpret_cast<const float4*>(kb + (long long)j * ldk + d);
                    b = *reinterpret_cast<const float4*>(vb + (long long)j * ldv + d);
                }
                *reinterpret_cast<float4*>(&sk[t][d]) = a;
                *reinterpret_cast<float4*>(&sv[t][d]) = b;
            }
            sbias[t] = (have && (keep == nullptr || keep[(long long)n * Sk + j] != 0)) ? 0.f : -INFINITY;
        }
        __syncthreads();
        const int nk = min(RL_KT, Sk - kt);
        for (int j0 = 0; j0 < nk; j0 += RL_CH) {
            float sc[RL_CH], cmax = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < RL_CH; ++jj) {
                float a = sbias[j0 + jj];
#pragma unroll
                for (int d = 0; d < DH; d += 4) {
                    const float4 kk = *reinterpret_cast<const float4*>(&sk[j0 + jj][d]);
                    a = fmaf(qr[d], kk.x, a); a = fmaf(qr[d + 1], kk.y, a);
                    a = fmaf(qr[d + 2], kk.z, a); a = fmaf(qr[d + 3], kk.w, a);
                }
                sc[jj] = a;
                cmax = fmaxf(cmax, a);
            }
            const float m_new = fmaxf(m, cmax);
            if (m_new == -INFINITY) continue;             // nothing attended so far
            const float corr = (m == -INFINITY) ? 0.f : expf(m - m_new);
            l *= corr;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[d] *= corr;
#pragma unroll
            for (int jj = 0; jj < RL_CH; ++jj) {
                const float p = (sc[jj] == -INFINITY) ? 0.f : expf(sc[jj] - m_new);
                l += p;
#pragma unroll
                for (int d = 0; d < DH; d += 4) {
                    const float4 vv = *reinterpret_cast<const float4*>(&sv[j0 + jj][d]);
                    acc[d] = fmaf(p, vv.x, acc[d]); acc[d + 1] = fmaf(p, vv.y, acc[d + 1]);
                    acc[d + 2] = fmaf(p, vv.z, acc[d + 2]); acc[d + 3] = fmaf(p, vv.w, acc[d + 3]);
                }
            }
            m = m_new;
        }
    }
    if (active) {
        const float inv = 1.0f / l;
        float* o = out + ((long long)n * Sq + row) * ((long long)H * DH) + h * DH;
#pragma unroll
        for (int d = 0; d < DH; d += 4)
            *reinterpret_cast<float4*>(o + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
    }
}

template <int DH>
cudaError_t launch(const float* q, const float* k, const float* v, const unsigned char* keep, float* out, int N,
                   int Sq, int Sk, int H, long long ldq, long long ldk, long long ldv, float scale, cudaStream_t st) {
    const size_t smem = sizeof(float) * (MHA_QT * DH + MHA_KT * (DH + 4) + MHA_KT * DH);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && !attr_set[dev]) {
        e = cudaFuncSetAttribute(mha_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    dim3 grid(N, H, (Sq + MHA_QT - 1) / MHA_QT);
    mha_kernel<DH><<<grid, MHA_WARPS * 32, smem, st>>>(q, k, v, keep, out, Sq, Sk, H, ldq, ldk, ldv, scale);
    return cudaGetLastError();
}

}  // namespace

// q [N, Sq, >=H*dh] rows ldq floats apart (head h at column h*dh), likewise k, v [N, Sk, .]; keep [N, Sk] bytes
// (non-zero = attend) or NULL; out [N, Sq, H*dh] contiguous.  dh in {16, 32, 64, 128}; 16-byte aligned rows.
cudaError_t rg_launch_mha(const float* q, const float* k, const float* v, const unsigned char* keep, float* out,
                          int N, int Sq, int Sk, int H, int dh, long long ldq, long long ldk, long long ldv,
                          float scale, cudaStream_t st) {
    if (N <= 0 || Sq <= 0 || Sk <= 0) return cudaSuccess;
    if (H > 65535 || (Sq + MHA_QT - 1) / MHA_QT > 65535) return cudaErrorInvalidValue;
    switch (dh) {
        case 16: {
            if ((Sq + RL_ROWS - 1) / RL_ROWS > 65535) return cudaErrorInvalidValue;
            dim3 grid(N, H, (Sq + RL_ROWS - 1) / RL_ROWS);
            mha_rowlane_kernel<16><<<grid, RL_ROWS, 0, st>>>(q, k, v, keep, out, Sq, Sk, H, ldq, ldk, ldv, scale);
            return cudaGetLastError();
        }
        case 32: return launch<32>(q, k, v, keep, out, N, Sq, Sk, H, ldq, ldk, ldv, scale, st);
        case 64: return launch<64>(q, k, v, keep, out, N, Sq, Sk, H, ldq, ldk, ldv, scale, st);
        case 128: return launch<128>(q, k, v, keep, out, N, Sq, Sk, H, ldq, ldk, ldv, scale, st);
        default: return cudaErrorInvalidValue;
    }
}
