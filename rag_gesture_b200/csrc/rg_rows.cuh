// Warp-per-row helpers shared by rowops.cu and attention.cu (header-only: no -rdc needed).
#pragma once
#include "rg_common.cuh"

__device__ __forceinline__ void load_row(const float* p, int lane, float4 v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(p + (lane + 32 * j) * 4);
}
__device__ __forceinline__ void store_row(float* p, int lane, const float4 v[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(p + (lane + 32 * j) * 4) = v[j];
}

// one 512-wide row (starting at column col0 of row `row`) to an RgRowOut destination
__device__ __forceinline__ void rg_store_row_out(const RgRowOut& o, long long row, int col0, int lane,
                                                 const float4 v[4]) {
    if (o.f32) {
        store_row(o.f32 + row * o.ld + col0, lane, v);
        return;
    }
    __nv_bfloat16* base = o.b16 + row * o.ld + col0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float f[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            h[k] = __float2bfloat16_rn(f[k]);
            l[k] = __float2bfloat16_rn(f[k] - __bfloat162float(h[k]));
        }
        *reinterpret_cast<uint2*>(base + (lane + 32 * j) * 4) = *reinterpret_cast<uint2*>(h);
        if (o.lo_off) *reinterpret_cast<uint2*>(base + o.lo_off + (lane + 32 * j) * 4) = *reinterpret_cast<uint2*>(l);
    }
}

// two-pass LayerNorm statistics over a 512-wide row spread across the warp (eps = 1e-5)
__device__ __forceinline__ void rg_ln_normalize(float4 v[4]) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = rg_warp_sum(s) * (1.0f / RG_D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = 1.0f / sqrtf(rg_warp_sum(q) * (1.0f / RG_D) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd; }
}

__device__ __forceinline__ float4 affine4(float4 v, float4 g, float4 b) {
    return make_float4(v.x * g.x + b.x, v.y * g.y + b.y, v.z * g.z + b.z, v.w * g.w + b.w);
}

// Stylization parameters of one block for this lane's 16 columns, kept in registers across rows
struct RgStylRegs { float4 g[4], b[4], sc[4], sh[4]; };
__device__ __forceinline__ void rg_styl_load(RgStylRegs& r, const RgStylParams& sp, int clip, int lane) {
    const float* ss = sp.ss + (long long)clip * sp.ss_clip_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        r.g[j] = __ldg(reinterpret_cast<const float4*>(sp.gamma) + lane + 32 * j);
        r.b[j] = __ldg(reinterpret_cast<const float4*>(sp.beta) + lane + 32 * j);
        r.sc[j] = __ldg(reinterpret_cast<const float4*>(ss) + lane + 32 * j);
        r.sh[j] = __ldg(reinterpret_cast<const float4*>(ss + RG_D) + lane + 32 * j);
    }
}
__device__ __forceinline__ void rg_styl_apply(float4 v[4], const RgStylRegs& r) {
    rg_ln_normalize(v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 h = affine4(v[j], r.g[j], r.b[j]);
        h.x = rg_silu(h.x * (1.0f + r.sc[j].x) + r.sh[j].x);
        h.y = rg_silu(h.y * (1.0f + r.sc[j].y) + r.sh[j].y);
        h.z = rg_silu(h.z * (1.0f + r.sc[j].z) + r.sh[j].z);
        h.w = rg_silu(h.w * (1.0f + r.sc[j].w) + r.sh[j].w);
        v[j] = h;
    }
}

// shared with attention.cu: LN -> affine -> *(1+scale)+shift -> SiLU on a row held across a warp
__device__ __forceinline__ void rg_styl_row(float4 v[4], const RgStylParams& sp, int clip, int lane) {
    rg_ln_normalize(v);
    const float* ss = sp.ss + (long long)clip * sp.ss_clip_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(sp.gamma) + lane + 32 * j);
        const float4 b = __ldg(reinterpret_cast<const float4*>(sp.beta) + lane + 32 * j);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(ss) + lane + 32 * j);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(ss + RG_D) + lane + 32 * j);
        float4 h = affine4(v[j], g, b);
        h.x = rg_silu(h.x * (1.0f + sc.x) + sh.x);
        h.y = rg_silu(h.y * (1.0f + sc.y) + sh.y);
        h.z = rg_silu(h.z * (1.0f + sc.z) + sh.z);
        h.w = rg_silu(h.w * (1.0f + sc.w) + sh.w);
        v[j] = h;
    }
}

