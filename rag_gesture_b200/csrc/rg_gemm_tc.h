// tcgen05 GEMM launcher interface (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

struct RgGemmTc {
    int M, N, K;              // K: reduction length of ONE operand plane
    int split;                // 0: bf16 x bf16;  1: three passes hi*hi + lo*hi + hi*lo (bf16x3)
    int a_lo_off, w_lo_off;   // column (K) offset of the lo plane inside A / W rows (split mode)
    int groups;               // blockIdx.z: A K-offset += a_goff, W row += w_goff, bias += b_goff, C col += c_goff
    int a_goff, w_goff, b_goff, c_goff, r_grouped;
    const float* bias;
    const float* R; int ldr;  // fp32 residual (RG_EPI_BIAS_RESIDUAL)
    const float* pos; int pos_T;
    float* C32; int ldc32;    // optional fp32 output
    void* C16_; int ldc16;    // optional bf16 output (+ lo plane at column c16_lo_off when non-zero)
    int c16_lo_off;
    int epi;                  // RgEpilogue
    int no_pdl;               // 1: W was produced by the preceding kernel -> plain (fully serialised) launch
    long long* trace;         // diagnostics: 10 clock64/globaltimer stamps per CTA (rg_probe_gemm_trace), else null
    // 2-CTA persistent kernel (gemm2_tc.cu): TMA-store maps of the outputs in use (rg_make_store_map; host pointers,
    // copied into the launch) and the first column of this GEMM's output inside each map
    const CUtensorMap* tmC32; const CUtensorMap* tmC16;
    int c32_col0, c16_col0;
    // pair128 kernel (gemm_pair128_kernel): the weight operand's map with a 64-row box (each CTA of a pair stages half
    // of a 128-row weight tile); null = not available for this weight
    const CUtensorMap* tmW64;
};

cudaError_t rg_make_tensor_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld,
                               int box_rows);
// Dispatch: the persistent 2-CTA kernel (gemm2_tc.cu) when the launch is eligible (store maps given, N % 256 == 0, one
// group) and large enough (rg_gemm_kernel_mode), else the 128x128 one-tile-per-CTA kernel (gemm_tc.cu).
cudaError_t rg_launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, cudaStream_t st);
cudaError_t rg_launch_gemm1_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, cudaStream_t st);
cudaError_t rg_launch_gemm2_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const RgGemmTc& p, cudaStream_t st);
bool rg_gemm2_eligible(const RgGemmTc& p);
cudaError_t rg_launch_gemm_pair128(const CUtensorMap& tmA, const RgGemmTc& p, cudaStream_t st);
bool rg_gemm_pair128_eligible(const RgGemmTc& p);
extern int rg_pair128_min_rows;   // automatic choice: launches of at least this many rows take the pair128 kernel
cudaError_t rg_make_store_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int elem_bytes);
// 0: automatic, 1: always the 128x128 kernel, 2: the 2-CTA 256x256 kernel whenever eligible, 3: the pair128 kernel whenever eligible
extern int rg_gemm_kernel_mode;
extern int rg_gemm2_min_rows;
extern int rg_gemm2_persist_tiles;   // 2-CTA kernel: from this many pair tiles on the persistent variant runs
// pdl = false: plain launch, i.e. the kernel starts only after everything enqueued before it has completed
cudaError_t rg_launch_split_bf16(const float* x, int ldx, void* out, int ldo, int lo_off, long long rows, int cols,
                                 cudaStream_t st, bool pdl = true);
