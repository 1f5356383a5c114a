// C ABI of the hot path (include/rg_b200.h): handle management, weight packing, the timestep
// table (K7), the per-clip cross-attention state (K6) and the per-step denoiser orchestration.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/rg_b200.h"
#include "rg_common.cuh"
#include "rg_internal.h"
#include "rg_gemm_tc.h"

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};     // the pipeline's worker thread launches concurrently

int rg_fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
void rg_count_launch(int n) { g_launches += n; }
void rg_keep_mempool() {
    static std::atomic<unsigned long long> done_mask{0};      // one bit per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || (done_mask.load() >> dev & 1)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_mask.fetch_or(1ull << dev);
}

#define CU(expr)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return rg_fail("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)
#define LAUNCH(expr)      \
    do {                  \
        CU(expr);         \
        ++g_launches;     \
    } while (0)

struct Layer {
    float *Wqkv, *bqkv;                   // [1536,512] sa q|k|v with sa_block.norm folded in
    float *sa_g, *sa_b, *sa_Wo, *sa_bo;   // sa_block.proj_out.{norm, out_layers.2}
    float *Wcaq, *bcaq;                   // [1536,512] 3 x ca query with ca norm folded in
    float *ca_g, *ca_b;                   // [3,512] ca proj_out.norm
    float *ca_Wo, *ca_bo;                 // [3,512,512], [3,512] ca proj_out.out_layers.2
    float *Wmix, *bmix;                   // [512,1536]
    float *W1, *b1, *W2, *b2;             // ffn
    float *ffn_g, *ffn_b, *ffn_Wo, *ffn_bo;
};

struct W16 {                 // bf16 operand planes [N, planes*K] (hi | lo) + its TMA descriptors
    void* ptr;
    CUtensorMap tm;          // box 64 x 128 rows
    CUtensorMap tm64;        // box 64 x 64 rows: half a weight tile per CTA of a pair (gemm_pair128_kernel)
    int N, K;
};
struct LayerTc { W16 qkv, sa_o, caq, fold, w1, w2, ffn_o; float* b_fold; };

#define RG_MAX_LANES 4
struct Ws {                  // one lane's activation workspace (rows = clips * n_tokens of that lane)
    long long rows;
    float *h, *a, *big, *o3, *g, *y;
    void *x16, *a16, *a16x, *h16, *g16;
    CUtensorMap tm_x16, tm_a16, tm_a16x, tm_h16, tm_g16;
    // TMA-store maps of the GEMM outputs (2-CTA persistent kernel, gemm2_tc.cu)
    CUtensorMap ts_h, ts_big, ts_y, ts_h16, ts_a16x, ts_g16;
};

struct rg_model {
    rg_config cfg;
    // tensor-core path (cfg.precision != RG_PREC_FP32)
    int planes;              // 1: bf16 operands, 2: hi|lo planes (bf16x3)
    int attn_mode_ca;
    int gemm_only;           // measurement probe (rg_probe_gemm_only): rg_denoise launches its GEMMs only
    int fuse_styl_max;       // largest batch (clips) that takes the attention kernels fused with the Stylization prologue
    int attn_mode;           // attention cores: 0 fp32 SIMT, 1 TF32 mma.sync (bf16 tier), 2 3xTF32 (bf16x3 tier)
    std::vector<LayerTc> tc;
    W16 tc_joint, tc_out, tc_kv[3], tc_text, tc_audio;
    void* kv_a16;
    int device;
    std::vector<void*> allocs;
    std::vector<Layer> layers;
    float *W_joint, *b_joint, *W_out, *b_out, *pos;
    float *W_text, *b_text, *W_audio, *b_audio, *spk_table;
    float *W_t0, *b_t0, *W_t2, *b_t2;     // time_embed
    float *We_all, *be_all;               // [L*5*1024, E] all emb_layers.1, block order sa,text,audio,spk,ffn
    float* Wkv_all[3];                    // [L*1024, 512] per cond: key|value with text_norm folded in
    float* bkv_all[3];
    // schedule
    int n_steps;
    std::vector<int> timestep_map;
    std::vector<float> coef;
    float* table;                         // [n_steps, L, 5, 1024]
    float* tau_row;                       // [L*5*1024] scratch for an off-table timestep
    int tau_cached;
    float* ss_rep;                        // [ss_rep_clips, L*5*1024] per-clip table rows (rg_denoise_groups)
    long long ss_rep_clips;
    // workspaces: rg_denoise cuts a batch into `lanes` clip ranges that run as concurrent kernel chains
    int lanes;                            // 0: automatic (2 lanes from auto_lane_clips clips on)
    int auto_lane_clips;
    Ws ws[RG_MAX_LANES];
    cudaStream_t lane_st[RG_MAX_LANES];   // [0] unused: lane 0 runs on the caller's stream
    cudaEvent_t ev_fork, ev_join[RG_MAX_LANES];
    long long kv_rows;
    float *kv_ln, *kv_buf;
    long long tt_rows;
    float *tt_emb, *tt_t1, *tt_e;
    // CUDA graphs of the evaluation chain (one per distinct set of buffer addresses; see run_eval)
    int use_graphs;
    float* ss_one;                        // [L*5*1024]: the current level's table row at a fixed address
    cudaStream_t cap_st;                  // capture happens here (the caller's stream may be the legacy stream)
    struct EvalGraph {
        const void *x, *src_mask, *qmask, *state, *x0, *ss;
        long long ss_stride, qm_stride;
        int B, gemm_only, kmode, kmin, kpt, kp128, lanes;
        cudaGraphExec_t exec;
        long long launches;
        unsigned long long last_use;
    };
    std::vector<EvalGraph> graphs;
    unsigned long long graph_clock;
};

static int dalloc(rg_model* m, void** p, size_t bytes) {
    CU(cudaMalloc(p, bytes ? bytes : 4));
    m->allocs.push_back(*p);
    return 0;
}
static void dfree_one(rg_model* m, void* p) {
    if (!p) return;
    for (size_t i = 0; i < m->allocs.size(); ++i)
        if (m->allocs[i] == p) { m->allocs.erase(m->allocs.begin() + i); break; }
    cudaFree(p);
}

struct TensorMap {
    std::map<std::string, std::pair<const float*, long long>> t;
    const float* get(const std::string& k, long long numel) const {
        auto it = t.find(k);
        if (it == t.end()) { rg_fail("missing tensor '%s'", k.c_str()); return nullptr; }
        if (it->second.second != numel) {
            rg_fail("tensor '%s': expected %lld elements, got %lld", k.c_str(), numel, it->second.second);
            return nullptr;
        }
        return it->second.first;
    }
};

// upload `numel` host floats of tensor `key` to dst (device)
static int up_to(const TensorMap& tm, const std::string& key, long long numel, float* dst) {
    const float* src = tm.get(key, numel);
    if (!src) return 1;
    CU(cudaMemcpy(dst, src, (size_t)numel * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
static int up_new(rg_model* m, const TensorMap& tm, const std::string& key, long long numel, float** dst) {
    if (dalloc(m, (void**)dst, (size_t)numel * sizeof(float))) return 1;
    return up_to(tm, key, numel, *dst);
}

static RgGemm mk_gemm(const float* A, int lda, const float* W, const float* bias, float* C, int ldc,
                      int M, int N, int K, int epi) {
    RgGemm g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.W = W; g.bias = bias; g.C = C;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = K; g.ldc = ldc; g.groups = 1; g.epi = epi;
    return g;
}

// fp32 [N,K] weight -> bf16 planes [N, planes*K] + TMA descriptor (box 64 x 128, 128B swizzle)
static int make_w16(rg_model* m, const float* W, int N, int K, W16* out) {
    const int P = m->planes;
    if (dalloc(m, &out->ptr, (size_t)N * K * P * 2)) return 1;
    CU(rg_launch_split_bf16(W, K, out->ptr, K * P, P == 2 ? K : 0, N, K, 0));
    ++g_launches;
    CU(rg_make_tensor_map(&out->tm, out->ptr, N, (long long)K * P, (long long)K * P, 128));
    CU(rg_make_tensor_map(&out->tm64, out->ptr, N, (long long)K * P, (long long)K * P, 64));
    out->N = N; out->K = K;
    return 0;
}

// ca_mix(cat_c(h + proj_c(s_c))) == W_f [s_text | s_audio | s_spk | h] + b_f  with
//   W_f[:, c] = Wmix[:, c] Wo_c,  W_f[:, 3] = sum_c Wmix[:, c],  b_f = bmix + sum_c Wmix[:, c] bo_c
// (diffusion_transformer.py:115-122 + stylization_block.py:39): the three projections and the mix become
// ONE K = 2048 contraction on the tensor-core tiers.  Folded once, in fp32, at model creation.
static int make_fold(rg_model* m, const Layer& ly, LayerTc* t) {
    const int D = RG_D;
    float *Wf = nullptr, *WoT = nullptr;
    if (dalloc(m, (void**)&Wf, (size_t)D * 4 * D * sizeof(float))) return 1;
    if (dalloc(m, (void**)&WoT, (size_t)D * D * sizeof(float))) return 1;
    if (dalloc(m, (void**)&t->b_fold, (size_t)D * sizeof(float))) return 1;
    for (int c = 0; c < 3; ++c) {
        CU(rg_launch_transpose_sq(ly.ca_Wo + (long long)c * D * D, WoT, D, 0));
        RgGemm g = mk_gemm(ly.Wmix + c * D, 3 * D, WoT, nullptr, Wf + c * D, 4 * D, D, D, D, RG_EPI_BIAS);
        CU(rg_launch_gemm_f32(g, 0));
        g_launches += 2;
    }
    CU(rg_launch_sum3_blocks(ly.Wmix, 3 * D, Wf + 3 * D, 4 * D, D, D, 0));
    CU(rg_launch_gemm_f32(mk_gemm(ly.ca_bo, 3 * D, ly.Wmix, ly.bmix, t->b_fold, D, 1, D, 3 * D, RG_EPI_BIAS), 0));
    g_launches += 2;
    if (make_w16(m, Wf, D, 4 * D, &t->fold)) return 1;
    CU(cudaStreamSynchronize(0));
    dfree_one(m, Wf); dfree_one(m, WoT);
    return 0;
}

extern "C" const char* rg_last_error(void) { return g_err; }
extern "C" int rg_abi_version(void) { return RG_ABI_VERSION; }
extern "C" int64_t rg_launch_count(void) { return g_launches; }

static void drop_graphs(rg_model* m) {
    for (auto& g : m->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    m->graphs.clear();
}

extern "C" int rg_destroy(rg_handle h) {
    if (!h) return 0;
    drop_graphs(h);
    if (h->cap_st) cudaStreamDestroy(h->cap_st);
    for (void* p : h->allocs) cudaFree(p);
    for (int i = 0; i < RG_MAX_LANES; ++i) {
        if (h->lane_st[i]) cudaStreamDestroy(h->lane_st[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    delete h;
    return 0;
}

extern "C" int rg_create(const rg_config* cfg, int n_tensors, const char* const* names,
                         const float* const* ptrs, const int64_t* numels, rg_handle* out) {
    if (!cfg || !out) return rg_fail("rg_create: null argument");
    if (cfg->latent_dim != RG_D || cfg->num_heads != RG_H)
        return rg_fail("rg_create: kernels are built for latent_dim=%d, num_heads=%d (got %d, %d)",
                       RG_D, RG_H, cfg->latent_dim, cfg->num_heads);
    if (cfg->n_tokens > RG_MAX_T || cfg->n_tokens != 4 * cfg->n_chunks + 3)
        return rg_fail("rg_create: n_tokens=%d must be 4*n_chunks+3 and <= %d", cfg->n_tokens, RG_MAX_T);
    if (cfg->ffn_dim % 128 || cfg->time_embed_dim % 128 || cfg->text_dim % 16)
        return rg_fail("rg_create: ffn_dim/time_embed_dim must be multiples of 128, text_dim of 16");
    if (cfg->precision != RG_PREC_FP32 && cfg->precision != RG_PREC_BF16 && cfg->precision != RG_PREC_BF16X3)
        return rg_fail("rg_create: unknown precision %d", cfg->precision);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return rg_fail("rg_create: no CUDA device (this library has no CPU path)");
    TensorMap tm;
    for (int i = 0; i < n_tensors; ++i) tm.t[names[i]] = {ptrs[i], (long long)numels[i]};

    rg_model* m = new rg_model();
    memset(&m->cfg, 0, sizeof(m->cfg));
    m->cfg = *cfg;
    cudaGetDevice(&m->device);
    m->n_steps = 0; m->table = nullptr; m->tau_row = nullptr; m->tau_cached = -1; m->ss_rep = nullptr; m->ss_rep_clips = 0;
    memset(m->ws, 0, sizeof(m->ws));
    m->lanes = 0; m->ev_fork = nullptr; m->auto_lane_clips = 64;
    if (const char* e = getenv("RG_AUTO_LANE_CLIPS")) m->auto_lane_clips = atoi(e);
    for (int i = 0; i < RG_MAX_LANES; ++i) { m->lane_st[i] = nullptr; m->ev_join[i] = nullptr; }
    m->kv_rows = 0; m->kv_ln = m->kv_buf = nullptr;
    m->tt_rows = 0; m->tt_emb = m->tt_t1 = m->tt_e = nullptr;
    m->use_graphs = 1; m->ss_one = nullptr; m->cap_st = nullptr; m->graph_clock = 0;
    if (const char* e = getenv("RG_GRAPHS")) m->use_graphs = atoi(e);
    m->planes = cfg->precision == RG_PREC_BF16X3 ? 2 : 1;
    m->attn_mode = cfg->precision == RG_PREC_BF16X3 ? 2 : 1;
    m->attn_mode_ca = m->attn_mode;
    if (const char* e = getenv("RG_ATTN_MODE")) m->attn_mode = m->attn_mode_ca = atoi(e);      // diagnostics only
    if (const char* e = getenv("RG_ATTN_MODE_CA")) m->attn_mode_ca = atoi(e);
    m->gemm_only = 0;
    m->fuse_styl_max = 128;
    if (const char* e = getenv("RG_FUSE_STYL_MAX")) m->fuse_styl_max = atoi(e);
    m->kv_a16 = nullptr;
    const int D = RG_D, E = cfg->time_embed_dim, F = cfg->ffn_dim, L = cfg->num_layers, T = cfg->n_tokens;
    const long long DD = (long long)D * D;
    cudaStream_t st = 0;
    int rc = 1;
    float *tmpW = nullptr, *tmpb = nullptr, *tmpg = nullptr, *tmpbe = nullptr, *seq = nullptr, *glob = nullptr;
#define TRY(x) do { if (x) goto fail; } while (0)
#define TRYCU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { rg_fail("%s -> %s", #x, cudaGetErrorString(_e)); goto fail; } } while (0)
    TRY(up_new(m, tm, "joint_embed.weight", DD, &m->W_joint));
    TRY(up_new(m, tm, "joint_embed.bias", D, &m->b_joint));
    TRY(up_new(m, tm, "out.weight", DD, &m->W_out));
    TRY(up_new(m, tm, "out.bias", D, &m->b_out));
    TRY(up_new(m, tm, "text_pre_proj.weight", (long long)D * cfg->text_dim, &m->W_text));
    TRY(up_new(m, tm, "text_pre_proj.bias", D, &m->b_text));
    TRY(up_new(m, tm, "audio_pre_proj.weight", (long long)D * cfg->text_dim, &m->W_audio));
    TRY(up_new(m, tm, "audio_pre_proj.bias", D, &m->b_audio));
    TRY(up_new(m, tm, "speaker_embedding.weight", (long long)cfg->num_speakers * D, &m->spk_table));
    TRY(up_new(m, tm, "time_embed.0.weight", (long long)E * D, &m->W_t0));
    TRY(up_new(m, tm, "time_embed.0.bias", E, &m->b_t0));
    TRY(up_new(m, tm, "time_embed.2.weight", (long long)E * E, &m->W_t2));
    TRY(up_new(m, tm, "time_embed.2.bias", E, &m->b_t2));
    // positional table: per-part sine + learned global
    TRY(up_new(m, tm, "sequence_embedding.pe", (long long)cfg->n_chunks * D, &seq));
    TRY(up_new(m, tm, "global_positional_embedding.pe", (long long)T * D, &glob));
    TRY(dalloc(m, (void**)&m->pos, (size_t)T * D * sizeof(float)));
    TRYCU(rg_launch_pos_table(seq, glob, m->pos, T, cfg->n_chunks, st));
    ++g_launches;

    TRY(dalloc(m, (void**)&m->We_all, (size_t)L * 5 * 2 * D * E * sizeof(float)));
    TRY(dalloc(m, (void**)&m->be_all, (size_t)L * 5 * 2 * D * sizeof(float)));
    for (int c = 0; c < 3; ++c) {
        TRY(dalloc(m, (void**)&m->Wkv_all[c], (size_t)L * 2 * DD * sizeof(float)));
        TRY(dalloc(m, (void**)&m->bkv_all[c], (size_t)L * 2 * D * sizeof(float)));
    }
    // scratch for folding
    TRY(dalloc(m, (void**)&tmpW, (size_t)3 * DD * sizeof(float)));
    TRY(dalloc(m, (void**)&tmpb, (size_t)3 * D * sizeof(float)));
    TRY(dalloc(m, (void**)&tmpg, (size_t)D * sizeof(float)));
    TRY(dalloc(m, (void**)&tmpbe, (size_t)D * sizeof(float)));

    m->layers.resize(L);
    for (int l = 0; l < L; ++l) {
        Layer& ly = m->layers[l];
        const std::string p = "temporal_decoder_blocks." + std::to_string(l);
        const char* conds[3] = {"xf_text", "xf_audio", "xf_spk"};
        // self-attention q|k|v, LayerNorm folded
        const char* qkv[3] = {"query", "key", "value"};
        for (int i = 0; i < 3; ++i) {
            TRY(up_to(tm, p + ".sa_block." + qkv[i] + ".weight", DD, tmpW + i * DD));
            TRY(up_to(tm, p + ".sa_block." + qkv[i] + ".bias", D, tmpb + i * D));
        }
        TRY(up_to(tm, p + ".sa_block.norm.weight", D, tmpg));
        TRY(up_to(tm, p + ".sa_block.norm.bias", D, tmpbe));
        TRY(dalloc(m, (void**)&ly.Wqkv, (size_t)3 * DD * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.bqkv, (size_t)3 * D * sizeof(float)));
        TRYCU(rg_launch_fold_ln(tmpW, tmpb, tmpg, tmpbe, ly.Wqkv, ly.bqkv, 3 * D, D, st));
        ++g_launches;
        auto styl = [&](const std::string& q, int blk, float** gmm, float** bta, float* Wo, float* bo) -> int {
            if (up_to(tm, q + ".emb_layers.1.weight", (long long)2 * D * E, m->We_all + ((long long)l * 5 + blk) * 2 * D * E)) return 1;
            if (up_to(tm, q + ".emb_layers.1.bias", 2 * D, m->be_all + ((long long)l * 5 + blk) * 2 * D)) return 1;
            if (up_to(tm, q + ".norm.weight", D, *gmm)) return 1;
            if (up_to(tm, q + ".norm.bias", D, *bta)) return 1;
            if (up_to(tm, q + ".out_layers.2.weight", DD, Wo)) return 1;
            if (up_to(tm, q + ".out_layers.2.bias", D, bo)) return 1;
            return 0;
        };
        TRY(dalloc(m, (void**)&ly.sa_g, D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.sa_b, D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.sa_Wo, DD * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.sa_bo, D * sizeof(float)));
        TRY(styl(p + ".sa_block.proj_out", 0, &ly.sa_g, &ly.sa_b, ly.sa_Wo, ly.sa_bo));
        // cross-attention
        TRY(dalloc(m, (void**)&ly.Wcaq, (size_t)3 * DD * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.bcaq, (size_t)3 * D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ca_g, (size_t)3 * D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ca_b, (size_t)3 * D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ca_Wo, (size_t)3 * DD * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ca_bo, (size_t)3 * D * sizeof(float)));
        for (int c = 0; c < 3; ++c) {
            const std::string q = p + ".ca_blocks." + conds[c];
            TRY(up_to(tm, q + ".query.weight", DD, tmpW));
            TRY(up_to(tm, q + ".query.bias", D, tmpb));
            TRY(up_to(tm, q + ".norm.weight", D, tmpg));
            TRY(up_to(tm, q + ".norm.bias", D, tmpbe));
            TRYCU(rg_launch_fold_ln(tmpW, tmpb, tmpg, tmpbe, ly.Wcaq + c * DD, ly.bcaq + c * D, D, D, st));
            ++g_launches;
            TRYCU(cudaStreamSynchronize(st));
            TRY(up_to(tm, q + ".key.weight", DD, tmpW));
            TRY(up_to(tm, q + ".key.bias", D, tmpb));
            TRY(up_to(tm, q + ".value.weight", DD, tmpW + DD));
            TRY(up_to(tm, q + ".value.bias", D, tmpb + D));
            TRY(up_to(tm, q + ".text_norm.weight", D, tmpg));
            TRY(up_to(tm, q + ".text_norm.bias", D, tmpbe));
            TRYCU(rg_launch_fold_ln(tmpW, tmpb, tmpg, tmpbe, m->Wkv_all[c] + (long long)l * 2 * DD,
                                    m->bkv_all[c] + (long long)l * 2 * D, 2 * D, D, st));
            ++g_launches;
            TRYCU(cudaStreamSynchronize(st));
            float* gq = ly.ca_g + c * D; float* bq = ly.ca_b + c * D;
            TRY(styl(q + ".proj_out", 1 + c, &gq, &bq, ly.ca_Wo + c * DD, ly.ca_bo + c * D));
        }
        TRY(up_new(m, tm, p + ".ca_mix.weight", 3 * DD, &ly.Wmix));
        TRY(up_new(m, tm, p + ".ca_mix.bias", D, &ly.bmix));
        TRY(up_new(m, tm, p + ".ffn.linear1.weight", (long long)F * D, &ly.W1));
        TRY(up_new(m, tm, p + ".ffn.linear1.bias", F, &ly.b1));
        TRY(up_new(m, tm, p + ".ffn.linear2.weight", (long long)D * F, &ly.W2));
        TRY(up_new(m, tm, p + ".ffn.linear2.bias", D, &ly.b2));
        TRY(dalloc(m, (void**)&ly.ffn_g, D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ffn_b, D * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ffn_Wo, DD * sizeof(float)));
        TRY(dalloc(m, (void**)&ly.ffn_bo, D * sizeof(float)));
        TRY(styl(p + ".ffn.proj_out", 4, &ly.ffn_g, &ly.ffn_b, ly.ffn_Wo, ly.ffn_bo));
        TRYCU(cudaStreamSynchronize(st));   // tmp buffers are reused by the next layer
    }
    TRYCU(cudaDeviceSynchronize());
    dfree_one(m, tmpW); dfree_one(m, tmpb); dfree_one(m, tmpg); dfree_one(m, tmpbe);
    dfree_one(m, seq); dfree_one(m, glob);
    TRY(dalloc(m, (void**)&m->tau_row, (size_t)L * 5 * 2 * D * sizeof(float)));
    TRY(dalloc(m, (void**)&m->ss_one, (size_t)L * 5 * 2 * D * sizeof(float)));
    if (cfg->precision != RG_PREC_FP32) {
        m->tc.resize(L);
        TRY(make_w16(m, m->W_joint, D, D, &m->tc_joint));
        TRY(make_w16(m, m->W_out, D, D, &m->tc_out));
        if (cfg->text_dim % 64 == 0) {          // condition pre-projections on the tensor cores too
            TRY(make_w16(m, m->W_text, D, cfg->text_dim, &m->tc_text));
            TRY(make_w16(m, m->W_audio, D, cfg->text_dim, &m->tc_audio));
        }
        for (int c = 0; c < 3; ++c) TRY(make_w16(m, m->Wkv_all[c], L * 2 * D, D, &m->tc_kv[c]));
        for (int l = 0; l < L; ++l) {
            const Layer& ly = m->layers[l];
            LayerTc& t = m->tc[l];
            TRY(make_w16(m, ly.Wqkv, 3 * D, D, &t.qkv));
            TRY(make_w16(m, ly.sa_Wo, D, D, &t.sa_o));
            TRY(make_w16(m, ly.Wcaq, 3 * D, D, &t.caq));
            TRY(make_fold(m, ly, &t));
            TRY(make_w16(m, ly.W1, F, D, &t.w1));
            TRY(make_w16(m, ly.W2, D, F, &t.w2));
            TRY(make_w16(m, ly.ffn_Wo, D, D, &t.ffn_o));
        }
        TRYCU(cudaDeviceSynchronize());
    }
    *out = m;
    rc = 0;
fail:
    if (rc) { std::string keep = g_err; rg_destroy(m); snprintf(g_err, sizeof(g_err), "%s", keep.c_str()); }
    return rc;
#undef TRY
#undef TRYCU
}

// ---- K7: (scale|shift) rows for arbitrary timesteps ------------------------------------------
// taus: n original-scale timesteps (host).  out [n, L*5*1024] (device).
static int time_rows(rg_model* m, const int* taus, int n, float* out, cudaStream_t st) {
    const int D = RG_D, E = m->cfg.time_embed_dim, L = m->cfg.num_layers;
    if (n > m->tt_rows) {
        dfree_one(m, m->tt_emb); dfree_one(m, m->tt_t1); dfree_one(m, m->tt_e);
        m->tt_emb = m->tt_t1 = m->tt_e = nullptr; m->tt_rows = 0;
        if (dalloc(m, (void**)&m->tt_emb, (size_t)n * D * sizeof(float))) return 1;
        if (dalloc(m, (void**)&m->tt_t1, (size_t)n * E * sizeof(float))) return 1;
        if (dalloc(m, (void**)&m->tt_e, (size_t)n * E * sizeof(float))) return 1;
        m->tt_rows = n;
    }
    // timestep_embedding (diffusion_transformer.py:36-43) in the reference's fp32 op order
    std::vector<float> temb((size_t)n * D);
    const int half = D / 2;
    const float neg_log = (float)(-log(10000.0));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < half; ++j) {
            const float freq = expf((neg_log * (float)j) / (float)half);
            const float arg = (float)taus[i] * freq;
            temb[(size_t)i * D + j] = cosf(arg);
            temb[(size_t)i * D + half + j] = sinf(arg);
        }
    CU(cudaMemcpyAsync(m->tt_emb, temb.data(), temb.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));   // temb is a stack-owned host buffer
    LAUNCH(rg_launch_gemm_f32(mk_gemm(m->tt_emb, D, m->W_t0, m->b_t0, m->tt_t1, E, n, E, D, RG_EPI_BIAS_SILU), st));
    // emb = time_embed.2(.)  then SiLU (the first op of every emb_layers)
    LAUNCH(rg_launch_gemm_f32(mk_gemm(m->tt_t1, E, m->W_t2, m->b_t2, m->tt_e, E, n, E, E, RG_EPI_BIAS_SILU), st));
    const int NT = L * 5 * 2 * D;
    LAUNCH(rg_launch_gemm_f32(mk_gemm(m->tt_e, E, m->We_all, m->be_all, out, NT, n, NT, E, RG_EPI_BIAS), st));
    return 0;
}

extern "C" int rg_set_schedule(rg_handle m, int n_steps, const int32_t* timestep_map,
                               const float* coef, void* stream) {
    if (!m || n_steps <= 0 || !timestep_map || !coef) return rg_fail("rg_set_schedule: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long NT = (long long)m->cfg.num_layers * 5 * 2 * RG_D;
    if (m->table) { dfree_one(m, m->table); m->table = nullptr; }
    if (dalloc(m, (void**)&m->table, (size_t)n_steps * NT * sizeof(float))) return 1;
    m->n_steps = n_steps;
    m->timestep_map.assign(timestep_map, timestep_map + n_steps);
    m->coef.assign(coef, coef + (size_t)n_steps * 8);
    std::vector<int> taus(timestep_map, timestep_map + n_steps);
    if (time_rows(m, taus.data(), n_steps, m->table, st)) return 1;
    CU(cudaStreamSynchronize(st));
    return 0;
}

// ---- workspaces ------------------------------------------------------------------------------
static int ensure_ws(rg_model* m, Ws& w, long long rows) {
    if (rows <= w.rows) return 0;
    if (!m->graphs.empty()) {               // captured launches point into the buffers that are about to move
        CU(cudaDeviceSynchronize());
        drop_graphs(m);
    }
    float** bufs[6] = {&w.h, &w.a, &w.big, &w.o3, &w.g, &w.y};
    for (auto b : bufs) { dfree_one(m, *b); *b = nullptr; }
    w.rows = 0;
    const int D = RG_D, F = m->cfg.ffn_dim;
    if (dalloc(m, (void**)&w.h, (size_t)rows * D * sizeof(float))) return 1;
    if (dalloc(m, (void**)&w.a, (size_t)rows * 3 * D * sizeof(float))) return 1;
    if (dalloc(m, (void**)&w.big, (size_t)rows * 3 * D * sizeof(float))) return 1;
    if (dalloc(m, (void**)&w.o3, (size_t)rows * 3 * D * sizeof(float))) return 1;
    if (dalloc(m, (void**)&w.g, (size_t)rows * F * sizeof(float))) return 1;
    if (dalloc(m, (void**)&w.y, (size_t)rows * D * sizeof(float))) return 1;
    if (m->cfg.precision != RG_PREC_FP32) {
        const int P = m->planes;
        void** b16[5] = {&w.x16, &w.a16, &w.a16x, &w.h16, &w.g16};
        const int width[5] = {D, D, 4 * D, D, F};
        CUtensorMap* tms[5] = {&w.tm_x16, &w.tm_a16, &w.tm_a16x, &w.tm_h16, &w.tm_g16};
        for (int i = 0; i < 5; ++i) {
            dfree_one(m, *b16[i]); *b16[i] = nullptr;
            if (dalloc(m, b16[i], (size_t)rows * width[i] * P * 2)) return 1;
            CU(cudaMemset(*b16[i], 0, (size_t)rows * width[i] * P * 2));
            CU(rg_make_tensor_map(tms[i], *b16[i], rows, (long long)width[i] * P, (long long)width[i] * P, 128));
        }
        CU(rg_make_store_map(&w.ts_h, w.h, rows, D, D, 4));
        CU(rg_make_store_map(&w.ts_big, w.big, rows, 3 * D, 3 * D, 4));
        CU(rg_make_store_map(&w.ts_y, w.y, rows, D, D, 4));
        CU(rg_make_store_map(&w.ts_h16, w.h16, rows, (long long)D * P, (long long)D * P, 2));
        CU(rg_make_store_map(&w.ts_a16x, w.a16x, rows, (long long)4 * D * P, (long long)4 * D * P, 2));
        CU(rg_make_store_map(&w.ts_g16, w.g16, rows, (long long)F * P, (long long)F * P, 2));
        // the memsets run on the legacy stream, which non-blocking lane streams do not wait for
        CU(cudaDeviceSynchronize());
    }
    w.rows = rows;
    return 0;
}

static int tc_gemm(rg_model* m, const CUtensorMap& tmA, int a_w, const W16& w, const float* bias, int M, int N,
                   int K, int epi, const float* R, float* C32, int ldc32, void* C16, int c16_w, cudaStream_t st,
                   const CUtensorMap* s32 = nullptr, const CUtensorMap* s16 = nullptr,
                   int groups = 1, int a_goff = 0, int w_goff = 0, int bc_goff = 0);

extern "C" int64_t rg_state_floats_per_clip(rg_handle m) {
    return m ? (int64_t)m->cfg.num_layers * 3 * RG_H * RG_HD * RG_HD : 0;
}

extern "C" int rg_encode_conditions(rg_handle m, const float* word, const float* audio,
                                    const int64_t* spk_ids, int B, int n_text, int n_audio,
                                    int n_spk, float* xf_text, float* xf_audio, float* xf_spk,
                                    void* stream) {
    if (!m) return rg_fail("rg_encode_conditions: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const int D = RG_D, TD = m->cfg.text_dim;
    // text_pre_proj / audio_pre_proj (diffusion_transformer.py:544-606): 2*rows*512*768 flop per condition, once per
    // clip.  Tensor-core tiers: the rows are split into bf16 planes and go through the tcgen05 GEMM (12 K blocks);
    // the fp32 tier keeps the FMA GEMM.
    const bool tc = m->cfg.precision != RG_PREC_FP32 && TD % 64 == 0;
    const float* src[2] = {word, audio};
    float* dst[2] = {xf_text, xf_audio};
    const int nrow[2] = {B * n_text, B * n_audio};
    const float* Wf[2] = {m->W_text, m->W_audio};
    const float* bf[2] = {m->b_text, m->b_audio};
    const W16* Wt[2] = {&m->tc_text, &m->tc_audio};
    for (int c = 0; c < 2; ++c) {
        if (!src[c] || !dst[c] || nrow[c] <= 0) continue;
        if (!tc) {
            LAUNCH(rg_launch_gemm_f32(mk_gemm(src[c], TD, Wf[c], bf[c], dst[c], D, nrow[c], D, TD, RG_EPI_BIAS), st));
            continue;
        }
        const int P = m->planes;
        rg_keep_mempool();
        void* a16 = nullptr;
        CU(cudaMallocAsync(&a16, (size_t)nrow[c] * TD * P * 2, st));
        int rc = 0;
        cudaError_t e = rg_launch_split_bf16(src[c], TD, a16, TD * P, P == 2 ? TD : 0, nrow[c], TD, st);
        CUtensorMap tmA;
        if (e == cudaSuccess) { ++g_launches; e = rg_make_tensor_map(&tmA, a16, nrow[c], (long long)TD * P, (long long)TD * P, 128); }
        if (e == cudaSuccess)
            rc = tc_gemm(m, tmA, TD, *Wt[c], bf[c], nrow[c], D, TD, RG_EPI_BIAS, nullptr, dst[c], D, nullptr, 0, st);
        cudaFreeAsync(a16, st);                 // stream-ordered: released after the GEMM that reads it
        if (e != cudaSuccess) return rg_fail("rg_encode_conditions: %s", cudaGetErrorString(e));
        if (rc) return 1;
    }
    if (spk_ids && xf_spk)
        LAUNCH(rg_launch_gather_rows(m->spk_table, (const long long*)spk_ids, xf_spk, (long long)B * n_spk,
                                     m->cfg.num_speakers, st));
    return 0;
}

extern "C" int rg_precompute_clip_state(rg_handle m, const float* xf_text, const float* xf_audio,
                                        const float* xf_spk, int n_text, int n_audio, int n_spk,
                                        int B, float* state, void* stream) {
    if (!m || !state) return rg_fail("rg_precompute_clip_state: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int D = RG_D, L = m->cfg.num_layers;
    const long long clip_stride = rg_state_floats_per_clip(m);
    const float* xf[3] = {xf_text, xf_audio, xf_spk};
    const int nt[3] = {n_text, n_audio, n_spk};
    const long long MAX_ROWS = 32768;
    for (int c = 0; c < 3; ++c) {
        if (!xf[c] || nt[c] <= 0) return rg_fail("rg_precompute_clip_state: condition %d missing", c);
        const int N = nt[c];
        int per = (int)(MAX_ROWS / N); if (per < 1) per = 1;
        const long long need = (long long)per * N;
        const bool tc = m->cfg.precision != RG_PREC_FP32;
        const int P = m->planes;
        if (need > m->kv_rows) {
            dfree_one(m, m->kv_ln); dfree_one(m, m->kv_buf); dfree_one(m, m->kv_a16);
            m->kv_ln = m->kv_buf = nullptr; m->kv_a16 = nullptr; m->kv_rows = 0;
            if (!tc && dalloc(m, (void**)&m->kv_ln, (size_t)need * D * sizeof(float))) return 1;
            if (tc && dalloc(m, &m->kv_a16, (size_t)need * D * P * 2)) return 1;
            if (dalloc(m, (void**)&m->kv_buf, (size_t)need * L * 2 * D * sizeof(float))) return 1;
            m->kv_rows = need;
        }
        for (int b0 = 0; b0 < B; b0 += per) {
            const int nb = (B - b0 < per) ? B - b0 : per;
            const int rows = nb * N;
            if (tc) {
                // normalised condition rows as bf16 planes, then all 8 layers' key|value in one tcgen05 GEMM
                LAUNCH(rg_launch_ln_rows(xf[c] + (long long)b0 * N * D, D, nullptr, nullptr,
                                         rg_out_b16(m->kv_a16, D * P, P == 2 ? D : 0), rows, st));
                CUtensorMap tmA;
                CU(rg_make_tensor_map(&tmA, m->kv_a16, rows, (long long)D * P, (long long)D * P, 128));
                if (tc_gemm(m, tmA, D, m->tc_kv[c], m->bkv_all[c], rows, L * 2 * D, D, RG_EPI_BIAS, nullptr, m->kv_buf,
                            L * 2 * D, nullptr, 0, st)) return 1;
            } else {
                LAUNCH(rg_launch_ln_rows(xf[c] + (long long)b0 * N * D, D, nullptr, nullptr, rg_out_f32(m->kv_ln, D), rows, st));
                LAUNCH(rg_launch_gemm_f32(mk_gemm(m->kv_ln, D, m->Wkv_all[c], m->bkv_all[c], m->kv_buf, L * 2 * D,
                                                  rows, L * 2 * D, D, RG_EPI_BIAS), st));
            }
            LAUNCH(rg_launch_kv_state(m->kv_buf, L * 2 * D, 0, D, N,
                                      state + (long long)b0 * clip_stride + (long long)c * RG_H * RG_HD * RG_HD,
                                      clip_stride, nb, L, 2 * D, (long long)3 * RG_H * RG_HD * RG_HD, st));
        }
    }
    return 0;
}

// one tcgen05 GEMM: A (bf16 planes, tensor map tmA, plane width a_w) x W16 -> fp32 and/or bf16 planes
static int tc_gemm(rg_model* m, const CUtensorMap& tmA, int a_w, const W16& w, const float* bias, int M, int N,
                   int K, int epi, const float* R, float* C32, int ldc32, void* C16, int c16_w, cudaStream_t st,
                   const CUtensorMap* s32, const CUtensorMap* s16, int groups, int a_goff, int w_goff, int bc_goff) {
    RgGemmTc p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.split = m->planes == 2; p.a_lo_off = a_w; p.w_lo_off = w.K;
    p.groups = groups; p.a_goff = a_goff; p.w_goff = w_goff; p.b_goff = bc_goff; p.c_goff = bc_goff;
    p.bias = bias; p.R = R; p.ldr = RG_D; p.pos = m->pos; p.pos_T = m->cfg.n_tokens;
    p.C32 = C32; p.ldc32 = ldc32; p.C16_ = C16; p.ldc16 = c16_w * m->planes;
    p.c16_lo_off = m->planes == 2 ? c16_w : 0; p.epi = epi;
    p.tmC32 = s32; p.tmC16 = s16; p.tmW64 = &w.tm64;
    LAUNCH(rg_launch_gemm_tc(tmA, w.tm, p, st));
    return 0;
}

// rg_denoise on the tensor cores: bf16 (or hi|lo bf16) operands, fp32 accumulation in TMEM, fp32
// residual stream / softmaxes / LayerNorm statistics / -1e6 masks exactly as on the fp32 path.
static int denoise_tc(rg_model* m, Ws& w, const float* x, int B, const float* ssrow, long long ss_stride,
                      const float* src_mask, const float* query_mask, long long qm_cond_stride, const float* state,
                      float* x0_out, cudaStream_t st) {
    const int D = RG_D, F = m->cfg.ffn_dim, L = m->cfg.num_layers, T = m->cfg.n_tokens, P = m->planes;
    const int M = B * T;
    const long long clip_stride = rg_state_floats_per_clip(m);
    const long long HS = (long long)RG_H * RG_HD * RG_HD;
    const int lo = P == 2;
    // gemm_only (roofline probe): the dense contractions of the evaluation as the same PDL chain, same weights,
    // shapes and epilogues, without the attention / row kernels between them (their inputs are then stale data)
    const bool all = !m->gemm_only;
    if (all) LAUNCH(rg_launch_split_bf16(x, D, w.x16, D * P, lo ? D : 0, M, D, st));
    if (tc_gemm(m, w.tm_x16, D, m->tc_joint, m->b_joint, M, D, D, RG_EPI_BIAS_POS, nullptr, w.h, D, nullptr, 0, st, &w.ts_h)) return 1;
    for (int l = 0; l < L; ++l) {
        const Layer& ly = m->layers[l];
        const LayerTc& t = m->tc[l];
        const float* ss = ssrow + (long long)l * 5 * 2 * D;
        // --- self-attention
        if (all) LAUNCH(rg_launch_ln_rows(w.h, D, nullptr, nullptr, rg_out_b16(w.a16, D * P, lo ? D : 0), M, st));
        if (tc_gemm(m, w.tm_a16, D, t.qkv, ly.bqkv, M, 3 * D, D, RG_EPI_BIAS, nullptr, w.big, 3 * D, nullptr, 0, st, &w.ts_big)) return 1;
        RgStylParams sp = {ly.sa_g, ly.sa_b, ss, ss_stride};
        // mma.sync core + Stylization prologue in one kernel (one CTA per clip): measured faster than core +
        // row kernel for the single-pass TF32 cores up to ~128 clips; the 3xTF32 variant is register-bound
        const bool fuse_styl = m->attn_mode == 1 && B <= m->fuse_styl_max;
        if (!all) {
        } else if (fuse_styl) {
            LAUNCH(rg_launch_sa_styl(w.big, src_mask, sp, rg_out_b16(w.a16, D * P, lo ? D : 0), B, T, m->attn_mode == 2, st));
        } else {
            LAUNCH(rg_launch_sa_core(w.big, src_mask, w.y, B, T, m->attn_mode, st));
            LAUNCH(rg_launch_styl_rows(w.y, D, sp, T, rg_out_b16(w.a16, D * P, lo ? D : 0), M, st));
        }
        // h1 = h + proj(...): fp32 residual stream + its bf16 planes as the 4th K-block of the folded ca GEMM
        {
            RgGemmTc p;
            memset(&p, 0, sizeof(p));
            p.M = M; p.N = D; p.K = D; p.split = lo; p.a_lo_off = D; p.w_lo_off = D; p.groups = 1;
            p.bias = ly.sa_bo; p.R = w.h; p.ldr = D; p.C32 = w.h; p.ldc32 = D;
            p.C16_ = reinterpret_cast<__nv_bfloat16*>(w.a16x) + 3 * D; p.ldc16 = 4 * D * P; p.c16_lo_off = lo ? 4 * D : 0;
            p.epi = RG_EPI_BIAS_RESIDUAL;
            p.tmC32 = &w.ts_h; p.tmC16 = &w.ts_a16x; p.c16_col0 = 3 * D; p.tmW64 = &t.sa_o.tm64;
            LAUNCH(rg_launch_gemm_tc(w.tm_a16, t.sa_o.tm, p, st));
        }
        // --- three cross-attentions on the same h, their projections and ca_mix folded into one GEMM
        if (all) LAUNCH(rg_launch_ln_rows(w.h, D, nullptr, nullptr, rg_out_b16(w.a16, D * P, lo ? D : 0), M, st));
        if (tc_gemm(m, w.tm_a16, D, t.caq, ly.bcaq, M, 3 * D, D, RG_EPI_BIAS, nullptr, w.big, 3 * D, nullptr, 0, st, &w.ts_big)) return 1;
        RgStylParams sp3[3];
        for (int c = 0; c < 3; ++c) sp3[c] = {ly.ca_g + c * D, ly.ca_b + c * D, ss + (1 + c) * 2 * D, ss_stride};
        if (!all) {
        } else if (m->attn_mode_ca == 1 && B <= m->fuse_styl_max) {
            LAUNCH(rg_launch_ca_styl(w.big, 3 * D, state + (long long)l * 3 * HS, clip_stride, HS, query_mask, qm_cond_stride,
                                     sp3, rg_out_b16(w.a16x, 4 * D * P, lo ? 4 * D : 0), B, T, m->attn_mode_ca == 2, st));
        } else {
            LAUNCH(rg_launch_ca_core(w.big, 3 * D, state + (long long)l * 3 * HS, clip_stride, HS, query_mask,
                                     qm_cond_stride, w.o3, 3 * D, B, T, m->attn_mode_ca, st));
            LAUNCH(rg_launch_styl_rows3(w.o3, 3 * D, sp3, T, rg_out_b16(w.a16x, 4 * D * P, lo ? 4 * D : 0), M, st));
        }
        if (tc_gemm(m, w.tm_a16x, 4 * D, t.fold, t.b_fold, M, D, 4 * D, RG_EPI_BIAS, nullptr, w.h, D, w.h16, D, st, &w.ts_h, &w.ts_h16)) return 1;
        // --- FFN
        if (tc_gemm(m, w.tm_h16, D, t.w1, ly.b1, M, F, D, RG_EPI_BIAS_GELU, nullptr, nullptr, 0, w.g16, F, st, nullptr, &w.ts_g16)) return 1;
        if (tc_gemm(m, w.tm_g16, F, t.w2, ly.b2, M, D, F, RG_EPI_BIAS, nullptr, w.y, D, nullptr, 0, st, &w.ts_y)) return 1;
        RgStylParams spf = {ly.ffn_g, ly.ffn_b, ss + 4 * 2 * D, ss_stride};
        if (all) LAUNCH(rg_launch_styl_rows(w.y, D, spf, T, rg_out_b16(w.a16, D * P, lo ? D : 0), M, st));
        if (tc_gemm(m, w.tm_a16, D, t.ffn_o, ly.ffn_bo, M, D, D, RG_EPI_BIAS_RESIDUAL, w.h, w.h, D, w.h16, D, st, &w.ts_h, &w.ts_h16)) return 1;
    }
    // x0_out is the caller's buffer: its store map is built per call (host-side encode, no device work)
    CUtensorMap ts_out;
    const bool out_map = (reinterpret_cast<uintptr_t>(x0_out) & 15) == 0;
    if (out_map) CU(rg_make_store_map(&ts_out, x0_out, M, D, D, 4));
    return tc_gemm(m, w.tm_h16, D, m->tc_out, m->b_out, M, D, D, RG_EPI_BIAS, nullptr, x0_out, D, nullptr, 0, st,
                   out_map ? &ts_out : nullptr);
}

// rg_denoise, fp32 SIMT tier (RG_PREC_FP32): one lane's clips.
static int denoise_f32(rg_model* m, Ws& w, const float* x, int B, const float* ssrow, long long ss_stride,
                       const float* src_mask, const float* query_mask, long long qm_cond_stride, const float* state,
                       float* x0_out, cudaStream_t st) {
    const int D = RG_D, F = m->cfg.ffn_dim, L = m->cfg.num_layers, T = m->cfg.n_tokens;
    const int M = B * T;
    const long long clip_stride = rg_state_floats_per_clip(m);
    const long long HS = (long long)RG_H * RG_HD * RG_HD;

    // h = joint_embed(x) + positional tables
    {
        RgGemm g = mk_gemm(x, D, m->W_joint, m->b_joint, w.h, D, M, D, D, RG_EPI_BIAS_POS);
        g.pos = m->pos; g.pos_T = T;
        LAUNCH(rg_launch_gemm_f32(g, st));
    }
    for (int l = 0; l < L; ++l) {
        const Layer& ly = m->layers[l];
        const float* ss = ssrow + (long long)l * 5 * 2 * D;
        // --- self-attention
        LAUNCH(rg_launch_ln_rows(w.h, D, nullptr, nullptr, rg_out_f32(w.a, D), M, st));
        LAUNCH(rg_launch_gemm_f32(mk_gemm(w.a, D, ly.Wqkv, ly.bqkv, w.big, 3 * D, M, 3 * D, D, RG_EPI_BIAS), st));
        RgStylParams sp = {ly.sa_g, ly.sa_b, ss, ss_stride};
        LAUNCH(rg_launch_sa_attention(w.big, src_mask, sp, nullptr, rg_out_f32(w.a, D), B, T, 1, st));
        {
            RgGemm g = mk_gemm(w.a, D, ly.sa_Wo, ly.sa_bo, w.h, D, M, D, D, RG_EPI_BIAS_RESIDUAL);
            g.R = w.h; g.ldr = D;
            LAUNCH(rg_launch_gemm_f32(g, st));
        }
        // --- three cross-attentions on the same h
        LAUNCH(rg_launch_ln_rows(w.h, D, nullptr, nullptr, rg_out_f32(w.a, D), M, st));
        LAUNCH(rg_launch_gemm_f32(mk_gemm(w.a, D, ly.Wcaq, ly.bcaq, w.big, 3 * D, M, 3 * D, D, RG_EPI_BIAS), st));
        RgStylParams sp3[3];
        for (int c = 0; c < 3; ++c) sp3[c] = {ly.ca_g + c * D, ly.ca_b + c * D, ss + (1 + c) * 2 * D, ss_stride};
        LAUNCH(rg_launch_ca_attention(w.big, 3 * D, state + (long long)l * 3 * HS, clip_stride, HS, query_mask,
                                      qm_cond_stride, sp3, rg_out_f32(w.a, 3 * D), B, T, 3, st));
        {
            RgGemm g = mk_gemm(w.a, 3 * D, ly.ca_Wo, ly.ca_bo, w.o3, 3 * D, M, D, D, RG_EPI_BIAS_RESIDUAL);
            g.groups = 3; g.a_g = D; g.w_g = (long long)D * D; g.b_g = D; g.c_g = D; g.R = w.h; g.ldr = D; g.r_g = 0;
            LAUNCH(rg_launch_gemm_f32(g, st));
        }
        LAUNCH(rg_launch_gemm_f32(mk_gemm(w.o3, 3 * D, ly.Wmix, ly.bmix, w.h, D, M, D, 3 * D, RG_EPI_BIAS), st));
        // --- FFN
        LAUNCH(rg_launch_gemm_f32(mk_gemm(w.h, D, ly.W1, ly.b1, w.g, F, M, F, D, RG_EPI_BIAS_GELU), st));
        LAUNCH(rg_launch_gemm_f32(mk_gemm(w.g, F, ly.W2, ly.b2, w.y, D, M, D, F, RG_EPI_BIAS), st));
        RgStylParams spf = {ly.ffn_g, ly.ffn_b, ss + 4 * 2 * D, ss_stride};
        LAUNCH(rg_launch_styl_rows(w.y, D, spf, T, rg_out_f32(w.a, D), M, st));
        {
            RgGemm g = mk_gemm(w.a, D, ly.ffn_Wo, ly.ffn_bo, w.h, D, M, D, D, RG_EPI_BIAS_RESIDUAL);
            g.R = w.h; g.ldr = D;
            LAUNCH(rg_launch_gemm_f32(g, st));
        }
    }
    LAUNCH(rg_launch_gemm_f32(mk_gemm(w.h, D, m->W_out, m->b_out, x0_out, D, M, D, D, RG_EPI_BIAS), st));
    return 0;
}

// broadcast one table row to `n_clips` consecutive per-clip rows
__global__ void __launch_bounds__(256) rep_rows_kernel(const float4* __restrict__ row, float4* __restrict__ out,
                                                      long long n4_row, long long n4_total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_total; i += stride)
        out[i] = __ldg(row + i % n4_row);
}

static int rep_rows(const float* row, float* out, long long n4_row, long long n4_total, cudaStream_t st) {
    if (n4_total <= 0) return 0;
    const int blocks = (int)std::min<long long>((n4_total + 255) / 256, 148 * 8);
    rep_rows_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(row), reinterpret_cast<float4*>(out), n4_row, n4_total);
    CU(cudaGetLastError());
    rg_count_launch(1);
    return 0;
}

// Lanes: clips are independent, so the batch can be cut into contiguous clip ranges whose kernel chains run
// concurrently on separate streams (fork / join with events; inside a stream capture the lane streams join the
// capture).  While one lane sits in a launch's fixed latency (PDL release, first operand stage, epilogue tail: ~6 us
// per GEMM) the other lane's kernels run.  Measured on B200, graph-replayed grouped evaluation, 1 / 2 / 3
// lanes (tools/diag_lanes.py): 64 clips 0.895 / 0.874 / 0.881 ms, 96: 1.037 / 1.029 / 1.057, 128: 1.331 / 1.213 /
// 1.242, 160: 1.626 / 1.478 / 1.501, 224: 2.433 / 1.902 / 2.031 -> automatic = 2 lanes from 64 clips on.
static int lanes_for(rg_model* m, int B) {
    int lanes = m->lanes > 0 ? m->lanes : (B >= m->auto_lane_clips ? 2 : 1);
    if (lanes > RG_MAX_LANES) lanes = RG_MAX_LANES;
    if (lanes > B) lanes = B;
    return lanes < 1 ? 1 : lanes;
}
static int eval_lanes(rg_model* m, int lanes, const float* x, int B, const float* ss, long long ss_stride,
                      const float* src_mask, const float* query_mask, long long qm_stride, const float* state,
                      float* x0_out, cudaStream_t st) {
    const bool tc = m->cfg.precision != RG_PREC_FP32;
    const int D = RG_D, T = m->cfg.n_tokens;
    if (lanes <= 1)
        return (tc ? denoise_tc : denoise_f32)(m, m->ws[0], x, B, ss, ss_stride, src_mask, query_mask, qm_stride, state,
                                               x0_out, st);
    if (!m->ev_fork) {
        CU(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
        for (int i = 1; i < RG_MAX_LANES; ++i) {
            CU(cudaStreamCreateWithFlags(&m->lane_st[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
        }
    }
    CU(cudaEventRecord(m->ev_fork, st));
    const long long clip_stride = rg_state_floats_per_clip(m);
    int rc = 0;
    for (int i = lanes - 1; i >= 0 && !rc; --i) {         // lane 0 (the caller's stream) is enqueued last
        const int b0 = (int)((long long)B * i / lanes), b1 = (int)((long long)B * (i + 1) / lanes);
        cudaStream_t ls = i ? m->lane_st[i] : st;
        if (i) CU(cudaStreamWaitEvent(ls, m->ev_fork, 0));
        const long long r0 = (long long)b0 * T;
        rc = (tc ? denoise_tc : denoise_f32)(m, m->ws[i], x + r0 * D, b1 - b0, ss + (long long)b0 * ss_stride, ss_stride,
                                             src_mask + r0, query_mask ? query_mask + r0 : nullptr, qm_stride,
                                             state + b0 * clip_stride, x0_out + r0 * D, ls);
        if (i) CU(cudaEventRecord(m->ev_join[i], ls));
    }
    for (int i = 1; i < lanes; ++i) CU(cudaStreamWaitEvent(st, m->ev_join[i], 0));
    return rc;
}

// One denoiser evaluation.  The ~100 launches of the chain depend only on buffer ADDRESSES (latents, masks, state,
// output, the (scale|shift) rows at `ss`), on B and on the lane count: the first call with a given set runs them
// directly, the second captures the same call sequence into a CUDA graph (PDL edges and the lanes' fork / join
// included), later calls replay it -- one cudaGraphLaunch instead of ~100 cudaLaunchKernelEx (host enqueue was the
// limit at B = 1 and at 8 ranks per host).  Graphs die with the workspace (ensure_ws) and are evicted
// LRU beyond 16.
static int run_eval(rg_model* m, const float* x, int B, const float* ss, long long ss_stride, const float* src_mask,
                    const float* query_mask, long long qm_stride, const float* state, float* x0_out, cudaStream_t st) {
    const int lanes = lanes_for(m, B), T = m->cfg.n_tokens;
    for (int i = 0; i < lanes; ++i) {               // before any capture: growing a workspace drops the graphs
        const int b0 = (int)((long long)B * i / lanes), b1 = (int)((long long)B * (i + 1) / lanes);
        if (ensure_ws(m, m->ws[i], (long long)(b1 - b0) * T)) return 1;
    }
    auto direct = [&]() { return eval_lanes(m, lanes, x, B, ss, ss_stride, src_mask, query_mask, qm_stride, state, x0_out, st); };
    if (!m->use_graphs) return direct();
    rg_model::EvalGraph key = {x, src_mask, query_mask, state, x0_out, ss, ss_stride, qm_stride, B, m->gemm_only,
                               rg_gemm_kernel_mode, rg_gemm2_min_rows, rg_gemm2_persist_tiles, rg_pair128_min_rows, lanes,
                               nullptr, 0, 0};
    rg_model::EvalGraph* hit = nullptr;
    for (auto& g : m->graphs)
        if (g.x == key.x && g.src_mask == key.src_mask && g.qmask == key.qmask && g.state == key.state && g.x0 == key.x0 &&
            g.ss == key.ss && g.ss_stride == key.ss_stride && g.qm_stride == key.qm_stride && g.B == key.B &&
            g.gemm_only == key.gemm_only && g.kmode == key.kmode && g.kmin == key.kmin && g.kpt == key.kpt &&
            g.kp128 == key.kp128 && g.lanes == key.lanes) { hit = &g; break; }
    if (!hit) {                                     // first sight: run directly (also performs every one-off init)
        if (m->graphs.size() >= 16) {
            size_t old = 0;
            for (size_t i = 1; i < m->graphs.size(); ++i)
                if (m->graphs[i].last_use < m->graphs[old].last_use) old = i;
            if (m->graphs[old].exec) {
                // the exec may still be running on the caller's stream
                CU(cudaStreamSynchronize(st));
                cudaGraphExecDestroy(m->graphs[old].exec);
            }
            m->graphs.erase(m->graphs.begin() + old);
        }
        key.last_use = ++m->graph_clock;
        m->graphs.push_back(key);
        return direct();
    }
    hit->last_use = ++m->graph_clock;
    if (!hit->exec) {
        if (!m->cap_st) CU(cudaStreamCreateWithFlags(&m->cap_st, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        const long long l0 = g_launches;
        cudaError_t e = cudaStreamBeginCapture(m->cap_st, cudaStreamCaptureModeThreadLocal);
        int rc = 1;
        if (e == cudaSuccess) {
            rc = eval_lanes(m, lanes, x, B, ss, ss_stride, src_mask, query_mask, qm_stride, state, x0_out, m->cap_st);
            e = cudaStreamEndCapture(m->cap_st, &graph);
        }
        hit->launches = g_launches - l0;
        g_launches -= hit->launches;                // nothing ran yet (other threads may have counted meanwhile)
        if (rc == 0 && e == cudaSuccess && graph) e = cudaGraphInstantiate(&hit->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != 0 || e != cudaSuccess || !hit->exec) {   // capture not possible here: stay on direct launches
            cudaGetLastError();
            hit->exec = nullptr;
            m->use_graphs = 0;
            return direct();
        }
    }
    CU(cudaGraphLaunch(hit->exec, st));
    g_launches += hit->launches;
    return 0;
}

extern "C" int rg_denoise(rg_handle m, const float* x, int B, int step_idx, int tau,
                          const float* src_mask, const float* query_mask, const float* state,
                          float* x0_out, void* stream) {
    if (!m || !x || !src_mask || !state || !x0_out) return rg_fail("rg_denoise: null argument");
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = RG_D, L = m->cfg.num_layers, T = m->cfg.n_tokens;
    const long long NT = (long long)L * 5 * 2 * D;
    const float* ssrow;
    if (step_idx >= 0) {
        if (step_idx >= m->n_steps) return rg_fail("rg_denoise: step_idx %d outside the schedule (%d steps)", step_idx, m->n_steps);
        ssrow = m->table + (long long)step_idx * NT;
    } else {
        if (m->tau_cached != tau) {
            if (time_rows(m, &tau, 1, m->tau_row, st)) return 1;
            m->tau_cached = tau;
        }
        ssrow = m->tau_row;
    }
    const float* row = ssrow;
    if (m->use_graphs && step_idx >= 0) {           // the level's row at an address that does not change per level
        if (rep_rows(ssrow, m->ss_one, NT / 4, NT / 4, st)) return 1;
        row = m->ss_one;
    }
    return run_eval(m, x, B, row, 0, src_mask, query_mask, (long long)B * T, state, x0_out, st);
}

extern "C" int rg_denoise_groups(rg_handle m, const float* x, int B, int n_groups, const int32_t* group_clips,
                                 const int32_t* group_step_idx, const float* src_mask, const float* query_mask,
                                 const float* state, float* x0_out, void* stream) {
    if (!m || !x || !src_mask || !state || !x0_out || !group_clips || !group_step_idx)
        return rg_fail("rg_denoise_groups: null argument");
    if (n_groups < 1 || n_groups > 8) return rg_fail("rg_denoise_groups: 1..8 groups");
    int total = 0;
    for (int g = 0; g < n_groups; ++g) {
        if (group_clips[g] < 0) return rg_fail("rg_denoise_groups: negative group size");
        if (group_step_idx[g] < 0 || group_step_idx[g] >= m->n_steps)
            return rg_fail("rg_denoise_groups: step_idx %d outside the schedule (%d steps)", group_step_idx[g], m->n_steps);
        total += group_clips[g];
    }
    if (total != B) return rg_fail("rg_denoise_groups: group sizes sum to %d, B = %d", total, B);
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = RG_D, L = m->cfg.num_layers, T = m->cfg.n_tokens;
    const long long NT = (long long)L * 5 * 2 * D;
    if (B > m->ss_rep_clips) {
        dfree_one(m, m->ss_rep);
        m->ss_rep = nullptr; m->ss_rep_clips = 0;
        if (dalloc(m, (void**)&m->ss_rep, (size_t)B * NT * sizeof(float))) return 1;
        m->ss_rep_clips = B;
    }
    long long b0 = 0;
    for (int g = 0; g < n_groups; ++g) {
        if (rep_rows(m->table + (long long)group_step_idx[g] * NT, m->ss_rep + b0 * NT, NT / 4,
                     (long long)group_clips[g] * NT / 4, st)) return 1;
        b0 += group_clips[g];
    }
    return run_eval(m, x, B, m->ss_rep, NT, src_mask, query_mask, (long long)B * T, state, x0_out, st);
}

extern "C" int rg_set_lanes(rg_handle m, int lanes) {
    if (!m) return rg_fail("rg_set_lanes: null handle");
    if (lanes < 0 || lanes > RG_MAX_LANES) return rg_fail("rg_set_lanes: lanes must be 0 (auto) .. %d", RG_MAX_LANES);
    m->lanes = lanes;
    return 0;
}

static int step_coef(rg_model* m, int step_idx, const float** c) {
    if (!m) return rg_fail("null handle");
    if (step_idx < 0 || step_idx >= m->n_steps) return rg_fail("step_idx %d outside the schedule (%d steps)", step_idx, m->n_steps);
    *c = m->coef.data() + (size_t)step_idx * 8;
    return 0;
}

extern "C" int rg_ddim_update(rg_handle m, const float* x, const float* x0, int step_idx,
                              int direction, float* out, int64_t n, void* stream) {
    const float* c;
    if (step_coef(m, step_idx, &c)) return 1;
    const float ca = direction < 0 ? c[2] : c[4], cb = direction < 0 ? c[3] : c[5];
    LAUNCH(rg_launch_ddim_update(x, x0, out, n, c[0], c[1], ca, cb, (cudaStream_t)stream));
    return 0;
}

extern "C" int rg_blend_in_seq(rg_handle m, const float* x, const float* in_seq, const float* noise,
                               int step_idx, float* out, int64_t rows, void* stream) {
    const float* c;
    if (step_coef(m, step_idx, &c)) return 1;
    LAUNCH(rg_launch_blend(x, in_seq, noise, out, rows, c[6], c[7], (cudaStream_t)stream));
    return 0;
}

extern "C" int rg_mix_branches(rg_handle m, const float* out2, int B, const float* coef, const float* joint_scale,
                               float* out, void* stream) {
    if (!m || !out2 || !coef || !joint_scale || !out) return rg_fail("rg_mix_branches: null argument");
    if (B <= 0) return 0;
    LAUNCH(rg_launch_mix_branches(out2, coef, joint_scale, out, (long long)B * m->cfg.n_tokens, m->cfg.n_tokens,
                                  (cudaStream_t)stream));
    return 0;
}

extern "C" int rg_guidance_steps(rg_handle, float* x, const float* in_seq, int64_t rows, int iters,
                                 float lr, int64_t numel, void* stream) {
    if (iters <= 0) return 0;
    LAUNCH(rg_launch_guidance(x, in_seq, rows, iters, (float)(2.0 * (double)lr / (double)numel), (cudaStream_t)stream));
    return 0;
}

// ---- whole loops in one call ------------------------------------------------------------------
extern "C" int rg_run_levels(rg_handle m, int S, float* xj, int B, int E, const float* in_seq0,
                             const float* inv_list, const float* noise, const int32_t* guidance_iters,
                             float guidance_lr, int run_dead_guidance, const float* src_mask,
                             const float* query_mask, const float* state, float* samples_out,
                             float* x0_scratch, void* stream) {
    if (!m || !xj || !src_mask || !state || !x0_scratch) return rg_fail("rg_run_levels: null argument");
    if (B < 0 || E < 0 || B + E <= 0) return rg_fail("rg_run_levels: empty batch");
    if (S < 1 || S > m->n_steps) return rg_fail("rg_run_levels: S = %d outside the schedule (%d steps)", S, m->n_steps);
    if (E > 0 && !samples_out) return rg_fail("rg_run_levels: samples_out is required with exemplars");
    if (B > 0 && (in_seq0 || inv_list) && !noise) return rg_fail("rg_run_levels: blend noise is required with in_seq");
    const int T = m->cfg.n_tokens, D = RG_D;
    const long long clip = (long long)T * D, nB = (long long)B * clip, nE = (long long)E * clip;
    float* xe = xj + nB;
    for (int j = 0; j < S; ++j) {
        const int i = S - 1 - j;                       // guided level (descending); j = inversion level (ascending)
        if (B > 0) {
            const float* in_seq = (inv_list && i != S - 1) ? inv_list + (long long)i * nB : in_seq0;
            if (in_seq) {
                if (inv_list && i != S - 1 && run_dead_guidance && guidance_iters && guidance_iters[i] > 0)
                    if (rg_guidance_steps(m, xj, in_seq, (long long)B * T, guidance_iters[i], guidance_lr, nB, stream)) return 1;
                if (rg_blend_in_seq(m, xj, in_seq, noise + (long long)j * nB, i, xj, (long long)B * T, stream)) return 1;
            }
        }
        if (B > 0 && E > 0) {
            const int32_t clips[2] = {B, E}, steps[2] = {i, j};
            if (rg_denoise_groups(m, xj, B + E, 2, clips, steps, src_mask, query_mask, state, x0_scratch, stream)) return 1;
        } else {
            if (rg_denoise(m, xj, B + E, B > 0 ? i : j, 0, src_mask, query_mask, state, x0_scratch, stream)) return 1;
        }
        if (B > 0 && rg_ddim_update(m, xj, x0_scratch, i, -1, xj, nB, stream)) return 1;
        if (E > 0) {
            float* sj = samples_out + (long long)j * nE;
            if (rg_ddim_update(m, xe, x0_scratch + nB, j, +1, sj, nE, stream)) return 1;
            CU(cudaMemcpyAsync(xe, sj, (size_t)nE * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        }
    }
    return 0;
}

// ---- op-level entry points -------------------------------------------------------------------
extern "C" int rg_op_linear(const float* x, int ldx, const float* W, const float* b,
                            const float* residual, float* out, int M, int N, int K, int epilogue,
                            void* stream) {
    int epi;
    switch (epilogue) {
        case RG_OP_NONE: epi = RG_EPI_BIAS; break;
        case RG_OP_RESIDUAL: epi = RG_EPI_BIAS_RESIDUAL; break;
        case RG_OP_GELU: epi = RG_EPI_BIAS_GELU; break;
        case RG_OP_SILU: epi = RG_EPI_BIAS_SILU; break;
        default: return rg_fail("rg_op_linear: unknown epilogue %d", epilogue);
    }
    if (epi == RG_EPI_BIAS_RESIDUAL && !residual) return rg_fail("rg_op_linear: residual epilogue without residual");
    RgGemm g = mk_gemm(x, ldx, W, b, out, N, M, N, K, epi);
    g.R = residual; g.ldr = N;
    LAUNCH(rg_launch_gemm_f32(g, (cudaStream_t)stream));
    return 0;
}
// x W^T on the tensor cores; `w16_pre`: the weight already converted by rg_op_split_bf16 (planes as `split` says),
// or NULL to convert W here.
static int linear_tc_impl(const float* x, const float* W, const void* w16_pre, const float* b, const float* residual,
                          float* out, void* out_bf16, int M, int N, int K, int epilogue, int split, cudaStream_t st,
                          const char* who) {
    int epi;
    switch (epilogue) {
        case RG_OP_NONE: epi = RG_EPI_BIAS; break;
        case RG_OP_RESIDUAL: epi = RG_EPI_BIAS_RESIDUAL; break;
        case RG_OP_GELU: epi = RG_EPI_BIAS_GELU; break;
        case RG_OP_SILU: epi = RG_EPI_BIAS_SILU; break;
        default: return rg_fail("%s: unknown epilogue %d", who, epilogue);
    }
    if (N % 128 || K % 64) return rg_fail("%s: need N %% 128 == 0 and K %% 64 == 0 (got N=%d K=%d)", who, N, K);
    if (epi == RG_EPI_BIAS_RESIDUAL && !residual) return rg_fail("%s: residual epilogue without residual", who);
    if (M <= 0) return 0;
    const int planes = split ? 2 : 1;
    rg_keep_mempool();
    void *a16 = nullptr, *w16 = nullptr;
    CU(cudaMallocAsync(&a16, (size_t)M * K * planes * 2, st));
    LAUNCH(rg_launch_split_bf16(x, K, a16, K * planes, split ? K : 0, M, K, st));
    if (!w16_pre) {
        CU(cudaMallocAsync(&w16, (size_t)N * K * planes * 2, st));
        LAUNCH(rg_launch_split_bf16(W, K, w16, K * planes, split ? K : 0, N, K, st));
    }
    void* wp = w16_pre ? const_cast<void*>(w16_pre) : w16;
    CUtensorMap tmA, tmW;
    CU(rg_make_tensor_map(&tmA, a16, M, (long long)K * planes, (long long)K * planes, 128));
    CU(rg_make_tensor_map(&tmW, wp, N, (long long)K * planes, (long long)K * planes, 128));
    RgGemmTc p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.split = split ? 1 : 0; p.a_lo_off = K; p.w_lo_off = K; p.groups = 1;
    p.bias = b; p.R = residual; p.ldr = N; p.C32 = out; p.ldc32 = N;
    p.C16_ = out_bf16; p.ldc16 = N * planes; p.c16_lo_off = split ? N : 0; p.epi = epi;
    p.no_pdl = 1;       // the operand planes were written by the split kernel just above
    CUtensorMap ts32, ts16, tmW64;
    CU(rg_make_tensor_map(&tmW64, wp, N, (long long)K * planes, (long long)K * planes, 64));
    p.tmW64 = &tmW64;
    if (out && (reinterpret_cast<uintptr_t>(out) & 15) == 0) { CU(rg_make_store_map(&ts32, out, M, N, N, 4)); p.tmC32 = &ts32; }
    if (out_bf16 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0) {
        CU(rg_make_store_map(&ts16, out_bf16, M, (long long)N * planes, (long long)N * planes, 2));
        p.tmC16 = &ts16;
    }
    LAUNCH(rg_launch_gemm_tc(tmA, tmW, p, st));
    CU(cudaFreeAsync(a16, st));
    if (w16) CU(cudaFreeAsync(w16, st));
    return 0;
}
extern "C" int rg_op_linear_tc(const float* x, const float* W, const float* b, const float* residual,
                               float* out, void* out_bf16, int M, int N, int K, int epilogue, int split,
                               void* stream) {
    return linear_tc_impl(x, W, nullptr, b, residual, out, out_bf16, M, N, K, epilogue, split, (cudaStream_t)stream,
                          "rg_op_linear_tc");
}
extern "C" int rg_op_split_bf16(const float* w, void* out16, int rows, int cols, int split, void* stream) {
    if (rows <= 0 || cols <= 0) return rg_fail("rg_op_split_bf16: bad shape %d x %d", rows, cols);
    const int planes = split ? 2 : 1;
    LAUNCH(rg_launch_split_bf16(w, cols, out16, cols * planes, split ? cols : 0, rows, cols, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_linear_tc_w16(const float* x, const void* w16, const float* b, const float* residual,
                                   float* out, void* out_bf16, int M, int N, int K, int epilogue, int split,
                                   void* stream) {
    if (!w16) return rg_fail("rg_op_linear_tc_w16: no weight planes");
    return linear_tc_impl(x, nullptr, w16, b, residual, out, out_bf16, M, N, K, epilogue, split, (cudaStream_t)stream,
                          "rg_op_linear_tc_w16");
}
extern "C" int rg_op_mha(const float* q, const float* k, const float* v, const unsigned char* keep, float* out,
                         int N, int Sq, int Sk, int H, int dh, int64_t ldq, int64_t ldk, int64_t ldv, void* stream) {
    if (dh != 16 && dh != 32 && dh != 64 && dh != 128) return rg_fail("rg_op_mha: head dim %d not in {16, 32, 64, 128}", dh);
    if ((ldq | ldk | ldv) & 3) return rg_fail("rg_op_mha: row strides must be multiples of 4 floats");
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) return rg_fail("rg_op_mha: q, k, v must be 16-byte aligned");
    LAUNCH(rg_launch_mha(q, k, v, keep, out, N, Sq, Sk, H, dh, ldq, ldk, ldv, 1.0f / sqrtf((float)dh), (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_probe_gemm_tc(const float* x, const float* W, const float* b, float* out, int M, int N, int K,
                                int split, int reps, void* flush_buf, int64_t flush_bytes, float* median_ms,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N % 128 || K % 64 || reps < 1 || reps > 64) return rg_fail("rg_probe_gemm_tc: bad shape or reps");
    const int planes = split ? 2 : 1;
    void *a16 = nullptr, *w16 = nullptr;
    CU(cudaMalloc(&a16, (size_t)M * K * planes * 2));
    CU(cudaMalloc(&w16, (size_t)N * K * planes * 2));
    LAUNCH(rg_launch_split_bf16(x, K, a16, K * planes, split ? K : 0, M, K, st));
    LAUNCH(rg_launch_split_bf16(W, K, w16, K * planes, split ? K : 0, N, K, st));
    CUtensorMap tmA, tmW;
    CU(rg_make_tensor_map(&tmA, a16, M, (long long)K * planes, (long long)K * planes, 128));
    CU(rg_make_tensor_map(&tmW, w16, N, (long long)K * planes, (long long)K * planes, 128));
    RgGemmTc p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.split = split ? 1 : 0; p.a_lo_off = K; p.w_lo_off = K; p.groups = 1;
    p.bias = b; p.C32 = out; p.ldc32 = N; p.epi = RG_EPI_BIAS;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    std::vector<float> ts;
    for (int i = 0; i < reps + 2; ++i) {
        if (flush_buf) CU(cudaMemsetAsync(flush_buf, i, (size_t)flush_bytes, st));
        CU(cudaEventRecord(e0, st));
        LAUNCH(rg_launch_gemm_tc(tmA, tmW, p, st));
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= 2) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    *median_ms = ts[ts.size() / 2];
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(a16); cudaFree(w16);
    return 0;
}
// L2 -> SM read bandwidth probe: every SM streams an L2-resident buffer with 128-bit loads, `passes` times
__global__ void __launch_bounds__(512) l2_read_kernel(const float4* __restrict__ p, long long n4, int passes, float* sink) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int it = 0; it < passes; ++it)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < n4) ? __ldcg(p + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    if (acc == 1.2345e-30f) *sink = acc;        // keeps the loads alive
}
extern "C" int rg_probe_l2_read(const void* buf, int64_t bytes, int passes, float* gb_per_s, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!buf || bytes < 1024 || passes < 1 || !gb_per_s) return rg_fail("rg_probe_l2_read: bad argument");
    float* sink = nullptr;
    CU(cudaMalloc((void**)&sink, 4));
    const long long n4 = bytes / 16;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    l2_read_kernel<<<148 * 2, 512, 0, st>>>((const float4*)buf, n4, 2, sink);       // warm: the buffer now sits in L2
    CU(cudaEventRecord(e0, st));
    l2_read_kernel<<<148 * 2, 512, 0, st>>>((const float4*)buf, n4, passes, sink);
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    *gb_per_s = (float)((double)n4 * 16.0 * passes / (ms * 1e-3) / 1e9);
    rg_count_launch(2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 0;
}
extern "C" int rg_set_graphs(rg_handle m, int on) {
    if (!m) return rg_fail("rg_set_graphs: null handle");
    m->use_graphs = on ? 1 : 0;
    return 0;
}
extern "C" int rg_set_gemm_kernel(int mode, int min_rows, int persist_tiles, int pair128_min_rows) {
    if (mode < 0 || mode > 3)
        return rg_fail("rg_set_gemm_kernel: mode must be 0 (auto), 1 (128x128 tiles), 2 (2-CTA 256x256 tiles) or 3 (pair128)");
    rg_gemm_kernel_mode = mode;
    if (min_rows > 0) rg_gemm2_min_rows = min_rows;
    if (persist_tiles > 0) rg_gemm2_persist_tiles = persist_tiles;
    if (pair128_min_rows > 0) rg_pair128_min_rows = pair128_min_rows;
    return 0;
}
extern "C" int rg_probe_gemm_only(rg_handle m, int on) {
    if (!m) return rg_fail("rg_probe_gemm_only: null handle");
    m->gemm_only = on ? 1 : 0;
    return 0;
}
extern "C" int rg_probe_gemm_trace(int M, int N, int K, int split, int epilogue, int64_t* trace_host,
                                   int64_t n_trace, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N % 128 || K % 64) return rg_fail("rg_probe_gemm_trace: bad shape");
    const int planes = split ? 2 : 1;
    const long long ctas = (long long)(N / 128) * ((M + 127) / 128);
    // 2-CTA kernel (rg_set_gemm_kernel mode 2): 16 globaltimer stamps per CTA, the last TWO launches of the chain
    const bool k2 = rg_gemm_kernel_mode == 2 && N % 256 == 0;
    const long long need = k2 ? 2 * ctas * 10 : ctas * 10;
    if (n_trace < need) return rg_fail("rg_probe_gemm_trace: trace buffer needs %lld entries", need);
    void *a16 = nullptr, *w16 = nullptr; float *out = nullptr, *res = nullptr; long long* tr = nullptr;
    CU(cudaMalloc(&a16, (size_t)M * K * planes * 2));
    CU(cudaMalloc(&w16, (size_t)N * K * planes * 2));
    CU(cudaMalloc(&out, (size_t)M * N * 4));
    CU(cudaMalloc(&res, (size_t)M * N * 4));
    CU(cudaMalloc(&tr, (size_t)need * 8));
    CU(cudaMemset(tr, 0, (size_t)need * 8));
    CU(cudaMemset(a16, 0, (size_t)M * K * planes * 2));
    CU(cudaMemset(w16, 0, (size_t)N * K * planes * 2));
    CU(cudaMemset(res, 0, (size_t)M * N * 4));
    CUtensorMap tmA, tmW, ts32, ts16;
    CU(rg_make_tensor_map(&tmA, a16, M, (long long)K * planes, (long long)K * planes, 128));
    CU(rg_make_tensor_map(&tmW, w16, N, (long long)K * planes, (long long)K * planes, 128));
    CU(rg_make_store_map(&ts32, out, M, N, N, 4));
    CU(rg_make_store_map(&ts16, res, M, (long long)N * planes, (long long)N * planes, 2));
    RgGemmTc p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.split = split ? 1 : 0; p.a_lo_off = K; p.w_lo_off = K; p.groups = 1;
    p.C32 = out; p.ldc32 = N; p.epi = RG_EPI_BIAS; p.tmC32 = &ts32;
    if (epilogue == 1) { p.R = res; p.ldr = N; p.epi = RG_EPI_BIAS_RESIDUAL; }
    if (epilogue == 2) {
        p.C32 = nullptr; p.tmC32 = nullptr; p.C16_ = res; p.tmC16 = &ts16; p.ldc16 = N * planes; p.c16_lo_off = split ? N : 0;
        p.epi = RG_EPI_BIAS_GELU;
    }
    for (int i = 0; i < 6; ++i) {
        p.trace = nullptr;
        if (k2 && i >= 4) p.trace = tr + (long long)(i - 4) * ctas * 10;
        if (!k2 && i == 5) p.trace = tr;
        LAUNCH(rg_launch_gemm_tc(tmA, tmW, p, st));
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(trace_host, tr, (size_t)need * 8, cudaMemcpyDeviceToHost));
    cudaFree(a16); cudaFree(w16); cudaFree(out); cudaFree(res); cudaFree(tr);
    return 0;
}
extern "C" int rg_op_layernorm(const float* x, const float* gamma, const float* beta, float* out,
                               int M, void* stream) {
    LAUNCH(rg_launch_ln_rows(x, RG_D, gamma, beta, rg_out_f32(out, RG_D), M, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_silu(const float* x, float* out, int64_t n, void* stream) {
    LAUNCH(rg_launch_silu(x, out, n, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_stylization_rows(const float* y, const float* gamma, const float* beta,
                                      const float* ss, int ss_per_clip, int rows_per_clip,
                                      float* out, int M, void* stream) {
    RgStylParams sp = {gamma, beta, ss, ss_per_clip ? 2 * RG_D : 0};
    LAUNCH(rg_launch_styl_rows(y, RG_D, sp, rows_per_clip, rg_out_f32(out, RG_D), M, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_self_attention(const float* qkv, const float* src_mask, const float* gamma,
                                    const float* beta, const float* ss, int ss_per_clip,
                                    const float* x_res, float* out, int B, int T, int with_styl,
                                    void* stream) {
    if (T > RG_MAX_T) return rg_fail("rg_op_self_attention: T=%d > %d", T, RG_MAX_T);
    if (!with_styl && !x_res) return rg_fail("rg_op_self_attention: x_res required when with_styl=0");
    RgStylParams sp = {gamma, beta, ss, ss_per_clip ? 2 * RG_D : 0};
    LAUNCH(rg_launch_sa_attention(qkv, src_mask, sp, x_res, rg_out_f32(out, RG_D), B, T, with_styl, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_cross_attention(const float* q, const float* state, const float* query_mask,
                                     const float* gamma, const float* beta, const float* ss,
                                     int ss_per_clip, float* out, int B, int T, void* stream) {
    if (T > RG_MAX_T) return rg_fail("rg_op_cross_attention: T=%d > %d", T, RG_MAX_T);
    RgStylParams sp[3];
    sp[0] = {gamma, beta, ss, ss_per_clip ? 2 * RG_D : 0};
    sp[1] = sp[0]; sp[2] = sp[0];
    LAUNCH(rg_launch_ca_attention(q, RG_D, state, (long long)RG_H * RG_HD * RG_HD, 0, query_mask, 0, sp,
                                  rg_out_f32(out, RG_D), B, T, 1, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_self_attention_core(const float* qkv, const float* src_mask, float* y, int B, int T, int mode,
                                         void* stream) {
    if (!qkv || !src_mask || !y) return rg_fail("rg_op_self_attention_core: null argument");
    if (mode < 0 || mode > 2) return rg_fail("rg_op_self_attention_core: mode must be 0, 1 or 2");
    LAUNCH(rg_launch_sa_core(qkv, src_mask, y, B, T, mode, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_cross_attention_core(const float* q3, const float* state, const float* query_mask, float* y,
                                          int B, int T, int mode, void* stream) {
    if (!q3 || !state || !y) return rg_fail("rg_op_cross_attention_core: null argument");
    if (mode < 0 || mode > 2) return rg_fail("rg_op_cross_attention_core: mode must be 0, 1 or 2");
    const long long HS = (long long)RG_H * RG_HD * RG_HD;
    LAUNCH(rg_launch_ca_core(q3, 3 * RG_D, state, 3 * HS, HS, query_mask, (long long)B * T, y, 3 * RG_D, B, T, mode,
                             (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_self_attention_tc(const float* qkv, const float* src_mask, const float* gamma, const float* beta,
                                       const float* ss, int ss_per_clip, float* out, int B, int T, int split,
                                       void* stream) {
    if (!qkv || !src_mask || !gamma || !beta || !ss || !out) return rg_fail("rg_op_self_attention_tc: null argument");
    RgStylParams sp = {gamma, beta, ss, ss_per_clip ? 2ll * RG_D : 0};
    LAUNCH(rg_launch_sa_styl(qkv, src_mask, sp, rg_out_f32(out, RG_D), B, T, split, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_cross_attention_tc(const float* q3, const float* state, const float* query_mask,
                                        const float* gamma3, const float* beta3, const float* ss3, int ss_per_clip,
                                        float* out, int B, int T, int split, void* stream) {
    if (!q3 || !state || !gamma3 || !beta3 || !ss3 || !out) return rg_fail("rg_op_cross_attention_tc: null argument");
    const long long HS = (long long)RG_H * RG_HD * RG_HD;
    RgStylParams sp3[3];
    for (int c = 0; c < 3; ++c) sp3[c] = {gamma3 + c * RG_D, beta3 + c * RG_D, ss3 + c * 2 * RG_D, ss_per_clip ? 6ll * RG_D : 0};
    LAUNCH(rg_launch_ca_styl(q3, 3 * RG_D, state, 3 * HS, HS, query_mask, (long long)B * T, sp3, rg_out_f32(out, 3 * RG_D), B, T,
                             split, (cudaStream_t)stream));
    return 0;
}
extern "C" int rg_op_kv_state(const float* kv, int n_tokens, int B, float* state, void* stream) {
    LAUNCH(rg_launch_kv_state(kv, 2 * RG_D, 0, RG_D, n_tokens, state, (long long)RG_H * RG_HD * RG_HD, B, 1, 0, 0,
                              (cudaStream_t)stream));
    return 0;
}
