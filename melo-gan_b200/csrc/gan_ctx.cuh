// gan_ctx.cuh -- context (workspaces, bound parameters) and layer helpers of the GAN hot path.
#pragma once
#include <string>
#include <vector>

#include "elem.cuh"
#include "gemm_tc.cuh"
#include "thin.cuh"
#include "banded.cuh"

// ---- cross-GPU sum of a small vector through peer memory (SyncBatchNorm statistics) ----
// Each rank owns one exchange buffer (cudaMalloc + CUDA IPC handle): kSyncSlots slots x 2 parities x kSyncMax floats, then one
// flag and one epoch counter per slot.  A sync point of slot s: write my vector into parity (epoch & 1) of my slot, publish
// flag = epoch with a system-scope release, spin until every peer's flag reached the epoch, add the peers' vectors in RANK
// order (every rank computes bit-identical sums).  Two parities suffice: a rank can reach epoch e + 2 of a slot only after
// every peer published e + 1, i.e. after they all finished reading epoch e.  One CTA, no NCCL, capturable in a CUDA graph.
constexpr int kMaxPeers = 8, kSyncSlots = 4, kSyncMax = 512;
constexpr int kSyncFlagOff = kSyncSlots * 2 * kSyncMax;                 // in 4-byte words
constexpr size_t kSyncBytes = (size_t)(kSyncFlagOff + 2 * kSyncSlots) * 4;
struct PeerSet { float* base[kMaxPeers]; int world, rank; };

static __global__ void __launch_bounds__(kSyncMax) peer_allreduce_kernel(const PeerSet P, float* __restrict__ data, int n, int slot) {
    __shared__ unsigned e_s;
    float* mine = P.base[P.rank];
    unsigned* flags = reinterpret_cast<unsigned*>(mine + kSyncFlagOff);
    const int tid = threadIdx.x;
    if (tid == 0) e_s = flags[kSyncSlots + slot] + 1u;                   // my epoch counter of this slot
    __syncthreads();
    const unsigned e = e_s;
    const int off = (slot * 2 + (int)(e & 1u)) * kSyncMax;
    const float v = tid < n ? data[tid] : 0.0f;
    if (tid < n) mine[off + tid] = v;
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        flags[kSyncSlots + slot] = e;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags + slot), "r"(e) : "memory");
    }
    if (tid < P.world && tid != P.rank) {
        const unsigned* pf = reinterpret_cast<const unsigned*>(P.base[tid] + kSyncFlagOff) + slot;
        unsigned seen = 0;
        for (unsigned spins = 0;; ++spins) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(pf) : "memory");
            if ((int)(seen - e) >= 0) break;
            if (spins > (1u << 28)) asm volatile("trap;");              // a lost peer must not hang the GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (tid < n) {
        float s = 0.0f;
        for (int r = 0; r < P.world; ++r) s += (r == P.rank) ? v : __ldcv(P.base[r] + off + tid);
        data[tid] = s;
    }
}

struct mg_gan;
inline int peer_allreduce(const PeerSet& P, float* data, int n, int slot, cudaStream_t st) {
    if (P.world <= 1) return MG_OK;
    MG_REQUIRE(n <= kSyncMax && slot >= 0 && slot < kSyncSlots, "peer_allreduce: bad size / slot");
    peer_allreduce_kernel<<<1, kSyncMax, 0, st>>>(P, data, n, slot);
    MG_LAUNCH_OK();
    return MG_OK;
}

struct mg_gan {
    mg_gan_config cfg;
    // derived sizes
    int B, T, L0, zin;             // batch, max_notes, max_notes/8, noise+embed
    bool bf16;
    bool bias_fused_c2 = false, bias_fused_c0 = false, bias_fused_c4 = false;   // disc_dgrad already summed these bias gradients
    // SyncBatchNorm over NVLink peer memory (mg_gan_sync_bn_*): every rank's exchange buffer, opened through CUDA IPC
    PeerSet sync{};
    void* sync_opened[kMaxPeers] = {};
    int sync_pending_world = 0;

    // ---- bound parameters / gradients (caller-owned device memory) ----
    struct EP { float *ln_w, *ln_b, *w1, *b1, *w2, *b2, *w3, *b3; } E{}, gE{};
    struct GP {
        float *a_w, *a_b, *l_w, *l_b, *p0_w, *p0_b, *p2_w, *p2_b, *d0_w, *d0_b, *bn1_w, *bn1_b, *d3_w, *d3_b,
            *bn2_w, *bn2_b, *d6_w, *d6_b;
        float *bn1_rm, *bn1_rv, *bn2_rm, *bn2_rv;
    } G{}, gG{};
    struct DP { float *c0_w, *c0_b, *c2_w, *c2_b, *c4_w, *c4_b, *fc_w, *fc_b, *rf_w, *rf_b; } D{}, gD{};
    struct EDP {
        struct { float *w, *b, *g, *be, *rm, *rv; } conv[4];
        float *pj_w, *pj_b, *c0_w, *c0_b, *c3_w, *c3_b, *hd_w, *hd_b;
    } ED{};
    struct EDG {   // gradients of the emotion discriminator (BASELINE config #3, training the classifier itself)
        struct { float *w, *b, *g, *be; } conv[4];
        float *pj_w, *pj_b, *c0_w, *c0_b, *c3_w, *c3_b, *hd_w, *hd_b;
    } gED{};
    bool bound[4] = {false, false, false, false};
    bool has_grads[4] = {false, false, false, false};
    bool ed_folded = false;
    bool weight_cache = false;     // mg_gan_weight_cache: packed tensor-core weights persist between launches

    // ---- workspaces (one arena) ----
    char* arena = nullptr;
    size_t arena_bytes = 0;
    struct Named { std::string name; void* ptr; size_t bytes; };
    std::vector<Named> named;

    // FeatureEncoder
    float *e_xhat, *e_ln, *e_z1, *e_h1, *e_z2, *e_h2, *e_emb, *e_d1, *e_d2, *e_dln;
    const float *e_mask1 = nullptr, *e_mask2 = nullptr;   // masks of the last train-mode forward (caller memory)
    const float* e_numeric = nullptr;
    // Generator
    float *g_xcat, *g_ha, *g_lat, *g_notes, *g_x1, *g_x2;
    void *g_hb, *g_y0, *g_y1, *g_y2;                 // activation dtype
    float *g_bn1_stats, *g_bn1_mean, *g_bn1_is, *g_bn2_stats, *g_bn2_mean, *g_bn2_is, *g_bn_sums;
    float *g_dy2, *g_dy1;                                   // float32 even in bf16 mode (BN backward)
    void *g_dx2, *g_dx1, *g_dy0;
    float *g_dhb, *g_dlat, *g_dha, *g_dxcat, *g_demb;
    // Critic (up to 3B rows)
    float *d_x3, *d_pool, *d_hf, *d_score, *d_seed, *d_dzf, *d_dp, *d_gx, *d_gp_ps, *d_q, *d_dnotes;
    void *d_h1, *d_h2, *d_h3, *d_dz1, *d_dz2, *d_dz3;
    int d_rows = 0;                                         // rows of the last critic forward
    const float* d_emb = nullptr;
    // Emotion discriminator
    void *ed_h[4], *ed_g[4], *ed_dzA, *ed_dzB;
    float *ed_pool, *ed_pj, *ed_c1, *ed_c1g, *ed_c2, *ed_c2g, *ed_logits, *ed_dlogits, *ed_d128, *ed_d256a,
        *ed_d256b, *ed_scale[4], *ed_shift[4];
    // banded tensor-core forms of the 4-channel layers (bf16 mode): zero-padded bf16 note copies + scratch
    __nv_bfloat16 *d_xp = nullptr, *g_np = nullptr, *d_dnp = nullptr;
    mg::banded::Scratch2 band{};
    // emotion-discriminator TRAINING workspaces (allocated on first use: mg_emotion_train_forward)
    char* ed_train_arena = nullptr;
    float *edt_z[4] = {nullptr, nullptr, nullptr, nullptr};   // pre-BatchNorm conv outputs (float32)
    float *edt_mean[4], *edt_is[4], *edt_stats, *edt_dyb;      // batch statistics, scratch, d(BN output) float32
    float *edt_z1, *edt_h1, *edt_z2, *edt_h2, *edt_d1, *edt_d2, *edt_dpj, *edt_dpool;
    const float *edt_mask1 = nullptr, *edt_mask2 = nullptr;
    float edt_drop_scale = 1.0f;
    // misc
    float *partial, *metrics, *seed_g;
    const float* g_cond = nullptr;     // 'conditioning' mode: encoder latent (B, cond_dim), caller-owned
    size_t partial_floats = 0;
    int fwd_state = 0;   // bit flags of completed forwards, for MG_ERR_STATE checks
};

namespace mg {

enum { FWD_E = 1, FWD_G = 2, FWD_D = 4, FWD_ED = 8 };

inline TapGemmArgs tap_defaults() {
    TapGemmArgs a{};
    a.alpha = 1.0f;
    a.ntaps = 1;
    return a;
}

// ---- Linear: Out[R, N] = act(A[R, K] W[N, K]^T + bias) ----
template <typename TA, typename TO>
int linear_fwd(const TA* A, TO* Out, const float* W, const float* bias, int R, int K, int N, int act, void* aux,
               cudaStream_t st, int n_perm_q = 0, int n_perm_p = 0) {
    TapGemmArgs a = tap_defaults();
    a.A = A; a.a_bstride = K; a.a_mstride = 0; a.a_valid = K; a.K = K;
    a.W = W; a.w_nstride = K; a.w_kstride = 1; a.n_perm_q = n_perm_q; a.n_perm_p = n_perm_p;
    a.Out = Out; a.o_bstride = N; a.o_mstride = 0; a.B = R; a.Mper = 1; a.N = N;
    a.bias = bias; a.act = act; a.aux = aux;
    return launch_tapgemm<TA, TO>(a, st);
}

// ---- Linear dgrad: dX[R, K] = (dZ[R, N] W[N, K]) * f'(ref) ----
template <typename TA, typename TO, typename TMSK = TO>
int linear_dgrad(const TA* dZ, TO* dX, const float* W, int R, int K, int N, const void* mul_src, int mul_mode,
                 cudaStream_t st, int k_perm_q = 0, int k_perm_p = 0) {
    TapGemmArgs a = tap_defaults();
    a.A = dZ; a.a_bstride = N; a.a_valid = N; a.K = N;
    a.W = W; a.w_nstride = 1; a.w_kstride = K; a.k_perm_q = k_perm_q; a.k_perm_p = k_perm_p;
    a.Out = dX; a.o_bstride = K; a.B = R; a.Mper = 1; a.N = K;
    a.mul_src = mul_src; a.mul_mode = mul_mode;
    return launch_tapgemm<TA, TO, TMSK>(a, st);
}

// ---- Linear wgrad: dW[N, K] += dZ[R, N]^T A[R, K] ----
template <typename TG, typename TA>
int linear_wgrad(const TG* dZ, const TA* A, float* dW, int r0, int r1, int K, int N, cudaStream_t st,
                 int n_perm_q = 0, int n_perm_p = 0) {
    WgradArgs w{};
    w.G = dZ; w.g_bstride = N; w.A = A; w.a_bstride = K; w.a_valid = K; w.ntaps = 1; w.K = K;
    w.dW = dW; w.w_nstride = K; w.w_kstride = 1; w.n_perm_q = n_perm_q; w.n_perm_p = n_perm_p;
    w.B = r1; w.Mper = 1; w.N = N; w.alpha = 1.0f; w.row_begin = r0; w.row_end = r1;
    return launch_wgrad<TG, TA>(w, st);
}

// ---- Conv1d forward (channels-last): in [R, Lin, Cin] -> out [R, Lin/stride, Cout]; W [Cout][Cin][ks] ----
template <typename TA, typename TO, typename TMSK = TO>
int conv_fwd(const TA* in, TO* out, const float* W, const float* bias, int R, int Lin, int Cin, int Cout, int ks,
             int stride, int pad, int act, const float* col_scale, void* aux, const void* mul_src, int mul_mode,
             cudaStream_t st, int w_nstride = -1, int w_kstride = -1, float* pool_out = nullptr, float pool_scale = 0.0f,
             int* pool_done = nullptr, int pool_only = 0) {
    TapGemmArgs a = tap_defaults();
    a.pool_out = pool_out; a.pool_scale = pool_scale; a.pool_done = pool_done; a.pool_only = pool_only;
    const int Lout = Lin / stride;
    a.A = in; a.a_bstride = (long long)Lin * Cin; a.a_mstride = stride * Cin; a.a_valid = Lin * Cin;
    a.ntaps = ks; a.K = Cin;
    for (int t = 0; t < ks; ++t) { a.a_toff[t] = (t - pad) * Cin; a.w_toff[t] = t; }
    a.W = W; a.w_nstride = w_nstride < 0 ? Cin * ks : w_nstride; a.w_kstride = w_kstride < 0 ? ks : w_kstride;
    a.Out = out; a.o_bstride = (long long)Lout * Cout; a.o_mstride = Cout; a.B = R; a.Mper = Lout; a.N = Cout;
    a.bias = bias; a.act = act; a.col_scale = col_scale; a.aux = aux; a.mul_src = mul_src; a.mul_mode = mul_mode;
    return launch_tapgemm<TA, TO, TMSK>(a, st);
}

// ---- stride-1 conv dgrad: dIn[R, L, Cin] = sum_t dOut[R, l + pad - t, Cout] W[Cout][Cin][ks] ----
template <typename TA, typename TO, typename TMSK = TO>
int conv_s1_dgrad(const TA* dOut, TO* dIn, const float* W, int R, int L, int Cin, int Cout, int ks, int pad,
                  const float* col_scale, const void* mul_src, int mul_mode, int accumulate, cudaStream_t st) {
    TapGemmArgs a = tap_defaults();
    a.A = dOut; a.a_bstride = (long long)L * Cout; a.a_mstride = Cout; a.a_valid = L * Cout;
    a.ntaps = ks; a.K = Cout;
    for (int t = 0; t < ks; ++t) { a.a_toff[t] = (pad - t) * Cout; a.w_toff[t] = t; }
    a.W = W; a.w_nstride = ks; a.w_kstride = Cin * ks;      // n = ci, k = co
    a.Out = dIn; a.o_bstride = (long long)L * Cin; a.o_mstride = Cin; a.B = R; a.Mper = L; a.N = Cin;
    a.col_scale = col_scale; a.mul_src = mul_src; a.mul_mode = mul_mode; a.accumulate = accumulate;
    return launch_tapgemm<TA, TO, TMSK>(a, st);
}

// ---- k5 s2 p2 op1 up-sampling contraction in two sub-pixel phases ----
//   out[r, 2m+ph, n] = sum_{t = ph (mod 2)} sum_k in[r, m + 1 - t/2, k] * W(t, n, k)
// ConvTranspose1d forward: W [Cin][Cout][5] -> (w_nstride, w_kstride) = (5, Cout*5)
// dgrad of a strided Conv1d: W [Cconv_out][Cconv_in][5], k = conv_out, n = conv_in -> (5, Cconv_in*5)
template <typename TA, typename TO, typename TMSK = TO>
int upsample2_fwd(const TA* in, TO* out, const float* W, const float* bias, int R, int Lin, int K, int N,
                  int w_nstride, int w_kstride, int act, const void* mul_src, int mul_mode, int accumulate,
                  cudaStream_t st, float* colsum_out = nullptr, long long colsum_rows = 0, int* colsum_done = nullptr,
                  float* stats_out = nullptr, int* stats_done = nullptr) {
    // fused column sums of the output (bias gradient of the layer below): both phases add into colsum_out; the caller's
    // fallback reduction runs unless BOTH phases did it
    int done[2] = {0, 0}, sdone[2] = {0, 0};
    for (int ph = 0; ph < 2; ++ph) {
        TapGemmArgs a = tap_defaults();
        if (stats_out && (ph == 0 || sdone[0])) { a.stats_out = stats_out; a.stats_done = &sdone[ph]; }
        if (colsum_out && (ph == 0 || done[0])) { a.colsum_out = colsum_out; a.colsum_rows = colsum_rows; a.colsum_done = &done[ph]; }
        a.A = in; a.a_bstride = (long long)Lin * K; a.a_mstride = K; a.a_valid = Lin * K; a.K = K;
        a.ntaps = 0;
        for (int t = ph; t < 5; t += 2) {
            a.a_toff[a.ntaps] = (1 - t / 2) * K;
            a.w_toff[a.ntaps] = t;
            ++a.ntaps;
        }
        a.W = W; a.w_nstride = w_nstride; a.w_kstride = w_kstride;
        a.Out = out; a.o_bstride = (long long)2 * Lin * N; a.o_mstride = 2 * N; a.o_off = ph * N;
        a.B = R; a.Mper = Lin; a.N = N;
        a.bias = bias; a.act = act; a.mul_src = mul_src; a.mul_mode = mul_mode; a.accumulate = accumulate;
        int rc = launch_tapgemm<TA, TO, TMSK>(a, st);
        if (rc != MG_OK) return rc;
    }
    if (colsum_done) *colsum_done = (done[0] && done[1]) ? 1 : (done[0] ? -1 : 0);   // -1: half done (cannot happen: same shape)
    if (stats_done) *stats_done = (sdone[0] && sdone[1]) ? 1 : (sdone[0] ? -1 : 0);
    return MG_OK;
}

// ---- wgrad of a strided/unit Conv1d: dW[Cout][Cin][ks] += sum dOut[r, l, co] * in[r, stride*l + t - pad, ci] ----
template <typename TG, typename TA>
int conv_wgrad(const TG* dOut, const TA* in, float* dW, long long row0, long long row1, int Lin, int Cin, int Cout,
               int ks, int stride, int pad, cudaStream_t st) {
    WgradArgs w{};
    const int Lout = Lin / stride;
    w.G = dOut; w.g_bstride = (long long)Lout * Cout; w.g_mstride = Cout;
    w.A = in; w.a_bstride = (long long)Lin * Cin; w.a_mstride = stride * Cin; w.a_valid = Lin * Cin;
    w.ntaps = ks; w.K = Cin;
    for (int t = 0; t < ks; ++t) { w.a_toff[t] = (t - pad) * Cin; w.w_toff[t] = t; }
    w.dW = dW; w.w_nstride = Cin * ks; w.w_kstride = ks;
    w.B = (int)((row1 + Lout - 1) / Lout); w.Mper = Lout; w.N = Cout; w.alpha = 1.0f;
    w.row_begin = (int)row0; w.row_end = (int)row1;
    return launch_wgrad<TG, TA>(w, st);
}

// ---- wgrad of ConvTranspose1d k5 s2: dW[Cin][Cout][5] += sum in[r, i, ci] * dOut[r, 2i + t - 2, co] ----
template <typename TG, typename TA>
int convT_wgrad(const TG* in, const TA* dOut, float* dW, int R, int Lin, int Cin, int Cout, cudaStream_t st) {
    WgradArgs w{};
    w.G = in; w.g_bstride = (long long)Lin * Cin; w.g_mstride = Cin;
    w.A = dOut; w.a_bstride = (long long)2 * Lin * Cout; w.a_mstride = 2 * Cout; w.a_valid = 2 * Lin * Cout;
    w.ntaps = 5; w.K = Cout;
    for (int t = 0; t < 5; ++t) { w.a_toff[t] = (t - 2) * Cout; w.w_toff[t] = t; }
    w.dW = dW; w.w_nstride = Cout * 5; w.w_kstride = 5;     // n = ci, k = co
    w.B = R; w.Mper = Lin; w.N = Cin; w.alpha = 1.0f; w.row_begin = 0; w.row_end = R * Lin;
    return launch_wgrad<TG, TA>(w, st);
}

// ---- column reductions ----
template <typename T, int OP, typename TY = T>
int colreduce(mg_gan* c, const T* x, int ldx, const void* y, int ldy, const float* mean, const float* invstd,
              const float* roww, int roww_div, long long r0, long long r1, int C, float* out, int out_kstride,
              int perm_q, int perm_p, float alpha, int accumulate, cudaStream_t st) {
    constexpr int NOUT = (OP == COL_SUM_SQ || OP == COL_BN_BWD) ? 2 : 1;
    const long long rows = r1 - r0;
    if (rows <= 0 || C <= 0) return MG_OK;
    long long maxchunks = (long long)(c->partial_floats / ((size_t)NOUT * C));
    const bool vec4 = (C % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)x) % (4 * sizeof(T)) == 0) &&
                      (OP != COL_BN_BWD || ((ldy % 4 == 0) && (((uintptr_t)y) % (4 * sizeof(TY)) == 0)));
    const bool flat = vec4 && C <= 1024 && ((C / 4) & (C / 4 - 1)) == 0;
    if (flat) {
        const long long want = (long long)num_sms() * 8;
        if (maxchunks > want) maxchunks = want;
    } else {   // enough CTAs to fill the machine: (column groups) x (row chunks) ~ 4 per SM
        long long want = (long long)num_sms() * 4 / ((C / 4 + 31) / 32 > 0 ? (C / 4 + 31) / 32 : 1);
        if (want < 16) want = 16;
        if (want > 256) want = 256;
        if (maxchunks > want) maxchunks = want;
    }
    MG_REQUIRE(maxchunks >= 1, "colreduce: scratch too small for C=%d", C);
    long long rpc = (rows + maxchunks - 1) / maxchunks;
    if (rpc < 64) rpc = 64;
    rpc = (rpc + 7) / 8 * 8;
    const int nchunk = (int)((rows + rpc - 1) / rpc);
    ProbeScope probe(PROBE_ELEM, 0.0, (double)rows * C * (sizeof(T) + (OP == COL_BN_BWD ? sizeof(TY) : 0)), st);
    ColReduceArgs a{};
    a.x = x; a.ldx = ldx; a.y = y; a.ldy = ldy; a.mean = mean; a.invstd = invstd; a.roww = roww;
    a.roww_div = roww_div > 0 ? roww_div : 1; a.r0 = r0; a.r1 = r1; a.C = C; a.partial = c->partial;
    a.rows_per_chunk = (int)rpc;
    if (flat) {
        colreduce_flat_kernel<T, TY, OP><<<dim3(1, nchunk), 256, 0, st>>>(a);
    } else if (vec4) {
        dim3 grid((C / 4 + 31) / 32, nchunk);
        colreduce_vec4_kernel<T, TY, OP><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid((C + 31) / 32, nchunk);
        colreduce_kernel<T, TY, OP><<<grid, 256, 0, st>>>(a);
    }
    MG_LAUNCH_OK();
    const int n = NOUT * C;
    colreduce_finish_kernel<<<(n + 7) / 8, 256, 0, st>>>(c->partial, nchunk, NOUT, C, out, out_kstride, perm_q,
                                                             perm_p, alpha, accumulate);
    MG_LAUNCH_OK();
    return MG_OK;
}

inline int grid_for(long long n, int threads = 256, int max_per_sm = 8) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)num_sms() * max_per_sm;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// ---- BatchNorm1d + ReLU over a float32 pre-activation [rows, C] (train: batch statistics + running update) ----
template <typename T>
int bn_train_or_eval(mg_gan* c, const float* x, T* y, long long rows, int C, float* stats, float* mean, float* invstd,
                     const float* gamma, const float* beta, float* rm, float* rv, int train, cudaStream_t st,
                     bool stats_ready = false,     // the producing epilogue already summed x and x^2 into `stats`
                     int sync_slot = -1) {         // SyncBatchNorm: statistics over all ranks (mg_gan_sync_bn_connect)
    if (train) {
        if (!stats_ready)
            MG_TRY((colreduce<float, COL_SUM_SQ>(c, x, C, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, rows, C, stats, C, 0, 0,
                                             1.0f, 0, st)));
        const int world = (sync_slot >= 0 && c->sync.world > 1) ? c->sync.world : 1;
        if (world > 1) MG_TRY(peer_allreduce(c->sync, stats, 2 * C, sync_slot, st));
        bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(stats, C, rows * world, (float)c->cfg.bn_eps,
                                                            (float)c->cfg.bn_momentum, mean, invstd, rm, rv, 1);
    } else {
        bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(rm, rv, (float)c->cfg.bn_eps, C, mean, invstd);
    }
    MG_LAUNCH_OK();
    const long long n4 = rows * C / 4;
    ProbeScope probe(PROBE_ELEM, 0.0, (double)rows * C * (4 + sizeof(T)), st);
    bn_relu_apply_kernel<float, T><<<grid_for(n4), 256, 0, st>>>(x, y, n4, C, mean, invstd, gamma, beta);
    MG_LAUNCH_OK();
    return MG_OK;
}


// ---- its backward: dy (float32, ReLU mask already applied) -> dx, accumulates d gamma / d beta ----
template <typename T>
int bn_backward(mg_gan* c, const float* x, const float* dy, T* dx, long long rows, int C, const float* mean,
                const float* invstd, const float* gamma, float* dgamma, float* dbeta, cudaStream_t st, int sync_slot = -1) {
    // sums[0..C) = sum dy, sums[C..2C) = sum dy*xhat
    MG_TRY((colreduce<float, COL_BN_BWD, float>(c, x, C, dy, C, mean, invstd, nullptr, 1, 0, rows, C, c->g_bn_sums, C, 0, 0, 1.0f,
                                     0, st)));
    add2_kernel<<<(C + 127) / 128, 128, 0, st>>>(dbeta, c->g_bn_sums, dgamma, c->g_bn_sums + C, C);
    MG_LAUNCH_OK();
    // SyncBatchNorm: the parameter gradients above stay local sums (the gradient all-reduce adds them up); dx needs the sums
    // over the GLOBAL batch -- this rank computes d(its own loss), the 1/world of the mean over ranks is Adam's grad_scale
    const int world = (sync_slot >= 0 && c->sync.world > 1) ? c->sync.world : 1;
    if (world > 1) MG_TRY(peer_allreduce(c->sync, c->g_bn_sums, 2 * C, sync_slot, st));
    const long long n4 = rows * C / 4;
    ProbeScope probe(PROBE_ELEM, 0.0, (double)rows * C * (8 + sizeof(T)), st);
    bn_bwd_apply_kernel<float, T, float><<<grid_for(n4), 256, 0, st>>>(x, dy, dx, n4, C, 1.0f / (float)(rows * world), mean, invstd, gamma,
                                                         c->g_bn_sums);
    MG_LAUNCH_OK();
    return MG_OK;
}


}  // namespace mg
