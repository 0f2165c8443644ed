// vae.cu -- A-12 (BASELINE config #2): the piano-roll VAE forward / backward of reference src/ae/model.py:4-148
// with the loss of src/ae/train_ae.py:35-51, on the same kernel set as the GAN step (channels-last, implicit GEMM).
//
//   encoder  3 x [Conv1d k5 s2 p2 + BatchNorm1d + ReLU]  4 -> 32 -> 64 -> 128,  Flatten, Linear 128*L0 -> 512, ReLU
//            fc_mu / fc_log_var 512 -> latent;  z = mu + eps * exp(0.5 * log_var)   (eps is an input)
//   decoder  Linear latent -> 512 ReLU, Linear 512 -> 128*L0 ReLU, view (128, L0),
//            2 x [ConvTranspose1d k5 s2 + BatchNorm1d + ReLU] 128 -> 64 -> 32, ConvTranspose1d 32 -> 4, Tanh
// The reference flattens channel-major (c*L0 + l); activations here are [l][c], so both big Linears address their
// weight through the (q = 128, p = L0) index permutation instead of moving data.
#include <math.h>

#include "gan_ctx.cuh"

#include <cstring>
using namespace mg;

__global__ void relu_mask_inplace_kernel(float* d, const float* h, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = h[i] > 0.f ? d[i] : 0.f;
}

struct mg_vae {
    mg_gan g;                 // only its reduction scratch (partial, g_bn_sums) and cfg.bn_* are used
    int B, T, L0, latent;
    bool bf16;
    char* arena = nullptr;
    size_t arena_bytes = 0;
    struct P {
        float *c0_w, *c0_b, *bn0_w, *bn0_b, *c1_w, *c1_b, *bn1_w, *bn1_b, *c2_w, *c2_b, *bn2_w, *bn2_b;
        float *lin_w, *lin_b, *mu_w, *mu_b, *lv_w, *lv_b, *p0_w, *p0_b, *p2_w, *p2_b;
        float *d0_w, *d0_b, *dbn0_w, *dbn0_b, *d3_w, *d3_b, *dbn1_w, *dbn1_b, *d6_w, *d6_b;
        float *bn0_rm, *bn0_rv, *bn1_rm, *bn1_rv, *bn2_rm, *bn2_rv, *dbn0_rm, *dbn0_rv, *dbn1_rm, *dbn1_rv;
    } W{}, G{};
    bool bound = false, has_grads = false, fwd_done = false;
    // activations
    float *e_x[3], *d_x[2];               // pre-BatchNorm conv outputs (float32)
    void *e_a[3], *d_y0, *d_y[2];         // post BN+ReLU (activation dtype)
    float *bn_mean[5], *bn_is[5], *bn_stats;
    float *h, *mu, *lv, *z, *d0, *pre_t, *recon;
    const float* eps = nullptr;
    // gradients
    float *dt, *dy_f[2], *dh, *dmu, *dlv, *dz, *dd0, *de_f[3];
    void *dxd[2], *dy0, *dxe[3];
    float *ext_dmu, *ext_dlv;
    float* metrics;
};

namespace {

void vae_layout(mg_vae* v, char* base) {
    size_t off = 0;
    const size_t B = v->B, T = v->T, L0 = v->L0, lat = v->latent;
    const size_t es = v->bf16 ? 2 : 4;
    auto F = [&](size_t n) { float* p = base ? reinterpret_cast<float*>(base + off) : nullptr; off += (n * 4 + 255) / 256 * 256; return p; };
    auto A = [&](size_t n) { void* p = base ? (void*)(base + off) : nullptr; off += (n * es + 255) / 256 * 256; return p; };
    const size_t ech[3] = {32, 64, 128}, elen[3] = {T / 2, T / 4, T / 8};
    for (int i = 0; i < 3; ++i) { v->e_x[i] = F(B * elen[i] * ech[i]); v->e_a[i] = A(B * elen[i] * ech[i]); }
    const size_t dch[2] = {64, 32}, dlen[2] = {2 * L0, 4 * L0};
    for (int i = 0; i < 2; ++i) { v->d_x[i] = F(B * dlen[i] * dch[i]); v->d_y[i] = A(B * dlen[i] * dch[i]); }
    v->d_y0 = A(B * L0 * 128);
    for (int i = 0; i < 5; ++i) { v->bn_mean[i] = F(128); v->bn_is[i] = F(128); }
    v->bn_stats = F(256);
    v->h = F(B * 512); v->mu = F(B * lat); v->lv = F(B * lat); v->z = F(B * lat); v->d0 = F(B * 512);
    v->pre_t = F(B * T * 4); v->recon = F(B * T * 4);
    v->dt = F(B * T * 4);
    for (int i = 0; i < 2; ++i) { v->dy_f[i] = F(B * dlen[i] * dch[i]); v->dxd[i] = A(B * dlen[i] * dch[i]); }
    v->dy0 = A(B * L0 * 128);
    v->dh = F(B * 512); v->dmu = F(B * lat); v->dlv = F(B * lat); v->dz = F(B * lat); v->dd0 = F(B * 512);
    v->ext_dmu = F(B * lat); v->ext_dlv = F(B * lat);
    for (int i = 0; i < 3; ++i) { v->de_f[i] = F(B * elen[i] * ech[i]); v->dxe[i] = A(B * elen[i] * ech[i]); }
    v->metrics = F(16);
    v->g.partial_floats = (size_t)4 << 20;
    v->g.partial = F(v->g.partial_floats);
    v->g.g_bn_sums = F(512);
    v->arena_bytes = off;
}

__global__ void reparam_kernel(const float* mu, const float* lv, const float* eps, float* z, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = mu[i] + eps[i] * expf(0.5f * lv[i]);          // model.py:127-133
}
__global__ void tanh_kernel(const float* x, float* y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = tanhf(x[i]);
}
// dt = drecon * (1 - recon^2)
__global__ void tanh_bwd_kernel(const float* drecon, const float* recon, float* dt, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dt[i] = drecon[i] * (1.0f - recon[i] * recon[i]);
}
// dmu += dz ; dlv += dz * eps * 0.5 * exp(0.5 lv)      (external dmu/dlv/dz may be null)
__global__ void reparam_bwd_kernel(const float* dz_in, const float* dz_ext, const float* dmu_ext, const float* dlv_ext,
                                   const float* eps, const float* lv, float* dmu, float* dlv, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float dz = dz_in[i] + (dz_ext ? dz_ext[i] : 0.f);
    dmu[i] = dz + (dmu_ext ? dmu_ext[i] : 0.f);
    dlv[i] = dz * eps[i] * 0.5f * expf(0.5f * lv[i]) + (dlv_ext ? dlv_ext[i] : 0.f);
}
// vae_loss (train_ae.py:35-51): metrics = [total, recon (MSE mean), kld]; drecon = 2 (recon - x) / N;
// dmu = beta * mu / M ; dlv = beta * 0.5 (exp(lv) - 1) / M   with M = B * latent
__global__ void vae_loss_kernel(const float* recon, const float* x, long long n, const float* mu, const float* lv, int m,
                                float beta, float* drecon, float* dmu, float* dlv, float* out) {
    __shared__ double red[2][32];
    double se = 0, kl = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = recon[i] - x[i];
        se += (double)d * d;
        drecon[i] = 2.0f * d / (float)n;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const float e = expf(lv[i]);
        kl += (double)(1.0f + lv[i] - mu[i] * mu[i] - e);
        dmu[i] = beta * mu[i] / (float)m;
        dlv[i] = beta * 0.5f * (e - 1.0f) / (float)m;
    }
    for (int o = 16; o; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); kl += __shfl_xor_sync(0xffffffffu, kl, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = se; red[1][threadIdx.x >> 5] = kl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        se = kl = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { se += red[0][i]; kl += red[1][i]; }
        const float r = (float)(se / (double)n), k = (float)(-0.5 * kl / (double)m);
        out[0] = r + beta * k; out[1] = r; out[2] = k;
    }
}

template <typename T>
int vae_forward(mg_vae* v, const float* x, const float* eps, int train, float* recon_out, float* z_out, float* mu_out,
                float* lv_out, cudaStream_t st) {
    mg_gan* g = &v->g;
    const int B = v->B, T4 = v->T, L0 = v->L0, lat = v->latent;
    auto& W = v->W;
    // ---- encoder ----
    MG_TRY((conv_fwd<float, float>(x, v->e_x[0], W.c0_w, W.c0_b, B, T4, 4, 32, 5, 2, 2, ACT_NONE, nullptr, nullptr, nullptr, MUL_NONE, st)));
    MG_TRY((bn_train_or_eval<T>(g, v->e_x[0], (T*)v->e_a[0], (long long)B * T4 / 2, 32, v->bn_stats, v->bn_mean[0], v->bn_is[0],
                                W.bn0_w, W.bn0_b, W.bn0_rm, W.bn0_rv, train, st)));
    MG_TRY((conv_fwd<T, float>((const T*)v->e_a[0], v->e_x[1], W.c1_w, W.c1_b, B, T4 / 2, 32, 64, 5, 2, 2, ACT_NONE, nullptr, nullptr, nullptr, MUL_NONE, st)));
    MG_TRY((bn_train_or_eval<T>(g, v->e_x[1], (T*)v->e_a[1], (long long)B * T4 / 4, 64, v->bn_stats, v->bn_mean[1], v->bn_is[1],
                                W.bn1_w, W.bn1_b, W.bn1_rm, W.bn1_rv, train, st)));
    MG_TRY((conv_fwd<T, float>((const T*)v->e_a[1], v->e_x[2], W.c2_w, W.c2_b, B, T4 / 4, 64, 128, 5, 2, 2, ACT_NONE, nullptr, nullptr, nullptr, MUL_NONE, st)));
    MG_TRY((bn_train_or_eval<T>(g, v->e_x[2], (T*)v->e_a[2], (long long)B * L0, 128, v->bn_stats, v->bn_mean[2], v->bn_is[2],
                                W.bn2_w, W.bn2_b, W.bn2_rm, W.bn2_rv, train, st)));
    // Flatten (channel-major in the reference) + Linear + ReLU: logical k = l*128 + c <-> weight column c*L0 + l
    {
        TapGemmArgs a = tap_defaults();
        const int K = 128 * L0;
        a.A = v->e_a[2]; a.a_bstride = K; a.a_valid = K; a.K = K;
        a.W = W.lin_w; a.w_nstride = K; a.w_kstride = 1; a.k_perm_q = 128; a.k_perm_p = L0;
        a.Out = v->h; a.o_bstride = 512; a.B = B; a.Mper = 1; a.N = 512; a.bias = W.lin_b; a.act = ACT_RELU;
        MG_TRY((launch_tapgemm<T, float>(a, st)));
    }
    float* mu = mu_out ? mu_out : v->mu;
    float* lv = lv_out ? lv_out : v->lv;
    float* z = z_out ? z_out : v->z;
    MG_TRY((linear_fwd<float, float>(v->h, mu, W.mu_w, W.mu_b, B, 512, lat, ACT_NONE, nullptr, st)));
    MG_TRY((linear_fwd<float, float>(v->h, lv, W.lv_w, W.lv_b, B, 512, lat, ACT_NONE, nullptr, st)));
    reparam_kernel<<<(B * lat + 255) / 256, 256, 0, st>>>(mu, lv, eps, z, B * lat);
    MG_LAUNCH_OK();
    if (mu != v->mu) MG_CUDA_OK(cudaMemcpyAsync(v->mu, mu, sizeof(float) * B * lat, cudaMemcpyDeviceToDevice, st));
    if (lv != v->lv) MG_CUDA_OK(cudaMemcpyAsync(v->lv, lv, sizeof(float) * B * lat, cudaMemcpyDeviceToDevice, st));
    if (z != v->z) MG_CUDA_OK(cudaMemcpyAsync(v->z, z, sizeof(float) * B * lat, cudaMemcpyDeviceToDevice, st));
    // ---- decoder ----
    MG_TRY((linear_fwd<float, float>(v->z, v->d0, W.p0_w, W.p0_b, B, lat, 512, ACT_RELU, nullptr, st)));
    MG_TRY((linear_fwd<float, T>(v->d0, (T*)v->d_y0, W.p2_w, W.p2_b, B, 512, 128 * L0, ACT_RELU, nullptr, st, 128, L0)));
    MG_TRY((upsample2_fwd<T, float>((const T*)v->d_y0, v->d_x[0], W.d0_w, W.d0_b, B, L0, 128, 64, 5, 64 * 5, ACT_NONE, nullptr, MUL_NONE, 0, st)));
    MG_TRY((bn_train_or_eval<T>(g, v->d_x[0], (T*)v->d_y[0], (long long)B * 2 * L0, 64, v->bn_stats, v->bn_mean[3], v->bn_is[3],
                                W.dbn0_w, W.dbn0_b, W.dbn0_rm, W.dbn0_rv, train, st)));
    MG_TRY((upsample2_fwd<T, float>((const T*)v->d_y[0], v->d_x[1], W.d3_w, W.d3_b, B, 2 * L0, 64, 32, 5, 32 * 5, ACT_NONE, nullptr, MUL_NONE, 0, st)));
    MG_TRY((bn_train_or_eval<T>(g, v->d_x[1], (T*)v->d_y[1], (long long)B * 4 * L0, 32, v->bn_stats, v->bn_mean[4], v->bn_is[4],
                                W.dbn1_w, W.dbn1_b, W.dbn1_rm, W.dbn1_rv, train, st)));
    MG_TRY((upsample2_fwd<T, float>((const T*)v->d_y[1], v->pre_t, W.d6_w, W.d6_b, B, 4 * L0, 32, 4, 5, 4 * 5, ACT_NONE, nullptr, MUL_NONE, 0, st)));
    const long long n = (long long)B * T4 * 4;
    tanh_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v->pre_t, v->recon, n);
    MG_LAUNCH_OK();
    if (recon_out) MG_CUDA_OK(cudaMemcpyAsync(recon_out, v->recon, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    v->eps = eps;
    v->fwd_done = true;
    return MG_OK;
}

template <typename T>
int vae_backward(mg_vae* v, const float* x, const float* drecon, const float* dz_ext, const float* dmu_ext,
                 const float* dlv_ext, cudaStream_t st) {
    mg_gan* g = &v->g;
    const int B = v->B, T4 = v->T, L0 = v->L0, lat = v->latent;
    auto& W = v->W;
    auto& G = v->G;
    auto colsum_f = [&](const float* d, int N, long long rows, float* gb) {
        return colreduce<float, COL_SUM>(g, d, N, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, rows, N, gb, 0, 0, 0, 1.0f, 1, st);
    };
    auto colsum_t = [&](const T* d, int N, long long rows, float* gb) {
        return colreduce<T, COL_SUM>(g, d, N, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, rows, N, gb, 0, 0, 0, 1.0f, 1, st);
    };
    const long long n = (long long)B * T4 * 4;
    tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(drecon, v->recon, v->dt, n);
    MG_LAUNCH_OK();
    // ---- decoder backward ----
    MG_TRY(colsum_f((const float*)v->dt, 4, (long long)B * T4, G.d6_b));
    MG_TRY((convT_wgrad<T, float>((const T*)v->d_y[1], v->dt, G.d6_w, B, 4 * L0, 32, 4, st)));
    MG_TRY((conv_fwd<float, float, T>(v->dt, v->dy_f[1], W.d6_w, nullptr, B, T4, 4, 32, 5, 2, 2, ACT_NONE, nullptr, nullptr,
                                      v->d_y[1], MUL_RELU_SIGN, st, 4 * 5, 5)));
    MG_TRY((bn_backward<T>(g, v->d_x[1], v->dy_f[1], (T*)v->dxd[1], (long long)B * 4 * L0, 32, v->bn_mean[4], v->bn_is[4],
                           W.dbn1_w, G.dbn1_w, G.dbn1_b, st)));
    MG_TRY(colsum_t((const T*)v->dxd[1], 32, (long long)B * 4 * L0, G.d3_b));
    MG_TRY((convT_wgrad<T, T>((const T*)v->d_y[0], (const T*)v->dxd[1], G.d3_w, B, 2 * L0, 64, 32, st)));
    MG_TRY((conv_fwd<T, float, T>((const T*)v->dxd[1], v->dy_f[0], W.d3_w, nullptr, B, 4 * L0, 32, 64, 5, 2, 2, ACT_NONE, nullptr,
                                  nullptr, v->d_y[0], MUL_RELU_SIGN, st, 32 * 5, 5)));
    MG_TRY((bn_backward<T>(g, v->d_x[0], v->dy_f[0], (T*)v->dxd[0], (long long)B * 2 * L0, 64, v->bn_mean[3], v->bn_is[3],
                           W.dbn0_w, G.dbn0_w, G.dbn0_b, st)));
    MG_TRY(colsum_t((const T*)v->dxd[0], 64, (long long)B * 2 * L0, G.d0_b));
    MG_TRY((convT_wgrad<T, T>((const T*)v->d_y0, (const T*)v->dxd[0], G.d0_w, B, L0, 128, 64, st)));
    MG_TRY((conv_fwd<T, T>((const T*)v->dxd[0], (T*)v->dy0, W.d0_w, nullptr, B, 2 * L0, 64, 128, 5, 2, 2, ACT_NONE, nullptr, nullptr,
                           v->d_y0, MUL_RELU_SIGN, st, 64 * 5, 5)));
    const int N2 = 128 * L0;
    MG_TRY((colreduce<T, COL_SUM>(g, (const T*)v->dy0, N2, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, B, N2, G.p2_b, 0, 128, L0,
                                  1.0f, 1, st)));
    MG_TRY((linear_wgrad<T, float>((const T*)v->dy0, v->d0, G.p2_w, 0, B, 512, N2, st, 128, L0)));
    MG_TRY((linear_dgrad<T, float>((const T*)v->dy0, v->dd0, W.p2_w, B, 512, N2, v->d0, MUL_RELU_SIGN, st, 128, L0)));
    MG_TRY(colsum_f((const float*)v->dd0, 512, B, G.p0_b));
    MG_TRY((linear_wgrad<float, float>(v->dd0, v->z, G.p0_w, 0, B, lat, 512, st)));
    MG_TRY((linear_dgrad<float, float>(v->dd0, v->dz, W.p0_w, B, lat, 512, nullptr, MUL_NONE, st)));
    // ---- reparameterisation ----
    reparam_bwd_kernel<<<(B * lat + 255) / 256, 256, 0, st>>>(v->dz, dz_ext, dmu_ext, dlv_ext, v->eps, v->lv, v->dmu, v->dlv, B * lat);
    MG_LAUNCH_OK();
    // ---- fc_mu / fc_log_var ----
    MG_TRY(colsum_f((const float*)v->dmu, lat, B, G.mu_b));
    MG_TRY(colsum_f((const float*)v->dlv, lat, B, G.lv_b));
    MG_TRY((linear_wgrad<float, float>(v->dmu, v->h, G.mu_w, 0, B, 512, lat, st)));
    MG_TRY((linear_wgrad<float, float>(v->dlv, v->h, G.lv_w, 0, B, 512, lat, st)));
    MG_TRY((linear_dgrad<float, float>(v->dmu, v->dh, W.mu_w, B, 512, lat, nullptr, MUL_NONE, st)));
    {   // dh += dlv W_lv, then the ReLU mask of h
        TapGemmArgs a = tap_defaults();
        a.A = v->dlv; a.a_bstride = lat; a.a_valid = lat; a.K = lat;
        a.W = W.lv_w; a.w_nstride = 1; a.w_kstride = 512;
        a.Out = v->dh; a.o_bstride = 512; a.B = B; a.Mper = 1; a.N = 512; a.accumulate = 1;
        MG_TRY((launch_tapgemm<float, float>(a, st)));
        const long long m = (long long)B * 512;
        relu_mask_inplace_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(v->dh, v->h, m);
        MG_LAUNCH_OK();
    }
    // ---- encoder._linear.1 (flatten permutation on the weight columns) ----
    MG_TRY(colsum_f((const float*)v->dh, 512, B, G.lin_b));
    {
        const int K = 128 * L0;
        // dW[n][c*L0 + l] += sum_b a3[b, l*128 + c] * dh[b, n]: the flattened activation plays the "G" role so that the
        // column permutation can ride on the kernel's output-row permutation
        WgradArgs w{};
        w.G = v->e_a[2]; w.g_bstride = K; w.A = v->dh; w.a_bstride = 512; w.a_valid = 512; w.ntaps = 1; w.K = 512;
        w.dW = G.lin_w; w.w_nstride = 1; w.w_kstride = K; w.n_perm_q = 128; w.n_perm_p = L0;
        w.B = B; w.Mper = 1; w.N = K; w.alpha = 1.0f; w.row_begin = 0; w.row_end = B;
        MG_TRY((launch_wgrad<T, float>(w, st)));
        // d a3[b, l*128 + c] = sum_n dh[b, n] W[n][c*L0 + l], masked by ReLU(a3); float32 because BatchNorm backward follows
        TapGemmArgs a = tap_defaults();
        a.A = v->dh; a.a_bstride = 512; a.a_valid = 512; a.K = 512;
        a.W = W.lin_w; a.w_nstride = 1; a.w_kstride = K; a.n_perm_q = 128; a.n_perm_p = L0;
        a.Out = v->de_f[2]; a.o_bstride = K; a.B = B; a.Mper = 1; a.N = K; a.mul_src = v->e_a[2]; a.mul_mode = MUL_RELU_SIGN;
        MG_TRY((launch_tapgemm<float, float, T>(a, st)));
    }
    // ---- encoder convs (BN + ReLU backward, then wgrad / bias / dgrad) ----
    MG_TRY((bn_backward<T>(g, v->e_x[2], v->de_f[2], (T*)v->dxe[2], (long long)B * L0, 128, v->bn_mean[2], v->bn_is[2], W.bn2_w,
                           G.bn2_w, G.bn2_b, st)));
    MG_TRY(colsum_t((const T*)v->dxe[2], 128, (long long)B * L0, G.c2_b));
    MG_TRY((conv_wgrad<T, T>((const T*)v->dxe[2], (const T*)v->e_a[1], G.c2_w, 0, (long long)B * L0, T4 / 4, 64, 128, 5, 2, 2, st)));
    MG_TRY((upsample2_fwd<T, float, T>((const T*)v->dxe[2], v->de_f[1], W.c2_w, nullptr, B, L0, 128, 64, 5, 64 * 5, ACT_NONE,
                                       v->e_a[1], MUL_RELU_SIGN, 0, st)));
    MG_TRY((bn_backward<T>(g, v->e_x[1], v->de_f[1], (T*)v->dxe[1], (long long)B * T4 / 4, 64, v->bn_mean[1], v->bn_is[1], W.bn1_w,
                           G.bn1_w, G.bn1_b, st)));
    MG_TRY(colsum_t((const T*)v->dxe[1], 64, (long long)B * T4 / 4, G.c1_b));
    MG_TRY((conv_wgrad<T, T>((const T*)v->dxe[1], (const T*)v->e_a[0], G.c1_w, 0, (long long)B * T4 / 4, T4 / 2, 32, 64, 5, 2, 2, st)));
    MG_TRY((upsample2_fwd<T, float, T>((const T*)v->dxe[1], v->de_f[0], W.c1_w, nullptr, B, T4 / 4, 64, 32, 5, 32 * 5, ACT_NONE,
                                       v->e_a[0], MUL_RELU_SIGN, 0, st)));
    MG_TRY((bn_backward<T>(g, v->e_x[0], v->de_f[0], (T*)v->dxe[0], (long long)B * T4 / 2, 32, v->bn_mean[0], v->bn_is[0], W.bn0_w,
                           G.bn0_w, G.bn0_b, st)));
    MG_TRY(colsum_t((const T*)v->dxe[0], 32, (long long)B * T4 / 2, G.c0_b));
    MG_TRY((conv_wgrad<T, float>((const T*)v->dxe[0], x, G.c0_w, 0, (long long)B * T4 / 2, T4, 4, 32, 5, 2, 2, st)));
    return MG_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int mg_vae_create(int batch, int max_notes, int latent_dim, int precision, mg_vae** out) {
    MG_REQUIRE(out && batch >= 1 && batch <= 16384, "vae_create: bad batch");
    MG_REQUIRE(max_notes >= 8 && max_notes % 8 == 0 && max_notes <= 4096, "vae_create: max_notes must be a multiple of 8");
    MG_REQUIRE(latent_dim >= 1 && latent_dim <= 512, "vae_create: bad latent_dim");
    MG_REQUIRE(precision == 0 || precision == 1, "vae_create: precision must be 0 or 1");
    int dev = 0, major = 0;
    MG_CUDA_OK(cudaGetDevice(&dev));
    MG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MG_REQUIRE(major == 10, "vae_create: built for sm_100a only (device has compute capability %d.x)", major);
    mg_vae* v = new mg_vae();
    v->B = batch; v->T = max_notes; v->L0 = max_notes / 8; v->latent = latent_dim; v->bf16 = precision == 1;
    v->g.cfg.bn_eps = 1e-5; v->g.cfg.bn_momentum = 0.1; v->g.B = batch;
    vae_layout(v, nullptr);
    if (cudaMalloc(&v->arena, v->arena_bytes) != cudaSuccess) { delete v; mg::set_error("vae_create: out of memory"); return MG_ERR_CUDA; }
    cudaMemset(v->arena, 0, v->arena_bytes);
    vae_layout(v, v->arena);
    if (v->bf16) {
        int rc = mg::tc::ensure_scratch((size_t)128 * v->L0 * 512);
        if (rc != MG_OK) { cudaFree(v->arena); delete v; return rc; }
    }
    *out = v;
    return MG_OK;
}

extern "C" void mg_vae_destroy(mg_vae* v) {
    if (!v) return;
    if (v->arena) cudaFree(v->arena);
    delete v;
}

extern "C" int mg_vae_bind(mg_vae* v, void* const* params, int nparams, void* const* grads, int ngrads) {
    MG_REQUIRE(v && params && nparams == 42, "vae_bind: expects 42 parameter pointers (32 trainable + 10 running statistics)");
    MG_REQUIRE(!grads || ngrads == 32, "vae_bind: expects 32 gradient pointers");
    for (int i = 0; i < 42; ++i) MG_REQUIRE(params[i], "vae_bind: parameter pointer %d is null", i);
    float** d = reinterpret_cast<float**>(&v->W);
    for (int i = 0; i < 42; ++i) d[i] = static_cast<float*>(params[i]);
    if (grads) {
        float** gq = reinterpret_cast<float**>(&v->G);
        for (int i = 0; i < 32; ++i) { MG_REQUIRE(grads[i], "vae_bind: gradient pointer %d is null", i); gq[i] = static_cast<float*>(grads[i]); }
    }
    v->bound = true;
    v->has_grads = grads != nullptr;
    return MG_OK;
}

extern "C" int mg_vae_buffer(mg_vae* v, const char* name, void** ptr, long long* nbytes) {
    MG_REQUIRE(v && name && ptr && nbytes, "vae_buffer: null argument");
    const size_t B = v->B, T = v->T, L0 = v->L0, lat = v->latent, es = v->bf16 ? 2 : 4;
    const struct { const char* n; void* p; size_t b; } tab[] = {
        {"e_x0", v->e_x[0], B * T / 2 * 32 * 4}, {"e_x1", v->e_x[1], B * T / 4 * 64 * 4}, {"e_x2", v->e_x[2], B * L0 * 128 * 4},
        {"e_a0", v->e_a[0], B * T / 2 * 32 * es}, {"e_a1", v->e_a[1], B * T / 4 * 64 * es}, {"e_a2", v->e_a[2], B * L0 * 128 * es},
        {"h", v->h, B * 512 * 4}, {"d0", v->d0, B * 512 * 4}, {"d_y0", v->d_y0, B * L0 * 128 * es},
        {"d_x0", v->d_x[0], B * 2 * L0 * 64 * 4}, {"d_x1", v->d_x[1], B * 4 * L0 * 32 * 4},
        {"d_y1", v->d_y[0], B * 2 * L0 * 64 * es}, {"d_y2", v->d_y[1], B * 4 * L0 * 32 * es},
        {"pre_t", v->pre_t, B * T * 4 * 4}, {"recon", v->recon, B * T * 4 * 4}, {"dt", v->dt, B * T * 4 * 4},
        {"dy_f1", v->dy_f[0], B * 2 * L0 * 64 * 4}, {"dy_f2", v->dy_f[1], B * 4 * L0 * 32 * 4},
        {"dxd1", v->dxd[0], B * 2 * L0 * 64 * es}, {"dxd2", v->dxd[1], B * 4 * L0 * 32 * es}, {"dy0", v->dy0, B * L0 * 128 * es},
        {"dd0", v->dd0, B * 512 * 4}, {"dz", v->dz, B * lat * 4}, {"dmu", v->dmu, B * lat * 4}, {"dlv", v->dlv, B * lat * 4},
        {"dh", v->dh, B * 512 * 4}, {"de_f2", v->de_f[2], B * L0 * 128 * 4}, {"de_f1", v->de_f[1], B * T / 4 * 64 * 4},
        {"de_f0", v->de_f[0], B * T / 2 * 32 * 4}, {"dxe2", v->dxe[2], B * L0 * 128 * es}, {"dxe1", v->dxe[1], B * T / 4 * 64 * es},
        {"dxe0", v->dxe[0], B * T / 2 * 32 * es}};
    for (const auto& e : tab)
        if (!strcmp(e.n, name)) { *ptr = e.p; *nbytes = (long long)e.b; return MG_OK; }
    mg::set_error("vae_buffer: unknown buffer '%s'", name);
    return MG_ERR_INVALID;
}

extern "C" int mg_vae_forward(mg_vae* v, const float* x, const float* eps, int train, float* recon_out, float* z_out,
                              float* mu_out, float* logvar_out, void* stream) {
    MG_REQUIRE(v && v->bound, "vae_forward: context not bound");
    MG_REQUIRE(x && eps, "vae_forward: null pointer");
    mg::tc::set_tf32(v->bf16); mg::tc::set_cache_mode(false);
    return v->bf16 ? vae_forward<__nv_bfloat16>(v, x, eps, train, recon_out, z_out, mu_out, logvar_out, as_stream(stream))
                   : vae_forward<float>(v, x, eps, train, recon_out, z_out, mu_out, logvar_out, as_stream(stream));
}

extern "C" int mg_vae_backward(mg_vae* v, const float* x, const float* drecon, const float* dz, const float* dmu,
                               const float* dlogvar, void* stream) {
    MG_REQUIRE(v && v->bound && v->has_grads, "vae_backward: context needs bound gradients");
    MG_REQUIRE(x && drecon, "vae_backward: null pointer");
    if (!v->fwd_done) { mg::set_error("vae_backward before vae_forward"); return MG_ERR_STATE; }
    mg::tc::set_tf32(v->bf16); mg::tc::set_cache_mode(false);
    return v->bf16 ? vae_backward<__nv_bfloat16>(v, x, drecon, dz, dmu, dlogvar, as_stream(stream))
                   : vae_backward<float>(v, x, drecon, dz, dmu, dlogvar, as_stream(stream));
}

extern "C" int mg_vae_loss_step(mg_vae* v, const float* x, const float* eps, double beta, float* metrics_out, void* stream) {
    MG_REQUIRE(v && v->bound && v->has_grads, "vae_loss_step: context needs bound gradients");
    MG_REQUIRE(x && eps, "vae_loss_step: null pointer");
    cudaStream_t st = as_stream(stream);
    int rc = mg_vae_forward(v, x, eps, 1, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc != MG_OK) return rc;
    // drecon -> v->dt is overwritten inside backward, so stage it in pre_t (no longer needed after tanh)
    vae_loss_kernel<<<1, 1024, 0, st>>>(v->recon, x, (long long)v->B * v->T * 4, v->mu, v->lv, v->B * v->latent, (float)beta,
                                        v->pre_t, v->ext_dmu, v->ext_dlv, metrics_out ? metrics_out : v->metrics);
    MG_LAUNCH_OK();
    return mg_vae_backward(v, x, v->pre_t, nullptr, v->ext_dmu, v->ext_dlv, stream);
}
