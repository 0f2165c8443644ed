// debug.cu -- per-layer harness of the contraction kernels (include/melogan_b200.h: mg_debug_*).
//
// Runs exactly the helper the step bodies of gan.cu call for one layer (conv_fwd, conv_s1_dgrad, upsample2_fwd,
// linear_*, *_wgrad of gan_ctx.cuh), so the dispatch -- tcgen05 weight-stationary / one-tile / CUDA-core -- and the kernel
// variant are the ones the training cycle gets for that shape.  tests/test_tc_layers_gpu.py pins every layer of the benched
// configuration against a float64 contraction of the same bf16-exact operands; scripts/bench_layers.py times them alone.
#include <string.h>

#include "gan_ctx.cuh"

using namespace mg;

namespace {

template <typename TA, typename TO, typename TMSK>
int run_tap(const mg_debug_layer& L, cudaStream_t st) {
    const TA* in = static_cast<const TA*>(L.in);
    TO* out = static_cast<TO*>(L.out);
    switch (L.op) {
        case 0:
            return conv_fwd<TA, TO, TMSK>(in, out, L.W, L.bias, L.R, L.Lin, L.Cin, L.Cout, L.ks, L.stride, L.pad, L.act,
                                          L.col_scale, L.aux, L.mul_src, L.mul_mode, st, L.w_nstride, L.w_kstride, L.pool_out,
                                          L.pool_scale, L.pool_done);
        case 1:
            return conv_s1_dgrad<TA, TO, TMSK>(in, out, L.W, L.R, L.Lin, L.Cin, L.Cout, L.ks, L.pad, L.col_scale, L.mul_src,
                                               L.mul_mode, L.accumulate, st);
        case 2:
            return upsample2_fwd<TA, TO, TMSK>(in, out, L.W, L.bias, L.R, L.Lin, L.Cin, L.Cout, L.w_nstride, L.w_kstride, L.act,
                                               L.mul_src, L.mul_mode, L.accumulate, st, L.colsum_out,
                                               (long long)L.colsum_samples * L.Lin, L.colsum_done, L.stats_out, L.stats_done);
        case 3:
            return linear_fwd<TA, TO>(in, out, L.W, L.bias, L.R, L.Cin, L.Cout, L.act, L.aux, st, L.n_perm_q, L.n_perm_p);
        case 4:
            return linear_dgrad<TA, TO, TMSK>(in, out, L.W, L.R, L.Cin, L.Cout, L.mul_src, L.mul_mode, st, L.n_perm_q,
                                              L.n_perm_p);
    }
    set_error("debug_layer: op %d is not a tap-GEMM", L.op);
    return MG_ERR_INVALID;
}

template <typename TG, typename TA>
int run_wgrad(const mg_debug_layer& L, cudaStream_t st) {
    const TG* g = static_cast<const TG*>(L.in);
    const TA* a = static_cast<const TA*>(L.in2);
    switch (L.op) {
        case 5:
            return conv_wgrad<TG, TA>(g, a, L.dW, 0, (long long)L.R * (L.Lin / L.stride), L.Lin, L.Cin, L.Cout, L.ks, L.stride,
                                      L.pad, st);
        case 6: return convT_wgrad<TG, TA>(g, a, L.dW, L.R, L.Lin, L.Cin, L.Cout, st);
        case 7: return linear_wgrad<TG, TA>(g, a, L.dW, 0, L.R, L.Cin, L.Cout, st, L.n_perm_q, L.n_perm_p);
    }
    set_error("debug_layer: op %d is not a wgrad", L.op);
    return MG_ERR_INVALID;
}

thread_local char g_line[320];

}  // namespace

extern "C" int mg_debug_layer_run(const mg_debug_layer* Lp, void* stream) {
    MG_REQUIRE(Lp, "debug_layer: null descriptor");
    const mg_debug_layer& L = *Lp;
    MG_REQUIRE(L.op >= 0 && L.op <= 7, "debug_layer: op must be 0..7");
    MG_REQUIRE(L.R > 0 && L.Cin > 0 && L.Cout > 0, "debug_layer: bad sizes");
    MG_REQUIRE(L.in && (L.op >= 5 ? (L.in2 && L.dW) : (L.out && L.W)), "debug_layer: null tensor");
    cudaStream_t st = as_stream(stream);
    if (L.op < 5) {   // packed-weight scratch of the tensor-core kernels (a GAN context sizes it in mg_gan_create)
        const size_t taps = L.op == 0 || L.op == 1 ? (size_t)L.ks : (L.op == 2 ? 3 : 1);
        size_t need = 2 * taps * (size_t)L.Cin * L.Cout;
        if (need < (size_t)256 * 64 * 512) need = (size_t)256 * 64 * 512;
        MG_TRY(tc::ensure_scratch(need));
    }
    tc::last_launch() = tc::LaunchInfo();
    tc::set_tf32(L.tf32 != 0); tc::set_cache_mode(false);
    using bf = __nv_bfloat16;
    if (L.op >= 5) {
        if (L.in_bf16 == 1 && L.out_bf16 == 1) return run_wgrad<bf, bf>(L, st);      // out_bf16 = dtype of in2 here
        if (L.in_bf16 == 1 && L.out_bf16 == 0) return run_wgrad<bf, float>(L, st);
        if (L.in_bf16 == 0 && L.out_bf16 == 0) return run_wgrad<float, float>(L, st);
        set_error("debug_layer: unsupported wgrad dtypes");
        return MG_ERR_INVALID;
    }
    const int key = L.in_bf16 * 4 + L.out_bf16 * 2 + (L.mul_src ? L.mask_bf16 : L.out_bf16);
    switch (key) {
        case 7: return run_tap<bf, bf, bf>(L, st);
        case 5: return run_tap<bf, float, bf>(L, st);
        case 4: return run_tap<bf, float, float>(L, st);
        case 0: return run_tap<float, float, float>(L, st);
        case 3: return run_tap<float, bf, bf>(L, st);
    }
    set_error("debug_layer: unsupported dtype combination (in %d, out %d, mask %d)", L.in_bf16, L.out_bf16, L.mask_bf16);
    return MG_ERR_INVALID;
}

extern "C" int mg_debug_set(const char* key, int value) {
    MG_REQUIRE(key, "debug_set: null key");
    tc::Tuning& t = tc::tuning();
    if (!strcmp(key, "reset")) { t = tc::Tuning(); return MG_OK; }
    if (!strcmp(key, "force_bn")) { t.force_bn = value; return MG_OK; }
    if (!strcmp(key, "max_stages")) { t.max_stages = value; return MG_OK; }
    if (!strcmp(key, "staging_bufs")) { t.staging_bufs = value; return MG_OK; }
    if (!strcmp(key, "mask_bufs")) { t.mask_bufs = value; return MG_OK; }
    if (!strcmp(key, "no_ws")) { t.no_ws = value; return MG_OK; }
    if (!strcmp(key, "no_pair")) { t.no_pair = value; return MG_OK; }
    if (!strcmp(key, "no_rot")) { t.no_rot = value; return MG_OK; }
    if (!strcmp(key, "fp32_tc")) { t.fp32_tc = value; return MG_OK; }
    if (!strcmp(key, "no_fuse")) { t.no_fuse = value; return MG_OK; }
    if (!strcmp(key, "ring2_stages")) { t.ring2_stages = value; return MG_OK; }
    if (!strcmp(key, "dbg")) { t.dbg = value; return MG_OK; }
    if (!strcmp(key, "reverse")) { t.reverse = value; return MG_OK; }
    if (!strcmp(key, "no_tma_store")) { t.no_tma_store = value; return MG_OK; }
    if (!strcmp(key, "no_tma_mask")) { t.no_tma_mask = value; return MG_OK; }
    if (!strcmp(key, "no_reuse")) { t.no_reuse = value; return MG_OK; }
    set_error("debug_set: unknown key '%s'", key);
    return MG_ERR_INVALID;
}

extern "C" const char* mg_debug_last_launch(void) {
    const tc::LaunchInfo& li = tc::last_launch();
    if (li.kind == 0) { g_line[0] = 0; return g_line; }
    if (li.kind == 2)
        snprintf(g_line, sizeof(g_line), "tc_wgrad rows=%lld N=%d K=%d taps=%d BNK=%d splits=%d fp32x6=%d flops=%.6e bytes=%.6e", li.rows,
                 li.N, li.K, li.taps, li.BN, li.splits, li.fp32x6, li.flops, li.bytes);
    else
        snprintf(g_line, sizeof(g_line),
                 "tc_tap rows=%lld N=%d K=%d taps=%d groups=%d halo=%d BN=%d out%d ws=%d stages=%d act=%d mul=%d aux=%d "
                 "tma_store=%d tma_mask=%d nsb=%d reverse=%d grid=%dx%d tf32=%d pair=%d pool=%d fp32x6=%d flops=%.6e bytes=%.6e",
                 li.rows, li.N, li.K, li.taps, li.groups, li.halo, li.BN, li.out_bytes, li.ws, li.stages, li.act, li.mul, li.aux,
                 li.tma_store, li.tma_mask, li.nsb, li.reverse, li.ctas_x, li.slabs, li.tf32, li.pair, li.pool, li.fp32x6, li.flops, li.bytes);
    return g_line;
}
