"""Every tensor-core contraction of the BENCHED configuration (bf16 mode, per-GPU batch 8192: the grids, slab widths,
halo tap groups, tile orders, staging rings and wgrad splits the bench actually runs) against a float64 contraction of
the same bf16-exact operands, through the C ABI (mg_debug_layer_run calls the very helper the step bodies call).

Bars: float32 results (pre-BatchNorm activations, TF32 Linears) within 1e-5 of the tensor's scale (x sqrt(K/1024) for
the 16384-long reduction of decoder.pre.2's dgrad), weight gradients (fp32 atomics over up to 3 M rows) within 1e-4 --
only the fp32 accumulation order differs; bf16 results within ONE bf16 ulp of the float64 value (measured: 0.50, i.e.
correctly rounded).  A wrong halo row, tap
or tile edge is an O(1) error and cannot hide here (the whole-step bf16 tests allow 4-8 % for mask flips)."""
import os

import pytest
import torch

import tc_layers as TL

pytestmark = pytest.mark.gpu

B = int(os.environ.get("MELOGAN_TEST_LAYER_BATCH", "8192"))
SPECS = TL.cycle_layers(B)
F32_TOL, WGRAD_TOL, ULP_TOL = 1e-5, 1e-4, 1.0


def _check(layer, info):
    s = layer.s
    if s["op"] >= 5:
        ref = layer.reference_wgrad()
        rel, _ = TL.compare(layer.dW, ref, False)
        assert rel <= WGRAD_TOL, f"{s['name']}: wgrad rel err {rel:.3e} ({info})"
        return rel, 0.0
    R = s["R"]
    chunk = max(1, min(R, (1 << 24) // max(1, layer.out[0].numel())))
    worst_rel = worst_ulp = 0.0
    stored = bool(s["out_bf16"])
    for r0 in range(0, R, chunk):
        r1 = min(R, r0 + chunk)
        y, aux = layer.reference(r0, r1)
        rel, ulp = TL.compare(layer.out[r0:r1], y, stored)
        worst_rel, worst_ulp = max(worst_rel, rel), max(worst_ulp, ulp)
        if aux is not None:
            rel, ulp = TL.compare(layer.aux[r0:r1], aux, stored)
            worst_rel, worst_ulp = max(worst_rel, rel), max(worst_ulp, ulp)
    if stored:
        assert worst_ulp <= ULP_TOL, f"{s['name']}: {worst_ulp:.2f} bf16 ulps off (rel {worst_rel:.2e}) ({info})"
    else:
        # fp32 accumulation noise grows with the reduction length (decoder.pre.2's dgrad sums 16384 products)
        ktot = s["Cin"] * max(s["ks"], 1) if s["op"] != 4 else s["Cout"]
        tol = F32_TOL * max(1.0, (ktot / 1024.0) ** 0.5)
        assert worst_rel <= tol, f"{s['name']}: rel err {worst_rel:.3e} > {tol:.1e} ({info})"
    return worst_rel, worst_ulp


@pytest.mark.parametrize("spec", SPECS, ids=[s["name"] for s in SPECS])
def test_layer_matches_float64(spec):
    TL.debug_set("reset", 0)
    layer = TL.Layer(spec, seed=11)
    for reverse in (0, 1):                      # both tile orders of the snake walk
        TL.debug_set("reverse", reverse)
        info = layer.run()
        torch.cuda.synchronize()
        assert info.startswith("tc_"), f"{spec['name']} did not run on the tensor cores: '{info}'"
        rel, ulp = _check(layer, info)
        print(f"{spec['name']:18s} reverse={reverse} rel={rel:.2e} ulp={ulp:.2f}  {info}")
    TL.debug_set("reset", 0)


@pytest.mark.parametrize("name,knobs", [
    ("D.conv4.fwd", {"force_bn": 64}), ("D.conv4.fwd", {"staging_bufs": 1}), ("D.conv2.fwd", {"staging_bufs": 1}),
    ("D.conv2.fwd", {"no_reuse": 1}), ("ED.conv2.fwd", {"force_bn": 64}), ("ED.conv1.fwd", {"no_tma_store": 1}),
    ("D.conv4.dgrad", {"no_tma_mask": 1}), ("D.conv2.dgrad", {"no_ws": 1}), ("ED.conv3.dgrad", {"max_stages": 3}),
    # one CTA per tile row where the heuristics now take CTA pairs (cta_group::2), and pairs with thin rings
    ("D.conv4.fwd", {"no_pair": 1}), ("ED.conv3.fwd", {"no_pair": 1}), ("D.conv4.dgrad", {"no_pair": 1}),
    ("G.pre2.fwd", {"no_pair": 1}), ("ED.conv2.fwd", {"max_stages": 2}), ("D.conv2.fwd", {"mask_bufs": 1, "staging_bufs": 1}),
    ("D.conv2.fwd", {"no_pair": 1}), ("ED.conv1.fwd", {"no_pair": 1}), ("D.conv2.adjoint", {"no_pair": 1}),
])
def test_kernel_variants_match_float64(name, knobs):
    """The other kernel variants the heuristics can pick at other batch sizes (narrow slabs, single staging tile, one
    tile per CTA, per-thread mask loads and stores, no tap reuse) on the same layers."""
    spec = dict(next(s for s in TL.cycle_layers(min(B, 2048)) if s["name"] == name))
    TL.debug_set("reset", 0)
    for k, v in knobs.items():
        TL.debug_set(k, v)
    try:
        layer = TL.Layer(spec, seed=5)
        info = layer.run()
        torch.cuda.synchronize()
        assert info.startswith("tc_"), info
        rel, ulp = _check(layer, info)
        print(f"{name:18s} {knobs} rel={rel:.2e} ulp={ulp:.2f}  {info}")
    finally:
        TL.debug_set("reset", 0)


@pytest.mark.parametrize("name", ["D.conv4.fwd", "D.conv4.adjoint", "ED.conv3.fwd", "D.conv4.fwd.B"])
def test_fused_pooling_matches_mean_of_stored_output(name):
    """The weight-stationary kernels also emit AdaptiveAvgPool1d(1) of the tile they store (ws_pool_*: column sums of the
    staging tile; complete per tile for the critic's 64-row samples, atomics over the 4 tiles of an ED sample): it must
    equal the float64 mean of the bf16 values the same launch stored, to float32 summation noise, in both tile orders and on
    CTA pairs as well as single CTAs."""
    spec = dict(next(s for s in SPECS if s["name"] == name))
    for knobs in ({}, {"no_pair": 1}):
        TL.debug_set("reset", 0)
        for k, v in knobs.items():
            TL.debug_set(k, v)
        try:
            layer = TL.Layer(spec, seed=13)
            layer.enable_pool()
            for reverse in (0, 1):
                TL.debug_set("reverse", reverse)
                layer.pool.fill_(float("nan"))
                layer.pool_done[0] = 0
                info = layer.run()
                torch.cuda.synchronize()
                if knobs and "pool=0" in info:           # a single CTA has no room for conv.4's staging tile / takes 64-wide
                    assert layer.pool_done[0] == 0, info  # slabs for ED conv.3: no fused pooling, the caller is told so and
                    continue                              # runs pool_rows_kernel
                assert "ws=1" in info and "pool=1" in info and layer.pool_done[0] == 1, info
                want = layer.out.double().mean(dim=1)
                err = ((layer.pool.double() - want).abs().max() / want.abs().max()).item()
                assert err < 2e-6, f"{name} {knobs} reverse={reverse}: fused pooling off by {err:.2e} ({info})"
                _check(layer, info)                      # and the stored tile itself is unchanged
        finally:
            TL.debug_set("reset", 0)


@pytest.mark.parametrize("name", ["D.conv4.dgrad", "D.conv2.dgrad"])
def test_fused_bias_gradient_column_sums(name):
    """The dgrad epilogues also add the column sums of the tensor they store, over the first Rb samples, into the bias
    gradient of the layer below (ws_colsum_*: per-CTA running sums of the staging tiles, one atomicAdd per column and CTA
    at the end; the 128-wide pair kernel and the 64-wide single-CTA kernel).  Against the float64 sum of the stored bf16
    values; accumulates (+=) like the colreduce it replaces."""
    spec = dict(next(s for s in SPECS if s["name"] == name))
    TL.debug_set("reset", 0)
    try:
        layer = TL.Layer(spec, seed=17)
        R = spec["R"]
        Rb = (2 * R) // 3                              # the critic step: real + fake rows of the 3B-row batch
        layer.enable_colsum(Rb)
        for reverse in (0, 1):
            TL.debug_set("reverse", reverse)
            layer.colsum.fill_(1.5)                    # += semantics
            layer.colsum_done[0] = 0
            info = layer.run()
            torch.cuda.synchronize()
            assert "ws=1" in info and "pool=2" in info and layer.colsum_done[0] == 1, info
            want = layer.out[:Rb].double().sum(dim=(0, 1)) + 1.5
            err = ((layer.colsum.double() - want).abs().max() / want.abs().max()).item()
            assert err < 5e-6, f"{name} reverse={reverse}: fused column sums off by {err:.2e} ({info})"
            _check(layer, info)
    finally:
        TL.debug_set("reset", 0)


@pytest.mark.parametrize("name", ["G.deconv0.fwd", "G.deconv3.fwd"])
def test_fused_batchnorm_statistics(name):
    """The generator's transposed-conv epilogues also sum x and x^2 per channel of the float32 tile they store (ws_stats_*)
    for the train-mode BatchNorm that follows: against float64 sums of the stored values (what colreduce COL_SUM_SQ gave)."""
    spec = dict(next(s for s in SPECS if s["name"] == name))
    TL.debug_set("reset", 0)
    try:
        layer = TL.Layer(spec, seed=19)
        layer.enable_stats()
        for reverse in (0, 1):
            TL.debug_set("reverse", reverse)
            layer.stats.zero_()
            layer.stats_done[0] = 0
            info = layer.run()
            torch.cuda.synchronize()
            assert "ws=1" in info and "pool=3" in info and layer.stats_done[0] == 1, info
            x = layer.out.double()
            want = torch.cat([x.sum(dim=(0, 1)), (x * x).sum(dim=(0, 1))])
            C = x.shape[2]
            got = layer.stats.double()
            e1 = ((got[:C] - want[:C]).abs().max() / want[:C].abs().max().clamp_min(1e-30)).item()
            e2 = ((got[C:] - want[C:]).abs().max() / want[C:].abs().max()).item()
            assert e2 < 2e-6 and e1 < 1e-4, f"{name} reverse={reverse}: fused statistics off by {e1:.2e} / {e2:.2e} ({info})"
            _check(layer, info)
    finally:
        TL.debug_set("reset", 0)
