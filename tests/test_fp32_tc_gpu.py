"""float32 contractions on the bf16 tensor cores (csrc/gemm_tc.cuh, try_split_tapgemm / try_split_wgrad: every float32
operand is split exactly into three bf16 parts, six part products are accumulated in the fp32 TMEM accumulator by the same
tcgen05 kernels bf16 mode uses).  Opt-in: mg_debug_set("fp32_tc", 1) or MELOGAN_FP32_TC=1.

Every layer of the training cycle in float32 storage with FULL-precision random operands against the float64 contraction,
the CUDA-core result of the same launch measured beside it (both land in gpurun_out/fp32_tc_layers.jsonl).  The operand
split is exact and the dropped part products are O(2^-26); what remains is the tensor core's own accumulation: tcgen05 adds
into the fp32 TMEM accumulator with truncation, up to one ulp OF THE ACCUMULATOR per instruction and always towards zero
(measured ~0.5 ulp each), where the CUDA cores' FFMA chain rounds to nearest.  The kernels therefore issue the five small-
term segments of all taps first, while the accumulator is 2^-8 of its final size, and the full-size hh segment last; the
error is then that of the hh chain alone (taps x K / 16 instructions).  Bar per layer: max(1e-5 of the tensor's scale,
n 2^-23) with n = 6 x taps x K / 16 (the worst case had ALL six segments run on a full-size accumulator; measured:
3e-7..2.7e-6, at or below the CUDA-core kernels, 1.9e-5 for the K = 16384 reduction of pre.2's dgrad); weight gradients
1e-4 (as for bf16 mode; measured <= 9e-6).  Then the whole critic and generator steps of the fp32 parity mode with the
switch on against the switch off: losses to 2e-5, gradients to 5e-3 (measured: critic step 4.5e-5 worst, generator step
5e-4 on its first layer, where BatchNorm backward's cancellations amplify the residual bias of the hh chains)."""
import json
import os

import pytest
import torch

import tc_layers as TL
from gan_testlib import assert_close_l2, cuda_batch, make_engine
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

B = int(os.environ.get("MELOGAN_TEST_FP32_TC_BATCH", "1024"))
SPECS = TL.fp32_layers(B)
F32_TOL, WGRAD_TOL = 1e-5, 1e-4


def _err(layer):
    s = layer.s
    if s["op"] >= 5:
        return TL.compare(layer.dW, layer.reference_wgrad(), False)[0]
    R = s["R"]
    chunk = max(1, min(R, (1 << 23) // max(1, layer.out[0].numel())))
    worst = 0.0
    for r0 in range(0, R, chunk):
        y, aux = layer.reference(r0, min(R, r0 + chunk))
        worst = max(worst, TL.compare(layer.out[r0:r0 + chunk], y, False)[0])
        if aux is not None:
            worst = max(worst, TL.compare(layer.aux[r0:r0 + chunk], aux, False)[0])
    return worst


def _time(layer, iters=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    layer.run()
    e0.record()
    for _ in range(iters):
        layer.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


@pytest.mark.parametrize("spec", SPECS, ids=[s["name"] for s in SPECS])
def test_fp32_layer_on_tensor_cores_matches_float64(spec):
    s = spec
    layer = TL.Layer(spec, seed=23)
    rec = {"name": s["name"], "B": B}
    try:
        for mode in (0, 1):
            TL.debug_set("reset", 0)
            TL.debug_set("fp32_tc", mode)
            info = layer.run()
            torch.cuda.synchronize()
            err = _err(layer)
            ms = _time(layer)
            rec["tc" if mode else "cuda_core"] = {"rel_err": err, "ms": ms, "info": info}
            if mode:
                assert info.startswith("tc_") and "fp32x6=1" in info, f"{s['name']} did not take the six-term path: '{info}'"
                if s["op"] >= 5:
                    tol = WGRAD_TOL
                else:
                    taps = {0: s["ks"], 1: s["ks"], 2: 3}.get(s["op"], 1)       # op 2: up to 3 taps per sub-pixel phase
                    kred = s["Cout"] if s["op"] in (1, 4) else s["Cin"]
                    tol = max(F32_TOL, (6 * taps * kred / 16) * 2.0 ** -23)      # one ulp per instruction: the worst case
                assert err <= tol, f"{s['name']}: rel err {err:.3e} > {tol:.1e} ({info})"
    finally:
        TL.debug_set("reset", 0)
        print(json.dumps(rec))
        if os.path.isdir("gpurun_out"):
            with open("gpurun_out/fp32_tc_layers.jsonl", "a") as f:
                f.write(json.dumps(rec) + "\n")


def _step(Bs, which, on):
    TL.debug_set("reset", 0)
    TL.debug_set("fp32_tc", 1 if on else 0)
    try:
        params = O.make_params(4, fan_in_scale=True)
        batch = O.make_batch(40, Bs)
        eng, cp, grads = make_engine(Bs, params, precision="fp32")
        cb = cuda_batch(batch)
        if which == "d":
            m = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"])
            g = {k: v.clone() for k, v in grads["D"].items()}
        else:
            m = eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"])
            g = {k: v.clone() for k, v in grads["G"].items()}
            g.update({"E." + k: v.clone() for k, v in grads["E"].items()})
        torch.cuda.synchronize()
        return m.clone(), g
    finally:
        TL.debug_set("reset", 0)


@pytest.mark.parametrize("which", ["d", "g"])
def test_fp32_step_on_tensor_cores_stays_close_to_cuda_core_step(which):
    """fp32 parity mode, B = 160: losses and every parameter gradient with the six-term tensor-core contractions against the
    CUDA-core FFMA kernels (see the module docstring for the bars)."""
    from gan_testlib import rel_l2
    m0, g0 = _step(160, which, False)
    m1, g1 = _step(160, which, True)
    # (the biases in front of a train-mode BatchNorm have a true gradient of zero: both paths return cancellation noise there)
    noise = ("decoder.deconv.0.bias", "decoder.deconv.3.bias")
    errs = {k: rel_l2(g1[k], g0[k]) for k in g0 if k not in noise}
    rec = {"step": which, "B": 160, "losses_cuda_core": m0.tolist(), "losses_tc": m1.tolist(),
           "worst_grad": max(errs, key=errs.get), "worst_grad_rel_l2": max(errs.values()),
           "median_grad_rel_l2": sorted(errs.values())[len(errs) // 2]}
    print(json.dumps(rec))
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/fp32_tc_layers.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
    assert torch.allclose(m1, m0, rtol=2e-5, atol=1e-6), (m1, m0)
    for k in errs:
        assert_close_l2(g1[k], g0[k], 5e-3, f"{which}: grad {k}")
