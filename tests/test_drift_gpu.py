"""Free-running parameter drift over 100 iterations (north_star: "parameter drift tracked over 100 steps").

The reference loop (src/gan/train_gan.py:159-251) at its own settings -- B = 32, CRITIC_ITERS = 5, Adam(0.5, 0.9) --
run for 100 batches = 100 critic steps + 20 generator steps, once by the oracle on the CPU and once by the CUDA path,
from the same initial parameters and the same per-step noise / alpha / dropout draws, WITHOUT resynchronising.
Adam divides by sqrt(v), so rounding differences in gradients that are themselves rounding noise become +-lr
updates: the reference drifts from ITSELF when only its thread count changes -- critic loss
relative difference 0 / 4.8e-7 / 9.7e-6 / 2.4e-5 at steps 0 / 24 / 49 / 99 and parameter relative L2 1.63e-2
(SURVEY.md appendix A), and it is 1.1e-8 / 8.1e-6 / 1.2e-3 away (steps 0 / 49 / 99) from the same model run in float64.
fp32 mode: the first ten steps (where the reference's self-drift is exactly zero) must agree to 1e-6, the parameters must
stay within 3x the self-drift, and the critic loss within the reference's own fp32-vs-float64 distance over the window;
the values at steps 0 / 24 / 49 / 99 are reported next to the self-drift numbers.  bf16 mode (north_star tolerance 1e-2 per
step) is tracked and has to stay within 1e-2 on the losses.  The numbers are written to gpurun_out/drift.json.
"""
import json
import os

import pytest
import torch

from gan_testlib import cuda_batch, make_flat_trainer
from melogan import engine as E
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELF_DRIFT = {0: 0.0, 24: 4.8e-7, 49: 9.7e-6, 99: 2.4e-5}      # the reference against itself (8 threads vs 1 thread)
SELF_PARAM_L2 = 1.63e-2
FP64_DRIFT = {0: 1.1e-8, 49: 8.1e-6, 99: 1.2e-3}                # the fp32 reference against the same model in float64
NOISE_BIASES = ("decoder.deconv.0.bias", "decoder.deconv.3.bias")   # zero true gradient in front of BatchNorm: pure noise


def _run_oracle(params, B, iters):
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    st_d, st_g, losses, g_losses = {}, {}, [], []
    for it in range(iters):
        batch = O.make_batch(2000 + it, B)
        losses.append(O.critic_step(params, batch, st_d)["loss_d"].item())
        if (it + 1) % O.CFG["CRITIC_ITERS"] == 0:
            o = O.generator_step(params, batch, st_g)
            g_losses.append((o["loss_g_adv"].item(), o["loss_g_emo"].item()))
    return losses, g_losses


def _run_cuda(params, B, iters, precision):
    T = make_flat_trainer(B, params, precision, lr_d=O.CFG["LR_D"], lr_g=O.CFG["LR_G"], betas=(O.CFG["BETA1"], O.CFG["BETA2"]))
    eng, losses, g_losses = T["eng"], [], []
    for it in range(iters):
        cb = cuda_batch(O.make_batch(2000 + it, B))
        T["optD"].zero_grad()
        m = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"])
        T["optD"].step()
        losses.append(m[0].item())
        if (it + 1) % O.CFG["CRITIC_ITERS"] == 0:
            T["optG"].zero_grad()
            g = eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"])
            T["optG"].step()
            g_losses.append((g[0].item(), g[1].item()))
    return T, losses, g_losses


def _param_drift(T, oparams):
    num = den = 0.0
    per = {}
    for mod, keys in (("D", E.D_KEYS), ("G", E.G_PARAM_KEYS), ("E", E.E_KEYS)):
        for k in keys:
            a, b = T[mod][k].detach().double().cpu(), oparams[mod][k].double()
            per[f"{mod}.{k}"] = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
            if k in NOISE_BIASES:
                continue
            num += (a - b).pow(2).sum().item()
            den += b.pow(2).sum().item()
    return (num / den) ** 0.5, per


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_hundred_step_drift(precision):
    B, iters = 32, 100
    params = O.make_params(1)
    oparams = O.clone_params(params)
    ref_losses, ref_g = _run_oracle(oparams, B, iters)
    T, losses, g_losses = _run_cuda(params, B, iters, precision)
    all_rel = [abs(a - b) / abs(b) for a, b in zip(losses, ref_losses)]
    rel = {s: all_rel[s] for s in (0, 9, 24, 49, 99)}
    g_rel = [abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(g_losses, ref_g)]
    l2, per = _param_drift(T, oparams)
    worst = sorted(per.items(), key=lambda kv: -kv[1])[:5]
    rec = {"precision": precision, "B": B, "iterations": iters, "loss_d_rel_diff": rel, "loss_d_rel_diff_max": max(all_rel), "g_emo_rel_diff_max": max(g_rel),
           "param_rel_l2": l2, "largest_per_tensor": worst, "reference_self_drift": {"loss_d": SELF_DRIFT, "param_rel_l2": SELF_PARAM_L2},
           "reference_fp32_vs_fp64": FP64_DRIFT,
           "loss_d_first_last": [losses[0], losses[-1]], "ref_loss_d_first_last": [ref_losses[0], ref_losses[-1]]}
    print(json.dumps(rec))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"drift_{precision}.json"), "w") as f:
            json.dump(rec, f, indent=1)
    except OSError:
        pass
    if precision == "fp32":
        # (a) no systematic error: before the dynamics amplify anything (the reference's self-drift is exactly 0 through
        #     step 9) every step agrees to fp32 rounding of the loss
        assert max(all_rel[:10]) <= 1e-6, all_rel[:10]
        # (b) chaotic regime: the CUDA path's own run-to-run variation (fp32 atomics in the weight-gradient and column
        #     reductions reorder sums) moves the value at a given step by up to ~1e-4, about 5x the reference's 8-vs-1
        #     thread self-drift at step 99, so single steps are REPORTED against SELF_DRIFT and the bound asserted is the
        #     reference's own distance from exact arithmetic over the window (fp32 vs float64: 1.2e-3 at step 99)
        assert max(all_rel) <= FP64_DRIFT[99], (max(all_rel), rel)
        # (c) parameters: within 3x the reference's self-drift (measured: below 1x)
        assert l2 <= 3 * SELF_PARAM_L2, l2
    else:
        assert max(all_rel) <= 1e-2, (max(all_rel), rel)
        assert max(g_rel) <= 1e-2, g_rel
        assert l2 <= 0.25, l2
