"""Step-level parity (fp32 mode): the two composites of train_gan.py:183-251 plus the fused Adam,
teacher-forced per step against the oracle (hard part 3)."""
import os

import numpy as np
import pytest
import torch

from gan_testlib import MODS, assert_close, cuda_batch, make_engine, to_double
from melogan import engine as E
from melogan.optim import FlatParams, FusedAdam
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5
GRAD_TOL = 5e-5
NOISE_BIASES = ("decoder.deconv.0.bias", "decoder.deconv.3.bias")


def assert_param_step(p_new, p_ref_new, g_ref, lr, what):
    """Post-Adam parameters.  Adam divides by sqrt(v): where |g| is at the level of its own rounding error
    the update is +-lr times noise on BOTH sides, so elements whose reference gradient is
    below 100x the gradient tolerance only have to stay within one Adam step (2.2*lr); the rest must agree
    to 2e-5 of the tensor's scale."""
    p_new, p_ref_new, g_ref = p_new.detach().cpu().double(), p_ref_new.detach().cpu().double(), g_ref.detach().cpu().double()
    diff = (p_new - p_ref_new).abs()
    strong = g_ref.abs() > 100 * GRAD_TOL * g_ref.abs().max()
    scale = p_ref_new.abs().max().item()
    assert diff.max().item() <= 2.2 * lr, (what, "bounded by one Adam step", diff.max().item())
    if strong.any() and not any(what.endswith(nb) for nb in NOISE_BIASES):   # those gradients are noise everywhere
        assert diff[strong].max().item() <= 2e-5 * scale + 0.02 * lr, (what, diff[strong].max().item(), scale)


@pytest.mark.parametrize("B,fan,pseed", [(8, False, 1), (8, True, 4), (32, False, 2), (3, True, 5)])
def test_critic_step_matches_oracle(B, fan, pseed):
    params = O.make_params(pseed, fan_in_scale=fan)
    batch = O.make_batch(10 * pseed, B)
    eng, cp, grads = make_engine(B, params)
    cb = cuda_batch(batch)
    ref = O.critic_step(O.clone_params(params), batch, {}, update=False)
    ref64 = O.critic_step(to_double(params), to_double(batch), {}, update=False)
    m = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"]).cpu()
    want = [ref["loss_d"].item(), ref["gp"].item(), ref["d_real"].mean().item(), ref["d_fake"].mean().item()]
    assert abs(m[0].item() - want[0]) <= LOSS_TOL * abs(want[0]), ("loss_d", m[0].item(), want[0])
    assert abs(m[1].item() - want[1]) <= LOSS_TOL * max(abs(want[1]), 1.0), ("gp", m[1].item(), want[1])
    sc = max(ref["d_real"].abs().max().item(), ref["d_fake"].abs().max().item())
    assert abs(m[2].item() - want[2]) <= LOSS_TOL * sc and abs(m[3].item() - want[3]) <= LOSS_TOL * sc
    assert_close(eng.buffer("g.notes").view(B, 512, 4), ref["fake"], 1e-5, "fake notes")
    for k, g in ref["grads"].items():
        if k == "real_fake.bias":
            assert grads["D"][k].abs().max().item() <= 1e-6
        elif k == "real_fake.weight":
            assert_close(grads["D"][k][:, :256], g[:, :256], GRAD_TOL, k, ref64["grads"][k][:, :256])
        else:
            assert_close(grads["D"][k], g, GRAD_TOL, "D grad " + k, ref64["grads"][k])


def test_golden_reference_losses_case_A(golden_dir):
    """Same inputs as the golden run of the REFERENCE modules (oracle/make_golden_gan.py, case A step 0)."""
    gold = np.load(os.path.join(golden_dir, "gan_golden.npz"))
    params = O.make_params(1)
    batch = O.make_batch(10, 8)
    eng, cp, grads = make_engine(8, params)
    cb = cuda_batch(batch)
    m = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"]).cpu().numpy()
    np.testing.assert_allclose(m, gold["A.d0.scalars"], rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(eng.buffer("g.notes").view(8, 512, 4)[0].cpu().numpy(), gold["A.d0.fake0"],
                               rtol=0, atol=1e-5 * np.abs(gold["A.d0.fake0"]).max())


@pytest.mark.parametrize("B,fan,pseed", [(8, False, 1), (8, True, 4), (32, False, 2)])
def test_generator_step_matches_oracle(B, fan, pseed):
    params = O.make_params(pseed, fan_in_scale=fan)
    batch = O.make_batch(10 * pseed + 3, B)
    eng, cp, grads = make_engine(B, params)
    cb = cuda_batch(batch)
    ref = O.generator_step(O.clone_params(params), batch, {}, update=False)
    ref64 = O.generator_step(to_double(params), to_double(batch), {}, update=False)
    m = eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"]).cpu()
    assert abs(m[0].item() - ref["loss_g_adv"].item()) <= LOSS_TOL * max(abs(ref["loss_g_adv"].item()), 1e-2)
    assert abs(m[1].item() - ref["loss_g_emo"].item()) <= LOSS_TOL * abs(ref["loss_g_emo"].item())
    assert_close(eng.buffer("g.notes").view(B, 512, 4), ref["notes"], 1e-5, "notes")
    assert_close(eng.buffer("ed.logits")[:B * 4].view(B, 4), ref["logits"], 1e-5, "logits")
    for k, g in ref["grads_G"].items():
        if k in NOISE_BIASES:
            continue
        assert_close(grads["G"][k], g, GRAD_TOL, "G grad " + k, ref64["grads_G"][k])
    for k, g in ref["grads_E"].items():
        assert_close(grads["E"][k], g, GRAD_TOL, "E grad " + k, ref64["grads_E"][k])


def test_full_cycle_with_fused_adam_teacher_forced():
    """5 critic steps + 1 generator step (CRITIC_ITERS=5), parameters resynchronised to the oracle
    before every step; after each step the updated parameters must agree."""
    B = 8
    params = O.make_params(6, fan_in_scale=True)
    oparams = O.clone_params(params)
    eng = E.GanEngine(B)
    cp = {m: {k: v.clone().cuda() for k, v in P.items()} for m, P in params.items()}
    # flat parameter buffers, as the trainer keeps them
    leafD = [torch.nn.Parameter(cp["D"][k]) for k in E.D_KEYS]
    leafG = [torch.nn.Parameter(cp["G"][k]) for k in E.G_PARAM_KEYS] + [torch.nn.Parameter(cp["E"][k]) for k in E.E_KEYS]
    flatD, flatG = FlatParams(leafD), FlatParams(leafG)
    optD = FusedAdam(flatD, lr=1e-4, betas=(0.5, 0.9))
    optG = FusedAdam(flatG, lr=2e-4, betas=(0.5, 0.9))
    Dp = {k: p.data for k, p in zip(E.D_KEYS, leafD)}
    Dg = {k: p.grad for k, p in zip(E.D_KEYS, leafD)}
    nG = len(E.G_PARAM_KEYS)
    Gp = {k: p.data for k, p in zip(E.G_PARAM_KEYS, leafG[:nG])}
    Gp.update({k: cp["G"][k] for k in E.G_BUFFER_KEYS})
    Gg = {k: p.grad for k, p in zip(E.G_PARAM_KEYS, leafG[:nG])}
    Ep = {k: p.data for k, p in zip(E.E_KEYS, leafG[nG:])}
    Eg = {k: p.grad for k, p in zip(E.E_KEYS, leafG[nG:])}
    eng.bind(E.MOD_E, Ep, Eg); eng.bind(E.MOD_G, Gp, Gg); eng.bind(E.MOD_D, Dp, Dg); eng.bind(E.MOD_ED, cp["ED"], None)
    st_d, st_g = {}, {}

    def resync():
        for k in E.D_KEYS: Dp[k].copy_(oparams["D"][k])
        for k in E.G_PARAM_KEYS + E.G_BUFFER_KEYS: Gp[k].copy_(oparams["G"][k])
        for k in E.E_KEYS: Ep[k].copy_(oparams["E"][k])

    for i in range(5):
        batch = O.make_batch(60 + i, B)
        cb = cuda_batch(batch)
        resync()
        # optimizer moments follow the oracle too (teacher forcing)
        for k, p in zip(E.D_KEYS, leafD):
            if k in st_d:
                o = flatD.offset_of(p)
                optD.exp_avg[o:o + p.numel()].copy_(st_d[k][0].flatten()); optD.exp_avg_sq[o:o + p.numel()].copy_(st_d[k][1].flatten())
        ref = O.critic_step(oparams, batch, st_d)
        optD.zero_grad()
        m = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"]).cpu()
        optD.step()
        assert abs(m[0].item() - ref["loss_d"].item()) <= LOSS_TOL * abs(ref["loss_d"].item()), i
        for k in E.D_KEYS:
            assert_param_step(Dp[k], oparams["D"][k], ref["grads"][k], 1e-4, f"step {i} D param {k}")
        for k in E.G_BUFFER_KEYS:
            assert_close(Gp[k], oparams["G"][k], 1e-5, f"step {i} running stat {k}")
    resync()
    ref = O.generator_step(oparams, batch, st_g)
    optG.zero_grad()
    m = eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"]).cpu()
    optG.step()
    assert abs(m[1].item() - ref["loss_g_emo"].item()) <= LOSS_TOL * abs(ref["loss_g_emo"].item())
    for k in E.G_PARAM_KEYS:
        assert_param_step(Gp[k], oparams["G"][k], ref["grads_G"][k], 2e-4, "G param " + k)
    for k in E.E_KEYS:
        assert_param_step(Ep[k], oparams["E"][k], ref["grads_E"][k], 2e-4, "E param " + k)
