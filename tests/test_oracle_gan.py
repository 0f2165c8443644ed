"""Pins oracle/gan_oracle.py against tests/golden/gan_golden.npz (= the reference's own modules and the
verbatim loop body of src/gan/train_gan.py:183-251, run by oracle/make_golden_gan.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import gan_oracle as O


def tstats(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, steps=min(8, t.numel())).long()
    return np.concatenate([[t.sum().item(), t.norm().item(), t.abs().max().item()], t[idx].numpy()])


def close_stats(got, want, rtol, what):
    scale = max(abs(want[2]), 1e-30)            # max |x| of the tensor: errors are judged against its scale
    assert abs(got[1] - want[1]) <= rtol * max(want[1], 1e-30), (what, "norm", got[1], want[1])
    assert np.all(np.abs(got[3:] - want[3:]) <= rtol * scale), (what, "samples", got[3:], want[3:])
    assert abs(got[0] - want[0]) <= rtol * scale * 2048, (what, "sum", got[0], want[0])


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "gan_golden.npz"))


@pytest.mark.parametrize("case,B,pseed,nsteps", [("A", 8, 1, 5), ("B", 32, 2, 1), ("C", 8, 4, 2)])
def test_cycle_matches_reference(gold, case, B, pseed, nsteps):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    params = O.make_params(pseed, fan_in_scale=(case == "C"))
    st_d, st_g = {}, {}
    for i in range(nsteps):
        batch = O.make_batch(10 * pseed + i, B)
        d = O.critic_step(params, batch, st_d)
        got = np.array([d["loss_d"].item(), d["gp"].item(), d["d_real"].mean().item(), d["d_fake"].mean().item()])
        np.testing.assert_allclose(got, gold[f"{case}.d{i}.scalars"], rtol=2e-5, atol=2e-6)
        for k, g in d["grads"].items():
            close_stats(tstats(g), gold[f"{case}.d{i}.grad.{k}"], 2e-4, (case, i, k))
        if i == 0:
            np.testing.assert_allclose(d["fake"][0].numpy(), gold[f"{case}.d0.fake0"], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(d["emb"][0].numpy(), gold[f"{case}.d0.emb0"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(d["d_real"].numpy(), gold[f"{case}.d0.d_real"], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(d["d_fake"].numpy(), gold[f"{case}.d0.d_fake"], rtol=1e-4, atol=1e-6)
    g = O.generator_step(params, batch, st_g)
    np.testing.assert_allclose([g["loss_g_adv"].item(), g["loss_g_emo"].item()], gold[f"{case}.g.scalars"],
                               rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(g["notes"][0].numpy(), gold[f"{case}.g.notes0"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(g["latent"].numpy(), gold[f"{case}.g.latent"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(g["logits"].numpy(), gold[f"{case}.g.logits"], rtol=1e-4, atol=2e-6)
    for k, t in g["grads_G"].items():
        close_stats(tstats(t), gold[f"{case}.g.gradG.{k}"], 5e-4, (case, "G", k))
    for k, t in g["grads_E"].items():
        close_stats(tstats(t), gold[f"{case}.g.gradE.{k}"], 5e-4, (case, "E", k))
    for key in ("E", "G", "D"):
        for k, t in params[key].items():
            close_stats(tstats(t), gold[f"{case}.final.{key}.{k}"], 5e-4, (case, "final", key, k))


def test_eval_mode_forwards_match_reference(gold):
    params = O.make_params(3)
    b = O.make_batch(77, 4)
    with torch.no_grad():
        emb = O.fe_forward(params["E"], b["numeric"], train=False)
        notes, lat = O.gen_forward(params["G"], b["noise_d"], emb, train=False)
        np.testing.assert_allclose(emb.numpy(), gold["eval.emb"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(notes.numpy(), gold["eval.notes"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(lat.numpy(), gold["eval.latent"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(O.disc_forward(params["D"], notes, emb).numpy(), gold["eval.score"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(O.ed_forward(params["ED"], notes).numpy(), gold["eval.logits"], rtol=1e-4, atol=1e-6)


def test_ed_training_steps_match_reference(gold):
    """A-13: two iterations of train_ed.run_epoch's training branch on the reference's own module."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    params = O.make_params(5)
    st = {}
    for i in range(2):
        eb = O.make_ed_batch(50 + i, 16)
        r = O.ed_train_step(params["ED"], eb, st)
        np.testing.assert_allclose([r["loss"].item(), r["acc"].item()], gold[f"ED.s{i}.scalars"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(r["logits"].numpy(), gold[f"ED.s{i}.logits"], rtol=1e-4, atol=2e-6)
        for k, g in r["grads"].items():
            close_stats(tstats(g), gold[f"ED.s{i}.grad.{k}"], 5e-4, ("ED", i, k))
    for k, t in params["ED"].items():
        close_stats(tstats(t), gold[f"ED.final.{k}"], 5e-4, ("ED final", k))


def test_vae_training_steps_match_reference(gold):
    """A-12: two iterations of train_ae.py:114-122 (forward, vae_loss, backward, clip, AdamW) on the reference's VAE."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    P = O.make_vae_params(6)
    st = {}
    for i in range(2):
        vb = O.make_vae_batch(70 + i, 8)
        r = O.vae_train_step(P, vb, st, beta=10.0)
        np.testing.assert_allclose([r["loss"].item(), r["recon_loss"].item(), r["kld"].item()], gold[f"VAE.s{i}.scalars"],
                                   rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(r["recon"][0].numpy(), gold[f"VAE.s{i}.recon0"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(r["mu"].numpy(), gold[f"VAE.s{i}.mu"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(r["log_var"].numpy(), gold[f"VAE.s{i}.log_var"], rtol=1e-4, atol=2e-6)
        for k, g in r["grads"].items():
            if k in O.VAE_NOISE_BIASES:        # bias in front of a train-mode BatchNorm: exact gradient is 0, fp noise only
                assert tstats(g)[1] < 1e-5
                continue
            close_stats(tstats(g), gold[f"VAE.s{i}.grad.{k}"], 5e-4, ("VAE", i, k))
    for k, t in P.items():
        close_stats(tstats(t), gold[f"VAE.final.{k}"], 2e-2 if k in O.VAE_NOISE_BIASES else 5e-4, ("VAE final", k))


def test_conditioning_mode_generator_matches_reference(gold):
    """8f-2: Generator(mode='conditioning') of the reference: forward and every gradient."""
    P = O.make_cond_params(8)
    r = O.cond_generator_grads(P, O.make_cond_batch(80, 8))
    np.testing.assert_allclose(r["notes"][0].numpy(), gold["COND.notes0"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(r["latent"].numpy(), gold["COND.latent"], rtol=1e-4, atol=2e-6)
    close_stats(tstats(r["demb"]), gold["COND.demb"], 5e-4, "COND demb")
    for k, g in r["grads"].items():
        if k in ("decoder.deconv.0.bias", "decoder.deconv.3.bias"):
            continue
        close_stats(tstats(g), gold[f"COND.grad.{k}"], 5e-4, ("COND", k))
