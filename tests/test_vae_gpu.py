"""A-12 (BASELINE config #2): VAE forward, vae_loss, backward and the train_ae.py iteration against the oracle
(reference src/ae/model.py:4-148, src/ae/train_ae.py:35-51,114-122), fp32 parity mode and bf16 mode."""
import os

import numpy as np
import pytest
import torch

from gan_testlib import assert_close, assert_close_l2, to_double
from melogan import engine as E
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gan_golden.npz")


def _engine(B, P0, precision, T=512, latent=8):
    eng = E.VaeEngine(B, T, latent, precision=precision)
    P = {k: v.clone().cuda() for k, v in P0.items()}
    G = {k: torch.zeros_like(P[k]) for k in E.VAE_PARAM_KEYS}
    eng.bind(P, G)
    return eng, P, G


def _ref(P0, vb, beta, dtype=torch.float32, masks=None):
    P = {k: v.to(dtype) for k, v in P0.items()}
    return O.vae_train_step(P, {k: v.to(dtype) for k, v in vb.items()}, {}, beta=beta, update=False, masks=masks)


def _gpu_relu_masks(eng, B, T):
    """The ReLU decisions the CUDA forward took, in the oracle's (B, C, L) layouts.  Every batch holds a few
    pre-activations within float32 rounding of zero (|v| ~ 1e-6 next to values of O(1)); on those the branch is decided
    by rounding, and one flipped element moves the small decoder gradients by ~1e-2.  The oracle takes the CUDA path's
    branch on exactly those elements (|v| < 1e-4) so that the gradients are compared like for like."""
    L0 = T // 8
    cl = lambda name, L, C: eng.buffer(name).view(B, L, C).permute(0, 2, 1).cpu() > 0
    m = {"e_a0": cl("e_a0", T // 2, 32), "e_a1": cl("e_a1", T // 4, 64), "e_a2": cl("e_a2", L0, 128),
         "h": eng.buffer("h").view(B, 512).cpu() > 0, "d0": eng.buffer("d0").view(B, 512).cpu() > 0,
         "d_y0": (eng.buffer("d_y0").view(B, L0, 128).permute(0, 2, 1).reshape(B, -1).cpu() > 0),
         "d_y1": cl("d_y1", 2 * L0, 64), "d_y2": cl("d_y2", 4 * L0, 32)}
    return m


@pytest.mark.parametrize("B,T,latent", [(8, 512, 8), (32, 512, 8), (5, 64, 16)])
def test_vae_forward_loss_backward_fp32(B, T, latent):
    P0 = O.make_vae_params(6, latent, T)
    vb = O.make_vae_batch(70, B, latent, T)
    eng, P, G = _engine(B, P0, "fp32", T, latent)
    recon, z, mu, lv = eng.forward(vb["x"].cuda(), vb["eps"].cuda(), train=True)
    masks = _gpu_relu_masks(eng, B, T)
    ref, ref64 = _ref(P0, vb, 10.0, masks=masks), _ref(P0, vb, 10.0, torch.float64, masks=masks)
    assert_close(recon, ref["recon"], 1e-5, "recon", ref64["recon"])
    assert_close(mu, ref["mu"], 1e-5, "mu", ref64["mu"])
    assert_close(lv, ref["log_var"], 1e-5, "log_var", ref64["log_var"])
    assert_close(z, ref["z"], 1e-5, "z", ref64["z"])
    o2 = {k: v.clone() for k, v in P0.items()}                  # running statistics advanced like nn.BatchNorm1d
    O.vae_forward(o2, vb["x"], vb["eps"], True, o2)
    for k in E.VAE_BUFFER_KEYS:
        assert_close(P[k], o2[k], 1e-5, k)
    # fused loss step: metrics + every parameter gradient of loss = MSE + beta * KLD
    for k in E.VAE_BUFFER_KEYS:
        P[k].copy_(P0[k])
    m = eng.loss_step(vb["x"].cuda(), vb["eps"].cuda(), 10.0).cpu()
    want = torch.stack([ref64["loss"], ref64["recon_loss"], ref64["kld"]])
    assert_close(m, want.float(), 2e-5, "vae_loss metrics")
    for k in E.VAE_PARAM_KEYS:
        if k in O.VAE_NOISE_BIASES:           # bias in front of a train-mode BatchNorm: exact gradient is zero
            assert G[k].abs().max().item() <= 1e-5 * ref["grads"][k.replace("bias", "weight")].abs().max().item() + 1e-7, k
            continue
        assert_close(G[k], ref["grads"][k], 5e-5, "VAE grad " + k, ref64["grads"][k])


def test_vae_eval_forward_fp32():
    P0 = O.make_vae_params(6)
    vb = O.make_vae_batch(72, 8)
    eng, P, G = _engine(8, P0, "fp32")
    recon, z, mu, lv = eng.forward(vb["x"].cuda(), vb["eps"].cuda(), train=False)
    r, rz, rmu, rlv = O.vae_forward(P0, vb["x"], vb["eps"], False)
    assert_close(recon, r, 1e-5, "recon (eval)")
    assert_close(mu, rmu, 1e-5, "mu (eval)")
    assert_close(lv, rlv, 1e-5, "log_var (eval)")
    for k in E.VAE_BUFFER_KEYS:
        assert torch.equal(P[k].cpu(), P0[k]), k


def test_vae_bf16_mode():
    B = 64
    P0 = O.make_vae_params(6)
    vb = O.make_vae_batch(73, B)
    ref = _ref(P0, vb, 10.0)
    eng, P, G = _engine(B, P0, "bf16")
    recon, z, mu, lv = eng.forward(vb["x"].cuda(), vb["eps"].cuda(), train=True)
    assert_close_l2(recon, ref["recon"], 1e-2, "recon (bf16)")
    assert_close_l2(mu, ref["mu"], 1e-2, "mu (bf16)")
    assert_close_l2(lv, ref["log_var"], 1e-2, "log_var (bf16)")
    for k in E.VAE_BUFFER_KEYS:
        P[k].copy_(P0[k])
    m = eng.loss_step(vb["x"].cuda(), vb["eps"].cuda(), 10.0).cpu()
    want = torch.stack([ref["loss"], ref["recon_loss"], ref["kld"]])
    assert_close(m, want, 1e-2, "vae_loss metrics (bf16)")
    for k in E.VAE_PARAM_KEYS:
        if k in O.VAE_NOISE_BIASES:
            continue
        assert_close_l2(G[k], ref["grads"][k], 0.1, "VAE grad (bf16) " + k)


def test_train_ae_iteration_on_the_dropin_module_matches_reference_golden():
    """The reference's loop body (train_ae.py:110-122) verbatim on the drop-in VAE with torch's AdamW, against the golden
    values recorded from the reference's own VAE."""
    from src.ae.model import VAE
    from src.ae.train_ae import vae_loss
    gold = np.load(GOLD)
    model = VAE({"LATENT_DIM": 8, "MAX_NOTES": 512}).cuda()
    with torch.no_grad():
        model.encoder(torch.zeros(1, 512, 4, device="cuda"))
    missing, unexpected = model.load_state_dict(O.make_vae_params(6), strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    model.train()
    real = torch.randn_like
    for i in range(2):
        vb = O.make_vae_batch(70 + i, 8)
        batch_notes = vb["x"].cuda()
        torch.randn_like = lambda t, _e=vb["eps"]: _e.clone().to(t.device)
        try:
            recon, z, mu, log_var = model(batch_notes)
        finally:
            torch.randn_like = real
        loss, recon_loss, kld_loss = vae_loss(recon, batch_notes, mu, log_var, 10.0)
        optimizer.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        optimizer.step()
        # step 0 starts from identical parameters: tight.  Step 1 follows an Adam step whose first update is lr * sign(g):
        # gradient elements that float32 rounding (ReLU branches of near-zero pre-activations, see _gpu_relu_masks) leaves
        # with an undetermined sign move a weight by 2 * lr, so the second forward agrees to ~1e-3 only.
        tight = i == 0
        np.testing.assert_allclose([loss.item(), recon_loss.item(), kld_loss.item()], gold[f"VAE.s{i}.scalars"],
                                   rtol=3e-5 if tight else 2e-3)
        np.testing.assert_allclose(recon[0].detach().cpu().numpy(), gold[f"VAE.s{i}.recon0"], rtol=1e-4,
                                   atol=1e-5 if tight else 5e-3)
        np.testing.assert_allclose(mu.detach().cpu().numpy(), gold[f"VAE.s{i}.mu"], rtol=1e-4, atol=1e-5 if tight else 1e-3)
        for k, g in grads.items():
            if k in O.VAE_NOISE_BIASES:
                continue
            want = gold[f"VAE.s{i}.grad.{k}"]
            got = g.double()
            tol = 2e-4 if (tight and not k.startswith("decoder")) else 5e-2
            assert abs(got.norm().item() - want[1]) <= tol * want[1], (i, k, got.norm().item(), want[1])
    sd = model.state_dict()
    assert int(sd["encoder.conv.1.num_batches_tracked"]) == 4      # the zero dummy pass counts twice (reference model.py:27-44), then 2 steps
    for k in E.VAE_PARAM_KEYS + E.VAE_BUFFER_KEYS:
        if k in O.VAE_NOISE_BIASES:
            continue
        want = gold[f"VAE.final.{k}"]
        assert abs(sd[k].double().norm().item() - want[1]) <= 1e-3 * max(want[1], 1e-12), (k, sd[k].double().norm().item(), want[1])


def test_encode_path_returns_eval_mode_mu():
    """src/ae/encode.py -- posterior means of an eval-mode VAE, ragged tail batch included."""
    from src.ae.encode import encode
    from src.ae.model import VAE
    P0 = O.make_vae_params(6)
    model = VAE({"LATENT_DIM": 8, "MAX_NOTES": 512}).cuda()
    with torch.no_grad():
        model.encoder(torch.zeros(1, 512, 4, device="cuda"))
    model.load_state_dict(P0, strict=False)
    notes = O.make_vae_batch(74, 11)["x"].numpy()
    got = encode(model, notes, batch_size=8)
    assert got.shape == (11, 8)
    _, _, mu, _ = O.vae_forward(P0, torch.from_numpy(notes), torch.zeros(11, 8), train=False)
    assert_close(torch.from_numpy(got), mu, 1e-5, "encode(): mu in eval mode")
