"""Module-level parity (fp32 mode) of the CUDA path against the oracle: forward activations and
every gradient, relative to the tensor's scale; tolerance 1e-5 on activations/losses (north_star)."""
import pytest
import torch
import torch.nn.functional as F

from gan_testlib import assert_close, cuda_batch, make_engine, to_double
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

ACT_TOL = 1e-5
GRAD_TOL = 5e-5


@pytest.fixture(scope="module", params=[(8, False), (5, True)])
def setup(request):
    B, fan = request.param
    params = O.make_params(11, fan_in_scale=fan)
    batch = O.make_batch(21, B)
    eng, cp, grads = make_engine(B, params)
    return B, params, batch, eng, cp, grads, cuda_batch(batch)


def test_feature_encoder_forward_backward(setup):
    B, params, batch, eng, cp, grads, cb = setup
    El = O._leaves(params["E"])
    emb_ref = O.fe_forward(El, batch["numeric"], batch["mask1_d"], batch["mask2_d"], train=True)
    emb = eng.encoder_forward(cb["numeric"], cb["mask1_d"], cb["mask2_d"], train=True)
    assert_close(emb, emb_ref, ACT_TOL, "emb(train)")
    demb = O.make_batch(5, B)["noise_d"][:, :128].contiguous()
    g_ref = torch.autograd.grad(emb_ref, list(El.values()), demb)
    for g in grads["E"].values():
        g.zero_()
    eng.encoder_backward(demb.cuda())
    for (k, _), gr in zip(El.items(), g_ref):
        assert_close(grads["E"][k], gr, GRAD_TOL, "E grad " + k)
    emb_eval = eng.encoder_forward(cb["numeric"], train=False)
    assert_close(emb_eval, O.fe_forward(params["E"], batch["numeric"], train=False), ACT_TOL, "emb(eval)")


def test_generator_forward_train_eval_and_backward(setup):
    B, params, batch, eng, cp, grads, cb = setup
    emb = O.fe_forward(params["E"], batch["numeric"], train=False).detach()
    Gl = O._leaves(params["G"])
    bn_state = {k: v.clone() for k, v in params["G"].items() if O.is_buffer(k)}
    emb_leaf = emb.clone().requires_grad_(True)
    keep = {}
    notes_ref, lat_ref = O.gen_forward(Gl, batch["noise_g"], emb_leaf, train=True, bn_state=bn_state, keep=keep)
    rs_before = {k: cp["G"][k].clone() for k in bn_state}
    notes, lat = eng.generator_forward(cb["noise_g"], emb.cuda(), train=True)
    assert_close(lat, lat_ref, ACT_TOL, "latent")
    assert_close(eng.buffer("g.y0").view(B, 64, 256), keep["pre2"].permute(0, 2, 1), ACT_TOL, "pre.2 (channels-last)")
    assert_close(eng.buffer("g.x1").view(B, 128, 128), keep["decoder.deconv.0"].permute(0, 2, 1), ACT_TOL, "deconv.0")
    assert_close(eng.buffer("g.y1").view(B, 128, 128), keep["decoder.deconv.1"].permute(0, 2, 1), ACT_TOL, "bn1+relu")
    assert_close(eng.buffer("g.y2").view(B, 256, 64), keep["decoder.deconv.4"].permute(0, 2, 1), ACT_TOL, "bn2+relu")
    assert_close(notes, notes_ref, ACT_TOL, "notes(train)")
    for k in bn_state:   # running statistics advanced like nn.BatchNorm1d
        assert_close(cp["G"][k], bn_state[k], 1e-6, "running stat " + k)
        assert not torch.equal(cp["G"][k], rs_before[k])
    dnotes = O.make_batch(6, B)["notes_real"] * 0.01
    dlat = O.make_batch(7, B)["noise_d"][:, :64].contiguous() * 0.01
    g_ref = torch.autograd.grad([notes_ref, lat_ref], list(Gl.values()) + [emb_leaf], [dnotes, dlat])
    for g in grads["G"].values():
        g.zero_()
    demb = eng.generator_backward(dnotes.cuda(), dlat.cuda())
    assert_close(demb, g_ref[-1], GRAD_TOL, "d emb")
    for (k, _), gr in zip(Gl.items(), g_ref[:-1]):
        if k in ("decoder.deconv.0.bias", "decoder.deconv.3.bias"):
            # mathematically zero (a bias in front of BatchNorm): both sides hold rounding noise only
            ref_scale = g_ref[list(Gl.keys()).index(k.replace("bias", "weight"))].abs().max().item()
            assert grads["G"][k].abs().max().item() <= 1e-4 * ref_scale + 1e-7, k
            continue
        assert_close(grads["G"][k], gr, GRAD_TOL, "G grad " + k)
    # eval mode (generation path, app.py:97-106): running statistics, no update
    rs = {k: cp["G"][k].clone() for k in bn_state}
    notes_e, _ = eng.generator_forward(cb["noise_g"], emb.cuda(), train=False)
    Pe = dict(params["G"]); Pe.update({k: v.cpu() for k, v in rs.items()})
    notes_e_ref, _ = O.gen_forward(Pe, batch["noise_g"], emb, train=False)
    assert_close(notes_e, notes_e_ref, ACT_TOL, "notes(eval)")
    for k in rs:
        assert torch.equal(cp["G"][k], rs[k])


def test_critic_forward_backward(setup):
    B, params, batch, eng, cp, grads, cb = setup
    emb = O.fe_forward(params["E"], batch["numeric"], train=False).detach()
    Dl = O._leaves(params["D"])
    x = batch["notes_real"].clone().requires_grad_(True)
    e = emb.clone().requires_grad_(True)
    keep = {}
    s_ref = O.disc_forward(Dl, x, e, keep=keep)
    s = eng.critic_forward(cb["notes_real"], emb.cuda())
    assert_close(eng.buffer("d.h1")[:B * 16384].view(B, 256, 64), keep["conv.0"].permute(0, 2, 1), ACT_TOL, "conv.0")
    assert_close(eng.buffer("d.h2")[:B * 16384].view(B, 128, 128), keep["conv.2"].permute(0, 2, 1), ACT_TOL, "conv.2")
    assert_close(eng.buffer("d.h3")[:B * 16384].view(B, 64, 256), keep["conv.4"].permute(0, 2, 1), ACT_TOL, "conv.4")
    assert_close(s, s_ref, ACT_TOL, "score")
    ds = O.make_batch(8, B)["alpha"] - 0.5
    g_ref = torch.autograd.grad(s_ref, list(Dl.values()) + [x, e], ds)
    for g in grads["D"].values():
        g.zero_()
    dnotes, demb = eng.critic_backward(ds.cuda(), param_grads=True, want_dnotes=True, want_demb=True)
    assert_close(dnotes, g_ref[-2], GRAD_TOL, "d notes")
    assert_close(demb, g_ref[-1], GRAD_TOL, "d emb")
    for (k, _), gr in zip(Dl.items(), g_ref[:-2]):
        assert_close(grads["D"][k], gr, GRAD_TOL, "D grad " + k)
    # no conditioning tensor (numeric_embedding=None branch of models.py:164)
    s2 = eng.critic_forward(cb["notes_real"], None)
    Dn = dict(params["D"]); Dn["real_fake.weight"] = params["D"]["real_fake.weight"][:, :256]
    assert_close(s2, O.disc_forward(Dn, batch["notes_real"], None), ACT_TOL, "score(no emb)")


def test_gradient_penalty_and_critic_loss(setup):
    B, params, batch, eng, cp, grads, cb = setup
    emb = O.fe_forward(params["E"], batch["numeric"], train=False).detach()
    fake = (O.make_batch(9, B)["notes_real"] * 0.7).contiguous()
    Dl = O._leaves(params["D"])
    d_real = O.disc_forward(Dl, batch["notes_real"], emb)
    d_fake = O.disc_forward(Dl, fake, emb)
    gp = O.gradient_penalty(Dl, batch["notes_real"], fake, emb, batch["alpha"])
    loss = d_fake.mean() - d_real.mean() + 10.0 * gp
    g_ref = torch.autograd.grad(loss, list(Dl.values()))
    # the same in float64: yardstick for cancellation-dominated gradients
    D64 = O._leaves(to_double(params["D"]))
    r64, f64, e64, a64 = batch["notes_real"].double(), fake.double(), emb.double(), batch["alpha"].double()
    loss64 = (O.disc_forward(D64, f64, e64).mean() - O.disc_forward(D64, r64, e64).mean()
              + 10.0 * O.gradient_penalty(D64, r64, f64, e64, a64))
    g64 = dict(zip(D64.keys(), torch.autograd.grad(loss64, list(D64.values()))))
    for g in grads["D"].values():
        g.zero_()
    m = eng.critic_loss_backward(cb["notes_real"], fake.cuda(), emb.cuda(), cb["alpha"]).cpu()
    assert abs(m[1].item() - gp.item()) <= ACT_TOL * max(1.0, abs(gp.item())), ("gp", m[1].item(), gp.item())
    assert abs(m[0].item() - loss.item()) <= ACT_TOL * abs(loss.item()), ("loss_d", m[0].item(), loss.item())
    assert abs(m[2].item() - d_real.mean().item()) <= ACT_TOL * max(d_real.abs().max().item(), 1e-3)
    assert abs(m[3].item() - d_fake.mean().item()) <= ACT_TOL * max(d_fake.abs().max().item(), 1e-3)
    for (k, _), gr in zip(Dl.items(), g_ref):
        if k == "real_fake.bias":
            assert grads["D"][k].abs().max().item() <= 1e-6
            continue
        if k == "real_fake.weight":   # the embedding half cancels exactly between real and fake
            assert_close(grads["D"][k][:, :256], gr[:, :256], GRAD_TOL, "D grad real_fake.weight[:256]")
            assert grads["D"][k][:, 256:].abs().max().item() <= 1e-6 * max(1.0, gr.abs().max().item())
            continue
        assert_close(grads["D"][k], gr, GRAD_TOL, "critic-loss grad " + k, g64[k])


def test_emotion_forward_and_input_gradient(setup):
    B, params, batch, eng, cp, grads, cb = setup
    x = (batch["notes_real"] * 0.8).clone().requires_grad_(True)
    keep = {}
    logits_ref = O.ed_forward(params["ED"], x, keep=keep)
    logits = eng.emotion_forward((batch["notes_real"] * 0.8).cuda().contiguous())
    for i, C in enumerate((64, 128, 256, 256)):
        assert_close(eng.buffer(f"ed.h{i}").view(B, 512, C), keep[f"conv{i}"].permute(0, 2, 1), ACT_TOL, f"ed conv{i}")
    assert_close(logits, logits_ref, ACT_TOL, "logits")
    loss = F.cross_entropy(logits_ref, batch["emot_idx"])
    dl, = torch.autograd.grad(loss, logits_ref, retain_graph=True)
    dx_ref, = torch.autograd.grad(loss, x)
    dx = eng.emotion_backward_input(dl.cuda().contiguous())
    assert_close(dx, dx_ref, GRAD_TOL, "d notes (ED)")
    dx2 = eng.emotion_backward_input(dl.cuda().contiguous(), out=(0.5 * dx).contiguous(), accumulate=True)
    assert_close(dx2, 1.5 * dx_ref, GRAD_TOL, "accumulate")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_generator_conditioning_mode(precision):
    """8f-2: INTEGRATION_MODE 'conditioning' (models.py:99-100,112-126): G's input row is
    [noise | numeric embedding | AE latent]; drop-in module forward + backward against the oracle."""
    import melogan.runtime as R
    from src.gan.models import Generator
    B = 8
    P0 = O.make_cond_params(8)
    cb = O.make_cond_batch(80, B)
    ref = O.cond_generator_grads(P0, cb)
    old = R.PRECISION
    R.set_precision(precision)
    try:
        G = Generator(noise_dim=128, latent_dim=64, mode="conditioning", max_notes=512, note_dim=4, numeric_embed_dim=128)
        assert G.input_dim == 128 + 128 + 64
        G.load_state_dict(P0, strict=False)
        G.cuda().train()
        emb = cb["emb"].cuda().requires_grad_(True)
        notes, latent = G(cb["noise"].cuda(), cb["cond"].cuda(), emb)
        loss = (notes * cb["w_notes"].cuda()).sum() + (latent * cb["w_latent"].cuda()).sum()
        loss.backward()
    finally:
        R.set_precision(old)
    if precision == "fp32":
        assert_close(notes, ref["notes"], 1e-5, "notes (conditioning)")
        assert_close(latent, ref["latent"], 1e-5, "latent (conditioning)")
        assert_close(emb.grad, ref["demb"], 5e-5, "d embedding (conditioning)")
        for k, p in G.named_parameters():
            if k in ("decoder.deconv.0.bias", "decoder.deconv.3.bias"):
                continue
            assert_close(p.grad, ref["grads"][k], 1e-4, "G grad (conditioning) " + k)
    else:
        from gan_testlib import assert_close_l2
        assert_close_l2(notes, ref["notes"], 1e-2, "notes (conditioning, bf16)")
        assert_close_l2(latent, ref["latent"], 1e-2, "latent (conditioning, bf16)")
        assert_close_l2(G.noise_to_latent.net[0].weight.grad, ref["grads"]["noise_to_latent.net.0.weight"], 0.3,
                        "first Linear grad (conditioning, bf16)")
    with pytest.raises(AssertionError):
        G(cb["noise"].cuda(), None, emb)
