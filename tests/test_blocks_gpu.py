"""Stand-alone inner blocks (melogan/blocks.py over mg_linear_* / mg_act_dropout_*): NoiseToLatent (reference
src/gan/models.py:20-29), GeneratorDecoder.pre (models.py:46-51), MLPClassifier and the emotion discriminator in
input_mode 'latent' (src/emotion_discriminator/ed_model.py:74-101,128-136,156-160) against a float64 torch restatement of
the same nn.Sequential on the same parameters and injected dropout masks: outputs and every gradient within 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from gan_testlib import assert_close

pytestmark = pytest.mark.gpu


def _ref_mlp(seq, x, masks, training):
    """float64 restatement of the reference's nn.Sequential (Linear / ReLU / GELU / Dropout with given keep-masks)."""
    di = 0
    for m in seq:
        if isinstance(m, torch.nn.Linear):
            x = F.linear(x, m.weight.double(), m.bias.double())
        elif isinstance(m, torch.nn.ReLU):
            x = F.relu(x)
        elif isinstance(m, torch.nn.GELU):
            x = F.gelu(x)
        elif isinstance(m, torch.nn.Dropout):
            if training:
                x = x * masks[di].double() / (1.0 - m.p)
            di += 1
    return x


def _grads(module, out, x):
    module.zero_grad()
    g = torch.Generator(device="cuda").manual_seed(9)
    w = torch.randn(out.shape, generator=g, device="cuda", dtype=out.dtype)
    (out * w).sum().backward()
    return x.grad.clone(), {k: p.grad.clone() for k, p in module.named_parameters()}, w


def _compare(module, seq_ref, x, masks, training, call):
    x.requires_grad_(True)
    out = call(x)
    dx, dP, w = _grads(module, out, x)
    xd = x.detach().double().requires_grad_(True)
    ref = seq_ref(xd)
    (ref * w.double()).sum().backward()
    assert_close(out, ref, 1e-5, "forward")
    assert_close(dx, xd.grad, 1e-5, "input gradient")
    return dP


@pytest.mark.parametrize("B", [5, 160])
def test_noise_to_latent_standalone(B):
    from src.gan.models import NoiseToLatent
    torch.manual_seed(3)
    m = NoiseToLatent(256, 64, hidden=512).cuda()
    x = torch.randn(B, 256, device="cuda")
    dP = _compare(m, lambda t: _ref_mlp(m.net, t, None, False), x, None, True, lambda t: m(t))
    # parameter gradients against float64 autograd of the restatement
    for p in m.parameters():
        p.grad = None
    xd = x.detach().double()
    g = torch.Generator(device="cuda").manual_seed(9)
    w = torch.randn((B, 64), generator=g, device="cuda")
    W0, b0, W1, b1 = (t.detach().double().requires_grad_(True) for t in (m.net[0].weight, m.net[0].bias, m.net[2].weight, m.net[2].bias))
    ref = F.linear(F.relu(F.linear(xd, W0, b0)), W1, b1)
    (ref * w.double()).sum().backward()
    for k, r in (("net.0.weight", W0), ("net.0.bias", b0), ("net.2.weight", W1), ("net.2.bias", b1)):
        assert_close(dP[k], r.grad, 2e-5, "NoiseToLatent d" + k)


def test_generator_decoder_pre_standalone():
    from src.gan.models import GeneratorDecoder
    torch.manual_seed(4)
    m = GeneratorDecoder(latent_dim=64, max_notes=64, out_channels=4).cuda()
    x = torch.randn(7, 64, device="cuda")
    out = m.pre_forward(x)
    assert out.shape == (7, 256, 8)
    ref = _ref_mlp(m.pre, x.double(), None, False).view(-1, 256, 8)
    assert_close(out, ref, 1e-5, "GeneratorDecoder.pre")


@pytest.mark.parametrize("training", [True, False])
def test_emotion_discriminator_latent_mode(training):
    """input_mode 'latent': the model is the MLPClassifier (reference ed_model.py:128-136)."""
    from src.emotion_discriminator.ed_model import EmotionDiscriminator
    torch.manual_seed(5)
    cfg = {"input_mode": "latent", "latent_dim": 128, "mlp_hidden": [256, 128], "n_classes": 4, "dropout": 0.2}
    m = EmotionDiscriminator(cfg).cuda()
    m.train(training)
    assert m.encoder is None and set(m.state_dict()) == {f"classifier.net.{i}.{s}" for i in (0, 3) for s in ("weight", "bias")} | {
        "classifier.head.weight", "classifier.head.bias"}
    B = 33
    x = torch.randn(B, 128, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(6)
    masks = [(torch.rand((B, h), generator=g, device="cuda") < 0.8).float() for h in (256, 128)]
    x.requires_grad_(True)
    logits = m(x, masks=masks)
    assert logits.shape == (B, 4)
    y = torch.arange(B, device="cuda") % 4
    loss = F.cross_entropy(logits, y)
    loss.backward()
    got = {k: p.grad.clone() for k, p in m.named_parameters()}
    P = {k: p.detach().double().requires_grad_(True) for k, p in m.named_parameters()}
    xd = x.detach().double().requires_grad_(True)
    h = xd
    for i, (li, mk) in enumerate(zip((0, 3), masks)):
        h = F.gelu(F.linear(h, P[f"classifier.net.{li}.weight"], P[f"classifier.net.{li}.bias"]))
        if training:
            h = h * mk.double() / 0.8
    ref = F.linear(h, P["classifier.head.weight"], P["classifier.head.bias"])
    F.cross_entropy(ref, y).backward()
    assert_close(logits, ref, 1e-5, "latent-mode logits")
    assert_close(x.grad, xd.grad, 2e-5, "latent-mode input gradient")
    for k in got:
        assert_close(got[k], P[k].grad, 2e-5, "latent-mode d" + k)
    with pytest.raises(ValueError):
        m(torch.zeros(2, 3, 128, device="cuda"))
    p = m.predict_proba(x.detach())
    assert torch.allclose(p.sum(-1), torch.ones(B, device="cuda"), atol=1e-5)


def test_cpu_tensor_raises():
    from src.gan.models import NoiseToLatent
    with pytest.raises(RuntimeError):
        NoiseToLatent(8, 4, hidden=16)(torch.zeros(2, 8))


# ---------------------------------------------------------------------------------------------------------------------
# conv-type inner blocks on the native conv unit (mg_convunit_*), against the SAME nn layers run by stock torch in float64
# ---------------------------------------------------------------------------------------------------------------------
import copy


def _check_grads(mod, twin, skip=()):
    for (k, p), (_, q) in zip(mod.named_parameters(), twin.named_parameters()):
        if any(s in k for s in skip):
            continue                                     # conv bias in front of a train-mode BatchNorm: exact gradient 0
        assert p.grad is not None, k
        assert_close(p.grad, q.grad, 1e-4, "d" + k)


def _buffers_close(mod, twin):
    for (k, b), (_, c) in zip(mod.named_buffers(), twin.named_buffers()):
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(c), k
        else:
            assert_close(b, c, 2e-5, k)


@pytest.mark.parametrize("training", [True, False])
def test_conv_block_1d_standalone(training):
    from src.emotion_discriminator.ed_model import ConvBlock1D
    torch.manual_seed(11)
    m = ConvBlock1D(64, 128, kernel_size=3, padding=1).cuda().train(training)
    with torch.no_grad():
        m.net[1].running_mean.normal_(0, 0.1); m.net[1].running_var.uniform_(0.5, 1.5)
        m.net[1].weight.uniform_(0.5, 1.5); m.net[1].bias.normal_(0, 0.1)
    twin = copy.deepcopy(m).double()
    x = torch.randn(6, 64, 96, device="cuda", requires_grad=True)
    xd = x.detach().double().requires_grad_(True)
    y, yd = m(x), twin.net(xd)
    assert y.shape == (6, 128, 96)
    assert_close(y, yd, 1e-5, "ConvBlock1D forward")
    w = torch.randn_like(y)
    (y * w).sum().backward(); (yd * w.double()).sum().backward()
    assert_close(x.grad, xd.grad, 5e-5, "ConvBlock1D input gradient")
    _check_grads(m, twin, skip=("net.0.bias",) if training else ())
    _buffers_close(m, twin)


def test_notes_encoder_standalone():
    from src.emotion_discriminator.ed_model import NotesEncoder
    torch.manual_seed(12)
    m = NotesEncoder(note_dim=4, hidden_dim=256, num_blocks=4).cuda().train()
    twin = copy.deepcopy(m).double()
    notes = (torch.rand(5, 128, 4, device="cuda") * 2 - 1).requires_grad_(True)
    nd = notes.detach().double().requires_grad_(True)
    y = m(notes)
    t = nd.permute(0, 2, 1)
    for blk in twin.conv:                                 # stock torch layers of each block (blk.forward is the native path)
        t = blk.net(t)
    yd = twin.project(twin.pool(t).squeeze(-1))
    assert_close(y, yd, 2e-5, "NotesEncoder forward")
    w = torch.randn_like(y)
    (y * w).sum().backward(); (yd * w.double()).sum().backward()
    assert_close(notes.grad, nd.grad, 2e-4, "NotesEncoder input gradient")
    _check_grads(m, twin, skip=("net.0.bias",))
    _buffers_close(m, twin)


def test_generator_decoder_standalone():
    from src.gan.models import GeneratorDecoder
    torch.manual_seed(13)
    m = GeneratorDecoder(latent_dim=64, max_notes=64, out_channels=4).cuda().train()
    twin = copy.deepcopy(m).double()
    z = torch.randn(9, 64, device="cuda", requires_grad=True)
    zd = z.detach().double().requires_grad_(True)
    y = m(z)
    yd = twin.deconv(twin.pre(zd).view(9, 256, 8)).permute(0, 2, 1)
    assert y.shape == (9, 64, 4)
    assert_close(y, yd, 2e-5, "GeneratorDecoder forward")
    w = torch.randn_like(y)
    (y * w).sum().backward(); (yd * w.double()).sum().backward()
    assert_close(z.grad, zd.grad, 2e-4, "GeneratorDecoder input gradient")
    _check_grads(m, twin, skip=("deconv.0.bias", "deconv.3.bias"))
    _buffers_close(m, twin)


def test_vae_conv_encoder_and_decoder_standalone():
    from src.ae.model import ConvDecoder, ConvEncoder
    torch.manual_seed(14)
    enc = ConvEncoder(in_channels=4, latent_dim=8, hidden_dim=64).cuda().train()
    with torch.no_grad():
        enc(torch.zeros(1, 64, 4, device="cuda"))            # the reference's dummy pass: creates _linear, 2 BN updates
    assert int(enc.conv[1].num_batches_tracked) == 2
    twin = copy.deepcopy(enc).double()
    x = (torch.rand(7, 64, 4, device="cuda") * 2 - 1).requires_grad_(True)
    xd = x.detach().double().requires_grad_(True)
    h = enc(x)
    hd = twin._linear(twin.conv(xd.permute(0, 2, 1)))
    assert_close(h, hd, 2e-5, "ConvEncoder forward")
    w = torch.randn_like(h)
    (h * w).sum().backward(); (hd * w.double()).sum().backward()
    assert_close(x.grad, xd.grad, 2e-4, "ConvEncoder input gradient")
    _check_grads(enc, twin, skip=("conv.0.bias", "conv.3.bias", "conv.6.bias"))
    _buffers_close(enc, twin)

    dec = ConvDecoder(out_channels=4, max_notes=64, latent_dim=8, hidden_dim=64).cuda().train()
    twin = copy.deepcopy(dec).double()
    z = torch.randn(7, 8, device="cuda", requires_grad=True)
    zd = z.detach().double().requires_grad_(True)
    y = dec(z)
    yd = twin.deconv(twin.pre(zd).view(7, 128, 8)).permute(0, 2, 1)
    assert y.shape == (7, 64, 4)
    assert_close(y, yd, 2e-5, "ConvDecoder forward")
    w = torch.randn_like(y)
    (y * w).sum().backward(); (yd * w.double()).sum().backward()
    assert_close(z.grad, zd.grad, 2e-4, "ConvDecoder input gradient")
    _check_grads(dec, twin, skip=("deconv.0.bias", "deconv.3.bias"))
    _buffers_close(dec, twin)


def test_emotion_discriminator_with_spectral_norm():
    """use_spectral_norm: true (reference ed_model.py:29-33,81-86: torch.nn.utils.spectral_norm on every Conv1d / hidden
    Linear): same wrapper and state_dict keys; runs unfused on the block operators.  Against the same layers run by stock
    torch in float64 from the same power-iteration state."""
    from src.emotion_discriminator.ed_model import EmotionDiscriminator
    torch.manual_seed(21)
    cfg = {"input_mode": "notes", "note_dim": 4, "notes_hidden": 256, "notes_blocks": 4, "mlp_hidden": [256, 128], "n_classes": 4,
           "dropout": 0.2, "use_spectral_norm": True}
    m = EmotionDiscriminator(cfg).cuda().train()
    keys = set(m.state_dict())
    assert "encoder.conv.0.net.0.weight_orig" in keys and "encoder.conv.0.net.0.weight_u" in keys
    assert "classifier.net.0.weight_orig" in keys and "classifier.head.weight" in keys
    twin = copy.deepcopy(m).double()
    B = 6
    x = (torch.rand(B, 64, 4, device="cuda") * 2 - 1).requires_grad_(True)
    xd = x.detach().double().requires_grad_(True)
    g = torch.Generator(device="cuda").manual_seed(3)
    masks = [(torch.rand((B, h), generator=g, device="cuda") < 0.8).float() for h in (256, 128)]
    y = m(x, masks=masks)
    t = xd.permute(0, 2, 1)
    for blk in twin.encoder.conv:
        t = blk.net(t)
    f = twin.encoder.project(twin.encoder.pool(t).squeeze(-1))
    h = f
    for i, mk in zip((0, 3), masks):
        h = F.gelu(twin.classifier.net[i](h)) * mk.double() / 0.8
    yd = twin.classifier.head(h)
    assert_close(y, yd, 5e-5, "spectral-norm ED logits")
    w = torch.randn_like(y)
    (y * w).sum().backward(); (yd * w.double()).sum().backward()
    assert_close(x.grad, xd.grad, 5e-4, "spectral-norm ED input gradient")
    for (k, p), (_, q) in zip(m.named_parameters(), twin.named_parameters()):
        if k.endswith("net.0.bias") and "encoder" in k:
            continue
        assert p.grad is not None, k
        assert_close(p.grad, q.grad, 5e-4, "d" + k)
    for (k, b), (_, c) in zip(m.named_buffers(), twin.named_buffers()):
        if k.endswith(("weight_u", "weight_v")):
            assert_close(b, c, 1e-4, k)                  # the power iteration advanced identically
