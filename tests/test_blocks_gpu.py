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
    with pytest.raises(NotImplementedError):
        m(x)


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
