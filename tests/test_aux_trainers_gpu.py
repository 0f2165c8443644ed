"""VaeTrainer / EdTrainer (melogan/aux_trainers.py): the fused, graph-captured training steps of BASELINE configs #2 / #3
against the oracle's restatement of the reference loops -- src/ae/train_ae.py:107-122 (forward, vae_loss, backward,
clip_grad_norm_(1.0), AdamW) and src/emotion_discriminator/train_ed.py:61-74 (forward, CrossEntropyLoss, backward,
AdamW) -- teacher-forced on the same eps / dropout masks."""
import os

import numpy as np
import pytest
import torch
import yaml

from gan_testlib import assert_close
from melogan import engine as E
from melogan.aux_trainers import EdTrainer, VaeTrainer
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu
CFG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "melo-gan_b200", "config")


def _vae_trainer(B, precision="fp32"):
    from src.ae.model import VAE
    cfg = dict(O.AE_CFG, BATCH_SIZE=B, SEED=3)
    model = VAE(cfg).cuda()
    with torch.no_grad():
        model.encoder(torch.zeros(1, 512, 4, device="cuda"))
    missing, unexpected = model.load_state_dict(O.make_vae_params(6), strict=False)
    assert not unexpected
    return VaeTrainer(cfg, batch=B, precision=precision, model=model), cfg


def test_vae_trainer_steps_match_oracle_incl_clip_and_adamw():
    B = 8
    tr, cfg = _vae_trainer(B)
    P, st = {k: v.clone() for k, v in O.make_vae_params(6).items()}, {}
    for i in range(3):
        vb = O.make_vae_batch(70 + i, B)
        ref = O.vae_train_step(P, vb, st, beta=10.0, update=True)
        norm_ref = torch.sqrt(sum((g.double() ** 2).sum() for g in ref["grads"].values())).item()
        m = tr.step(vb["x"].cuda(), eps=vb["eps"].cuda(), beta=10.0).cpu()
        # step 0 starts from identical parameters; later steps follow Adam updates of lr * sign(g) on elements whose sign
        # float32 rounding decides (tests/test_vae_gpu.py), so they agree to ~1e-3 only
        tol = 3e-5 if i == 0 else 3e-3
        np.testing.assert_allclose(m.numpy(), [ref["loss"].item(), ref["recon_loss"].item(), ref["kld"].item()], rtol=tol)
        assert abs(tr.opt.grad_norm.item() - norm_ref) <= (2e-4 if i == 0 else 2e-2) * norm_ref, (i, tr.opt.grad_norm.item(), norm_ref)
        if i == 0:
            assert norm_ref > 1.0                          # the clip is active: the update depends on it
    sd = tr.model.state_dict()
    assert int(sd["encoder.conv.1.num_batches_tracked"]) == 5      # the zero dummy pass counts twice (reference model.py:27-44), then 3 steps
    for k in E.VAE_PARAM_KEYS + E.VAE_BUFFER_KEYS:
        if k in O.VAE_NOISE_BIASES:
            continue
        got, want = sd[k].double().cpu(), P[k].double()
        assert (got - want).norm().item() <= 2e-3 * max(want.norm().item(), 1e-12), (k, (got - want).norm().item(), want.norm().item())
    loss, recon, kld = tr.epoch_means()
    assert np.isfinite([loss, recon, kld]).all() and tr.loss_acc.abs().sum().item() == 0.0


def test_vae_trainer_graph_replay_equals_eager_steps():
    B = 8
    a, _ = _vae_trainer(B)
    b, _ = _vae_trainer(B)
    xs = [O.make_vae_batch(80 + i, B)["x"].cuda() for i in range(4)]
    a.step(xs[0]); b.step(xs[0])                               # eager warm-up step on both (same device RNG stream)
    sx = b.capture()
    for x in xs[1:]:
        a.step(x)
        sx.copy_(x); b.replay()
    torch.cuda.synchronize()
    for k in E.VAE_PARAM_KEYS:
        if k in O.VAE_NOISE_BIASES:            # exact gradient 0 in front of BatchNorm: Adam turns atomics-order noise into +-lr
            continue
        pa, pb = dict(a.model.named_parameters())[k], dict(b.model.named_parameters())[k]
        assert_close(pb, pa, 1e-3, "graph replay vs eager: " + k)      # fp32 atomics in the weight gradients + Adam on near-zero g: not bit-equal
    assert_close(b.loss_acc, a.loss_acc, 1e-4, "loss accumulators")
    assert int(b.opt.step_dev.item()) == 4


def test_vae_trainer_bf16_mode_tracks_fp32():
    B = 8
    a, _ = _vae_trainer(B, "fp32")
    b, _ = _vae_trainer(B, "bf16")
    vb = O.make_vae_batch(70, B)
    ma = a.step(vb["x"].cuda(), eps=vb["eps"].cuda()).cpu()
    mb = b.step(vb["x"].cuda(), eps=vb["eps"].cuda()).cpu()
    np.testing.assert_allclose(mb.numpy(), ma.numpy(), rtol=1e-2)


def _ed_trainer(B, precision="fp32"):
    from src.emotion_discriminator.ed_model import EmotionDiscriminator
    cfg = yaml.safe_load(open(os.path.join(CFG_DIR, "ed_config.yaml")))
    params = O.make_params(5)
    model = EmotionDiscriminator(cfg)
    model.load_state_dict(params["ED"], strict=False)
    return EdTrainer(cfg, batch=B, precision=precision, model=model.cuda()), params


def test_ed_trainer_steps_match_oracle():
    B = 16
    tr, params = _ed_trainer(B)
    oparams, st = O.clone_params(params)["ED"], {}
    for i in range(3):
        eb = O.make_ed_batch(50 + i, B)
        ref = O.ed_train_step(oparams, eb, st)
        m = tr.step(eb["x"].cuda(), eb["y"].cuda(), masks=(eb["mask1"].cuda(), eb["mask2"].cuda())).cpu()
        assert abs(m[0].item() - ref["loss"].item()) <= (2e-5 if i == 0 else 1e-3) * ref["loss"].item(), (i, m, ref["loss"])
        assert abs(m[1].item() - ref["acc"].item()) < 1e-6
    sd = tr.model.state_dict()
    for k in ("classifier.head.weight", "encoder.project.weight", "encoder.conv.3.net.0.weight", "encoder.conv.1.net.1.weight",
              "encoder.conv.2.net.1.running_var"):
        assert_close(sd[k], oparams[k], 1e-3, "after 3 fused AdamW steps: " + k)
    assert int(sd["encoder.conv.0.net.1.num_batches_tracked"]) == 3
    ev = tr.evaluate(O.make_ed_batch(50, B)["x"].cuda(), O.make_ed_batch(50, B)["y"].cuda()).cpu()
    lg = O.ed_forward(oparams, O.make_ed_batch(50, B)["x"])
    want = torch.nn.functional.cross_entropy(lg, O.make_ed_batch(50, B)["y"]).item()
    assert abs(ev[0].item() - want) <= 5e-3 * want


def test_ed_trainer_graph_replay_equals_eager_steps():
    B = 16
    a, _ = _ed_trainer(B)
    b, _ = _ed_trainer(B)
    bs = [O.make_ed_batch(60 + i, B) for i in range(4)]
    x0, y0 = bs[0]["x"].cuda(), bs[0]["y"].cuda()
    a.step(x0, y0); b.step(x0, y0)
    sx, sy = b.capture()
    for e in bs[1:]:
        x, y = e["x"].cuda(), e["y"].cuda()
        a.step(x, y)
        sx.copy_(x); sy.copy_(y); b.replay()
    torch.cuda.synchronize()
    for k in E.ED_GRAD_KEYS:
        if k.endswith("net.0.bias"):           # conv bias in front of a train-mode BatchNorm: exact gradient 0, noise only
            continue
        assert_close(dict(b.model.named_parameters())[k], dict(a.model.named_parameters())[k], 1e-4, "graph replay vs eager: " + k)
    assert_close(b.loss_acc, a.loss_acc, 1e-4, "loss accumulators")


def test_label_validation_raises_like_cross_entropy():
    with pytest.raises(IndexError):
        EdTrainer.check_labels(torch.tensor([0, 1, -1]), 4)
    with pytest.raises(IndexError):
        EdTrainer.check_labels(torch.tensor([0, 4]), 4)
    EdTrainer.check_labels(torch.tensor([0, 3]), 4)
