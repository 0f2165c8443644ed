"""A-13 (BASELINE config #3): EmotionDiscriminator training step -- train-mode BatchNorm + dropout forward, full
backward, AdamW -- against the oracle (reference train_ed.py:61-74), fp32 parity mode and bf16 mode."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from gan_testlib import assert_close, assert_close_l2, to_double
from melogan import engine as E
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu


def _engine(B, params, precision):
    eng = E.GanEngine(B, precision=precision)
    P = {k: v.clone().cuda() for k, v in params["ED"].items()}
    G = {k: torch.zeros_like(P[k]) for k in E.ED_GRAD_KEYS}
    eng.bind(E.MOD_ED, P, G)
    return eng, P, G


@pytest.mark.parametrize("B", [16, 64])
def test_ed_train_forward_backward_fp32(B):
    params = O.make_params(5)
    eb = O.make_ed_batch(50, B)
    ref = O.ed_train_step(O.clone_params(params)["ED"], eb, {}, update=False)
    ref64 = O.ed_train_step({k: (v.double() if v.is_floating_point() else v) for k, v in params["ED"].items()},
                            to_double(eb), {}, update=False)
    eng, P, G = _engine(B, params, "fp32")
    logits = eng.emotion_train_forward(eb["x"].cuda(), eb["mask1"].cuda(), eb["mask2"].cuda(), 0.2)
    assert_close(logits, ref["logits"], 1e-5, "logits (train mode)")
    # running statistics advanced like nn.BatchNorm1d (momentum 0.1, unbiased variance)
    o2 = O.clone_params(params)["ED"]
    O.ed_train_forward(o2, eb["x"], eb["mask1"], eb["mask2"], o2)
    for k in E.ED_KEYS:
        if k.endswith(("running_mean", "running_var")):
            assert_close(P[k], o2[k], 1e-5, k)
    lg = ref["logits"].clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(lg, eb["y"])
    dlogits, = torch.autograd.grad(loss, lg)
    eng.emotion_train_backward(dlogits.cuda().contiguous())
    for k in E.ED_GRAD_KEYS:
        if k.startswith("encoder.conv.") and k.endswith("net.0.bias"):   # conv bias in front of BatchNorm: mathematically zero
            assert G[k].abs().max().item() <= 1e-5 * max(ref["grads"][k.replace("bias", "weight")].abs().max().item(), 1e-30) + 1e-7
            continue
        assert_close(G[k], ref["grads"][k], 5e-5, "ED grad " + k, ref64["grads"][k])


def test_ed_train_bf16_mode():
    B = 64
    params = O.make_params(5)
    eb = O.make_ed_batch(51, B)
    ref = O.ed_train_step(O.clone_params(params)["ED"], eb, {}, update=False)
    eng, P, G = _engine(B, params, "bf16")
    logits = eng.emotion_train_forward(eb["x"].cuda(), eb["mask1"].cuda(), eb["mask2"].cuda(), 0.2)
    assert_close_l2(logits, ref["logits"], 1e-2, "logits (bf16)")
    lg = ref["logits"].clone().requires_grad_(True)
    dlogits, = torch.autograd.grad(torch.nn.functional.cross_entropy(lg, eb["y"]), lg)
    eng.emotion_train_backward(dlogits.cuda().contiguous())
    for k in E.ED_GRAD_KEYS:
        if k.startswith("encoder.conv.") and k.endswith("net.0.bias"):
            continue
        assert_close_l2(G[k], ref["grads"][k], 0.1, "ED grad (bf16) " + k)


def test_run_epoch_of_the_dropin_trainer_matches_oracle_steps():
    """The reference's run_epoch flow (train_ed.py:51-82) on the drop-in module with torch's AdamW."""
    from src.emotion_discriminator.ed_model import EmotionDiscriminator
    from src.emotion_discriminator import train_ed
    import yaml
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "melo-gan_b200", "config")
    cfg = yaml.safe_load(open(os.path.join(root, "ed_config.yaml")))
    params = O.make_params(5)
    model = EmotionDiscriminator(cfg)
    model.load_state_dict(params["ED"], strict=False)
    model.cuda()
    opt = train_ed.build_optimizer(model, cfg)
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["betas"] == (0.5, 0.999)
    oparams, st = O.clone_params(params)["ED"], {}
    B = 16
    batches = [O.make_ed_batch(50 + i, B) for i in range(2)]
    refs = [O.ed_train_step(oparams, b, st) for b in batches]

    class Loader:
        def __iter__(self):
            return iter({"x": b["x"], "y": b["y"]} for b in batches)

    masks = iter([(b["mask1"].cuda(), b["mask2"].cuda()) for b in batches])
    fwd = model.forward
    model.forward = lambda x: fwd(x, masks=next(masks)) if model.training else fwd(x)
    loss, acc = train_ed.run_epoch(model, Loader(), nn.CrossEntropyLoss(), opt, torch.device("cuda"), is_train=True)
    want = np.mean([r["loss"].item() for r in refs])
    assert abs(loss - want) <= 2e-5 * want
    assert abs(acc - np.mean([r["acc"].item() for r in refs])) < 1e-6
    sd = model.state_dict()
    for k in ("classifier.head.weight", "encoder.project.weight", "encoder.conv.3.net.0.weight", "encoder.conv.1.net.1.weight",
              "encoder.conv.2.net.1.running_var"):
        assert_close(sd[k], oparams[k], 1e-3, "after 2 AdamW steps: " + k)
    assert int(sd["encoder.conv.0.net.1.num_batches_tracked"]) == 2
    model.eval()
    with torch.no_grad():
        lg = model(batches[0]["x"].cuda())
    assert_close(lg, O.ed_forward(oparams, batches[0]["x"]), 2e-3, "eval logits after training")
