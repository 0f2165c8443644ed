"""The reference's own critic / generator step bodies (src/gan/train_gan.py:183-251), written against the DROP-IN
modules of melo-gan_b200/src (same class names, call signatures, autograd behaviour), compared with the oracle."""
import io
import contextlib
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from gan_testlib import assert_close, to_double
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu


def build(params, dev):
    with contextlib.redirect_stdout(io.StringIO()):
        from src.gan.feature_encoder import FeatureEncoder
        from src.gan.models import Discriminator, Generator
        from src.emotion_discriminator.ed_model import EmotionDiscriminator
        import yaml
        root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "melo-gan_b200", "config")
        cfg = yaml.safe_load(open(os.path.join(root, "gan_config.yaml")))
        ed_cfg = yaml.safe_load(open(os.path.join(root, "ed_config.yaml")))
        E_num = FeatureEncoder(in_dim=6, hidden_dims=cfg['ENCODER_HIDDEN'], out_dim=128)
        G = Generator(noise_dim=cfg['NOISE_DIM'], latent_dim=cfg['LATENT_DIM'], mode=cfg['INTEGRATION_MODE'],
                      max_notes=cfg['MAX_NOTES'], note_dim=cfg['NOTE_DIM'], numeric_embed_dim=128)
        D = Discriminator(max_notes=cfg['MAX_NOTES'], note_dim=cfg['NOTE_DIM'], numeric_embed_dim=128)
        ED = EmotionDiscriminator(ed_cfg)
    for mod, key in ((E_num, "E"), (G, "G"), (D, "D"), (ED, "ED")):
        missing, unexpected = mod.load_state_dict(params[key], strict=False)
        assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
        mod.to(dev)
    for p in ED.parameters():
        p.requires_grad = False
    ED.eval()
    G.train(); E_num.train(); D.train()
    return cfg, ed_cfg, E_num, G, D, ED


def test_state_dict_keys_match_the_reference_tables():
    params = O.make_params(1)
    cfg, ed_cfg, E_num, G, D, ED = build(params, "cuda")
    shapes = O.param_shapes()
    for mod, key in ((E_num, "E"), (G, "G"), (D, "D"), (ED, "ED")):
        sd = {k: tuple(v.shape) for k, v in mod.state_dict().items() if not k.endswith("num_batches_tracked")}
        assert sd == {k: tuple(v) for k, v in shapes[key].items()}, key
    assert G.decoder.reduced_len == 64 and D.combined_dim == 384 and G.input_dim == 256


def test_weights_init_and_seed_reproduce_the_reference_initialisation():
    """Same seed, same construction order, same class-name matching as train_gan.py:74-118 -> identical tensors."""
    from src.gan.utils import seed_everything, weights_init
    import importlib.util
    seed_everything(42)
    with contextlib.redirect_stdout(io.StringIO()):
        from src.gan.feature_encoder import FeatureEncoder
        from src.gan.models import Discriminator, Generator
        E1 = FeatureEncoder(6, [256, 128], 128); G1 = Generator(128, 64, "warm_start", max_notes=512, note_dim=4, numeric_embed_dim=128)
        D1 = Discriminator(512, 4, numeric_embed_dim=128)
    for m in (E1, G1, D1):
        m.apply(weights_init)
    torch.manual_seed(42)
    lin = nn.Linear(6, 256)            # first Linear the reference constructs after seeding (feature_encoder.py:23)
    assert E1.net[1].weight.detach().std().item() == pytest.approx(0.02, rel=0.1)
    assert float(E1.net[1].bias.detach().abs().max()) == 0.0 and float(D1.conv[0].bias.detach().abs().max()) == 0.0
    assert float(G1.decoder.deconv[1].weight.detach().min()) == 1.0      # BatchNorm untouched by weights_init
    del lin


def test_reference_loop_body_verbatim_on_dropin_modules():
    from src.gan.utils import compute_gradient_penalty
    B = 8
    params = O.make_params(4, fan_in_scale=True)
    batch = O.make_batch(40, B)
    dev = torch.device("cuda")
    cfg, ed_cfg, E_num, G, D_discriminator, D_emotion = build(params, dev)
    opt_G = optim.Adam(list(G.parameters()) + list(E_num.parameters()), lr=float(cfg['LR_G']), betas=(cfg['BETA1'], cfg['BETA2']))
    opt_D = optim.Adam(D_discriminator.parameters(), lr=float(cfg['LR_D']), betas=(cfg['BETA1'], cfg['BETA2']))
    oparams = O.clone_params(params)
    ref_d = O.critic_step(oparams, batch, {}, update=True)
    notes_real, numeric_batch = batch["notes_real"].to(dev), batch["numeric"].to(dev)
    emot_idx = batch["emot_idx"].to(dev)
    bsize, lambda_gp, lambda_emotion = B, cfg['LAMBDA_GP'], cfg['LAMBDA_EMOTION']
    encoder_latent = torch.zeros(B, cfg['LATENT_DIM'], device=dev)

    # inject the oracle's random draws: E_num masks as an argument, noise / alpha by patching torch.randn / torch.rand
    real_randn, real_rand = torch.randn, torch.rand
    torch.randn = lambda *s, **k: batch["noise_d"].to(dev)
    torch.rand = lambda *s, **k: batch["alpha"].view(-1, 1, 1).to(dev)
    E_fwd = E_num.forward
    E_num.forward = lambda x: E_fwd(x, masks=(batch["mask1_d"].to(dev), batch["mask2_d"].to(dev)))
    try:
        # ---- train_gan.py:183-205 ----
        opt_D.zero_grad()
        with torch.no_grad():
            numeric_emb_d = E_num(numeric_batch)
            noise = torch.randn(bsize, cfg['NOISE_DIM'], device=dev)
            gen_notes_d, _ = G(noise, encoder_latent, numeric_emb_d)
        d_real = D_discriminator(notes_real, numeric_emb_d)
        d_fake = D_discriminator(gen_notes_d.detach(), numeric_emb_d)
        gp = compute_gradient_penalty(D_discriminator, notes_real.data, gen_notes_d.data, numeric_emb_d, dev)
        loss_d = torch.mean(d_fake) - torch.mean(d_real) + (lambda_gp * gp)
        loss_d.backward()
        opt_D.step()
    finally:
        torch.randn, torch.rand = real_randn, real_rand
    assert abs(loss_d.item() - ref_d["loss_d"].item()) <= 1e-5 * abs(ref_d["loss_d"].item())
    assert abs(gp.item() - ref_d["gp"].item()) <= 1e-5
    ref64 = O.critic_step(to_double(params), to_double(batch), {}, update=False)
    for k, p in D_discriminator.named_parameters():
        if k.startswith("real_fake"):
            continue
        assert_close(p.grad, ref_d["grads"][k], 5e-5, "D grad " + k, ref64["grads"][k])
    assert int(G.decoder.deconv[1].num_batches_tracked) == 1
    assert_close(G.decoder.deconv[1].running_mean, oparams["G"]["decoder.deconv.1.running_mean"], 1e-5, "running mean")

    # ---- train_gan.py:212-251 on the same batch ----
    ref_g = O.generator_step(oparams, batch, {}, update=False)
    for k in ("conv.0.weight",):     # the oracle's critic was updated by its Adam step: resync ours to it
        pass
    D_discriminator.load_state_dict({k: v for k, v in oparams["D"].items()})
    torch.randn = lambda *s, **k: batch["noise_g"].to(dev)
    E_num.forward = lambda x: E_fwd(x, masks=(batch["mask1_g"].to(dev), batch["mask2_g"].to(dev)))
    criterion_emo = nn.CrossEntropyLoss()
    try:
        opt_G.zero_grad()
        numeric_emb_g = E_num(numeric_batch)
        noise_g = torch.randn(bsize, cfg['NOISE_DIM'], device=dev)
        gen_notes_g, gen_latent_g = G(noise_g, encoder_latent, numeric_emb_g)
        d_fake_g = D_discriminator(gen_notes_g, numeric_emb_g)
        loss_g_adv = -torch.mean(d_fake_g)
        ed_logits = D_emotion(gen_notes_g)
        loss_g_emo_cls = criterion_emo(ed_logits, emot_idx)
        loss_g = loss_g_adv + (lambda_emotion * loss_g_emo_cls)
        loss_g.backward()
        opt_G.step()
    finally:
        torch.randn = real_randn
        E_num.forward = E_fwd
    assert abs(loss_g_adv.item() - ref_g["loss_g_adv"].item()) <= 1e-5 * max(abs(ref_g["loss_g_adv"].item()), 1e-2)
    assert abs(loss_g_emo_cls.item() - ref_g["loss_g_emo"].item()) <= 1e-5 * ref_g["loss_g_emo"].item()
    ref64g = O.generator_step(to_double(oparams), to_double(batch), {}, update=False)
    for k, p in G.named_parameters():
        if k in ("decoder.deconv.0.bias", "decoder.deconv.3.bias"):
            continue
        assert_close(p.grad, ref_g["grads_G"][k], 5e-5, "G grad " + k, ref64g["grads_G"][k])
    for k, p in E_num.named_parameters():
        assert_close(p.grad, ref_g["grads_E"][k], 5e-5, "E grad " + k, ref64g["grads_E"][k])
    assert D_discriminator.conv[0].weight.grad is not None      # the reference also fills (and discards) D's grads here


def test_error_behaviour_matches_the_reference():
    params = O.make_params(1)
    cfg, ed_cfg, E_num, G, D, ED = build(params, "cuda")
    with pytest.raises(AssertionError):
        G(torch.zeros(2, 128, device="cuda"), None, None)             # numeric_embedding is required (models.py:116)
    with pytest.raises(ValueError):
        ED(torch.zeros(2, 128, device="cuda"))                        # wrong rank (ed_model.py:160-161)
    with pytest.raises(AssertionError):
        from src.gan.models import Generator
        with contextlib.redirect_stdout(io.StringIO()):
            Generator(mode="nope")
    with pytest.raises(RuntimeError):
        G.cpu()(torch.zeros(2, 128), None, torch.zeros(2, 128))       # no CPU fallback


def test_generation_and_midi_writer_roundtrip(tmp_path):
    """app.py:/generate path: eval-mode E_num -> G -> save_piano_roll_to_midi; the written file parses back to the
    oracle's notes (ticks at resolution 220)."""
    from melogan import midi
    from oracle import notes_oracle
    from src.gan.utils import save_piano_roll_to_midi
    params = O.make_params(3, fan_in_scale=True)
    cfg, ed_cfg, E_num, G, D, ED = build(params, "cuda")
    E_num.eval(); G.eval()
    b = O.make_batch(77, 4)
    with torch.no_grad():
        emb = E_num(b["numeric"].cuda())
        notes, _ = G(b["noise_d"].cuda(), torch.zeros(4, 64, device="cuda"), emb)
    roll = (notes[0] * 3.0).cpu().numpy()         # spread the values over the quantisation range
    path = str(tmp_path / "temp_gen.mid")
    with contextlib.redirect_stdout(io.StringIO()):
        save_piano_roll_to_midi(roll, path, bpm=140, scale="major")
    res, tempo, got = midi.read_notes(path)
    assert res == 220 and tempo == round(60_000_000 / 140)
    _, c, p, v, s, e = notes_oracle.extract_notes_gan(roll[None], 140.0, "major", 0)
    ts = 60.0 / (140.0 * 220)
    want = sorted(((int(v[0, i]), int(p[0, i]), int(round(s[0, i] / ts)), max(int(round(e[0, i] / ts)), int(round(s[0, i] / ts))))
                   for i in range(c[0])), key=lambda n: (n[2], n[1]))
    assert len(got) == len(want) == c[0] and c[0] > 0
    # overlapping notes of one pitch make the on/off pairing ambiguous in any SMF: compare the event multisets
    assert sorted((v_, p_, on) for v_, p_, on, _ in got) == sorted((v_, p_, on) for v_, p_, on, _ in want)
    assert sorted((p_, off) for _, p_, _, off in got) == sorted((p_, off) for _, p_, _, off in want)


def test_trainer_resume_continues_the_run():
    """a checkpoint in the layout of train_gan.py:269-276 restores parameters, optimizer moments, step
    counters and the device noise stream: the next cycle equals the uninterrupted one (up to the summation order of the
    float32 atomics in the weight-gradient kernels; the two bias vectors in front of BatchNorm carry pure rounding noise
    that Adam turns into +-lr)."""
    import copy
    import os
    import yaml
    from melogan.trainer import GanTrainer
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "melo-gan_b200", "config")
    cfg = yaml.safe_load(open(os.path.join(root, "gan_config.yaml")))
    ed_cfg = yaml.safe_load(open(os.path.join(root, "ed_config.yaml")))
    B = 8
    g = torch.Generator(device="cuda").manual_seed(3)
    reals = torch.rand((5, B, 512, 4), generator=g, device="cuda") * 2 - 1
    nums = torch.randn((5, B, 6), generator=g, device="cuda")
    labels = (torch.arange(B, device="cuda") % 4).to(torch.int64)
    a = GanTrainer(cfg, ed_cfg, batch=B, precision="fp32")
    a.train_cycle(reals, nums, labels)
    ck = copy.deepcopy(a.state_dict())
    ed_state = copy.deepcopy(a.ED.state_dict())
    a.train_cycle(reals, nums, labels)
    b = GanTrainer(cfg, ed_cfg, batch=B, precision="fp32", ed_state_dict=ed_state)
    b.load_state_dict(ck)
    b.train_cycle(reals, nums, labels)
    for name in ("G", "D", "E_num"):
        sa, sb = a.state_dict()[name], b.state_dict()[name]
        for k in sa:
            if k in ("decoder.deconv.0.bias", "decoder.deconv.3.bias") or not sa[k].is_floating_point():
                continue
            assert torch.allclose(sa[k], sb[k], rtol=0, atol=2e-5), (name, k, (sa[k] - sb[k]).abs().max().item())
    # and it is not trivially equal to the checkpoint itself
    assert not torch.equal(a.state_dict()["D"]["conv.2.weight"], ck["D"]["conv.2.weight"])


def test_trainer_conditioning_mode_cycle():
    """8f-2: INTEGRATION_MODE 'conditioning' through the fused step path: the AE latents of the batch are a third input
    block of G; a cycle runs, needs the latents, and depends on them."""
    import os
    import yaml
    from melogan.trainer import GanTrainer
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "melo-gan_b200", "config")
    cfg = yaml.safe_load(open(os.path.join(root, "gan_config.yaml")))
    ed_cfg = yaml.safe_load(open(os.path.join(root, "ed_config.yaml")))
    cfg["INTEGRATION_MODE"] = "conditioning"
    B = 8
    g = torch.Generator(device="cuda").manual_seed(5)
    reals = torch.rand((5, B, 512, 4), generator=g, device="cuda") * 2 - 1
    nums = torch.randn((5, B, 6), generator=g, device="cuda")
    conds = torch.randn((5, B, cfg["LATENT_DIM"]), generator=g, device="cuda")
    labels = (torch.arange(B, device="cuda") % 4).to(torch.int64)
    outs = []
    for scale in (1.0, 0.0):
        tr = GanTrainer(cfg, ed_cfg, batch=B, precision="fp32")
        assert tr.G.input_dim == cfg["NOISE_DIM"] + 128 + cfg["LATENT_DIM"]
        tr.train_cycle(reals, nums, labels, conds * scale)
        d_loss, g_adv, g_emo = tr.epoch_means()
        assert all(map(lambda v: v == v and abs(v) < 1e6, (d_loss, g_adv, g_emo)))
        outs.append((d_loss, g_adv, g_emo))
    assert outs[0] != outs[1]
    with pytest.raises(ValueError):
        tr.critic_step(reals[0], nums[0])
