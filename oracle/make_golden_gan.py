"""Generates tests/golden/gan_golden.npz by running the REFERENCE's own modules and the verbatim loop
body of src/gan/train_gan.py:183-251 (D-step) / :212-251 (G-step) on seed-reproducible inputs.

    python oracle/make_golden_gan.py        (build container only: needs /root/reference)

The reference draws noise / alpha / dropout masks from torch's RNG; here torch.randn, torch.rand and
F.dropout are patched for the duration of a step so that the reference code consumes the injected
tensors of gan_oracle.make_batch (RNG row).  Everything else is the reference.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "melo-gan_b200"))
sys.path.insert(0, ROOT)

from oracle import gan_oracle as O  # noqa: E402


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


ref_models = _load("ref_models", "src/gan/models.py")
ref_fe = _load("ref_fe", "src/gan/feature_encoder.py")
ref_ed = _load("ref_ed", "src/emotion_discriminator/ed_model.py")
ref_utils = _load("ref_utils", "src/gan/utils.py")


class Inject:
    """Feeds queued tensors to torch.randn / torch.rand / F.dropout while active."""

    def __init__(self, randn=(), rand=(), masks=()):
        self.q = {"randn": list(randn), "rand": list(rand), "masks": list(masks)}

    def __enter__(self):
        self.saved = (torch.randn, torch.rand, torch.nn.functional.dropout)
        q = self.q

        def randn(*shape, **kw):
            t = q["randn"].pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        def rand(*shape, **kw):
            t = q["rand"].pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        def dropout(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            m = q["masks"].pop(0)
            return x * (m * (1.0 / (1.0 - p)))

        torch.randn, torch.rand, torch.nn.functional.dropout = randn, rand, dropout
        return self

    def __exit__(self, *a):
        torch.randn, torch.rand, torch.nn.functional.dropout = self.saved
        assert not any(self.q.values()), "injected tensors left unused"


def build_reference(params):
    with open(os.path.join(REF, "config/gan_config.yaml")) as f:
        cfg = yaml.safe_load(f)
    with open(os.path.join(REF, "config/ed_config.yaml")) as f:
        ed_cfg = yaml.safe_load(f)
    device = torch.device("cpu")
    # train_gan.py:85-133
    numeric_input_dim = cfg.get('NUMERIC_INPUT_DIM', 6)
    numeric_embed_dim = cfg.get('ENCODER_OUT_DIM', 128)
    E_num = ref_fe.FeatureEncoder(in_dim=numeric_input_dim, hidden_dims=cfg.get('ENCODER_HIDDEN', [256, 128]),
                                  out_dim=numeric_embed_dim).to(device)
    G = ref_models.Generator(noise_dim=cfg['NOISE_DIM'], latent_dim=cfg['LATENT_DIM'],
                             mode=cfg.get('INTEGRATION_MODE', 'conditioning'), max_notes=cfg['MAX_NOTES'],
                             note_dim=cfg['NOTE_DIM'], numeric_embed_dim=numeric_embed_dim).to(device)
    D = ref_models.Discriminator(max_notes=cfg['MAX_NOTES'], note_dim=cfg['NOTE_DIM'],
                                 numeric_embed_dim=numeric_embed_dim).to(device)
    ED = ref_ed.EmotionDiscriminator(ed_cfg).to(device)
    for mod, key in ((E_num, "E"), (G, "G"), (D, "D"), (ED, "ED")):
        missing, unexpected = mod.load_state_dict(params[key], strict=False)
        assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    for p in ED.parameters():
        p.requires_grad = False
    ED.eval()
    opt_G = optim.Adam(list(G.parameters()) + list(E_num.parameters()), lr=float(cfg['LR_G']),
                       betas=(cfg.get('BETA1', 0.5), cfg.get('BETA2', 0.9)))
    opt_D = optim.Adam(D.parameters(), lr=float(cfg['LR_D']), betas=(cfg.get('BETA1', 0.5), cfg.get('BETA2', 0.9)))
    G.train(); E_num.train(); D.train()
    return cfg, ed_cfg, E_num, G, D, ED, opt_G, opt_D


def ref_d_step(cfg, E_num, G, D_discriminator, opt_D, batch):
    device = torch.device("cpu")
    notes_real, numeric_batch = batch["notes_real"], batch["numeric"]
    bsize = notes_real.size(0)
    encoder_latent = torch.zeros(bsize, cfg['LATENT_DIM'])
    lambda_gp = cfg.get('LAMBDA_GP', 10.0)
    with Inject(randn=[batch["noise_d"]], rand=[batch["alpha"].view(-1, 1, 1)],
                masks=[batch["mask1_d"], batch["mask2_d"]]):
        # ---- verbatim train_gan.py:183-205 ----
        opt_D.zero_grad()
        with torch.no_grad():
            numeric_emb_d = E_num(numeric_batch)
            noise = torch.randn(bsize, cfg['NOISE_DIM'], device=device)
            gen_notes_d, _ = G(noise, encoder_latent, numeric_emb_d)
        d_real = D_discriminator(notes_real, numeric_emb_d)
        d_fake = D_discriminator(gen_notes_d.detach(), numeric_emb_d)
        gp = ref_utils.compute_gradient_penalty(D_discriminator, notes_real.data, gen_notes_d.data, numeric_emb_d, device)
        loss_d = torch.mean(d_fake) - torch.mean(d_real) + (lambda_gp * gp)
        loss_d.backward()
        grads = {k: p.grad.detach().clone() for k, p in D_discriminator.named_parameters()}
        opt_D.step()
    return {"loss_d": loss_d.item(), "gp": gp.item(), "d_real": d_real.detach(), "d_fake": d_fake.detach(),
            "fake": gen_notes_d.detach().contiguous(), "emb": numeric_emb_d, "grads": grads}


def ref_g_step(cfg, ed_cfg, E_num, G, D_discriminator, D_emotion, opt_G, batch):
    device = torch.device("cpu")
    numeric_batch, emot_idx = batch["numeric"], batch["emot_idx"]
    bsize = numeric_batch.size(0)
    encoder_latent = torch.zeros(bsize, cfg['LATENT_DIM'])
    criterion_emo = nn.CrossEntropyLoss()
    lambda_emotion = cfg.get('LAMBDA_EMOTION', 1.0)
    with Inject(randn=[batch["noise_g"]], masks=[batch["mask1_g"], batch["mask2_g"]]):
        # ---- verbatim train_gan.py:212-251 ----
        opt_G.zero_grad()
        numeric_emb_g = E_num(numeric_batch)
        noise_g = torch.randn(bsize, cfg['NOISE_DIM'], device=device)
        gen_notes_g, gen_latent_g = G(noise_g, encoder_latent, numeric_emb_g)
        d_fake_g = D_discriminator(gen_notes_g, numeric_emb_g)
        loss_g_adv = -torch.mean(d_fake_g)
        ed_input_mode = ed_cfg.get('input_mode', 'notes')
        ed_input = gen_latent_g if ed_input_mode == 'latent' else gen_notes_g
        ed_logits = D_emotion(ed_input)
        loss_g_emo_cls = criterion_emo(ed_logits, emot_idx)
        loss_g = loss_g_adv + (lambda_emotion * loss_g_emo_cls)
        loss_g.backward()
        gG = {k: p.grad.detach().clone() for k, p in G.named_parameters()}
        gE = {k: p.grad.detach().clone() for k, p in E_num.named_parameters()}
        opt_G.step()
    return {"loss_g_adv": loss_g_adv.item(), "loss_g_emo": loss_g_emo_cls.item(),
            "notes": gen_notes_g.detach().contiguous(), "latent": gen_latent_g.detach(), "logits": ed_logits.detach(),
            "grads_G": gG, "grads_E": gE}


def tstats(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, steps=min(8, t.numel())).long()
    return np.concatenate([[t.sum().item(), t.norm().item(), t.abs().max().item()], t[idx].numpy()])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {}
    # ---- case A: B=8, one full cycle (5 D-steps + 1 G-step), free-running, fresh batch per step ----
    for case, B, pseed in (("A", 8, 1), ("B", 32, 2), ("C", 8, 4)):
        params = O.make_params(pseed, fan_in_scale=(case == "C"))
        cfg, ed_cfg, E_num, G, D, ED, opt_G, opt_D = build_reference(params)
        nsteps = {"A": 5, "B": 1, "C": 2}[case]
        for i in range(nsteps):
            batch = O.make_batch(10 * pseed + i, B)
            d = ref_d_step(cfg, E_num, G, D, opt_D, batch)
            out[f"{case}.d{i}.scalars"] = np.array([d["loss_d"], d["gp"], d["d_real"].mean().item(), d["d_fake"].mean().item()])
            for k, g in d["grads"].items():
                out[f"{case}.d{i}.grad.{k}"] = tstats(g)
            if i == 0:
                out[f"{case}.d0.fake0"] = d["fake"][0].numpy().copy()
                out[f"{case}.d0.emb0"] = d["emb"][0].numpy().copy()
                out[f"{case}.d0.d_real"] = d["d_real"].numpy().copy()
                out[f"{case}.d0.d_fake"] = d["d_fake"].numpy().copy()
        g = ref_g_step(cfg, ed_cfg, E_num, G, D, ED, opt_G, batch)
        out[f"{case}.g.scalars"] = np.array([g["loss_g_adv"], g["loss_g_emo"]])
        out[f"{case}.g.notes0"] = g["notes"][0].numpy().copy()
        out[f"{case}.g.latent"] = g["latent"].numpy().copy()
        out[f"{case}.g.logits"] = g["logits"].numpy().copy()
        for k, t in g["grads_G"].items():
            out[f"{case}.g.gradG.{k}"] = tstats(t)
        for k, t in g["grads_E"].items():
            out[f"{case}.g.gradE.{k}"] = tstats(t)
        for mod, key in ((E_num, "E"), (G, "G"), (D, "D")):
            for k, t in mod.state_dict().items():
                if not k.endswith("num_batches_tracked"):
                    out[f"{case}.final.{key}.{k}"] = tstats(t)
    # ---- module forwards in eval mode (generation path, app.py:97-106) ----
    params = O.make_params(3)
    cfg, ed_cfg, E_num, G, D, ED, _, _ = build_reference(params)
    E_num.eval(); G.eval(); D.eval()
    b = O.make_batch(77, 4)
    with torch.no_grad():
        emb = E_num(b["numeric"])
        notes, lat = G(b["noise_d"], torch.zeros(4, 64), emb)
        out["eval.emb"] = emb.numpy().copy()
        out["eval.notes"] = notes.contiguous().numpy().copy()
        out["eval.latent"] = lat.numpy().copy()
        out["eval.score"] = D(notes, emb).numpy().copy()
        out["eval.logits"] = ED(notes).numpy().copy()
    # ---- A-13: the reference's ED training iteration (train_ed.py:61-74) on its own module, 2 steps ----
    params = O.make_params(5)
    with open(os.path.join(REF, "config/ed_config.yaml")) as f:
        ed_cfg = yaml.safe_load(f)
    model = ref_ed.EmotionDiscriminator(ed_cfg)
    missing, unexpected = model.load_state_dict(params["ED"], strict=False)
    assert not unexpected
    o = ed_cfg["optimizer"]
    optimizer = optim.AdamW(model.parameters(), lr=float(o["lr"]), weight_decay=o.get("weight_decay", 0), betas=tuple(o["betas"]))
    criterion = nn.CrossEntropyLoss()
    model.train()
    for i in range(2):
        eb = O.make_ed_batch(50 + i, 16)
        with Inject(masks=[eb["mask1"], eb["mask2"]]):
            optimizer.zero_grad()
            logits = model(eb["x"])
            loss = criterion(logits, eb["y"])
            loss.backward()
            grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            optimizer.step()
        out[f"ED.s{i}.scalars"] = np.array([loss.item(), (logits.argmax(1) == eb["y"]).float().mean().item()])
        out[f"ED.s{i}.logits"] = logits.detach().numpy().copy()
        for k, g in grads.items():
            out[f"ED.s{i}.grad.{k}"] = tstats(g)
    for k, t in model.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            out[f"ED.final.{k}"] = tstats(t)
    # ---- 8f-2: the reference's Generator in 'conditioning' mode (models.py:99-100,121-123): forward + all gradients ----
    Gc = ref_models.Generator(noise_dim=128, latent_dim=64, mode="conditioning", max_notes=512, note_dim=4,
                              numeric_embed_dim=128)
    pc = O.make_cond_params(8)
    missing, unexpected = Gc.load_state_dict(pc, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    Gc.train()
    cb = O.make_cond_batch(80, 8)
    emb_c = cb["emb"].clone().requires_grad_(True)
    notes_c, latent_c = Gc(cb["noise"], cb["cond"], emb_c)
    loss_c = (notes_c * cb["w_notes"]).sum() + (latent_c * cb["w_latent"]).sum()
    loss_c.backward()
    out["COND.notes0"] = notes_c[0].detach().numpy().copy()
    out["COND.latent"] = latent_c.detach().numpy().copy()
    out["COND.demb"] = tstats(emb_c.grad)
    for k, p_ in Gc.named_parameters():
        out[f"COND.grad.{k}"] = tstats(p_.grad)

    # ---- A-12: the reference's VAE (src/ae/model.py) and training iteration (train_ae.py:114-122), 2 steps ----
    ref_ae = _load("ref_ae_model", "src/ae/model.py")
    vp = O.make_vae_params(6)
    vae = ref_ae.VAE({"LATENT_DIM": 8, "MAX_NOTES": 512})
    with torch.no_grad():
        vae(torch.zeros(2, 512, 4))                       # materialises the lazy encoder._linear (train_ae.py:75-77)
    missing, unexpected = vae.load_state_dict(vp, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    opt = optim.AdamW(vae.parameters(), lr=1e-4, weight_decay=1e-5)
    vae.train()
    real_randn_like = torch.randn_like
    for i in range(2):
        vb = O.make_vae_batch(70 + i, 8)
        torch.randn_like = lambda t, _e=vb["eps"]: _e.clone()
        try:
            recon, z, mu, log_var = vae(vb["x"])
        finally:
            torch.randn_like = real_randn_like
        recon_loss = torch.nn.functional.mse_loss(recon, vb["x"])
        kld_loss = -0.5 * torch.mean(1 + log_var - mu.pow(2) - log_var.exp())
        loss = recon_loss + 10.0 * kld_loss                 # vae_loss (train_ae.py:35-51), beta = BETA of ae_config.yaml
        opt.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in vae.named_parameters()}
        torch.nn.utils.clip_grad_norm_(vae.parameters(), max_norm=1.0)
        opt.step()
        out[f"VAE.s{i}.scalars"] = np.array([loss.item(), recon_loss.item(), kld_loss.item()])
        out[f"VAE.s{i}.recon0"] = recon[0].detach().numpy().copy()
        out[f"VAE.s{i}.mu"] = mu.detach().numpy().copy()
        out[f"VAE.s{i}.log_var"] = log_var.detach().numpy().copy()
        for k, g in grads.items():
            out[f"VAE.s{i}.grad.{k}"] = tstats(g)
    for k, t in vae.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            out[f"VAE.final.{k}"] = tstats(t)
    path = os.path.join(ROOT, "tests", "golden", "gan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
    for k in ("A.d0.scalars", "A.d4.scalars", "A.g.scalars", "B.d0.scalars", "B.g.scalars", "C.d0.scalars",
              "C.d1.scalars", "C.g.scalars"):
        print(k, out[k])


if __name__ == "__main__":
    main()
