"""gan_oracle.py -- CPU restatement of the reference's GAN training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under melo-gan_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker
(never as the thing shipped or measured as the product).

The reference's arithmetic for this path lives in a third-party dependency that is not under
/root/reference: PyTorch (unpinned by the reference; torch 2.11.0 in this image).  This file
restates the reference's *call sites* functionally -- same ATen ops, same layouts, same order --
over an explicit parameter dict that uses the reference's state_dict key names:

  A-1  FeatureEncoder.forward            src/gan/feature_encoder.py:17-45
  A-2  NoiseToLatent.forward             src/gan/models.py:20-29
  A-3  GeneratorDecoder.forward          src/gan/models.py:46-83
  A-4  Generator.forward                 src/gan/models.py:108-130
  A-5  Discriminator.forward             src/gan/models.py:140-169
  A-6  compute_gradient_penalty          src/gan/utils.py:75-90
  A-7  critic (D) step                   src/gan/train_gan.py:183-205
  A-8  EmotionDiscriminator.forward      src/emotion_discriminator/ed_model.py:35-69,92-95,147-165
  A-9  generator (G) step                src/gan/train_gan.py:212-251
  A-10 torch.optim.Adam.step             src/gan/train_gan.py:136-145,204,248

Random draws (noise, alpha, dropout keep-masks) are explicit inputs, never drawn here, so that the
reference, this oracle and the CUDA path can be fed identical values (RNG).

Parity: PINNED.  tests/test_oracle_gan.py checks this file against tests/golden/gan_golden.npz,
which oracle/make_golden_gan.py produced in the build container by running the reference's OWN
modules and the verbatim loop body of train_gan.py:183-251 on the same seeded inputs.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# gan_config.yaml values the hot path reads (train_gan.py:136-156)
CFG = dict(NOISE_DIM=128, LATENT_DIM=64, MAX_NOTES=512, NOTE_DIM=4, NUMERIC_INPUT_DIM=6, ENCODER_HIDDEN=(256, 128),
           ENCODER_OUT_DIM=128, GEN_HIDDEN_USED=512, LR_G=2e-4, LR_D=1e-4, BETA1=0.5, BETA2=0.9, LAMBDA_GP=10.0,
           LAMBDA_EMOTION=5.0, CRITIC_ITERS=5, ENC_DROPOUT=0.2, ADAM_EPS=1e-8)


# ------------------------------------------------------------------------------------------------
# parameter tables (reference state_dict key names and shapes)
# ------------------------------------------------------------------------------------------------
def param_shapes(cfg=CFG):
    nd, ld, T, C = cfg["NOISE_DIM"], cfg["LATENT_DIM"], cfg["MAX_NOTES"], cfg["NOTE_DIM"]
    ed, nin = cfg["ENCODER_OUT_DIM"], cfg["NUMERIC_INPUT_DIM"]
    h1, h2 = cfg["ENCODER_HIDDEN"]
    hid = cfg["GEN_HIDDEN_USED"]
    L0 = max(1, T // 8)
    E = {"net.0.weight": (nin,), "net.0.bias": (nin,), "net.1.weight": (h1, nin), "net.1.bias": (h1,),
         "net.4.weight": (h2, h1), "net.4.bias": (h2,), "net.7.weight": (ed, h2), "net.7.bias": (ed,)}
    G = {"noise_to_latent.net.0.weight": (hid, nd + ed), "noise_to_latent.net.0.bias": (hid,),
         "noise_to_latent.net.2.weight": (ld, hid), "noise_to_latent.net.2.bias": (ld,),
         "decoder.pre.0.weight": (512, ld), "decoder.pre.0.bias": (512,),
         "decoder.pre.2.weight": (256 * L0, 512), "decoder.pre.2.bias": (256 * L0,),
         "decoder.deconv.0.weight": (256, 128, 5), "decoder.deconv.0.bias": (128,),
         "decoder.deconv.1.weight": (128,), "decoder.deconv.1.bias": (128,),
         "decoder.deconv.1.running_mean": (128,), "decoder.deconv.1.running_var": (128,),
         "decoder.deconv.3.weight": (128, 64, 5), "decoder.deconv.3.bias": (64,),
         "decoder.deconv.4.weight": (64,), "decoder.deconv.4.bias": (64,),
         "decoder.deconv.4.running_mean": (64,), "decoder.deconv.4.running_var": (64,),
         "decoder.deconv.6.weight": (64, C, 5), "decoder.deconv.6.bias": (C,)}
    D = {"conv.0.weight": (64, C, 5), "conv.0.bias": (64,), "conv.2.weight": (128, 64, 5), "conv.2.bias": (128,),
         "conv.4.weight": (256, 128, 5), "conv.4.bias": (256,), "fc.1.weight": (256, 256), "fc.1.bias": (256,),
         "real_fake.weight": (1, 256 + ed), "real_fake.bias": (1,)}
    ED = {}
    chans = [(C, 64, 5), (64, 128, 3), (128, 256, 3), (256, 256, 3)]  # ed_model.py:52-58 with notes_hidden=256
    for i, (ci, co, k) in enumerate(chans):
        ED[f"encoder.conv.{i}.net.0.weight"] = (co, ci, k)
        ED[f"encoder.conv.{i}.net.0.bias"] = (co,)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            ED[f"encoder.conv.{i}.net.1.{nm}"] = (co,)
    ED.update({"encoder.project.weight": (256, 256), "encoder.project.bias": (256,),
               "classifier.net.0.weight": (256, 256), "classifier.net.0.bias": (256,),
               "classifier.net.3.weight": (128, 256), "classifier.net.3.bias": (128,),
               "classifier.head.weight": (4, 128), "classifier.head.bias": (4,)})
    return {"E": E, "G": G, "D": D, "ED": ED}


def is_buffer(name):
    return name.endswith("running_mean") or name.endswith("running_var")


def make_params(seed, cfg=CFG, weight_std=0.02, nontrivial=True, fan_in_scale=False):
    """Seed-reproducible test parameters (machine independent, see melogan.synth).

    Weight matrices ~ zero-mean with std `weight_std` (weights_init's scale, utils.py:37-45).
    With nontrivial=True the biases, norm affine terms and running statistics are random too, so
    every term of every layer is exercised (weights_init's zero biases would hide bias bugs).
    fan_in_scale=True gives every module O(1) activations and an O(1) critic gradient norm, so the
    gradient-penalty double backward is exercised away from the |grad| ~ 0 regime of a fresh init."""
    from melogan import synth
    out = {}
    k = 0
    for mod, table in param_shapes(cfg).items():
        P = {}
        for name, shape in table.items():
            k += 1
            s = seed * 1000 + k
            if name.endswith("running_var"):
                v = synth.uniform(s, shape, 0.5, 1.5) if nontrivial else np.ones(shape, np.float32)
            elif name.endswith("running_mean"):
                v = synth.uniform(s, shape, -0.1, 0.1) if nontrivial else np.zeros(shape, np.float32)
            elif len(shape) == 1 and name.endswith("weight"):      # LayerNorm / BatchNorm gamma
                v = synth.uniform(s, shape, 0.8, 1.2) if nontrivial else np.ones(shape, np.float32)
            elif name.endswith("bias"):
                v = synth.uniform(s, shape, -0.05, 0.05) if nontrivial else np.zeros(shape, np.float32)
            else:
                std = weight_std
                if mod == "ED" or fan_in_scale:                       # "trained-like": default-init scale
                    fan_in = int(np.prod(shape[1:]))
                    std = 1.0 / math.sqrt(3.0 * fan_in) * 1.7
                v = synth.pseudo_normal(s, shape, std)
            P[name] = torch.from_numpy(np.ascontiguousarray(v))
        out[mod] = P
    return out


def make_batch(seed, B, cfg=CFG):
    """Synthetic step inputs per SURVEY.md 8(d): notes U(-1,1), numeric ~N(0,1) with column 5 == 0,
    balanced labels, noise ~N(0,1)-like, alpha U[0,1), dropout keep-masks Bernoulli(0.8)."""
    from melogan import synth
    h1, h2 = cfg["ENCODER_HIDDEN"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    keep = 1.0 - cfg["ENC_DROPOUT"]
    return {
        "notes_real": t(synth.uniform(seed * 100 + 1, (B, cfg["MAX_NOTES"], cfg["NOTE_DIM"]))),
        "numeric": t(synth.numeric_features(seed * 100 + 2, B)),
        "emot_idx": t(synth.emotion_labels(B)),
        "noise_d": t(synth.pseudo_normal(seed * 100 + 3, (B, cfg["NOISE_DIM"]))),
        "alpha": t(synth.uniform(seed * 100 + 4, (B,), 0.0, 1.0)),
        "mask1_d": t((synth.uniform(seed * 100 + 5, (B, h1), 0.0, 1.0) < keep).astype(np.float32)),
        "mask2_d": t((synth.uniform(seed * 100 + 6, (B, h2), 0.0, 1.0) < keep).astype(np.float32)),
        "noise_g": t(synth.pseudo_normal(seed * 100 + 7, (B, cfg["NOISE_DIM"]))),
        "mask1_g": t((synth.uniform(seed * 100 + 8, (B, h1), 0.0, 1.0) < keep).astype(np.float32)),
        "mask2_g": t((synth.uniform(seed * 100 + 9, (B, h2), 0.0, 1.0) < keep).astype(np.float32)),
    }


# ------------------------------------------------------------------------------------------------
# forward passes (functional; P maps reference key -> tensor)
# ------------------------------------------------------------------------------------------------
def fe_forward(P, x, mask1=None, mask2=None, train=True, p_drop=0.2):
    """FeatureEncoder (feature_encoder.py:17-45).  Dropout keep-masks are inputs: y = x * mask / (1-p)."""
    h = F.layer_norm(x, (x.shape[1],), P["net.0.weight"], P["net.0.bias"], 1e-5)
    h = F.gelu(F.linear(h, P["net.1.weight"], P["net.1.bias"]))
    if train:
        h = h * (mask1 * (1.0 / (1.0 - p_drop)))
    h = F.gelu(F.linear(h, P["net.4.weight"], P["net.4.bias"]))
    if train:
        h = h * (mask2 * (1.0 / (1.0 - p_drop)))
    return F.linear(h, P["net.7.weight"], P["net.7.bias"])


def gen_forward(P, noise, emb, train=True, bn_state=None, momentum=0.1, keep=None, cond=None):
    """Generator (models.py:108-130): cat[noise, emb (, encoder latent in 'conditioning' mode)] -> NoiseToLatent -> decoder.
    bn_state (dict of running_mean/var tensors) is updated in place in train mode like nn.BatchNorm1d."""
    x = torch.cat([noise, emb] + ([cond] if cond is not None else []), dim=1)
    h = F.relu(F.linear(x, P["noise_to_latent.net.0.weight"], P["noise_to_latent.net.0.bias"]))
    latent = F.linear(h, P["noise_to_latent.net.2.weight"], P["noise_to_latent.net.2.bias"])
    y = F.relu(F.linear(latent, P["decoder.pre.0.weight"], P["decoder.pre.0.bias"]))
    y = F.relu(F.linear(y, P["decoder.pre.2.weight"], P["decoder.pre.2.bias"]))
    b = y.shape[0]
    y = y.view(b, 256, -1)
    if keep is not None:
        keep["pre2"] = y
    for conv, bn in (("decoder.deconv.0", "decoder.deconv.1"), ("decoder.deconv.3", "decoder.deconv.4")):
        y = F.conv_transpose1d(y, P[conv + ".weight"], P[conv + ".bias"], stride=2, padding=2, output_padding=1)
        if keep is not None:
            keep[conv] = y
        st = bn_state if bn_state is not None else P
        y = F.batch_norm(y, st[bn + ".running_mean"], st[bn + ".running_var"], P[bn + ".weight"], P[bn + ".bias"],
                         training=train, momentum=momentum, eps=1e-5)
        y = F.relu(y)
        if keep is not None:
            keep[bn] = y
    y = F.conv_transpose1d(y, P["decoder.deconv.6.weight"], P["decoder.deconv.6.bias"], stride=2, padding=2,
                           output_padding=1)
    out = y.permute(0, 2, 1)
    return out, latent


def disc_forward(P, notes, emb, keep=None):
    """Discriminator / WGAN critic (models.py:158-169)."""
    h = notes.permute(0, 2, 1)
    for i, name in enumerate(("conv.0", "conv.2", "conv.4")):
        h = F.leaky_relu(F.conv1d(h, P[name + ".weight"], P[name + ".bias"], stride=2, padding=2), 0.2)
        if keep is not None:
            keep[name] = h
    h = F.adaptive_avg_pool1d(h, 1)
    feat = F.leaky_relu(F.linear(h.view(h.size(0), -1), P["fc.1.weight"], P["fc.1.bias"]), 0.2)
    if emb is not None:
        feat = torch.cat([feat, emb], dim=1)
    return F.linear(feat, P["real_fake.weight"], P["real_fake.bias"]).squeeze(1)


def ed_forward(P, notes, keep=None):
    """EmotionDiscriminator, input_mode 'notes', eval mode (BN running stats, dropout off)."""
    x = notes.permute(0, 2, 1)
    for i in range(4):
        pre = f"encoder.conv.{i}.net."
        x = F.conv1d(x, P[pre + "0.weight"], P[pre + "0.bias"], stride=1, padding=2 if i == 0 else 1)
        x = F.batch_norm(x, P[pre + "1.running_mean"], P[pre + "1.running_var"], P[pre + "1.weight"],
                         P[pre + "1.bias"], training=False, eps=1e-5)
        x = F.gelu(x)
        if keep is not None:
            keep[f"conv{i}"] = x
    x = F.adaptive_avg_pool1d(x, 1).squeeze(-1)
    x = F.linear(x, P["encoder.project.weight"], P["encoder.project.bias"])
    x = F.gelu(F.linear(x, P["classifier.net.0.weight"], P["classifier.net.0.bias"]))
    x = F.gelu(F.linear(x, P["classifier.net.3.weight"], P["classifier.net.3.bias"]))
    return F.linear(x, P["classifier.head.weight"], P["classifier.head.bias"])


def gradient_penalty(PD, real, fake, emb, alpha):
    """compute_gradient_penalty (utils.py:75-90) with alpha (B,) injected."""
    a = alpha.view(-1, 1, 1).expand_as(real)
    interp = (a * real + (1 - a) * fake).requires_grad_(True)
    d_interp = disc_forward(PD, interp, emb)
    grads = torch.autograd.grad(outputs=d_interp, inputs=interp, grad_outputs=torch.ones_like(d_interp),
                                create_graph=True, retain_graph=True, only_inputs=True)[0]
    grads = grads.reshape(grads.size(0), -1)
    return ((grads.norm(2, dim=1) - 1) ** 2).mean()


# ------------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam single-tensor arithmetic, restated so that state is an explicit dict)
# ------------------------------------------------------------------------------------------------
def adam_update(params, grads, state, lr, beta1, beta2, eps=1e-8):
    """In-place on params/state; state = {"step": int, name: (exp_avg, exp_avg_sq)}."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    for name, p in params.items():
        g = grads[name]
        if name not in state:
            state[name] = (torch.zeros_like(p), torch.zeros_like(p))
        m, v = state[name]
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def _leaves(P, skip_buffers=True):
    return {k: v.detach().clone().requires_grad_(True) for k, v in P.items() if not (skip_buffers and is_buffer(k))}


# ------------------------------------------------------------------------------------------------
# the two step bodies
# ------------------------------------------------------------------------------------------------
def critic_step(params, batch, opt_state_d, cfg=CFG, update=True):
    """train_gan.py:183-205.  Mutates params["D"], params["G"] running stats and opt_state_d when update."""
    PE, PG, PD = params["E"], params["G"], params["D"]
    B = batch["notes_real"].shape[0]
    with torch.no_grad():
        emb = fe_forward(PE, batch["numeric"], batch["mask1_d"], batch["mask2_d"], train=True,
                         p_drop=cfg["ENC_DROPOUT"])
        fake, _ = gen_forward(PG, batch["noise_d"], emb, train=True, bn_state=PG if update else
                              {k: v.clone() for k, v in PG.items() if is_buffer(k)})
    Dl = _leaves(PD)
    d_real = disc_forward(Dl, batch["notes_real"], emb)
    d_fake = disc_forward(Dl, fake.detach(), emb)
    gp = gradient_penalty(Dl, batch["notes_real"], fake, emb, batch["alpha"])
    loss_d = torch.mean(d_fake) - torch.mean(d_real) + cfg["LAMBDA_GP"] * gp
    grads = dict(zip(Dl.keys(), torch.autograd.grad(loss_d, list(Dl.values()))))
    out = {"loss_d": loss_d.detach(), "gp": gp.detach(), "d_real": d_real.detach(), "d_fake": d_fake.detach(),
           "fake": fake, "emb": emb, "grads": grads}
    if update:
        with torch.no_grad():
            adam_update(PD, grads, opt_state_d, cfg["LR_D"], cfg["BETA1"], cfg["BETA2"], cfg["ADAM_EPS"])
    return out


def generator_step(params, batch, opt_state_g, cfg=CFG, update=True, skip_wasted_d_grads=True):
    """train_gan.py:212-251.  opt_G covers G and E_num (train_gan.py:136-140); D receives grads that the
    reference discards at the next opt_D.zero_grad() (:183) -- not produced here."""
    PE, PG, PD, PED = params["E"], params["G"], params["D"], params["ED"]
    El, Gl = _leaves(PE), _leaves(PG)
    bn_state = PG if update else {k: v.clone() for k, v in PG.items() if is_buffer(k)}
    emb = fe_forward(El, batch["numeric"], batch["mask1_g"], batch["mask2_g"], train=True, p_drop=cfg["ENC_DROPOUT"])
    notes, latent = gen_forward(Gl, batch["noise_g"], emb, train=True, bn_state=bn_state)
    d_fake = disc_forward(PD, notes, emb)
    loss_adv = -torch.mean(d_fake)
    logits = ed_forward(PED, notes)
    loss_emo = F.cross_entropy(logits, batch["emot_idx"])
    loss_g = loss_adv + cfg["LAMBDA_EMOTION"] * loss_emo
    leaves = list(Gl.values()) + list(El.values())
    gl = torch.autograd.grad(loss_g, leaves)
    gG = dict(zip(Gl.keys(), gl[:len(Gl)]))
    gE = dict(zip(El.keys(), gl[len(Gl):]))
    out = {"loss_g_adv": loss_adv.detach(), "loss_g_emo": loss_emo.detach(), "notes": notes.detach(),
           "latent": latent.detach(), "logits": logits.detach(), "grads_G": gG, "grads_E": gE}
    if update:
        with torch.no_grad():
            # parameter order of the reference's opt_G: list(G.parameters()) + list(E_num.parameters())
            joint_p = {("G." + k): v for k, v in PG.items() if not is_buffer(k)}
            joint_p.update({("E." + k): v for k, v in PE.items()})
            joint_g = {("G." + k): v for k, v in gG.items()}
            joint_g.update({("E." + k): v for k, v in gE.items()})
            adam_update(joint_p, joint_g, opt_state_g, cfg["LR_G"], cfg["BETA1"], cfg["BETA2"], cfg["ADAM_EPS"])
    return out


def train_cycle(params, batches, opt_state_d, opt_state_g, cfg=CFG):
    """CRITIC_ITERS D-steps, the last one followed by a G-step on the same batch (train_gan.py:168-251)."""
    outs = []
    for i, b in enumerate(batches):
        o = {"d": critic_step(params, b, opt_state_d, cfg)}
        if (i + 1) % cfg["CRITIC_ITERS"] == 0:
            o["g"] = generator_step(params, b, opt_state_g, cfg)
        outs.append(o)
    return outs


def clone_params(params):
    return {m: {k: v.clone() for k, v in P.items()} for m, P in params.items()}


# ------------------------------------------------------------------------------------------------
# A-13: EmotionDiscriminator training step (BASELINE config #3)
#       reference src/emotion_discriminator/train_ed.py:61-74 with ed_model.py in train mode
# ------------------------------------------------------------------------------------------------
ED_CFG = dict(lr=2e-4, betas=(0.5, 0.999), weight_decay=0.0, dropout=0.2, batch_size=64)


def ed_train_forward(P, notes, mask1, mask2, bn_state, p_drop=0.2, keep=None):
    """EmotionDiscriminator.forward in train mode: BatchNorm batch statistics (running stats in bn_state are
    advanced in place), MLP dropout with injected keep-masks."""
    x = notes.permute(0, 2, 1)
    for i in range(4):
        pre = f"encoder.conv.{i}.net."
        x = F.conv1d(x, P[pre + "0.weight"], P[pre + "0.bias"], stride=1, padding=2 if i == 0 else 1)
        x = F.batch_norm(x, bn_state[pre + "1.running_mean"], bn_state[pre + "1.running_var"], P[pre + "1.weight"],
                         P[pre + "1.bias"], training=True, momentum=0.1, eps=1e-5)
        x = F.gelu(x)
        if keep is not None:
            keep[f"conv{i}"] = x
    x = F.adaptive_avg_pool1d(x, 1).squeeze(-1)
    x = F.linear(x, P["encoder.project.weight"], P["encoder.project.bias"])
    x = F.gelu(F.linear(x, P["classifier.net.0.weight"], P["classifier.net.0.bias"])) * (mask1 * (1.0 / (1.0 - p_drop)))
    x = F.gelu(F.linear(x, P["classifier.net.3.weight"], P["classifier.net.3.bias"])) * (mask2 * (1.0 / (1.0 - p_drop)))
    return F.linear(x, P["classifier.head.weight"], P["classifier.head.bias"])


def adamw_update(params, grads, state, lr, beta1, beta2, weight_decay, eps=1e-8):
    """torch.optim.AdamW single-tensor arithmetic (decoupled decay first, then the Adam update)."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    for name, p in params.items():
        g = grads[name]
        if name not in state:
            state[name] = (torch.zeros_like(p), torch.zeros_like(p))
        m, v = state[name]
        p.mul_(1 - lr * weight_decay)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


def make_ed_batch(seed, B, cfg=CFG):
    from melogan import synth
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return {"x": t(synth.uniform(seed * 100 + 1, (B, cfg["MAX_NOTES"], cfg["NOTE_DIM"]))),
            "y": t(synth.emotion_labels(B)),
            "mask1": t((synth.uniform(seed * 100 + 2, (B, 256), 0.0, 1.0) < 0.8).astype(np.float32)),
            "mask2": t((synth.uniform(seed * 100 + 3, (B, 128), 0.0, 1.0) < 0.8).astype(np.float32))}


def ed_train_step(PED, batch, opt_state, update=True, cfg=ED_CFG):
    """One train_ed.run_epoch iteration: logits -> CE -> backward -> AdamW.  Mutates PED (incl. running stats)."""
    leaves = _leaves(PED)
    bn_state = PED if update else {k: v.clone() for k, v in PED.items() if is_buffer(k)}
    logits = ed_train_forward(leaves, batch["x"], batch["mask1"], batch["mask2"], bn_state, p_drop=cfg["dropout"])
    loss = F.cross_entropy(logits, batch["y"])
    grads = dict(zip(leaves.keys(), torch.autograd.grad(loss, list(leaves.values()))))
    acc = (logits.argmax(dim=1) == batch["y"]).float().mean()
    if update:
        with torch.no_grad():
            adamw_update({k: v for k, v in PED.items() if not is_buffer(k)}, grads, opt_state, cfg["lr"], cfg["betas"][0],
                         cfg["betas"][1], cfg["weight_decay"])
    return {"loss": loss.detach(), "acc": acc, "logits": logits.detach(), "grads": grads}


# ------------------------------------------------------------------------------------------------
# A-12: VAE forward / loss / training step (BASELINE config #2)
#       reference src/ae/model.py:4-148, src/ae/train_ae.py:35-51,114-122
# ------------------------------------------------------------------------------------------------
AE_CFG = dict(LATENT_DIM=8, MAX_NOTES=512, LR=1e-4, WEIGHT_DECAY=1e-5, BETA=10.0, BATCH_SIZE=32)


VAE_NOISE_BIASES = {f"encoder.conv.{c}.bias" for c in (0, 3, 6)} | {f"decoder.deconv.{c}.bias" for c in (0, 3)}


def vae_param_shapes(latent=8, T=512):
    L0 = max(1, T // 8)
    S = {}
    for (c, b), (ci, co) in zip(((0, 1), (3, 4), (6, 7)), ((4, 32), (32, 64), (64, 128))):
        S[f"encoder.conv.{c}.weight"] = (co, ci, 5); S[f"encoder.conv.{c}.bias"] = (co,)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            S[f"encoder.conv.{b}.{nm}"] = (co,)
    S.update({"encoder._linear.1.weight": (512, 128 * L0), "encoder._linear.1.bias": (512,),
              "fc_mu.weight": (latent, 512), "fc_mu.bias": (latent,), "fc_log_var.weight": (latent, 512),
              "fc_log_var.bias": (latent,), "decoder.pre.0.weight": (512, latent), "decoder.pre.0.bias": (512,),
              "decoder.pre.2.weight": (128 * L0, 512), "decoder.pre.2.bias": (128 * L0,)})
    for (c, b), (ci, co) in zip(((0, 1), (3, 4)), ((128, 64), (64, 32))):
        S[f"decoder.deconv.{c}.weight"] = (ci, co, 5); S[f"decoder.deconv.{c}.bias"] = (co,)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            S[f"decoder.deconv.{b}.{nm}"] = (co,)
    S["decoder.deconv.6.weight"] = (32, 4, 5); S["decoder.deconv.6.bias"] = (4,)
    return S


def make_vae_params(seed, latent=8, T=512):
    from melogan import synth
    P = {}
    for k, (name, shape) in enumerate(vae_param_shapes(latent, T).items()):
        s = seed * 1000 + 500 + k
        if name.endswith("running_var"):
            v = synth.uniform(s, shape, 0.5, 1.5)
        elif name.endswith("running_mean"):
            v = synth.uniform(s, shape, -0.1, 0.1)
        elif len(shape) == 1 and name.endswith("weight"):
            v = synth.uniform(s, shape, 0.8, 1.2)
        elif name.endswith("bias"):
            v = synth.uniform(s, shape, -0.05, 0.05)
        else:
            v = synth.pseudo_normal(s, shape, 1.7 / math.sqrt(3.0 * int(np.prod(shape[1:]))))
        P[name] = torch.from_numpy(np.ascontiguousarray(v))
    return P


def make_vae_batch(seed, B, latent=8, T=512):
    from melogan import synth
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return {"x": t(synth.uniform(seed * 100 + 1, (B, T, 4))), "eps": t(synth.pseudo_normal(seed * 100 + 2, (B, latent)))}


def vae_forward(P, x, eps, train=True, bn_state=None, masks=None, snap=1e-4, margins=None):
    """masks (test aid): name -> bool tensor of the ReLU decisions another implementation took.  Where a pre-activation is
    within `snap` of zero -- where float32 rounding alone decides the branch -- that decision is used instead of this
    forward's own, so that gradients can be compared on identical branches.  margins: list collecting min |pre-activation|."""
    st = bn_state if bn_state is not None else P

    def relu(name, t):
        if margins is not None:
            margins.append(t.detach().abs().min().item())
        if masks is None:
            return F.relu(t)
        fragile = t.detach().abs() < snap
        keep = torch.where(fragile, masks[name], t.detach() > 0)
        return t * keep.to(t.dtype)

    h = x.permute(0, 2, 1)
    for i, (c, b) in enumerate(((0, 1), (3, 4), (6, 7))):
        h = F.conv1d(h, P[f"encoder.conv.{c}.weight"], P[f"encoder.conv.{c}.bias"], stride=2, padding=2)
        h = relu(f"e_a{i}", F.batch_norm(h, st[f"encoder.conv.{b}.running_mean"], st[f"encoder.conv.{b}.running_var"],
                                         P[f"encoder.conv.{b}.weight"], P[f"encoder.conv.{b}.bias"], training=train, momentum=0.1,
                                         eps=1e-5))
    h = relu("h", F.linear(h.reshape(h.size(0), -1), P["encoder._linear.1.weight"], P["encoder._linear.1.bias"]))
    mu = F.linear(h, P["fc_mu.weight"], P["fc_mu.bias"])
    lv = F.linear(h, P["fc_log_var.weight"], P["fc_log_var.bias"])
    z = mu + eps * torch.exp(0.5 * lv)
    y = relu("d0", F.linear(z, P["decoder.pre.0.weight"], P["decoder.pre.0.bias"]))
    y = relu("d_y0", F.linear(y, P["decoder.pre.2.weight"], P["decoder.pre.2.bias"]))
    y = y.view(y.size(0), 128, -1)
    for i, (c, b) in enumerate(((0, 1), (3, 4))):
        y = F.conv_transpose1d(y, P[f"decoder.deconv.{c}.weight"], P[f"decoder.deconv.{c}.bias"], stride=2, padding=2,
                               output_padding=1)
        y = relu(f"d_y{i + 1}", F.batch_norm(y, st[f"decoder.deconv.{b}.running_mean"], st[f"decoder.deconv.{b}.running_var"],
                                             P[f"decoder.deconv.{b}.weight"], P[f"decoder.deconv.{b}.bias"], training=train,
                                             momentum=0.1, eps=1e-5))
    y = torch.tanh(F.conv_transpose1d(y, P["decoder.deconv.6.weight"], P["decoder.deconv.6.bias"], stride=2, padding=2,
                                      output_padding=1))
    return y.permute(0, 2, 1), z, mu, lv


def vae_loss(recon, target, mu, log_var, beta):
    recon_loss = F.mse_loss(recon, target)
    kld = -0.5 * torch.mean(1 + log_var - mu.pow(2) - log_var.exp())
    return recon_loss + beta * kld, recon_loss, kld


def vae_train_step(P, batch, opt_state, beta=10.0, update=True, cfg=AE_CFG, masks=None, margins=None):
    """train_ae.py:114-122: forward, vae_loss, backward, clip_grad_norm_(1.0), AdamW(lr, weight_decay)."""
    leaves = _leaves(P)
    bn_state = P if update else {k: v.clone() for k, v in P.items() if is_buffer(k)}
    recon, z, mu, lv = vae_forward(leaves, batch["x"], batch["eps"], True, bn_state, masks=masks, margins=margins)
    loss, rl, kl = vae_loss(recon, batch["x"], mu, lv, beta)
    grads = dict(zip(leaves.keys(), torch.autograd.grad(loss, list(leaves.values()))))
    out = {"loss": loss.detach(), "recon_loss": rl.detach(), "kld": kl.detach(), "recon": recon.detach(), "mu": mu.detach(),
           "log_var": lv.detach(), "z": z.detach(), "grads": {k: g.clone() for k, g in grads.items()}}
    if update:
        with torch.no_grad():
            total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
            coef = torch.clamp(1.0 / (total + 1e-6), max=1.0)
            clipped = {k: g * coef for k, g in grads.items()}
            adamw_update({k: v for k, v in P.items() if not is_buffer(k)}, clipped, opt_state, cfg["LR"], 0.9, 0.999,
                         cfg["WEIGHT_DECAY"])
    return out


# ------------------------------------------------------------------------------------------------
# 8f-2: Generator in 'conditioning' mode (models.py:99-100,112-126): the AE latent is a third input block
# ------------------------------------------------------------------------------------------------
def make_cond_params(seed, cfg=CFG):
    """Generator parameters with the wider first Linear (noise + embedding + AE latent columns)."""
    from melogan import synth
    P = make_params(seed, cfg, fan_in_scale=True)["G"]
    zin = cfg["NOISE_DIM"] + cfg.get("ENCODER_OUT_DIM", 128) + cfg["LATENT_DIM"]
    shape = (P["noise_to_latent.net.0.weight"].shape[0], zin)
    P["noise_to_latent.net.0.weight"] = torch.from_numpy(np.ascontiguousarray(
        synth.pseudo_normal(seed * 1000 + 901, shape, 1.7 / math.sqrt(3.0 * zin))))
    return P


def make_cond_batch(seed, B, cfg=CFG):
    from melogan import synth
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return {"noise": t(synth.pseudo_normal(seed * 100 + 1, (B, cfg["NOISE_DIM"]))),
            "emb": t(synth.pseudo_normal(seed * 100 + 2, (B, cfg.get("ENCODER_OUT_DIM", 128)), 0.5)),
            "cond": t(synth.pseudo_normal(seed * 100 + 3, (B, cfg["LATENT_DIM"]))),
            "w_notes": t(synth.pseudo_normal(seed * 100 + 4, (B, cfg["MAX_NOTES"], cfg["NOTE_DIM"]), 1.0 / B)),
            "w_latent": t(synth.pseudo_normal(seed * 100 + 5, (B, cfg["LATENT_DIM"]), 1.0 / B))}


def cond_generator_grads(P, batch):
    """Forward in train mode and the gradients of sum(notes * w_notes) + sum(latent * w_latent)."""
    leaves = _leaves(P)
    emb = batch["emb"].clone().requires_grad_(True)
    bn_state = {k: v.clone() for k, v in P.items() if is_buffer(k)}
    notes, latent = gen_forward(leaves, batch["noise"], emb, True, bn_state, cond=batch["cond"])
    loss = (notes * batch["w_notes"]).sum() + (latent * batch["w_latent"]).sum()
    g = torch.autograd.grad(loss, list(leaves.values()) + [emb])
    return {"notes": notes.detach(), "latent": latent.detach(), "demb": g[-1],
            "grads": dict(zip(leaves.keys(), g[:-1]))}
