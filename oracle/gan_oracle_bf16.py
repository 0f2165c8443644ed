"""bf16-storage emulation of the GAN-step oracle (TEST INFRASTRUCTURE ONLY, same rules as gan_oracle.py).

The CUDA path's bf16 mode stores activations (and the tensor-core layers' weights) in bfloat16 and accumulates in
float32.  This module restates gan_oracle's forwards with a straight-through bfloat16 rounding `q` at exactly those
storage points, everything else in float32 on the CPU.  It answers "what does bf16 storage
alone do to the step?" so that the CUDA bf16 path can be held to a TIGHT tolerance against it, while its distance
to the float32 oracle documents the accepted bf16-mode error (north_star: 1e-2 on activations and losses).
"""
import torch
import torch.nn.functional as F

from oracle import gan_oracle as O


def q(t):
    """Round to bfloat16 storage, identity for autograd (also under double backward)."""
    return t + (t.to(torch.bfloat16).to(torch.float32) - t).detach()


def t32(t):
    """What tcgen05 kind::tf32 reads of a float32 operand: the low 13 mantissa bits are ignored.  Identity for autograd."""
    tt = (t.detach().contiguous().view(torch.int32) & -8192).view(torch.float32)
    return t + (tt - t).detach()


def lin(x, w, b):
    """The float32 Linears of bf16 mode run on the tensor cores as TF32 when they are tile-shaped (gemm_tc.cuh
    try_tc_tapgemm: batch >= 128, K % 32 == 0, N % 64 == 0); otherwise on the CUDA cores in full float32."""
    if x.shape[0] >= 128 and w.shape[1] % 32 == 0 and w.shape[0] % 64 == 0:
        return F.linear(t32(x), t32(w), b)
    return F.linear(x, w, b)


def fe_forward(P, x, mask1=None, mask2=None, train=True, p_drop=0.2):
    h = F.layer_norm(x, (x.shape[1],), P["net.0.weight"], P["net.0.bias"], 1e-5)
    h = F.gelu(lin(h, P["net.1.weight"], P["net.1.bias"]))
    if train:
        h = h * (mask1 * (1.0 / (1.0 - p_drop)))
    h = F.gelu(lin(h, P["net.4.weight"], P["net.4.bias"]))
    if train:
        h = h * (mask2 * (1.0 / (1.0 - p_drop)))
    return lin(h, P["net.7.weight"], P["net.7.bias"])


def gen_forward(P, noise, emb, bn_state):
    x = torch.cat([noise, emb], 1)
    h = F.relu(lin(x, P["noise_to_latent.net.0.weight"], P["noise_to_latent.net.0.bias"]))
    latent = lin(h, P["noise_to_latent.net.2.weight"], P["noise_to_latent.net.2.bias"])
    y = q(F.relu(lin(latent, P["decoder.pre.0.weight"], P["decoder.pre.0.bias"])))
    y = q(F.relu(F.linear(y, q(P["decoder.pre.2.weight"]), P["decoder.pre.2.bias"])))
    y = y.view(y.shape[0], 256, -1)
    for conv, bn in (("decoder.deconv.0", "decoder.deconv.1"), ("decoder.deconv.3", "decoder.deconv.4")):
        y = F.conv_transpose1d(y, q(P[conv + ".weight"]), P[conv + ".bias"], stride=2, padding=2, output_padding=1)
        y = F.batch_norm(y, bn_state[bn + ".running_mean"], bn_state[bn + ".running_var"], P[bn + ".weight"],
                         P[bn + ".bias"], training=True, momentum=0.1, eps=1e-5)     # pre-BN output stays float32
        y = q(F.relu(y))
    y = F.conv_transpose1d(y, q(P["decoder.deconv.6.weight"]), P["decoder.deconv.6.bias"], stride=2, padding=2,
                           output_padding=1)                                          # banded tensor-core form: bf16 weights
    return y.permute(0, 2, 1), latent


def disc_forward(P, notes, emb):
    h = q(notes).permute(0, 2, 1)                 # conv.0 reads a zero-padded bf16 copy of the notes (banded.cuh)
    for i, name in enumerate(("conv.0", "conv.2", "conv.4")):
        h = q(F.leaky_relu(F.conv1d(h, q(P[name + ".weight"]), P[name + ".bias"], stride=2, padding=2), 0.2))
    h = F.adaptive_avg_pool1d(h, 1)
    feat = F.leaky_relu(lin(h.view(h.size(0), -1), P["fc.1.weight"], P["fc.1.bias"]), 0.2)
    feat = torch.cat([feat, emb], 1)
    return F.linear(feat, P["real_fake.weight"], P["real_fake.bias"]).squeeze(1)


def ed_forward(P, notes):
    x = q(notes).permute(0, 2, 1)
    for i in range(4):
        pre = f"encoder.conv.{i}.net."
        x = F.conv1d(x, q(P[pre + "0.weight"]), P[pre + "0.bias"], stride=1, padding=2 if i == 0 else 1)
        x = F.batch_norm(x, P[pre + "1.running_mean"], P[pre + "1.running_var"], P[pre + "1.weight"], P[pre + "1.bias"],
                         training=False, eps=1e-5)
        x = q(F.gelu(x))
    x = F.adaptive_avg_pool1d(x, 1).squeeze(-1)
    x = lin(x, P["encoder.project.weight"], P["encoder.project.bias"])
    x = F.gelu(lin(x, P["classifier.net.0.weight"], P["classifier.net.0.bias"]))
    x = F.gelu(lin(x, P["classifier.net.3.weight"], P["classifier.net.3.bias"]))
    return F.linear(x, P["classifier.head.weight"], P["classifier.head.bias"])


def generator_step(params, batch, cfg=O.CFG):
    PE, PG, PD, PED = params["E"], params["G"], params["D"], params["ED"]
    El, Gl = O._leaves(PE), O._leaves(PG)
    bn_state = {k: v.clone() for k, v in PG.items() if O.is_buffer(k)}
    emb = fe_forward(El, batch["numeric"], batch["mask1_g"], batch["mask2_g"], train=True, p_drop=cfg["ENC_DROPOUT"])
    notes, latent = gen_forward(Gl, batch["noise_g"], emb, bn_state)
    loss_adv = -disc_forward(PD, notes, emb).mean()
    logits = ed_forward(PED, notes)
    loss_emo = F.cross_entropy(logits, batch["emot_idx"])
    g = torch.autograd.grad(loss_adv + cfg["LAMBDA_EMOTION"] * loss_emo, list(Gl.values()) + list(El.values()))
    return {"loss_g_adv": loss_adv.detach(), "loss_g_emo": loss_emo.detach(), "notes": notes.detach(),
            "grads_G": dict(zip(Gl.keys(), g[:len(Gl)])), "grads_E": dict(zip(El.keys(), g[len(Gl):]))}


def critic_step(params, batch, cfg=O.CFG):
    PE, PG, PD = params["E"], params["G"], params["D"]
    with torch.no_grad():
        emb = fe_forward(PE, batch["numeric"], batch["mask1_d"], batch["mask2_d"], train=True, p_drop=cfg["ENC_DROPOUT"])
        fake, _ = gen_forward(PG, batch["noise_d"], emb, {k: v.clone() for k, v in PG.items() if O.is_buffer(k)})
        fake = fake.contiguous()
    Dl = O._leaves(PD)
    real = batch["notes_real"]
    a = batch["alpha"].view(-1, 1, 1).expand_as(real)
    interp = (a * real + (1 - a) * fake).requires_grad_(True)
    gr = torch.autograd.grad(disc_forward(Dl, interp, emb).sum(), interp, create_graph=True)[0].reshape(real.size(0), -1)
    gp = ((gr.norm(2, dim=1) - 1) ** 2).mean()
    loss = disc_forward(Dl, fake, emb).mean() - disc_forward(Dl, real, emb).mean() + cfg["LAMBDA_GP"] * gp
    return {"loss_d": loss.detach(), "gp": gp.detach(), "fake": fake,
            "grads": dict(zip(Dl.keys(), torch.autograd.grad(loss, list(Dl.values()))))}
