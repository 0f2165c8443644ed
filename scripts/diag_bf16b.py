import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from gan_testlib import cuda_batch, make_engine, rel_err
from oracle import gan_oracle as O
import torch.nn.functional as F

B = 8
params = O.make_params(4, fan_in_scale=True)
batch = O.make_batch(43, B)
ref = O.generator_step(O.clone_params(params), batch, {}, update=False)
PE, PG, PD, PED = params["E"], params["G"], params["D"], params["ED"]
notes = ref["notes"].clone().requires_grad_(True)
emb = O.fe_forward(PE, batch["numeric"], batch["mask1_g"], batch["mask2_g"], train=True).detach()
keepd = {}
s = O.disc_forward(PD, notes, emb, keep=keepd)
la = -s.mean()
keep = {}
logits = O.ed_forward(PED, notes, keep=keep)
le = 5.0 * F.cross_entropy(logits, batch["emot_idx"])
dl, = torch.autograd.grad(le, logits, retain_graph=True)
dn_d, = torch.autograd.grad(la, notes, retain_graph=True)
dn_e, = torch.autograd.grad(le, notes, retain_graph=True)
inter = [keep[f"conv{i}"] for i in range(4)]
g_inter = torch.autograd.grad(le, inter, retain_graph=True)
interd = [keepd[k] for k in ("conv.0", "conv.2", "conv.4")]
g_interd = torch.autograd.grad(la, interd)
for prec in ("fp32", "bf16"):
    eng, cp, grads = make_engine(B, params, precision=prec)
    n = ref["notes"].cuda().contiguous()
    sc = eng.critic_forward(n, emb.cuda())
    seed = torch.full((B,), -1.0 / B, device="cuda")
    dnd, _ = eng.critic_backward(seed, param_grads=False)
    print(prec, "score", rel_err(sc, s), "dnotes_D", rel_err(dnd, dn_d))
    dt = torch.float32 if prec == "fp32" else torch.bfloat16
    for nm, gi, shp in (("d.dz3", g_interd[2], (B, 64, 256)), ("d.dz2", g_interd[1], (B, 128, 128)), ("d.dz1", g_interd[0], (B, 256, 64))):
        # reference grads are w.r.t. post-activation h; ours are w.r.t. pre-activation z -> multiply by mask
        h = {"d.dz3": interd[2], "d.dz2": interd[1], "d.dz1": interd[0]}[nm]
        mask = torch.where(h > 0, torch.ones_like(h), torch.full_like(h, 0.2))
        want = (gi * mask).permute(0, 2, 1)
        got = eng.buffer(nm, dt)[:B * 16384].view(shp).float()
        print("   ", prec, nm, rel_err(got, want))
    lg = eng.emotion_forward(n)
    dne = eng.emotion_backward_input(dl.cuda().contiguous())
    print(prec, "logits", rel_err(lg, logits), "dnotes_ED", rel_err(dne, dn_e))
    for i, C in enumerate((64, 128, 256, 256)):
        print("   ", prec, f"ed.h{i}", rel_err(eng.buffer(f"ed.h{i}", dt).view(B, 512, C).float(), keep[f"conv{i}"].permute(0, 2, 1)))
