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
for prec in ("fp32", "bf16"):
    eng, cp, grads = make_engine(B, params, precision=prec)
    cb = cuda_batch(batch)
    ref = O.generator_step(O.clone_params(params), batch, {}, update=False)
    eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"])
    # reference dnotes
    PE, PG, PD, PED = params["E"], params["G"], params["D"], params["ED"]
    notes = ref["notes"].clone().requires_grad_(True)
    emb = O.fe_forward(PE, batch["numeric"], batch["mask1_g"], batch["mask2_g"], train=True).detach()
    la = -O.disc_forward(PD, notes, emb).mean()
    le = 5.0 * F.cross_entropy(O.ed_forward(PED, notes), batch["emot_idx"])
    dn_d, = torch.autograd.grad(la, notes, retain_graph=True)
    dn_e, = torch.autograd.grad(le, notes)
    print(prec, "dnotes total", rel_err(eng.buffer("d.dnotes").view(B, 512, 4), dn_d + dn_e),
          "| scale D part", dn_d.abs().max().item(), "ED part", dn_e.abs().max().item())
    for k, g in ref["grads_G"].items():
        print(f"   {prec} G {k:40s} {rel_err(grads['G'][k], g):.3e}")
    for k, g in ref["grads_E"].items():
        print(f"   {prec} E {k:40s} {rel_err(grads['E'][k], g):.3e}")
