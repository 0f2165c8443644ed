import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import gan_oracle as O
from melogan import engine as E
from gan_testlib import rel_err
from src.ae.model import VAE
from src.ae.train_ae import vae_loss
torch.set_num_threads(8)
model = VAE({"LATENT_DIM": 8, "MAX_NOTES": 512}).cuda()
model.encoder.build_linear(512)
model.load_state_dict(O.make_vae_params(6), strict=False)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
P = O.make_vae_params(6); st = {}
P64 = {k: v.double() for k, v in P.items()}
real = torch.randn_like
for i in range(2):
    vb = O.make_vae_batch(70 + i, 8)
    r64 = O.vae_train_step({k: v.double() for k, v in P.items()}, {k: v.double() for k, v in vb.items()}, {}, 10.0, update=False)
    r = O.vae_train_step(P, vb, st, 10.0)
    torch.randn_like = lambda t, _e=vb["eps"]: _e.clone().to(t.device)
    recon, z, mu, lv = model(vb["x"].cuda())
    torch.randn_like = real
    loss, rl, kl = vae_loss(recon, vb["x"].cuda(), mu, lv, 10.0)
    opt.zero_grad(); loss.backward()
    print("step", i, "recon", rel_err(recon, r["recon"]), "oracle32 vs 64", rel_err(r["recon"], r64["recon"]), "mu", rel_err(mu, r["mu"]))
    for k, p in model.named_parameters():
        print("  grad %-28s got-vs-64 %.2e  oracle32-vs-64 %.2e   max %.2e" % (k, rel_err(p.grad, r64["grads"][k]), rel_err(r["grads"][k], r64["grads"][k]), r64["grads"][k].abs().max()))
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0); opt.step()
    sd = model.state_dict()
    for k in E.VAE_PARAM_KEYS:
        d = (sd[k].cpu() - P[k]).abs().max().item()
        if d > 2e-5: print("  param %-28s maxabs diff %.2e" % (k, d))
