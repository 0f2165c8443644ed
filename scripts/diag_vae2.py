import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import gan_oracle as O
from melogan import engine as E
from gan_testlib import rel_err
torch.set_num_threads(8)
P0 = O.make_vae_params(6)
vb = O.make_vae_batch(70, 8)
r64 = O.vae_train_step({k: v.double() for k, v in P0.items()}, {k: v.double() for k, v in vb.items()}, {}, 10.0, update=False)
eng = E.VaeEngine(8, 512, 8, precision="fp32")
x, eps = vb["x"].cuda(), vb["eps"].cuda()
KEYS = ["decoder.deconv.6.weight", "decoder.deconv.4.bias", "decoder.deconv.3.weight", "decoder.pre.2.weight", "fc_mu.weight"]
def run(tag, fn):
    P = {k: v.clone().cuda() for k, v in P0.items()}
    G = {k: torch.zeros_like(P[k]) for k in E.VAE_PARAM_KEYS}
    eng.bind(P, G)
    fn(P, G)
    torch.cuda.synchronize()
    print(tag, " ".join("%s %.1e" % (k.split("decoder.")[-1], rel_err(G[k], r64["grads"][k])) for k in KEYS))
run("loss_step", lambda P, G: eng.loss_step(x, eps, 10.0))
def manual(P, G, second_fwd=False, ext=True):
    recon, z, mu, lv = eng.forward(x, eps, True)
    if second_fwd:
        recon, z, mu, lv = eng.forward(x, eps, True)
    n = recon.numel()
    drecon = (2.0 * (recon - x) / n).contiguous()
    m = mu.numel()
    dmu = 10.0 * mu / m
    dlv = 10.0 * 0.5 * (lv.exp() - 1) / m
    eng.backward(drecon, torch.zeros_like(z) if ext else None, dmu, dlv)
run("manual   ", manual)
run("manual 2f", lambda P, G: manual(P, G, True))
run("manual nz", lambda P, G: manual(P, G, False, False))
def auto(P, G):
    recon, z, mu, lv = eng.forward(x, eps, True)
    rc = recon.clone().requires_grad_(True)
    loss = torch.nn.functional.mse_loss(rc, x)
    loss.backward()
    m = mu.numel()
    print("   drecon autograd vs manual", rel_err(rc.grad, 2.0 * (recon - x) / recon.numel()))
    eng.backward(rc.grad.contiguous(), None, 10.0 * mu / m, 10.0 * 0.5 * (lv.exp() - 1) / m)
run("autograd ", auto)
