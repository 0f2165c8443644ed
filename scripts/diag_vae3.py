import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
from oracle import gan_oracle as O
from melogan import engine as E
from gan_testlib import rel_err
torch.set_num_threads(8)
B = 8
P0 = O.make_vae_params(6)
vb = O.make_vae_batch(70, B)
P = {k: v.double().requires_grad_(not O.is_buffer(k)) for k, v in P0.items()}
x, eps = vb["x"].double(), vb["eps"].double()
I = {}
def keep(name, t):
    t.retain_grad(); I[name] = t; return t
h = x.permute(0, 2, 1)
for i, (c, b) in enumerate(((0, 1), (3, 4), (6, 7))):
    h = keep(f"e_x{i}", F.conv1d(h, P[f"encoder.conv.{c}.weight"], P[f"encoder.conv.{c}.bias"], stride=2, padding=2))
    h = keep(f"e_bn{i}", F.batch_norm(h, None, None, P[f"encoder.conv.{b}.weight"], P[f"encoder.conv.{b}.bias"], training=True))
    h = keep(f"e_a{i}", F.relu(h))
h = keep("h", F.relu(F.linear(h.reshape(B, -1), P["encoder._linear.1.weight"], P["encoder._linear.1.bias"])))
mu = keep("mu", F.linear(h, P["fc_mu.weight"], P["fc_mu.bias"])); lv = keep("lv", F.linear(h, P["fc_log_var.weight"], P["fc_log_var.bias"]))
z = keep("z", mu + eps * torch.exp(0.5 * lv))
y = keep("d0", F.relu(F.linear(z, P["decoder.pre.0.weight"], P["decoder.pre.0.bias"])))
y = keep("d_y0", F.relu(F.linear(y, P["decoder.pre.2.weight"], P["decoder.pre.2.bias"])).view(B, 128, -1))
for i, (c, b) in enumerate(((0, 1), (3, 4))):
    y = keep(f"d_x{i}", F.conv_transpose1d(y, P[f"decoder.deconv.{c}.weight"], P[f"decoder.deconv.{c}.bias"], stride=2, padding=2, output_padding=1))
    y = keep(f"d_bn{i}", F.batch_norm(y, None, None, P[f"decoder.deconv.{b}.weight"], P[f"decoder.deconv.{b}.bias"], training=True))
    y = keep(f"d_y{i+1}", F.relu(y))
pre = keep("pre_t", F.conv_transpose1d(y, P["decoder.deconv.6.weight"], P["decoder.deconv.6.bias"], stride=2, padding=2, output_padding=1))
recon = torch.tanh(pre).permute(0, 2, 1)
loss = F.mse_loss(recon, x) + 10.0 * (-0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp()))
loss.backward()
cl = lambda t: t.permute(0, 2, 1).contiguous().flatten()       # (B, C, L) -> channels-last flat
eng = E.VaeEngine(B, 512, 8, precision="fp32")
Pc = {k: v.clone().cuda() for k, v in P0.items()}
G = {k: torch.zeros_like(Pc[k]) for k in E.VAE_PARAM_KEYS}
eng.bind(Pc, G)
eng.loss_step(vb["x"].cuda(), vb["eps"].cuda(), 10.0)
def show(name, want):
    got = eng.buffer(name)
    print("%-8s err %.2e   max %.2e" % (name, rel_err(got, want), want.abs().max()))
show("d_x0", cl(I["d_x0"])); show("d_y1", cl(I["d_y1"])); show("d_x1", cl(I["d_x1"])); show("d_y2", cl(I["d_y2"]))
show("dt", cl(I["pre_t"].grad))
show("dy_f2", cl(I["d_bn1"].grad)); show("dxd2", cl(I["d_x1"].grad))
show("dy_f1", cl(I["d_bn0"].grad)); show("dxd1", cl(I["d_x0"].grad))
show("dy0", cl(I["d_y0"].grad * (I["d_y0"] > 0)))
show("dd0", I["d0"].grad.flatten() * (I["d0"].flatten() > 0)); show("dz", None if False else (I["z"].grad.flatten()))
show("dh", I["h"].grad.flatten() * (I["h"].flatten() > 0))
show("de_f2", cl(I["e_bn2"].grad)); show("dxe2", cl(I["e_x2"].grad)); show("de_f1", cl(I["e_bn1"].grad)); show("dxe1", cl(I["e_x1"].grad))
show("de_f0", cl(I["e_bn0"].grad)); show("dxe0", cl(I["e_x0"].grad))
# where is the dy_f2 error
got = eng.buffer("dy_f2").double().cpu().view(B, 256, 32); want = I["d_bn1"].grad.permute(0, 2, 1)
d = (got - want).abs()
print("dy_f2 err by position (first/last 4):", d.amax(dim=(0, 2))[:4].tolist(), d.amax(dim=(0, 2))[-4:].tolist())
print("dy_f2 err mean", d.mean().item(), "sum diff per channel (first 4)", (got - want).sum(dim=(0, 1))[:4].tolist(), "want sum", want.sum(dim=(0, 1))[:4].tolist())
flat = d.flatten(); top = flat.topk(8).indices
bn = I["d_bn1"].permute(0, 2, 1).flatten(); gy = eng.buffer("d_y2").double().cpu()
gpost = I["d_y2"].grad.permute(0, 2, 1).flatten()
for i in top.tolist():
    print(i, "b,pos,ch", i // (256 * 32), (i // 32) % 256, i % 32, "err %.2e got %.3e want %.3e  bn_out %.3e  gpu_y %.3e  grad_post %.3e" % (flat[i], got.flatten()[i], want.flatten()[i], bn[i], gy[i], gpost[i]))
print("n elements with err > 1e-8:", (flat > 1e-8).sum().item(), " > 1e-6:", (flat > 1e-6).sum().item())
