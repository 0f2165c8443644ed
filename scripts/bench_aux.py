"""Device-side timing of the two auxiliary training steps (BASELINE configs #2 and #3; parity-test configurations, not bench
lines): VAE forward + loss + backward (mg_vae_loss_step) and emotion-discriminator train forward + cross-entropy +
backward (mg_emotion_train_*), CUDA events, synthetic inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200")]
import torch
from melogan import engine as E
from oracle import gan_oracle as O          # parameter / batch generators only (test infrastructure)


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for precision in ("fp32", "bf16"):
    for B in (32, 1024):
        P0 = O.make_vae_params(6)
        eng = E.VaeEngine(B, 512, 8, precision=precision)
        P = {k: v.cuda() for k, v in P0.items()}
        G = {k: torch.zeros_like(P[k]) for k in E.VAE_PARAM_KEYS}
        eng.bind(P, G)
        x = torch.rand(B, 512, 4, device="cuda") * 2 - 1
        eps = torch.randn(B, 8, device="cuda")
        ms = timed(lambda: eng.loss_step(x, eps, 10.0))
        print(f"VAE  step {precision} B={B:5d}: {ms:7.3f} ms  {B / ms * 1e3:10.0f} rolls/s")
        eng.close()
        params = O.make_params(5)
        g = E.GanEngine(B, precision=precision)
        Pe = {k: v.cuda() for k, v in params["ED"].items()}
        Ge = {k: torch.zeros_like(Pe[k]) for k in E.ED_GRAD_KEYS}
        g.bind(E.MOD_ED, Pe, Ge)
        m1 = (torch.rand(B, 256, device="cuda") > 0.2).float()
        m2 = (torch.rand(B, 128, device="cuda") > 0.2).float()
        dl = torch.randn(B, 4, device="cuda") / B

        def ed_step():
            g.emotion_train_forward(x, m1, m2, 0.2)
            g.emotion_train_backward(dl)
        ms = timed(ed_step)
        print(f"ED   step {precision} B={B:5d}: {ms:7.3f} ms  {B / ms * 1e3:10.0f} rolls/s")
        g.close()
