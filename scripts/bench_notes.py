"""Config #5: conditional generation (E_num.eval + G.eval) and note extraction N-1 / N-2 on the device,
CUDA-event timing, against the HBM roofline (algorithmic bytes of SURVEY 8d / DESIGN section 4)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200")]
import torch
from melogan import engine as E, notes as N
from oracle import gan_oracle as O          # parameter generator only

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


R = 262144                                              # 2 GiB of rolls, >> 126 MB L2
rolls = torch.rand(R, 512, 4, device="cuda") * 2 - 1
ms = timed(lambda: N.extract_notes_gan(rolls, 140.0, "major", 0, check=False))
out = N.extract_notes_gan(rolls, 140.0, "major", 0, check=False)
emitted = int(out.counts.sum().item())
byt = R * (8192 + 4) + emitted * 18
print(f"N-1 extract_notes_gan: {ms:.3f} ms for {R} rolls = {R / ms * 1e3 / 1e6:.1f} M rolls/s, {byt / ms / 1e6:.0f} GB/s algorithmic "
      f"= {byt / ms / 1e6 / peak:.2f} of HBM peak ({emitted / R:.0f} notes per roll)")
rolls_abs = torch.rand(R, 512, 4, device="cuda") * 100
ms = timed(lambda: N.extract_notes_abs(rolls_abs, check=False))
byt = R * (8192 + 9216)
print(f"N-2 extract_notes_abs: {ms:.3f} ms for {R} rolls = {R / ms * 1e3 / 1e6:.1f} M rolls/s, {byt / ms / 1e6:.0f} GB/s algorithmic "
      f"= {byt / ms / 1e6 / peak:.2f} of HBM peak")
del rolls, rolls_abs, out
B = 8192
params = O.make_params(4, fan_in_scale=True)
for precision in ("fp32", "bf16"):
    eng = E.GanEngine(B, precision=precision)
    Pe = {k: v.cuda() for k, v in params["E"].items()}
    Pg = {k: v.cuda() for k, v in params["G"].items()}
    eng.bind(E.MOD_E, Pe, None)
    eng.bind(E.MOD_G, Pg, None)
    feats = torch.randn(B, 6, device="cuda")
    noise = torch.randn(B, 128, device="cuda")

    def gen():
        emb = eng.encoder_forward(feats, None, None, train=False)
        return eng.generator_forward(noise, emb, train=False)
    ms = timed(gen)
    print(f"generation (E_num.eval + G.eval) {precision}: {ms:.3f} ms for {B} rolls = {B / ms * 1e3 / 1e6:.2f} M rolls/s "
          f"({49.4e6 * B / ms / 1e9:.0f} TFLOP/s of the 49.4 MFLOP/roll)")
    eng.close()
