import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200")]
import torch
from melogan import notes as N
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
R = 262144
rolls = torch.rand(R, 512, 4, device="cuda") * 2 - 1
ms = timed(lambda: N.extract_notes_gan(rolls, 140.0, "major", 0, check=False))
out = N.extract_notes_gan(rolls, 140.0, "major", 0, check=False)
emitted = int(out.counts.sum().item())
byt = R * (8192 + 4) + emitted * 18
print(f"N-1 [{'3phase' if os.environ.get('MELOGAN_NOTES_3PHASE') else 'pipe'}]: {ms:.3f} ms = {R / ms * 1e3 / 1e6:.1f} M rolls/s, {byt / ms / 1e6:.0f} GB/s = {byt / ms / 1e6 / peak:.3f} of HBM peak ({emitted / R:.0f} notes/roll)")
