import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from gan_testlib import cuda_batch, make_engine, rel_l2
from melogan import _native
from oracle import gan_oracle as O
B, fan, pseed = 32, False, 2
params = O.make_params(pseed, fan_in_scale=fan)
batch = O.make_batch(10 * pseed + 3, B)
ref = O.generator_step(O.clone_params(params), batch, {}, update=False)
res = {}
for name, prec, tc in (("fp32", "fp32", 1), ("bf16-simt", "bf16", 0), ("bf16-tc", "bf16", 1)):
    _native.lib().mg_tc_enable(tc)
    eng, cp, grads = make_engine(B, params, precision=prec)
    cb = cuda_batch(batch)
    eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"])
    res[name] = {k: rel_l2(grads["G"][k], g) for k, g in ref["grads_G"].items()}
for k in ref["grads_G"]:
    print(f"{k:40s}" + "  ".join(f"{n}={res[n][k]:.2e}" for n in res))
