#!/usr/bin/env python3
"""Runs a few float32 layers of the cycle once with the six-term tensor-core form (mg_debug_set("fp32_tc", 1)) for an ncu
capture of the operand-split pass and of the tcgen05 launches it feeds:

    ncu --set full --clock-control none -k regex:'tc_|split3|pack_weight' -o /tmp/fp32tc python scripts/fp32_tc_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import tc_layers as TL  # noqa: E402

B = int(os.environ.get("MELOGAN_TEST_FP32_TC_BATCH", "1024"))
ONLY = ("ED.conv3.fwd.fp32", "D.conv4.fwd.fp32", "D.conv4.wgrad.fp32")
TL.debug_set("reset", 0)
TL.debug_set("fp32_tc", 1)
for spec in TL.fp32_layers(B):
    if spec["name"] not in ONLY:
        continue
    layer = TL.Layer(spec, seed=3)
    info = layer.run()
    torch.cuda.synchronize()
    print(spec["name"], info)
TL.debug_set("reset", 0)
