#!/usr/bin/env python3
"""Per-layer timing of the tensor-core contractions of the benched cycle (tests/tc_layers.py), each launched alone
through mg_debug_layer_run and timed by the library's per-launch CUDA events.  Prints, per layer: time, achieved
TFLOP/s and algorithmic GB/s, the fraction of the tensor roof and of the HBM roof (MEASURED_PEAKS.json), and the
kernel variant that ran; with --variants also the ablations (no stores / no MMA / no loads) and forced variants.

    python scripts/bench_layers.py [--batch 8192] [--variants] [--only PATTERN] > gpurun_out/layers.txt
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import tc_layers as TL  # noqa: E402

VARIANTS = [("base", {}), ("no pair", {"no_pair": 1}), ("no rotation", {"no_rot": 1}), ("1 staging tile", {"staging_bufs": 1}), ("1 mask tile", {"mask_bufs": 1}), ("BN=64", {"force_bn": 64}), ("BN=128", {"force_bn": 128}),
            ("no stores", {"dbg": 1}), ("no MMA", {"dbg": 2}), ("no loads", {"dbg": 4}), ("no MMA, no loads", {"dbg": 6}),
            ("no mask loads", {"dbg": 8})]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--variants", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default="")
    ap.add_argument("--knob", action="append", default=[], help="k=v tuning knob (mg_debug_set) applied to every run")
    a = ap.parse_args()
    peaks = {"bf16_tflops_sustained": 1379.6, "hbm_gbs": 6546.6}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    except Exception:
        pass
    ptf, pbw = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")       # > L2: every timed launch starts cold
    rows, total = [], 0.0
    for spec in TL.cycle_layers(a.batch):
        if a.only and a.only not in spec["name"]:
            continue
        layer = TL.Layer(spec, seed=3)
        fam = 4 if spec["op"] >= 5 else 3
        for vname, knobs in (VARIANTS if a.variants else VARIANTS[:1]):
            if spec["op"] >= 5 and vname != "base":
                continue
            TL.debug_set("reset", 0)
            for kv_ in a.knob:
                TL.debug_set(kv_.split("=")[0], int(kv_.split("=")[1]))
            for k, v in knobs.items():
                TL.debug_set(k, v)

            def once():
                flush.zero_()
                layer.run()
            ms, launches = TL.probe_time(once, a.iters, fam)
            info = TL.last_launch()
            kv = dict(x.split("=") for x in info.split()[1:] if "=" in x)
            flops, byts = float(kv.get("flops", 0)), float(kv.get("bytes", 0))   # both sub-pixel phases of an up-sampling layer
            tf, gbs = flops / ms / 1e9 if ms else 0, byts / ms / 1e6 if ms else 0
            rec = {"layer": spec["name"], "variant": vname, "per_cycle": spec["count"], "ms": ms, "tflops": tf, "gbs": gbs,
                   "frac_tensor": tf / ptf, "frac_hbm": gbs / pbw, "launches": launches, "info": info}
            rows.append(rec)
            if vname == "base":
                total += ms * spec["count"]
            print(f"{spec['name']:18s} {vname:18s} x{spec['count']} {ms * 1e3:9.1f} us  {tf:7.1f} TF/s ({tf / ptf:5.1%})  "
                  f"{gbs:7.1f} GB/s ({gbs / pbw:5.1%})  bound {max(tf / ptf, gbs / pbw):5.1%} | {info}", flush=True)
        TL.debug_set("reset", 0)
        del layer
        torch.cuda.empty_cache()
    print(f"sum over the cycle (base): {total:.2f} ms")
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
