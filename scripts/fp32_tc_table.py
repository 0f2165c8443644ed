#!/usr/bin/env python3
"""Table of gpurun_out/fp32_tc_layers.jsonl (written by tests/test_fp32_tc_gpu.py): per layer, error against float64 and time of
the CUDA-core kernel and of the six-term tensor-core form.    python scripts/fp32_tc_table.py IN.jsonl > profiles/rNN_fp32_tc_layers.txt"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1])]
out = ["fp32 contractions of the cycle at per-GPU batch %d, float32 storage, full-precision random operands, against float64" % rows[0]["B"],
       "(tests/test_fp32_tc_gpu.py on B200; times are whole launches incl. the operand-split pass and weight packing, CUDA events, 3 runs)",
       "",
       f"{'layer':22s} {'CUDA cores':>24s} {'six bf16 terms on tcgen05':>30s} {'speed-up':>9s}",
       f"{'':22s} {'rel err':>12s} {'ms':>11s} {'rel err':>16s} {'ms':>13s}"]
tc = tt = 0.0
steps = []
for r in rows:
    if "step" in r:
        steps.append(f"whole {r['step']}-step at B={r['B']}, switch on vs off: losses {r['losses_tc']} vs {r['losses_cuda_core']}; "
                     f"gradients rel L2 median {r['median_grad_rel_l2']:.2e}, worst {r['worst_grad_rel_l2']:.2e} ({r['worst_grad']})")
        continue
    c, t = r["cuda_core"], r["tc"]
    tc += c["ms"]; tt += t["ms"]
    out.append(f"{r['name']:22s} {c['rel_err']:12.2e} {c['ms']:11.3f} {t['rel_err']:16.2e} {t['ms']:13.3f} {c['ms'] / t['ms']:9.2f}")
out += ["", f"sum over the listed launches: {tc:.2f} ms -> {tt:.2f} ms ({tc / tt:.2f}x)", ""] + steps
print("\n".join(out))
