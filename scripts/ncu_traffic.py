#!/usr/bin/env python3
"""Joins the per-layer list of scripts/bench_layers.py (--iters 1 --json) with the raw CSV of the `ncu --set full` capture of
the same command (scripts/ncu_layers.sh) by launch order, and writes

  profiles/<tag>_ncu_layers.txt   per layer: ncu duration, DRAM read+write bytes next to the algorithmic bytes of that launch,
                                   DRAM bytes / duration as % of the measured HBM peak, tensor-pipe active %, issue active %, warps active %
  profiles/r02_tc_traffic.json     cycle-weighted DRAM bytes per tap-GEMM launch next to the algorithmic bytes of the SAME
                                   launches (bench.py reads it for roofline.traffic)

    python scripts/ncu_traffic.py gpurun_out/r02s_layers.json gpurun_out/r02s_ncu_raw.csv [tag]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    layers = [r for r in json.load(open(sys.argv[1])) if r["variant"] == "base"]
    tag = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(sys.argv[2]).split("_")[0]
    rows = list(csv.reader(open(sys.argv[2])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, n in enumerate(hdr):                 # plain metric columns win over the section-prefixed copies
        if ".TriageCompute." not in n:
            col.setdefault(n, i)
    for i, n in enumerate(hdr):
        if ".TriageCompute." in n:
            col.setdefault(n.split(".TriageCompute.")[-1], i)

    def val(r, name, scale=True):
        i = col.get(name)
        if i is None or r[i] == "":
            return float("nan")
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        return v * UNIT.get(units[i], 1.0) if scale else v

    pos, out, wsum_d, wsum_a, wl = 0, [], 0.0, 0.0, 0.0
    lines = [f"# ncu --set full --clock-control none of `python scripts/bench_layers.py --iters 1` ({tag}); per layer the SECOND launch "
             f"(first = cold instruction cache / tensor maps).  dram = dram__bytes_read.sum + dram__bytes_write.sum, alg = algorithmic",
             f"# bytes of the same launch(es) (activation once + output (+ derivative tile) + mask tile); tensor = "
             f"sm__pipe_tensor_cycles_active (% of elapsed), issue = smsp__issue_active, warps = sm__warps_active",
             f"{'layer':18s} {'x/cycle':>7s} {'ncu us':>9s} {'dram MB':>9s} {'alg MB':>9s} {'dram/alg':>8s} {'dram %':>7s} {'tensor %':>8s} "
             f"{'issue %':>7s} {'warps %':>7s} {'regs':>5s} {'cluster':>7s}  kernel"]
    for L in layers:
        n = int(round(L["launches"]))
        if n < 1:
            continue
        chunk = data[pos:pos + 2 * n]
        pos += 2 * n
        if len(chunk) < 2 * n:
            break
        kept = chunk[n:]                                   # the second run of this layer
        us = sum(val(r, "gpu__time_duration.sum") for r in kept)
        dram = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in kept)
        kv = dict(x.split("=") for x in L["info"].split()[1:] if "=" in x)
        alg = float(kv.get("bytes", 0))
        pick = lambda name: sum(val(r, name, False) for r in kept) / len(kept)
        tensor = pick("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
        rec = {"layer": L["layer"], "per_cycle": L["per_cycle"], "launches": n, "ncu_us": us, "dram_bytes": dram, "alg_bytes": alg,
               "dram_pct": dram / max(us, 1e-9) / 1e3 / 6546.6 * 100.0, "tensor_pct": tensor,   # of the measured 6546.6 GB/s
               "issue_pct": pick("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_pct": pick("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "kernel": kept[0][col["Kernel Name"]].split("(")[0][:70]}
        out.append(rec)
        is_tap = "wgrad" not in L["layer"]
        if is_tap:
            wsum_d += dram * L["per_cycle"]; wsum_a += alg * L["per_cycle"]; wl += n * L["per_cycle"]
        regs = kept[0][col["launch__registers_per_thread"]] if "launch__registers_per_thread" in col else ""
        cl = kept[0][col["launch__cluster_size"]] if "launch__cluster_size" in col else ""
        lines.append(f"{L['layer']:18s} {L['per_cycle']:7d} {us:9.1f} {dram / 1e6:9.1f} {alg / 1e6:9.1f} {dram / max(alg, 1):8.2f} "
                     f"{rec['dram_pct']:7.1f} {tensor:8.1f} {rec['issue_pct']:7.1f} {rec['warps_pct']:7.1f} {regs:>5s} {cl:>7s}  {rec['kernel']}")
    tj = {"source": f"profiles/{tag}_ncu_layers.txt", "launches_per_cycle": wl,
          "dram_bytes_per_launch_cycle_weighted": wsum_d / wl, "algorithmic_bytes_per_launch_cycle_weighted": wsum_a / wl,
          "ratio": wsum_d / wsum_a, "layers": out}
    lines.append(f"# tap-GEMM launches, cycle-weighted: dram {wsum_d / 1e9:.2f} GB vs algorithmic {wsum_a / 1e9:.2f} GB per cycle "
                 f"(ratio {wsum_d / wsum_a:.3f}) over {wl:.0f} launches")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_layers.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(tj, open(os.path.join(ROOT, "profiles", "r02_tc_traffic.json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
