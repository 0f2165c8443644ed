#!/bin/bash
# One `ncu --set full` capture of every tensor-core contraction of the benched cycle (scripts/bench_layers.py --iters 1: each
# layer launches twice, the second launch is kept), exported as raw CSV on the box.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash scripts/ncu_layers.sh r02s'
# then locally: python scripts/ncu_traffic.py gpurun_out/r02s_layers.json gpurun_out/r02s_ncu_raw.csv
set -e
TAG=${1:-r02s}
mkdir -p gpurun_out
python scripts/bench_layers.py --iters 1 --json gpurun_out/${TAG}_layers.json > gpurun_out/${TAG}_layers.txt 2>&1   # plain run first (exit 0)
ncu --set full --clock-control none -k regex:tc_ -o /tmp/${TAG}_ncu -f \
    python scripts/bench_layers.py --iters 1 --json /tmp/${TAG}_layers_under_ncu.json > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_ncu.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw_full.csv
# keep the columns the summary needs (the full raw page is ~2300 columns per launch)
python - "$TAG" <<'PY'
import csv, sys
tag = sys.argv[1]
keep = ("ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__inst_executed.sum")
rows = list(csv.reader(open(f"gpurun_out/{tag}_ncu_raw_full.csv")))
hdr = rows[0]
idx = [i for i, n in enumerate(hdr) if any(n == k or n.endswith("." + k) for k in keep)]
with open(f"gpurun_out/{tag}_ncu_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] if i < len(r) else "" for i in idx])
PY
rm -f gpurun_out/${TAG}_ncu_raw_full.csv
ls -la gpurun_out/${TAG}_*
