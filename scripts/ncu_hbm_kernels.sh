#!/bin/bash
# `ncu --set full` of the HBM-class kernels north_star names: fused Adam, the remaining element-wise / reduction kernels of one
# training cycle at the bench batch, and the two note-extraction kernels; selected metrics exported as CSV on the box.
#   gpurun --timeout 900 -- 'bash scripts/ncu_hbm_kernels.sh r03q'
set -e
TAG=${1:-r03q}
mkdir -p gpurun_out
python scripts/profile_cycle.py --batch 8192 > /dev/null 2>&1                      # plain runs first (exit 0)
python scripts/bench_notes.py > gpurun_out/${TAG}_notes_plain.txt 2>&1
ncu --set full --clock-control none --profile-from-start off \
    -k regex:'adam_kernel|bcast_rows_mul|bn_relu_apply|bn_bwd_apply|colreduce_flat|pad_convert|assemble_critic' --launch-count 40 \
    -o /tmp/${TAG}_elem -f python scripts/profile_cycle.py --batch 8192 > gpurun_out/${TAG}_elem_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'extract_notes' --launch-skip 12 --launch-count 6 \
    -o /tmp/${TAG}_notes -f python scripts/bench_notes.py > gpurun_out/${TAG}_notes_ncu.log 2>&1
for w in elem notes; do
  ncu -i /tmp/${TAG}_${w}.ncu-rep --page raw --csv > /tmp/${TAG}_${w}_full.csv
  python - "$TAG" "$w" <<'PY'
import csv, sys
tag, w = sys.argv[1], sys.argv[2]
keep = ("ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_bytes.sum")
rows = list(csv.reader(open(f"/tmp/{tag}_{w}_full.csv")))
hdr = rows[0]
idx = [i for i, n in enumerate(hdr) if any(n == k or n.endswith("." + k) for k in keep)]
with open(f"gpurun_out/{tag}_{w}_ncu_raw.csv", "w", newline="") as f:
    wr = csv.writer(f)
    for r in rows:
        wr.writerow([r[i] if i < len(r) else "" for i in idx])
PY
done
ls -la gpurun_out/${TAG}_*
