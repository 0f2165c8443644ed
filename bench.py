#!/usr/bin/env python3
"""bench.py -- GAN train-step throughput of the B200-native Melo-GAN hot path.

    python bench.py --gpus 1 --steps 10 --warmup 3            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                        (the reference's CPU path, oracle port)

One "step" is one training CYCLE of the reference loop (src/gan/train_gan.py:168-251 with
CRITIC_ITERS=5): 5 critic steps on 5 fresh batches of B real rolls each + 1 generator step, Adam
included, i.e. 5*B real rolls consumed per rank and step.  Metric: real rolls per second, whole job.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "melo-gan_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "GAN train-step real rolls/sec (cycle = 5 critic steps + 1 generator step, Adam included)"
MFLOP_PER_ROLL = 621.9          # algorithmic FLOPs of the reference's cycle per real roll


def load_cfgs():
    import yaml
    with open(os.path.join(PKG, "config", "gan_config.yaml")) as f:
        cfg = yaml.safe_load(f)
    with open(os.path.join(PKG, "config", "ed_config.yaml")) as f:
        ed_cfg = yaml.safe_load(f)
    return cfg, ed_cfg


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu_index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_cycle_baseline(budget_s=18.0, B=32):
    """The reference's CPU training cycle (oracle port: same ATen ops, same order, see oracle/gan_oracle.py)
    on all host cores, bounded to about `budget_s` seconds."""
    import torch
    from oracle import gan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = O.make_params(1)
    batches = [O.make_batch(900 + i, B) for i in range(5)]
    st_d, st_g = {}, {}
    O.train_cycle(params, batches, st_d, st_g)          # warm-up cycle
    n, t0 = 0, time.perf_counter()
    while True:
        O.train_cycle(params, batches, st_d, st_g)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 50:
            break
    return {"value": 5 * B * n / el, "unit": "rolls/s", "cores": cores, "kind": "port",
            "sample": f"{n} cycles of 5 D-steps + 1 G-step at B={B} (config/gan_config.yaml BATCH_SIZE), "
                      f"torch {torch.__version__} CPU, {cores} threads, {el:.1f} s", "ms_per_cycle": 1e3 * el / n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 32
    import torch
    from oracle import gan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = O.make_params(1)
    batches = [O.make_batch(900 + i, B) for i in range(5)]
    st_d, st_g = {}, {}
    for _ in range(max(1, min(args.warmup, 2))):
        O.train_cycle(params, batches, st_d, st_g)
    steps = max(1, min(args.steps, 12))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_cycle(params, batches, st_d, st_g)
    el = time.perf_counter() - t0
    v = 5 * B * steps / el
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rolls/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GAN cycle (5 D + 1 G), config/gan_config.yaml, B={B}, reference CPU path (oracle port)"},
            "cpu_baseline": {"value": v, "unit": "rolls/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} cycles at B={B}, {cores} threads"},
            "e2e": {"value": v, "unit": "rolls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MELOGAN_BENCH_BATCH", "8192")),
                    help="per-GPU batch B of every critic/generator step")
    ap.add_argument("--precision", default=os.environ.get("MELOGAN_PRECISION", "bf16"), choices=["fp32", "bf16"],
                    help="bf16 = tcgen05 tensor-core mode (north_star tolerance 1e-2); fp32 = CUDA-core parity mode (1e-5)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from melogan import _native, synth
    from melogan.trainer import GanTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    W = max(3, args.warmup)
    cfg, ed_cfg = load_cfgs()
    B, K = args.batch, int(cfg.get("CRITIC_ITERS", 5))
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):     # the drop-in modules print like the reference's; stdout is ONE JSON line
        tr = GanTrainer(cfg, ed_cfg, batch=B, precision=args.precision, device=dev, process_group=pg, seed_offset=rank)

    # synthetic inputs: several resident cycles so consecutive steps read different data
    NSETS = 3
    T, F = cfg["MAX_NOTES"], cfg.get("NUMERIC_INPUT_DIM", 6)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    reals = [torch.rand((K, B, T, 4), generator=g, device=dev) * 2 - 1 for _ in range(NSETS)]
    numerics = []
    for _ in range(NSETS):
        x = torch.randn((K, B, F), generator=g, device=dev)
        x[..., 5] = 0.0
        numerics.append(x)
    labels = (torch.arange(B, device=dev) % 4).to(torch.int64)
    # pinned host copies for the end-to-end arm
    h_reals = [r.cpu().pin_memory() for r in reals]
    h_numerics = [x.cpu().pin_memory() for x in numerics]
    h_labels = labels.cpu().pin_memory()
    h_metrics = torch.empty(8, dtype=torch.float32).pin_memory()

    L = _native.lib()
    # eager warm-up (also allocates lazily created scratch), then count this library's launches per cycle
    tr.train_cycle(reals[0], numerics[0], labels)
    torch.cuda.synchronize(dev)
    n0 = L.mg_launch_count()
    tr.train_cycle(reals[1], numerics[1], labels)
    torch.cuda.synchronize(dev)
    launches_per_cycle = int(L.mg_launch_count() - n0)

    use_graph = not args.no_graph
    if use_graph:
        try:
            s_reals, s_numerics, s_labels = tr.capture_cycle()
            s_labels.copy_(labels)
        except Exception as e:   # capture of NCCL or anything else refused: fall back to eager launches, and say so
            use_graph = False
            if rank == 0:
                print(f"[bench] CUDA-graph capture unavailable ({type(e).__name__}: {e}); running eager", file=sys.stderr)

    def device_step(i):
        if use_graph:
            s_reals.copy_(reals[i % NSETS]); s_numerics.copy_(numerics[i % NSETS])
            tr.replay_cycle()
        else:
            tr.train_cycle(reals[i % NSETS], numerics[i % NSETS], labels)

    def e2e_step(i):
        if use_graph:
            s_reals.copy_(h_reals[i % NSETS], non_blocking=True); s_numerics.copy_(h_numerics[i % NSETS], non_blocking=True)
            s_labels.copy_(h_labels, non_blocking=True)
            tr.replay_cycle()
        else:
            r = h_reals[i % NSETS].to(dev, non_blocking=True); x = h_numerics[i % NSETS].to(dev, non_blocking=True)
            lb = h_labels.to(dev, non_blocking=True)
            tr.train_cycle(r, x, lb)
        h_metrics.copy_(tr.loss_acc, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()      # the host reads the losses of this step
        return float(h_metrics[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        for i in range(W):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(W + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(device_step, args.steps)
    if sampler:
        sampler.stop_flag.set(); sampler.join(timeout=3)
    ms_e2e = timed(e2e_step, args.steps)

    # per-kernel roofline of the dominant kernel family, measured live with CUDA events (eager pass)
    family = 3 if args.precision == "bf16" else 1
    L.mg_probe_begin(family)
    tr.train_cycle(reals[0], numerics[0], labels)
    import ctypes
    pr = (ctypes.c_double * 4)()
    L.mg_probe_end(pr)
    probe_launches, probe_ms, probe_flops, probe_bytes = pr[0], pr[1], pr[2], pr[3]
    if family == 3 and probe_launches == 0:      # bf16 mode without tensor-core kernels yet
        family = 1
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json, bf16 sustained)" if peaks else "fallback (B200_PROFILING.md, 1.4 PF sustained)"
    achieved_tf = (probe_flops / (probe_ms * 1e-3) / 1e12) if probe_ms > 0 else 0.0
    # the same launches against the HBM roofline: algorithmic bytes (activation once + output + mask/derivative tiles)
    peak_gbs = float(peaks.get("hbm_gbs", 6546.6))
    achieved_gbs = (probe_bytes / (probe_ms * 1e-3) / 1e9) if probe_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_tc_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch_avg")
    except Exception:
        pass

    rolls_per_step = K * B * world
    value = rolls_per_step / (ms_dev / args.steps * 1e-3)
    e2e_value = rolls_per_step / (ms_e2e / args.steps * 1e-3)
    h2d = K * B * T * 4 * 4 + K * B * F * 4 + B * 8
    line = {
        "metric": METRIC, "value": value, "unit": "rolls/s", "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"GAN cycle (5 critic steps + 1 generator step + Adam), config/gan_config.yaml shapes "
                               f"(512x4 rolls, noise 128, latent 64, 4 emotion classes), per-GPU batch B={B}",
                   "per_gpu_batch": B, "rolls_per_step_per_gpu": K * B, "precision": args.precision,
                   "cuda_graph": use_graph, "parallelism": f"dp{world}" if world > 1 else "single",
                   "bn": "local" if world > 1 else "n/a",
                   "l2": f"{NSETS} rotating input sets; per-step activation working set {tr.engine.workspace_bytes() / 1e6:.0f} MB >> 126 MB L2",
                   "algorithmic_mflop_per_roll": MFLOP_PER_ROLL},
        "e2e": {"value": e2e_value, "unit": "rolls/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_cycle * args.steps,
        "clocks": sampler.summary() if sampler else None,
        "roofline": {"bound": "tensor", "kernel": {1: "tapgemm_kernel (CUDA-core fp32 implicit GEMM)",
                                                   3: "tc_gemm_kernel (tcgen05 bf16 implicit GEMM)"}[family],
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "peak_source": peak_src,
                     "traffic_note": "dram read+write bytes per launch, mean over the ncu --set full capture in profiles/ "
                                     "(r01_tc_traffic.json); algorithmic bytes per launch = "
                                     f"{probe_bytes / max(probe_launches, 1):.3e}",
                     "launches_per_step": probe_launches, "kernel_ms_per_step": probe_ms,
                     "share_of_step": probe_ms / (ms_dev / args.steps) if ms_dev > 0 else None,
                     "hbm_view": {"achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s",
                                  "frac": achieved_gbs / peak_gbs if peak_gbs else None}},
        "whole_step_model_tflops": value * MFLOP_PER_ROLL * 1e6 / 1e12 / max(world, 1),
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_cycle_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
