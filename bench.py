#!/usr/bin/env python3
"""bench.py -- GAN train-step throughput of the B200-native Melo-GAN hot path.

    python bench.py --gpus 1 --steps 10 --warmup 3            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                        (the reference's CPU path: unmodified modules under
                                                                 baseline/_ref, else the oracle port)

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


def _ref_cycle():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_cycle
    return ref_cycle


def cpu_reference_cycles(budget_s, max_cycles, B=32, warmup=1):
    """The reference's CPU training cycle on all host cores, bounded to about `budget_s` seconds.
    kind "reference": the UNMODIFIED reference modules under baseline/_ref (staged by __graft_entry__.build()) driven
    by the loop body of src/gan/train_gan.py:183-251 (baseline/ref_cycle.py); kind "port": the oracle's restatement
    of the same ATen calls (oracle/gan_oracle.py) when the reference sources did not travel."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rcm = _ref_cycle()
    if rcm.available():
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            rc = rcm.ReferenceCycle("cpu")
        batches, labels = rc.batches(B)
        step = lambda: rc.cycle(batches, labels)
        kind = "reference"
    else:
        from oracle import gan_oracle as O
        params = O.make_params(1)
        batches = [O.make_batch(900 + i, B) for i in range(5)]
        st_d, st_g = {}, {}
        step = lambda: O.train_cycle(params, batches, st_d, st_g)
        kind = "port"
    for _ in range(max(1, warmup)):
        step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= max_cycles:
            break
    return {"value": 5 * B * n / el, "unit": "rolls/s", "cores": cores, "kind": kind,
            "sample": f"{n} cycles of 5 D-steps + 1 G-step at B={B} (config/gan_config.yaml BATCH_SIZE), "
                      f"torch {torch.__version__} CPU, {cores} threads, {el:.1f} s", "ms_per_cycle": 1e3 * el / n}, n, el


def cpu_cycle_baseline(budget_s=18.0, B=32):
    return cpu_reference_cycles(budget_s, 50, B)[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 32
    cb, steps, el = cpu_reference_cycles(60.0, max(1, min(args.steps, 40)), B, warmup=max(1, min(args.warmup, 3)))
    v = cb["value"]
    what = "unmodified reference modules, baseline/_ref" if cb["kind"] == "reference" else "oracle port"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rolls/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GAN cycle (5 D + 1 G), config/gan_config.yaml, B={B}, reference CPU path ({what})"},
            "cpu_baseline": {"value": v, "unit": "rolls/s", "cores": cb["cores"], "kind": cb["kind"],
                             "sample": f"{steps} cycles at B={B}, {cb['cores']} threads"},
            "e2e": {"value": v, "unit": "rolls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# extra records of the default single-GPU line
# ------------------------------------------------------------------------------------------------
def probe_family(L, family, fn):
    import ctypes
    L.mg_probe_begin(family)
    fn()
    pr = (ctypes.c_double * 4)()
    L.mg_probe_end(pr)
    return {"launches": pr[0], "ms": pr[1], "flops": pr[2], "bytes": pr[3]}


def hbm_record(name, pr, peak_gbs, note):
    gbs = pr["bytes"] / (pr["ms"] * 1e-3) / 1e9 if pr["ms"] > 0 else 0.0
    return {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s",
            "frac": gbs / peak_gbs if peak_gbs else None, "launches_per_step": pr["launches"], "kernel_ms_per_step": pr["ms"],
            "algorithmic_bytes_per_step": pr["bytes"], "note": note}


def notes_records(L, dev, peak_gbs, R=262144):
    """Config #5 kernels: N-1 (src/gan/utils.py:130-155) and N-2 (tools/roll_to_midi.py:10-21) over R resident rolls
    (2 GiB >> L2), CUDA events around 10 launches each."""
    import torch
    from melogan import notes as N

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    rolls = torch.rand(R, 512, 4, device=dev) * 2 - 1
    out = N.extract_notes_gan(rolls, 140.0, "major", 0, check=False)
    emitted = int(out.counts.sum().item())
    del out
    # per-launch output buffers are allocated by the wrapper (cached by torch's allocator after the warm-up)
    ms1 = timed(lambda: N.extract_notes_gan(rolls, 140.0, "major", 0, check=False))
    b1 = R * (8192 + 4) + emitted * 18
    rolls.mul_(50).add_(50)
    ms2 = timed(lambda: N.extract_notes_abs(rolls, check=False))
    b2 = R * (8192 + 9216)
    del rolls
    torch.cuda.empty_cache()
    rec = lambda name, ms, b, note: {"kernel": name, "rolls": R, "ms": ms, "rolls_per_s": R / ms * 1e3, "bound": "hbm",
                                     "achieved": b / ms / 1e6, "peak": peak_gbs, "unit": "GB/s", "frac": b / ms / 1e6 / peak_gbs,
                                     "algorithmic_bytes": b, "note": note}
    return [rec("extract_notes_gan_kernel (N-1)", ms1, b1, f"8192 B read + 4 B count per roll + 18 B per emitted note "
                                                             f"({emitted / R:.0f} notes per roll on U(-1,1) rolls)"),
            rec("extract_notes_abs_kernel (N-2)", ms2, b2, "8192 B read + 9216 B written per roll")]


def aux_records(dev, steps=30):
    """BASELINE configs #2 / #3 through their fast-path trainers (melogan/aux_trainers.py): one step = forward + loss +
    backward + (clip) + AdamW, replayed as ONE CUDA graph over a static batch that is refreshed from a rotating set of
    resident batches before every replay; device RNG for eps / dropout.  Reference loops: src/ae/train_ae.py:100-122,
    src/emotion_discriminator/train_ed.py:61-74.  Returns one record per (config, precision, batch)."""
    import contextlib
    import torch
    import yaml
    from melogan.aux_trainers import EdTrainer, VaeTrainer

    with open(os.path.join(PKG, "config", "ae_config.yaml")) as f:
        ae_cfg = yaml.safe_load(f)
    with open(os.path.join(PKG, "config", "ed_config.yaml")) as f:
        ed_cfg = yaml.safe_load(f)
    out = []

    def timed(fn, n):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(5 + i)
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    g = torch.Generator(device=dev).manual_seed(11)
    for precision in ("fp32", "bf16"):
        for B in (int(ae_cfg.get("BATCH_SIZE", 32)), 1024):
            with contextlib.redirect_stdout(sys.stderr):
                tr = VaeTrainer(dict(ae_cfg), batch=B, precision=precision, device=dev)
            xs = [torch.rand((B, tr.T, 4), generator=g, device=dev) * 2 - 1 for _ in range(3)]
            tr.step(xs[0])
            sx = tr.capture()

            def vstep(i):
                sx.copy_(xs[i % 3]); tr.replay()
            ms = timed(vstep, steps)
            loss = tr.epoch_means()[0]
            out.append({"config": "#2 VAE train step (ae_config.yaml)", "precision": precision, "batch": B, "ms_per_step": ms,
                        "rolls_per_s": B / ms * 1e3, "cuda_graph": True, "loss_finite": bool(loss == loss and abs(loss) < 1e30)})
            del tr, sx, xs
            torch.cuda.empty_cache()
        for B in (int(ed_cfg.get("batch_size", 64)), 1024):
            with contextlib.redirect_stdout(sys.stderr):
                tr = EdTrainer(dict(ed_cfg), batch=B, precision=precision, device=dev)
            xs = [torch.rand((B, tr.T, 4), generator=g, device=dev) * 2 - 1 for _ in range(3)]
            y = (torch.arange(B, device=dev) % tr.n_classes).to(torch.int64)
            tr.step(xs[0], y)
            sx, sy = tr.capture()
            sy.copy_(y)

            def estep(i):
                sx.copy_(xs[i % 3]); tr.replay()
            ms = timed(estep, steps)
            loss = tr.epoch_means()[0]
            out.append({"config": "#3 emotion-discriminator train step (ed_config.yaml)", "precision": precision, "batch": B,
                        "ms_per_step": ms, "rolls_per_s": B / ms * 1e3, "cuda_graph": True,
                        "loss_finite": bool(loss == loss and abs(loss) < 1e30)})
            del tr, sx, sy, xs
            torch.cuda.empty_cache()
    return out


def run_aux_workload(args):
    """`--workload ae|ed`: one training step (forward + loss + backward + clip + AdamW) of BASELINE config #2 / #3 per "step",
    replayed as one CUDA graph; `value` with the batch resident on the device, `e2e` with the batch copied from pinned host
    memory and the loss read back on the host every step.  Batch: --batch if given on the command line, else 1024."""
    import contextlib
    import torch
    import yaml
    from melogan import _native
    from melogan.aux_trainers import EdTrainer, VaeTrainer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    B = args.batch if "--batch" in sys.argv else 1024
    W = max(3, args.warmup)
    name = {"ae": "ae_config.yaml", "ed": "ed_config.yaml"}[args.workload]
    with open(os.path.join(PKG, "config", name)) as f:
        cfg = yaml.safe_load(f)
    with contextlib.redirect_stdout(sys.stderr):
        tr = (VaeTrainer if args.workload == "ae" else EdTrainer)(dict(cfg), batch=B, precision=args.precision, device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    xs = [torch.rand((B, tr.T, 4), generator=g, device=dev) * 2 - 1 for _ in range(3)]
    hx = [x.cpu().pin_memory() for x in xs]
    y = (torch.arange(B, device=dev) % 4).to(torch.int64)
    L = _native.lib()
    step = (lambda x: tr.step(x)) if args.workload == "ae" else (lambda x: tr.step(x, y))
    step(xs[0])
    torch.cuda.synchronize(dev)
    n0 = L.mg_launch_count()
    step(xs[1])
    torch.cuda.synchronize(dev)
    launches = int(L.mg_launch_count() - n0)
    cap = tr.capture()
    sx = cap if args.workload == "ae" else cap[0]
    if args.workload == "ed":
        cap[1].copy_(y)
    h_loss = torch.empty_like(tr.metrics, device="cpu").pin_memory()

    def dev_step(i):
        sx.copy_(xs[i % 3]); tr.replay()

    def e2e_step(i):
        sx.copy_(hx[i % 3], non_blocking=True); tr.replay()
        h_loss.copy_(tr.metrics, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    def timed(fn):
        for i in range(W):
            fn(i)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.steps):
            fn(W + i)
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / args.steps
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(dev_step)
    sampler.stop_flag.set(); sampler.join(timeout=3)
    ms_e2e = timed(e2e_step)
    what = {"ae": "VAE train step (config #2: forward, vae_loss, backward, clip_grad_norm_(1.0), AdamW)",
            "ed": "emotion-discriminator train step (config #3: forward, cross-entropy, backward, AdamW)"}[args.workload]
    print(json.dumps({
        "metric": what + " rolls/sec", "value": B / ms * 1e3, "unit": "rolls/s", "n_gpus": 1, "steps": args.steps, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"{what}, config/{name} shapes, batch {B}", "per_gpu_batch": B, "precision": args.precision,
                   "cuda_graph": True, "l2": "3 rotating input batches"},
        "e2e": {"value": B / ms_e2e * 1e3, "unit": "rolls/s", "h2d_bytes_per_step": B * tr.T * 16,
                "d2h_bytes_per_step": int(tr.metrics.numel()) * 4, "ms_per_step": ms_e2e},
        "gpu_launches": launches * args.steps, "clocks": sampler.summary()}))


def parity_check(tr, cfg, ed_cfg, reals, numerics, labels, dev):
    """One critic step and one generator step of the BENCHED configuration (bf16 mode, bench batch, trained-for-a-few-
    steps parameters) against this engine's own fp32 parity mode (CUDA-core kernels, pinned to the oracle at 1e-5 by
    tests/) on identical parameters, inputs and injected noise / alpha / dropout masks.  north_star: 1e-2 in bf16 mode."""
    import torch
    from melogan.trainer import GanTrainer
    B = tr.B
    g = torch.Generator(device=dev).manual_seed(1234)
    h = tr.mask1.shape[1], tr.mask2.shape[1]
    noise = torch.randn((B, cfg['NOISE_DIM']), generator=g, device=dev)
    alpha = torch.rand(B, generator=g, device=dev)
    m1 = (torch.rand((B, h[0]), generator=g, device=dev) < tr.keep).float()
    m2 = (torch.rand((B, h[1]), generator=g, device=dev) < tr.keep).float()
    real, numeric = reals[0][0].contiguous(), numerics[0][0].contiguous()
    bns = (tr.G.decoder.deconv[1], tr.G.decoder.deconv[4])
    saved = [(bn.running_mean.clone(), bn.running_var.clone()) for bn in bns]

    def restore():
        for bn, (rm, rv) in zip(bns, saved):
            bn.running_mean.copy_(rm); bn.running_var.copy_(rv)

    def one(engine):
        md = engine.critic_step(real, numeric, noise, alpha, m1, m2).clone()
        restore()
        mg = engine.generator_step(numeric, noise, labels, m1, m2).clone()
        restore()
        torch.cuda.synchronize(dev)
        return [float(x) for x in md.cpu()] + [float(x) for x in mg.cpu()]

    got = one(tr.engine)
    tr.engine.close()                                 # the fp32 context needs the memory
    torch.cuda.empty_cache()
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        tr32 = GanTrainer(cfg, ed_cfg, batch=B, precision="fp32", device=dev, modules=(tr.E_num, tr.G, tr.D, tr.ED))
    want = one(tr32.engine)
    tr32.engine.close()
    torch.cuda.empty_cache()
    names = ["loss_d", "gp", "d_real", "d_fake", "g_adv", "g_emo"]
    rec, ok = {}, True
    for n, a, b in zip(names, got, want):
        denom = max(abs(b), 1e-2 if n in ("d_real", "d_fake", "g_adv") else 1e-6)   # the critic scores start near 0
        r = abs(a - b) / denom
        rec[n] = {"bf16": a, "fp32": b, "rel": r}
        if n in ("loss_d", "gp", "g_emo") and not r <= 1e-2:
            ok = False
    rec["tolerance"] = 1e-2
    rec["ok"] = ok
    rec["what"] = f"one critic + one generator step at B={B}: bf16 tensor-core mode vs the engine's fp32 parity mode, same inputs"
    return rec


def fp32_tc_record(cfg, ed_cfg, dev, B=2048, cycles=3):
    """The fp32 parity mode (1e-5 against the oracle) with its contractions on the CUDA cores (default) and as six bf16
    tensor-core terms per product (mg_debug_set("fp32_tc", 1): float32 operands split exactly into three bf16 parts, the same
    tcgen05 kernels as bf16 mode, fp32 TMEM accumulation): eager cycles of 5 critic + 1 generator step at per-GPU batch B,
    identical parameters and inputs, losses of the first cycle compared."""
    import ctypes
    import torch
    from melogan import _native
    from melogan.trainer import GanTrainer
    L = _native.lib()
    L.mg_debug_set.argtypes = [ctypes.c_char_p, ctypes.c_int]
    K = int(cfg.get("CRITIC_ITERS", 5))
    T, F = cfg["MAX_NOTES"], cfg.get("NUMERIC_INPUT_DIM", 6)
    g = torch.Generator(device=dev).manual_seed(99)
    reals = torch.rand((K, B, T, 4), generator=g, device=dev) * 2 - 1
    numerics = torch.randn((K, B, F), generator=g, device=dev)
    numerics[..., 5] = 0.0
    labels = (torch.arange(B, device=dev) % 4).to(torch.int64)
    rec = {"per_gpu_batch": B, "cycles_timed": cycles}
    import contextlib
    for name, on in (("cuda_cores", 0), ("tensor_cores_bf16x6", 1)):
        L.mg_debug_set(b"fp32_tc", on)
        try:
            torch.manual_seed(4321)
            with contextlib.redirect_stdout(sys.stderr):
                tr = GanTrainer(cfg, ed_cfg, batch=B, precision="fp32", device=dev, seed_offset=0)
            tr.train_cycle(reals, numerics, labels)              # also the lazily allocated scratch
            torch.cuda.synchronize(dev)
            first = [float(x) for x in tr.loss_acc.cpu()]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(cycles):
                tr.train_cycle(reals, numerics, labels)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / cycles
            rec[name] = {"ms_per_cycle": ms, "rolls_per_s": K * B / (ms * 1e-3), "first_cycle_loss_sums": first}
            tr.engine.close()
        finally:
            L.mg_debug_set(b"fp32_tc", -1)
        torch.cuda.empty_cache()
    a, b = rec["tensor_cores_bf16x6"]["first_cycle_loss_sums"], rec["cuda_cores"]["first_cycle_loss_sums"]
    rec["max_rel_diff_of_losses"] = max(abs(x - y) / max(abs(y), 1e-3) for x, y in zip(a, b))
    rec["speedup"] = rec["cuda_cores"]["ms_per_cycle"] / rec["tensor_cores_bf16x6"]["ms_per_cycle"]
    # whole-cycle float32 FLOP rate against the effective peak of the six-term form: one float32 product costs six bf16 products
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf = float(json.load(f).get("bf16_tflops_sustained", 1400.0))
    except Exception:
        peak_tf = 1400.0
    for k in ("cuda_cores", "tensor_cores_bf16x6"):
        rec[k]["whole_cycle_fp32_tflops"] = rec[k]["rolls_per_s"] * MFLOP_PER_ROLL * 1e6 / 1e12
    rec["effective_fp32_peak_tflops"] = peak_tf / 6.0
    rec["frac_of_effective_peak"] = rec["tensor_cores_bf16x6"]["whole_cycle_fp32_tflops"] / (peak_tf / 6.0)
    rec["note"] = ("opt-in (MELOGAN_FP32_TC=1): operands split exactly into three bf16 parts, six part products per float32 product on "
                   "the bf16 tcgen05 kernels; effective peak = measured sustained bf16 peak / 6; the whole-cycle rate includes the "
                   "operand-split passes, the 4-channel layers, element-wise kernels and Adam (DESIGN.md 5)")
    return rec


def stock_torch_yardstick(dev, B, budget_cycles=3):
    """The reference modules run by stock PyTorch eager (cuDNN / cuBLAS) on THIS GPU, same cycle, same batch:
    the "reference's Blackwell kernels" this framework has to beat (SURVEY.md 8d / BASELINE.md 3.4)."""
    import torch
    rcm = _ref_cycle()
    if not rcm.available():
        return {"unavailable": "reference sources not staged under baseline/_ref"}
    out = {"B": B, "torch": torch.__version__}
    import contextlib
    for name, ac in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        b = B
        while True:
            try:
                with contextlib.redirect_stdout(sys.stderr):
                    r = rcm.time_cycles(str(dev), b, steps=budget_cycles, warmup=1, autocast=ac)
                out[name] = {"rolls_per_s": r["rolls_per_s"], "ms_per_cycle": r["ms_per_cycle"], "B": b}
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                b //= 2
                if b < 256:
                    out[name] = {"unavailable": "out of memory"}
                    break
            except Exception as e:       # e.g. an op without a bf16 double-backward kernel
                out[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
                break
        torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--workload", default="gan", choices=["gan", "ae", "ed"],
                    help="gan = the headline cycle (BASELINE config #4); ae / ed = one training step of BASELINE config #2 / #3 "
                         "through melogan.aux_trainers (single GPU)")
    ap.add_argument("--sync-bn", action="store_true",
                    help="N > 1: BatchNorm statistics over all ranks through NVLink peer memory (default: local, like torch DDP)")
    ap.add_argument("--peer-allreduce", action="store_true",
                    help="N > 1: gradient exchange as kernels over NVLink peer memory (one graph per cycle) instead of NCCL")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the HBM-kernel, note-extraction, parity and yardstick records")
    ap.add_argument("--no-yardstick", action="store_true", help="skip the stock-PyTorch-eager run of the reference modules")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "gan":
        return run_aux_workload(args)

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
        tr = GanTrainer(cfg, ed_cfg, batch=B, precision=args.precision, device=dev, process_group=pg, seed_offset=rank,
                        sync_bn=args.sync_bn, peer_allreduce=True if args.peer_allreduce else None)

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
        # The call a user makes (GanTrainer.prefetch / replay_cycle_prefetched): this step's inputs were put on the copy
        # stream from pinned host memory while the previous step computed (the first one by the warm-up below); the
        # step waits for them, replays the cycle, starts the H2D of the NEXT step's inputs and reads this step's losses
        # on the host.  Every step's 337 MB H2D and its D2H are inside the timed region.
        if use_graph:
            tr.replay_cycle_prefetched()
            j = (i + 1) % NSETS
            tr.prefetch(h_reals[j], h_numerics[j], h_labels)
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
    if use_graph:
        tr.prefetch(h_reals[0], h_numerics[0], h_labels)  # inputs of the first (warm-up) step
    ms_e2e = timed(e2e_step, args.steps)

    # per-kernel roofline of the dominant kernel family, measured live with CUDA events (eager pass)
    import ctypes
    family = 3 if args.precision == "bf16" else 1
    eager_cycle = lambda: tr.train_cycle(reals[0], numerics[0], labels)
    pr = probe_family(L, family, eager_cycle)
    probe_launches, probe_ms, probe_flops, probe_bytes = pr["launches"], pr["ms"], pr["flops"], pr["bytes"]
    # the same family with the epilogue reductions (pooling, bias-gradient sums, BatchNorm statistics) switched off, i.e. the
    # contractions alone as in round 1 (the 35 separate passes then run in the element-wise family instead)
    pr_plain = None
    if family == 3:
        try:
            L.mg_debug_set(b"no_fuse", 7)
            eager_cycle()
            pr_plain = probe_family(L, family, eager_cycle)
        finally:
            L.mg_debug_set(b"no_fuse", 0)
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
    # DRAM traffic of the SAME launches: profiles/r02_tc_traffic.json holds, per layer of the cycle, the dram read+write
    # bytes of one `ncu --set full` capture of scripts/bench_layers.py next to that launch's algorithmic bytes
    traffic, traffic_note = None, "no ncu capture committed"
    try:
        with open(os.path.join(ROOT, "profiles", "r02_tc_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch_cycle_weighted")
        traffic_note = (f"dram read+write bytes per launch, cycle-weighted over the {tj.get('launches_per_cycle')} tap-GEMM launches of "
                        f"one cycle (ncu --set full of scripts/bench_layers.py, profiles/r02_tc_traffic.json); algorithmic bytes of the "
                        f"same launches = {tj.get('algorithmic_bytes_per_launch_cycle_weighted'):.3e} per launch "
                        f"(ratio {tj.get('ratio'):.2f}); live probe of this run: {probe_bytes / max(probe_launches, 1):.3e}")
    except Exception:
        pass
    extra = {}
    if rank == 0 and world == 1 and not args.no_extras:
        # HBM-class kernels north_star names: fused Adam, the element-wise / reduction family, note extraction
        hbm = []
        pa = probe_family(L, 6, eager_cycle)
        hbm.append(hbm_record("adam_kernel (fused Adam over the flat groups)", pa, peak_gbs,
                              "28 B per parameter and step; 5 launches on the critic's 0.27 M parameters (7.6 MB: latency-bound) + "
                              "1 on the generator's 8.85 M (248 MB), see largest_group"))
        try:       # the generator group alone: the one Adam launch that is large enough to be bandwidth-bound
            pg_ = probe_family(L, 6, lambda: tr.generator_step(numerics[0][K - 1], labels))
            hbm[-1]["largest_group"] = hbm_record("adam_kernel, generator + encoder group", pg_, peak_gbs, "8.85 M parameters")
        except Exception as e:
            hbm[-1]["largest_group"] = {"error": f"{type(e).__name__}: {e}"}
        pe = probe_family(L, 7, eager_cycle)
        hbm.append(hbm_record("element-wise / reduction family (remaining column sums, BatchNorm apply + backward, row broadcast x mask; pooling, bias-gradient sums and "
                              "BatchNorm statistics now ride in the tap-GEMM epilogues)",
                              pe, peak_gbs, "one read (+ one write) of the activation per kernel"))
        extra["roofline_hbm"] = hbm
        try:
            extra["notes"] = notes_records(L, dev, peak_gbs)
        except Exception as e:
            extra["notes"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            extra["parity_check"] = parity_check(tr, cfg, ed_cfg, reals, numerics, labels, dev)
        except Exception as e:
            extra["parity_check"] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        del reals, numerics
        torch.cuda.empty_cache()
        try:
            extra["aux_configs"] = aux_records(dev)
        except Exception as e:
            extra["aux_configs"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_yardstick:
            try:
                extra["stock_torch_eager_same_gpu"] = stock_torch_yardstick(dev, B)
            except Exception as e:
                extra["stock_torch_eager_same_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
        try:       # last: an opt-in kernel path; nothing after it depends on the device
            extra["fp32_mode"] = fp32_tc_record(cfg, ed_cfg, dev)
        except Exception as e:
            extra["fp32_mode"] = {"error": f"{type(e).__name__}: {e}"}

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
                   "gradient_exchange": ("peer memory (NVLink P2P kernels, one graph per cycle)" if tr._peer is not None
                                         else "NCCL all-reduce between two graphs per step") if world > 1 else "n/a",
                   "bn": ("sync (peer memory)" if tr.sync_bn else "local") if world > 1 else "n/a",
                   "l2": f"{NSETS} rotating input sets; per-step activation working set {tr.engine.workspace_bytes() / 1e6:.0f} MB >> 126 MB L2",
                   "algorithmic_mflop_per_roll": MFLOP_PER_ROLL},
        "e2e": {"value": e2e_value, "unit": "rolls/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                "ms_per_step": ms_e2e / args.steps,
                "pipeline": "H2D of step i+1 on a copy stream under the compute of step i (GanTrainer.prefetch), one D2D into the "
                            "graph's static buffers per step, losses read on the host every step"},
        "gpu_launches": launches_per_cycle * args.steps,
        "clocks": sampler.summary() if sampler else None,
        "roofline": {"bound": "tensor", "kernel": {1: "tapgemm_kernel (CUDA-core fp32 implicit GEMM)",
                                                   3: "tc_gemm_kernel (tcgen05 bf16 implicit GEMM)"}[family],
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "peak_source": peak_src,
                     "traffic_note": traffic_note,
                     "launches_per_step": probe_launches, "kernel_ms_per_step": probe_ms,
                     "flops_note": "algorithmic FLOPs per launch (2 x rows x N x taps x K); the banded 4-channel layers count their "
                                   "zero-padded K = 64 windows (20 useful taps x channels), about 1 % of the family total",
                     "fused_reductions_note": "these launches also do the pooling, bias-gradient and BatchNorm-statistics reductions "
                                              "of 35 former element-wise passes in their epilogues: kernel_ms includes that work, "
                                              "the FLOP count does not",
                     "contractions_only": ({"kernel_ms_per_step": pr_plain["ms"],
                                            "achieved": pr_plain["flops"] / (pr_plain["ms"] * 1e-3) / 1e12,
                                            "frac": pr_plain["flops"] / (pr_plain["ms"] * 1e-3) / 1e12 / peak_tf,
                                            "what": "same launches with mg_debug_set('no_fuse', 7): reductions back in separate passes"}
                                           if pr_plain and pr_plain["ms"] > 0 else None),
                     "share_of_step": probe_ms / (ms_dev / args.steps) if ms_dev > 0 else None,
                     "hbm_view": {"achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s",
                                  "frac": achieved_gbs / peak_gbs if peak_gbs else None}},
        "whole_step_model_tflops": value * MFLOP_PER_ROLL * 1e6 / 1e12 / max(world, 1),
    }
    line.update(extra)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_cycle_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
