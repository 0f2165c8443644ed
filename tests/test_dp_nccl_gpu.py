"""Two-GPU test of the data-parallel CUDA path over NCCL (skips on a one-GPU box).

Each rank runs GanTrainer.critic_step / generator_step on its shard with injected noise / alpha / dropout masks; the
all-reduced flat gradient must equal the SUM of the gradients a single-GPU trainer computes for the two shards one after
the other (local-BN semantics: G's BatchNorm sees the shard, as under torch DDP), and the parameters after the fused Adam
(grad_scale = 1/world) must be identical on both ranks.  Then the captured cycle (two graphs per step with the eager
all-reduce between them) must reproduce eager data-parallel steps, and the checkpoint's BatchNorm running statistics
must be the same on every rank.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(seed, B, cfg):
    g = torch.Generator().manual_seed(seed)
    h = cfg.get('ENCODER_HIDDEN', [256, 128])
    return dict(real=torch.rand((B, cfg['MAX_NOTES'], 4), generator=g) * 2 - 1,
                numeric=torch.randn((B, cfg.get('NUMERIC_INPUT_DIM', 6)), generator=g),
                noise=torch.randn((B, cfg['NOISE_DIM']), generator=g), alpha=torch.rand(B, generator=g),
                mask1=(torch.rand((B, h[0]), generator=g) < 0.8).float(),
                mask2=(torch.rand((B, h[1]), generator=g) < 0.8).float(),
                labels=torch.randint(0, 4, (B,), generator=g))


def _worker(rank, world, port, peer, out):
    for p in (ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import contextlib
    import io
    import yaml
    from melogan.trainer import GanTrainer
    cfgd = os.path.join(ROOT, "melo-gan_b200", "config")
    cfg = yaml.safe_load(open(os.path.join(cfgd, "gan_config.yaml")))
    ed_cfg = yaml.safe_load(open(os.path.join(cfgd, "ed_config.yaml")))
    cfg["INTEGRATION_MODE"] = "numeric_only" if cfg.get("INTEGRATION_MODE") == "conditioning" else cfg.get("INTEGRATION_MODE")
    Bl = 8                                   # per-rank shard
    with contextlib.redirect_stdout(io.StringIO()):
        tr = GanTrainer(cfg, ed_cfg, batch=Bl, precision="fp32", device=dev, process_group=dist.group.WORLD, seed_offset=rank,
                        peer_allreduce=peer)
        ref = GanTrainer(cfg, ed_cfg, batch=Bl, precision="fp32", device=dev) if rank == 0 else None
    assert (tr._peer is not None) == bool(peer)
    res = {}
    full = _inputs(5, Bl * world, cfg)
    mine = {k: v[rank * Bl:(rank + 1) * Bl].to(dev) for k, v in full.items()}
    # ---- critic step ----
    tr.critic_step(mine["real"], mine["numeric"], noise=mine["noise"], alpha=mine["alpha"], mask1=mine["mask1"], mask2=mine["mask2"])
    gd = tr.flat_d.grad.clone()
    if rank == 0:
        want = torch.zeros_like(gd)
        for r in range(world):
            sh = {k: v[r * Bl:(r + 1) * Bl].to(dev) for k, v in full.items()}
            ref.opt_D.lr, lr0 = 0.0, ref.opt_D.lr           # keep the reference trainer's parameters fixed between the shards
            ref.critic_step(sh["real"], sh["numeric"], noise=sh["noise"], alpha=sh["alpha"], mask1=sh["mask1"], mask2=sh["mask2"])
            ref.opt_D.lr = lr0
            want += ref.flat_d.grad
        res["critic_grad_err"] = float((gd - want).abs().max() / want.abs().max())
    # ---- generator step ----
    tr.generator_step(mine["numeric"], mine["labels"], noise=mine["noise"], mask1=mine["mask1"], mask2=mine["mask2"])
    gg = tr.flat_g.grad.clone()
    # parameters identical on every rank after the fused Adam steps
    for name, flat in (("d", tr.flat_d), ("g", tr.flat_g)):
        mx, mn = flat.data.clone(), flat.data.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        res[f"param_spread_{name}"] = float((mx - mn).abs().max())
    mx = gg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    res["g_grad_same_on_ranks"] = float((mx - gg).abs().max())
    # ---- captured cycle == eager data-parallel cycle ----
    K = tr.critic_iters
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    reals = torch.rand((K, Bl, cfg['MAX_NOTES'], 4), generator=g, device=dev) * 2 - 1
    nums = torch.randn((K, Bl, cfg.get('NUMERIC_INPUT_DIM', 6)), generator=g, device=dev)
    labels = (torch.arange(Bl, device=dev) % 4).to(torch.int64)
    snap = (tr.flat_d.data.clone(), tr.flat_g.data.clone(), tr.rng_counter.clone(), tr.opt_D.state_dict(), tr.opt_G.state_dict())
    import copy
    snap = copy.deepcopy(snap)
    tr.train_cycle(reals, nums, labels)                  # eager (also the warm-up the capture needs)
    tr.train_cycle(reals, nums, labels)
    eager_d, eager_g = tr.flat_d.data.clone(), tr.flat_g.data.clone()
    tr.flat_d.data.copy_(snap[0]); tr.flat_g.data.copy_(snap[1]); tr.rng_counter.copy_(snap[2])
    tr.opt_D.load_state_dict(snap[3]); tr.opt_G.load_state_dict(snap[4])
    tr.engine.weight_cache(True)
    s_reals, s_nums, s_labels = tr.capture_cycle()
    s_reals.copy_(reals); s_nums.copy_(nums); s_labels.copy_(labels)
    tr.replay_cycle(); tr.replay_cycle()
    torch.cuda.synchronize(dev)
    res["graph_vs_eager_d"] = float((tr.flat_d.data - eager_d).abs().max() / eager_d.abs().max())
    res["graph_vs_eager_g"] = float((tr.flat_g.data - eager_g).abs().max() / eager_g.abs().max())
    res["one_graph"] = bool(getattr(tr, "_one_graph", False))
    ck = tr.state_dict()                                 # collective: BatchNorm running stats averaged over ranks
    rm = ck["G"]["decoder.deconv.1.running_mean"].clone()
    mx = rm.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    res["bn_running_mean_spread"] = float((mx - rm).abs().max())
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def _run(peer):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + os.getpid() % 200 + (50 if peer else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, peer, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not alive, "data-parallel worker hung"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = out.get(timeout=10)
    print(res)
    assert res["critic_grad_err"] < 2e-5, res
    assert res["param_spread_d"] == 0.0 and res["param_spread_g"] == 0.0 and res["g_grad_same_on_ranks"] == 0.0, res
    assert res["graph_vs_eager_d"] < 2e-3 and res["graph_vs_eager_g"] < 2e-3, res     # fp32 atomics + Adam on near-zero g
    assert res["bn_running_mean_spread"] == 0.0, res
    assert res["one_graph"] == bool(peer), res           # peer-memory exchange: the whole data-parallel cycle is ONE graph
    return res


@pytest.mark.timeout(600)
def test_dp_two_gpus_gradients_and_captured_cycle():
    _run(False)


@pytest.mark.timeout(600)
def test_dp_two_gpus_gradient_exchange_over_peer_memory_one_graph():
    """The same checks with the gradient all-reduce as kernels over NVLink peer memory (csrc/peer.cu) instead of NCCL: no
    collective library call in the step, the data-parallel cycle is captured as ONE CUDA graph."""
    _run(True)


def _worker_syncbn(rank, world, port, precision, out):
    for p in (ROOT, os.path.join(ROOT, "melo-gan_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import contextlib
    import io
    import yaml
    from melogan.trainer import GanTrainer
    cfgd = os.path.join(ROOT, "melo-gan_b200", "config")
    cfg = yaml.safe_load(open(os.path.join(cfgd, "gan_config.yaml")))
    ed_cfg = yaml.safe_load(open(os.path.join(cfgd, "ed_config.yaml")))
    Bl = 16
    with contextlib.redirect_stdout(io.StringIO()):
        tr = GanTrainer(cfg, ed_cfg, batch=Bl, precision=precision, device=dev, process_group=dist.group.WORLD, seed_offset=rank,
                        sync_bn=True)
        ref = GanTrainer(cfg, ed_cfg, batch=Bl * world, precision=precision, device=dev) if rank == 0 else None
    assert tr.sync_bn
    full = _inputs(9, Bl * world, cfg)
    mine = {k: v[rank * Bl:(rank + 1) * Bl].to(dev) for k, v in full.items()}
    res = {}
    # one generator step on the shard, statistics over the global batch
    tr.generator_step(mine["numeric"], mine["labels"], noise=mine["noise"], mask1=mine["mask1"], mask2=mine["mask2"])
    notes = tr.engine.buffer("g.notes")[:Bl * 512 * 4].clone()
    gg = tr.flat_g.grad.clone()                         # all-reduced SUM over ranks; the mean is grad_scale = 1/world in Adam
    bn = [b.clone() for b in (tr.G.decoder.deconv[1].running_mean, tr.G.decoder.deconv[1].running_var,
                              tr.G.decoder.deconv[4].running_mean, tr.G.decoder.deconv[4].running_var)]
    gathered = [torch.empty_like(notes) for _ in range(world)]
    dist.all_gather(gathered, notes)
    for name, t in zip(("rm1", "rv1", "rm2", "rv2"), bn):
        mx, mn = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        res[f"spread_{name}"] = float((mx - mn).abs().max())
    if rank == 0:
        f = {k: v.to(dev) for k, v in full.items()}
        ref.generator_step(f["numeric"], f["labels"], noise=f["noise"], mask1=f["mask1"], mask2=f["mask2"])
        want_notes = ref.engine.buffer("g.notes")[:Bl * world * 512 * 4]
        got_notes = torch.cat(gathered)
        res["notes_err"] = float((got_notes - want_notes).abs().max() / want_notes.abs().max())
        want_g = ref.flat_g.grad
        res["grad_err"] = float(((gg / world) - want_g).norm() / want_g.norm())
        rbn = (ref.G.decoder.deconv[1].running_mean, ref.G.decoder.deconv[1].running_var,
               ref.G.decoder.deconv[4].running_mean, ref.G.decoder.deconv[4].running_var)
        res["bn_err"] = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(bn, rbn))
    # the peer exchange inside captured step graphs: a captured cycle reproduces an eager one
    K = tr.critic_iters
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    reals = torch.rand((K, Bl, cfg['MAX_NOTES'], 4), generator=g, device=dev) * 2 - 1
    nums = torch.randn((K, Bl, cfg.get('NUMERIC_INPUT_DIM', 6)), generator=g, device=dev)
    labels = (torch.arange(Bl, device=dev) % 4).to(torch.int64)
    snap = (tr.flat_d.data.clone(), tr.flat_g.data.clone(), tr.rng_counter.clone(), tr.opt_D.state_dict(), tr.opt_G.state_dict())
    tr.train_cycle(reals, nums, labels)
    eager_g = tr.flat_g.data.clone()
    tr.flat_d.data.copy_(snap[0]); tr.flat_g.data.copy_(snap[1]); tr.rng_counter.copy_(snap[2])
    tr.opt_D.load_state_dict(snap[3]); tr.opt_G.load_state_dict(snap[4])
    tr.engine.weight_cache(True)
    s_reals, s_nums, s_labels = tr.capture_cycle()
    s_reals.copy_(reals); s_nums.copy_(nums); s_labels.copy_(labels)
    tr.replay_cycle()
    torch.cuda.synchronize(dev)
    res["graph_vs_eager_g"] = float((tr.flat_g.data - eager_g).abs().max() / eager_g.abs().max())
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sync_batchnorm_over_peer_memory_equals_single_gpu_full_batch(precision):
    """SyncBatchNorm (mg_gan_sync_bn_*: statistics summed over the ranks through NVLink peer memory, no NCCL in the step):
    two ranks with 16 samples each must reproduce ONE GPU running the 32-sample batch -- generated notes, BatchNorm running
    statistics and the averaged generator gradient -- which local BatchNorm cannot (its statistics see 16 samples)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29900 + os.getpid() % 90 + (5 if precision == "bf16" else 0)
    procs = [ctx.Process(target=_worker_syncbn, args=(r, 2, port, precision, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not alive, "SyncBatchNorm worker hung"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = out.get(timeout=10)
    print(precision, res)
    assert all(res[f"spread_{n}"] == 0.0 for n in ("rm1", "rv1", "rm2", "rv2")), res      # rank-order sums: bit-identical
    tight = precision == "fp32"
    assert res["notes_err"] < (2e-5 if tight else 2e-2), res
    assert res["bn_err"] < (2e-5 if tight else 1e-3), res
    assert res["grad_err"] < (2e-4 if tight else 1.5e-1), res        # bf16: mask flips between a 16- and a 32-sample tiling
    assert res["graph_vs_eager_g"] < 2e-3, res


def _worker_peer(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "melo-gan_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from melogan.dist import PeerAllReduce
    sizes = [8848000, 272384, 4096 * world, 1000, 8848000, 12, 272384]      # both algorithms, odd tails, repeated epochs
    peer = PeerAllReduce(max(sizes), device=dev)
    worst, spread = 0.0, 0.0
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    for it, n in enumerate(sizes * 2):
        x = torch.randn(n, generator=g, device=dev)
        want = x.double().clone()
        dist.all_reduce(want)
        got = x.clone()
        peer.allreduce_sum_(got)
        worst = max(worst, float((got.double() - want).abs().max() / want.abs().max()))
        mx, mn = got.clone(), got.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        spread = max(spread, float((mx - mn).abs().max()))
    # inside a CUDA graph, replayed
    x = torch.randn(sizes[0], generator=g, device=dev)
    buf = x.clone()
    torch.cuda.synchronize(dev)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        peer.allreduce_sum_(buf)
    for _ in range(3):
        buf.copy_(x)
        gr.replay()
    want = x.double().clone()
    dist.all_reduce(want)
    graph_err = float((buf.double() - want).abs().max() / want.abs().max())
    torch.cuda.synchronize(dev)
    if rank == 0:
        out.put({"worst": worst, "spread": spread, "graph_err": graph_err})
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4])
def test_peer_memory_allreduce_equals_nccl(world):
    """mg_peer_allreduce_sum (csrc/peer.cu) against NCCL's all-reduce in float64 on random vectors of the flat-gradient
    sizes: 2 ranks take the read-everything form, 4 ranks the reduce-scatter + all-gather form; results must be bit-identical
    on all ranks (rank-order sums), also from inside a replayed CUDA graph."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world})")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 30100 + os.getpid() % 90 + world
    procs = [ctx.Process(target=_worker_peer, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(200)
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not alive, "peer all-reduce worker hung"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = out.get(timeout=10)
    print(world, res)
    assert res["worst"] < 2e-6 and res["graph_err"] < 2e-6 and res["spread"] == 0.0, res
