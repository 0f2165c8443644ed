"""World-size-2 gloo test of the data-parallel host logic (melogan.dist): sharding + flat-gradient all-reduce
reproduce the single-process critic gradients and loss (the critic has no BatchNorm, so the identity is exact
up to float summation order).  Runs on CPU; the arithmetic here is the ORACLE's (test infrastructure)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "melo-gan_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from melogan import dist as D_
    from oracle import gan_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    params = O.make_params(4, fan_in_scale=True)
    B = 8
    batch = O.make_batch(40, B)
    # fake notes and embedding come from the FULL batch (G's BatchNorm would otherwise see a different population:
    # that is the local-BN semantics of the real run, documented in DESIGN.md; here we isolate the critic exchange)
    with torch.no_grad():
        emb = O.fe_forward(params["E"], batch["numeric"], batch["mask1_d"], batch["mask2_d"], train=True)
        fake, _ = O.gen_forward(params["G"], batch["noise_d"], emb, train=True,
                                bn_state={k: v.clone() for k, v in params["G"].items() if O.is_buffer(k)})

    def critic_loss(P, sl):
        real, fk, e, a = batch["notes_real"][sl], fake[sl], emb[sl], batch["alpha"][sl]
        return (O.disc_forward(P, fk, e).mean() - O.disc_forward(P, real, e).mean()
                + 10.0 * O.gradient_penalty(P, real, fk, e, a))

    lo, hi = D_.shard_bounds(B, world, rank)
    Dl = O._leaves(params["D"])
    loss = critic_loss(Dl, slice(lo, hi))
    grads = torch.autograd.grad(loss, list(Dl.values()))
    flat = torch.cat([g.flatten() for g in grads])
    scale = D_.allreduce_sum_(flat)
    flat *= scale
    mean_loss = D_.mean_of_rank_means(loss.detach())
    if rank == 0:
        Df = O._leaves(params["D"])
        full = critic_loss(Df, slice(0, B))
        gfull = torch.cat([g.flatten() for g in torch.autograd.grad(full, list(Df.values()))])
        out.put((float((flat - gfull).abs().max() / gfull.abs().max()), float(mean_loss), float(full), scale))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_critic_gradients_equal_full_batch_over_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, mean_loss, full_loss, scale = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale == 0.5
    assert err < 1e-5, err
    assert abs(mean_loss - full_loss) < 1e-5 * abs(full_loss)


def test_shard_bounds():
    from melogan import dist as D_
    assert D_.shard_bounds(32, 4, 1) == (8, 16)
    with pytest.raises(ValueError):
        D_.shard_bounds(30, 4, 0)
    t = torch.arange(12).view(6, 2)
    assert torch.equal(D_.shard(t, 3, 2), t[4:6])
    assert D_.allreduce_sum_(torch.ones(3)) == 1.0     # no process group: identity
