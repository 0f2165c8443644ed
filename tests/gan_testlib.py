"""Shared helpers of the GPU parity tests: run the oracle on the CPU and the CUDA engine on the same
seeded inputs and compare tensors relative to their scale."""
import torch

from melogan import engine as E
from oracle import gan_oracle as O

MODS = {"E": E.MOD_E, "G": E.MOD_G, "D": E.MOD_D, "ED": E.MOD_ED}


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    denom = want.abs().max().item()
    return (got - want).abs().max().item() / max(denom, 1e-30)


def assert_close(got, want, tol, what, want64=None):
    """max|got - want| / max|want| <= tol.  With want64 (the same oracle evaluated in float64) the
    comparison is made against float64 and the allowance is widened to 3x the float32 oracle's own
    error: where the reference's result is dominated by cancellation noise (sums of +1/B and -1/B
    weighted terms), nobody can be closer to it than it is to the exact value."""
    assert tuple(got.shape) == tuple(want.shape), (what, got.shape, want.shape)
    if want64 is not None:
        yard = rel_err(want, want64)
        e = rel_err(got, want64)
        allowed = max(tol, 3.0 * yard)
        assert e <= allowed, f"{what}: err vs fp64 = {e:.3e} > {allowed:.1e} (fp32 oracle's own error {yard:.1e})"
        return e
    e = rel_err(got, want)
    assert e <= tol, f"{what}: max|diff|/max|ref| = {e:.3e} > {tol:.1e}"
    return e


def rel_l2(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-300)).item()


def assert_close_l2(got, want, tol, what):
    """Relative L2 error.  Used for bf16-mode gradients: a bf16 forward flips the sign of a few near-zero
    pre-activations relative to the fp32 oracle, which changes single gradient elements by O(1) (LeakyReLU /
    ReLU masks) while the tensor as a whole stays within rounding."""
    assert tuple(got.shape) == tuple(want.shape), (what, got.shape, want.shape)
    e = rel_l2(got, want)
    assert e <= tol, f"{what}: ||diff||/||ref|| = {e:.3e} > {tol:.1e}"
    return e


def to_double(obj):
    if isinstance(obj, dict):
        return {k: to_double(v) for k, v in obj.items()}
    if isinstance(obj, torch.Tensor) and obj.is_floating_point():
        return obj.double()
    return obj


def cuda_params(params):
    return {m: {k: v.clone().cuda() for k, v in P.items()} for m, P in params.items()}


def zero_grads(P, keys):
    return {k: torch.zeros_like(P[k]) for k in keys}


def make_engine(B, params, precision="fp32"):
    """Engine with all four modules bound; returns (engine, cuda params, grads dict per module)."""
    eng = E.GanEngine(B, precision=precision)
    cp = cuda_params(params)
    grads = {m: zero_grads(cp[m], E.GRAD_KEYS[MODS[m]]) for m in ("E", "G", "D")}
    for m in ("E", "G", "D"):
        eng.bind(MODS[m], cp[m], grads[m])
    eng.bind(E.MOD_ED, cp["ED"], None)
    return eng, cp, grads


def cuda_batch(batch):
    return {k: v.cuda() for k, v in batch.items()}


def make_flat_trainer(B, params, precision="fp32", lr_d=1e-4, lr_g=1e-4, betas=(0.5, 0.9)):
    """Engine + flat parameter groups + fused Adam wired like melogan.trainer.GanTrainer, but fed with explicit
    noise / alpha / masks so that a run can follow the oracle's random draws step by step."""
    from melogan.optim import FlatParams, FusedAdam
    eng = E.GanEngine(B, precision=precision)
    cp = cuda_params(params)
    leafD = [torch.nn.Parameter(cp["D"][k]) for k in E.D_KEYS]
    leafG = [torch.nn.Parameter(cp["G"][k]) for k in E.G_PARAM_KEYS] + [torch.nn.Parameter(cp["E"][k]) for k in E.E_KEYS]
    flatD, flatG = FlatParams(leafD), FlatParams(leafG)
    optD, optG = FusedAdam(flatD, lr=lr_d, betas=betas), FusedAdam(flatG, lr=lr_g, betas=betas)
    nG = len(E.G_PARAM_KEYS)
    Dp = {k: p.data for k, p in zip(E.D_KEYS, leafD)}
    Dg = {k: p.grad for k, p in zip(E.D_KEYS, leafD)}
    Gp = {k: p.data for k, p in zip(E.G_PARAM_KEYS, leafG[:nG])}
    Gp.update({k: cp["G"][k] for k in E.G_BUFFER_KEYS})
    Gg = {k: p.grad for k, p in zip(E.G_PARAM_KEYS, leafG[:nG])}
    Ep = {k: p.data for k, p in zip(E.E_KEYS, leafG[nG:])}
    Eg = {k: p.grad for k, p in zip(E.E_KEYS, leafG[nG:])}
    eng.bind(E.MOD_E, Ep, Eg); eng.bind(E.MOD_G, Gp, Gg); eng.bind(E.MOD_D, Dp, Dg); eng.bind(E.MOD_ED, cp["ED"], None)
    return {"eng": eng, "optD": optD, "optG": optG, "D": Dp, "G": Gp, "E": Ep, "keep": (leafD, leafG, flatD, flatG, cp)}
