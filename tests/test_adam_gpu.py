"""Fused Adam/AdamW (mg_adam_step) against torch.optim.Adam/AdamW on the CPU, teacher-forced per step."""
import numpy as np
import pytest
import torch

from melogan import synth

pytestmark = pytest.mark.gpu


def _run(n, steps, kind, lr, betas, wd):
    from melogan.optim import FlatParams, FusedAdam
    p0 = torch.from_numpy(synth.pseudo_normal(1, (n,), 0.02))
    ref_p = torch.nn.Parameter(p0.clone())
    opt_cls = torch.optim.Adam if kind == "adam" else torch.optim.AdamW
    ref = opt_cls([ref_p], lr=lr, betas=betas, weight_decay=wd)
    ours_p = torch.nn.Parameter(p0.clone().cuda())
    flat = FlatParams([ours_p])
    ours = FusedAdam(flat, lr=lr, betas=betas, weight_decay=wd, decoupled=(kind == "adamw"))
    for t in range(steps):
        g = torch.from_numpy(synth.pseudo_normal(100 + t, (n,), 1e-3))
        # teacher forcing: both sides start the step from the reference's state
        ours_p.data.copy_(ref_p.data)
        st = ref.state.get(ref_p)
        if st:
            ours.exp_avg[:n].copy_(st["exp_avg"]); ours.exp_avg_sq[:n].copy_(st["exp_avg_sq"])
        before = ref_p.data.clone()
        ref_p.grad = g.clone()
        ref.step()
        flat.grad[:n].copy_(g)
        ours.step()
        torch.cuda.synchronize()
        upd_ref = (ref_p.data - before).double()
        upd = (ours_p.data.cpu() - before).double()
        scale = upd_ref.abs().max().item()
        # p itself is rounded to float32 after the update: allow one ulp of |p| on top of the update error
        tol = 2e-6 * scale + 1.2e-7 * before.abs().double() + 1e-12
        assert bool(((upd - upd_ref).abs() <= tol).all()), (t, kind)
        st = ref.state[ref_p]
        assert torch.allclose(ours.exp_avg[:n].cpu(), st["exp_avg"], rtol=1e-6, atol=1e-10)
        assert torch.allclose(ours.exp_avg_sq[:n].cpu(), st["exp_avg_sq"], rtol=1e-6, atol=1e-16)
        assert int(ours.step_dev.item()) == t + 1


@pytest.mark.parametrize("n", [1, 7, 272705, 1048579])
def test_adam_matches_torch_gan_hyperparameters(n):
    _run(n, 4, "adam", 2e-4, (0.5, 0.9), 0.0)          # gan_config.yaml:50-55


def test_adamw_matches_torch_ae_and_ed_hyperparameters():
    _run(100003, 3, "adamw", 1e-4, (0.9, 0.999), 1e-5)  # ae_config.yaml / train_ae.py:79
    _run(100003, 3, "adamw", 2e-4, (0.5, 0.999), 0.0)   # ed_config.yaml optimizer block


def test_flat_params_keep_state_dict_semantics():
    from melogan.optim import FlatParams
    lin = torch.nn.Linear(5, 3).cuda()
    w0 = lin.weight.detach().clone()
    flat = FlatParams(list(lin.parameters()))
    assert torch.equal(lin.weight.detach(), w0) and lin.weight.data_ptr() == flat.data.data_ptr()
    assert set(lin.state_dict()) == {"weight", "bias"}
    flat.data.mul_(2.0)
    assert torch.equal(lin.weight.detach(), 2 * w0)


def test_bf16_compute_copy_is_written():
    from melogan.optim import FlatParams, FusedAdam
    p = torch.nn.Parameter(torch.from_numpy(synth.pseudo_normal(3, (4099,), 0.02)).cuda())
    flat = FlatParams([p])
    opt = FusedAdam(flat, lr=1e-3, betas=(0.5, 0.9))
    opt.bf16_copy = torch.zeros(flat.numel, dtype=torch.bfloat16, device="cuda")
    flat.grad.copy_(torch.from_numpy(synth.pseudo_normal(4, (flat.numel,), 1e-3)).cuda())
    opt.step()
    assert torch.equal(opt.bf16_copy, flat.data.to(torch.bfloat16))
