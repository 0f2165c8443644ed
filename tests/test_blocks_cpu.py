"""Host logic of melogan.blocks that needs no GPU: how the reference's nn.Sequential stacks are parsed into native operator
groups, which layers are refused, that nothing is computed on the CPU, and that spectral-norm weights come through the
wrapper's pre-hooks."""
import pytest
import torch
import torch.nn as nn

from melogan import blocks as B_


def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError):
        B_.linear(torch.zeros(2, 8), torch.zeros(4, 8), torch.zeros(4))
    with pytest.raises(RuntimeError):
        B_.act_dropout(torch.zeros(2, 8), B_.ACT_RELU)
    with pytest.raises(RuntimeError):
        B_.run_conv_stack(nn.Sequential(nn.Conv1d(4, 8, 3, padding=1), nn.BatchNorm1d(8), nn.GELU()), torch.zeros(1, 16, 4))


def test_unsupported_layers_raise_before_any_launch():
    with pytest.raises(NotImplementedError):
        B_.run_conv_stack(nn.Sequential(nn.MaxPool1d(2)), torch.zeros(1, 16, 4))
    with pytest.raises(NotImplementedError):
        B_.run_mlp(nn.Sequential(nn.Softmax(dim=-1)), torch.zeros(2, 8), False)
    with pytest.raises(NotImplementedError):
        B_.run_mlp(nn.Sequential(nn.LeakyReLU(0.1)), torch.zeros(2, 8), False)
    with pytest.raises(NotImplementedError):
        B_.run_mlp(nn.Sequential(nn.GELU(approximate="tanh")), torch.zeros(2, 8), False)


def test_effective_weight_runs_the_spectral_norm_pre_hook():
    from torch.nn.utils import spectral_norm
    torch.manual_seed(0)
    lin = spectral_norm(nn.Linear(16, 8))
    x = torch.randn(3, 16)
    u0 = lin.weight_u.clone()
    W = B_.effective_weight(lin, x)                      # train mode: one power iteration, weight = weight_orig / sigma
    assert not torch.equal(lin.weight_u, u0)
    sigma = torch.dot(lin.weight_u, torch.mv(lin.weight_orig.detach(), lin.weight_v))
    assert torch.allclose(W.detach(), lin.weight_orig.detach() / sigma, rtol=1e-5, atol=1e-6)
    assert W.requires_grad and W.grad_fn is not None     # gradients reach weight_orig
    W.sum().backward()
    assert lin.weight_orig.grad is not None
    plain = nn.Linear(16, 8)
    assert B_.effective_weight(plain, x) is plain.weight


def test_spectral_norm_models_keep_the_reference_state_dict_keys():
    from src.emotion_discriminator.ed_model import EmotionDiscriminator
    cfg = {"input_mode": "notes", "use_spectral_norm": True}
    keys = set(EmotionDiscriminator(cfg).state_dict())
    for k in ("encoder.conv.0.net.0.weight_orig", "encoder.conv.0.net.0.weight_u", "encoder.conv.0.net.0.weight_v",
              "encoder.conv.3.net.0.weight_orig", "classifier.net.0.weight_orig", "classifier.net.3.weight_u",
              "classifier.head.weight", "encoder.project.weight"):
        assert k in keys, k
    plain = set(EmotionDiscriminator({"input_mode": "notes"}).state_dict())
    assert "encoder.conv.0.net.0.weight" in plain and not any("weight_orig" in k for k in plain)
