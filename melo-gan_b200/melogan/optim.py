"""Fused Adam/AdamW over flat parameter buffers (host side of mg_adam_step).

Mirrors how the reference drives torch.optim.Adam (src/gan/train_gan.py:136-145,204,248): same
hyper-parameter names, `step()` / `zero_grad()` / `state_dict()`; the arithmetic is one CUDA launch
over the flat (param, grad, exp_avg, exp_avg_sq) buffers the parameters are views of.
"""
import torch

from . import _native


class FlatParams:
    """Re-points the given parameters at views of one flat float32 buffer (and one flat grad buffer).

    state_dict()/load_state_dict() keep working because the views are ordinary tensors; a single
    buffer gives one Adam launch and one gradient all-reduce per optimizer group.
    """

    def __init__(self, params, align=4):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatParams: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise ValueError("FlatParams: parameters must live on a CUDA device (no CPU path)")
        self.offsets, n = [], 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("FlatParams: float32 parameters only")
            self.offsets.append(n)
            n += (p.numel() + align - 1) // align * align
        self.numel = n
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            view = self.data[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view_as(p)

    def offset_of(self, p):
        for q, o in zip(self.params, self.offsets):
            if q is p:
                return o
        raise KeyError("parameter not in this flat buffer")


class FusedAdam:
    def __init__(self, flat: FlatParams, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 max_norm=None):
        """decoupled=True is the AdamW form.  max_norm: torch.nn.utils.clip_grad_norm_(params, max_norm) over the whole
        flat group fused in front of the update (reference src/ae/train_ae.py:121); the pre-clip norm of the last step
        is left in self.grad_norm (device)."""
        self.flat = flat
        self.max_norm = None if max_norm is None else float(max_norm)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.data.device)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.weight_decay, self.decoupled = float(weight_decay), bool(decoupled)
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=flat.data.device)
        self.grad_scale = 1.0
        self.bf16_copy = None

    def zero_grad(self, set_to_none=False):
        self.flat.grad.zero_()

    def step(self, stream=None):
        f = self.flat
        s = (stream if stream is not None else torch.cuda.current_stream(f.data.device)).cuda_stream
        bf = self.bf16_copy.data_ptr() if self.bf16_copy is not None else None
        with torch.cuda.device(f.data.device):
            if self.max_norm is not None:
                _native.call("mg_adam_step_clipped", f.data.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
                             self.exp_avg_sq.data_ptr(), f.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                             self.weight_decay, int(self.decoupled), float(self.grad_scale), self.max_norm,
                             self.grad_norm.data_ptr(), 0, self.step_dev.data_ptr(), bf, s)
                return
            _native.call("mg_adam_step", f.data.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
                         self.exp_avg_sq.data_ptr(), f.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                         self.weight_decay, int(self.decoupled), float(self.grad_scale), 0,
                         self.step_dev.data_ptr(), self.bf16_copy.data_ptr() if self.bf16_copy is not None else None,
                         s)

    def state_dict(self):
        """torch.optim-shaped: state[i] = {step, exp_avg, exp_avg_sq} per parameter, one param_group."""
        state = {}
        for i, (p, o) in enumerate(zip(self.flat.params, self.flat.offsets)):
            state[i] = {"step": self.step_dev.to(torch.float32).reshape(()).clone(),
                        "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "params": list(range(len(self.flat.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        for i, (p, o) in enumerate(zip(self.flat.params, self.flat.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
            self.step_dev.fill_(int(st["step"]))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
