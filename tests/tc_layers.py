"""Per-layer harness shared by tests/test_tc_layers_gpu.py and scripts/bench_layers.py.

Every contraction the bf16-mode training cycle launches at per-GPU batch B (the `MELOGAN_TRACE=1` list of the
benched configuration: critic conv.2/conv.4 forward, dgrad, adjoint and wgrad over 3B rows; generator pre.2,
deconv.0/3 forward, dgrad, wgrad; emotion-discriminator conv.1-3 forward with folded BatchNorm + GELU and their
dgrads; the TF32 Linears) is described here ONCE as a `mg_debug_layer` plus the float64 torch expression of the
same contraction.  Operands are drawn on the bf16 grid, so the tensor-core result differs from float64 only by
fp32 accumulation order (and one bf16 rounding where the layer stores bf16).
Reference call sites: src/gan/models.py:46-83,140-169, src/emotion_discriminator/ed_model.py:35-69.
"""
import ctypes
import math

import torch
import torch.nn.functional as F

from melogan import _native

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU = 0, 1, 2, 3
MUL_NONE, MUL_LRELU_SIGN, MUL_RELU_SIGN, MUL_VALUE = 0, 1, 2, 3


class DebugLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("op", "in_bf16", "out_bf16", "mask_bf16", "tf32", "R", "Lin", "Cin", "Cout", "ks", "stride", "pad", "act",
                 "mul_mode", "accumulate", "w_nstride", "w_kstride", "n_perm_q", "n_perm_p")] + \
               [(n, ctypes.c_void_p) for n in ("inp", "in2", "out", "aux", "mul_src", "W", "bias", "col_scale", "dW")] + \
               [("pool_out", ctypes.c_void_p), ("pool_scale", ctypes.c_float), ("pool_done", ctypes.c_void_p),
                ("colsum_out", ctypes.c_void_p), ("colsum_samples", ctypes.c_int), ("colsum_done", ctypes.c_void_p),
                ("stats_out", ctypes.c_void_p), ("stats_done", ctypes.c_void_p)]


def _lib():
    L = _native.lib()
    L.mg_debug_layer_run.argtypes = [ctypes.POINTER(DebugLayer), ctypes.c_void_p]
    L.mg_debug_layer_run.restype = ctypes.c_int
    L.mg_debug_set.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.mg_debug_set.restype = ctypes.c_int
    L.mg_debug_last_launch.argtypes = []
    L.mg_debug_last_launch.restype = ctypes.c_char_p
    return L


def debug_set(key, value):
    _native.check(_lib().mg_debug_set(key.encode(), int(value)))


def last_launch():
    return _lib().mg_debug_last_launch().decode()


def cycle_layers(B):
    """The tensor-core contractions of one training cycle at per-GPU batch B (bf16 mode), by name."""
    R, L0 = 3 * B, 64
    S = []

    def add(name, count, **kw):
        d = dict(op=0, in_bf16=1, out_bf16=1, mask_bf16=1, tf32=0, ks=1, stride=1, pad=0, act=ACT_NONE, mul=MUL_NONE,
                 bias=False, scale=False, aux=False, w_nstride=-1, w_kstride=-1, perm=(0, 0), inplace_mask=False)
        d.update(kw)
        d["name"], d["count"] = name, count
        S.append(d)

    # ---- critic over 3B rows (real | fake | x_hat): src/gan/models.py:140-169 and their autograd ----
    add("D.conv2.fwd", 5, op=0, R=R, Lin=4 * L0, Cin=64, Cout=128, ks=5, stride=2, pad=2, act=ACT_LRELU, bias=True)
    add("D.conv4.fwd", 5, op=0, R=R, Lin=2 * L0, Cin=128, Cout=256, ks=5, stride=2, pad=2, act=ACT_LRELU, bias=True)
    add("D.conv4.dgrad", 5, op=2, R=R, Lin=L0, Cin=256, Cout=128, w_nstride=5, w_kstride=128 * 5, mul=MUL_LRELU_SIGN)
    add("D.conv2.dgrad", 5, op=2, R=R, Lin=2 * L0, Cin=128, Cout=64, w_nstride=5, w_kstride=64 * 5, mul=MUL_LRELU_SIGN)
    add("D.conv2.adjoint", 5, op=0, R=B, Lin=4 * L0, Cin=64, Cout=128, ks=5, stride=2, pad=2, mul=MUL_LRELU_SIGN,
        inplace_mask=True)
    add("D.conv4.adjoint", 5, op=0, R=B, Lin=2 * L0, Cin=128, Cout=256, ks=5, stride=2, pad=2, mul=MUL_LRELU_SIGN,
        inplace_mask=True)
    add("D.conv4.wgrad", 5, op=5, R=R, Lin=2 * L0, Cin=128, Cout=256, ks=5, stride=2, pad=2)
    add("D.conv2.wgrad", 5, op=5, R=R, Lin=4 * L0, Cin=64, Cout=128, ks=5, stride=2, pad=2)
    # generator-step critic pass over B rows
    add("D.conv4.fwd.B", 1, op=0, R=B, Lin=2 * L0, Cin=128, Cout=256, ks=5, stride=2, pad=2, act=ACT_LRELU, bias=True)
    add("D.conv4.dgrad.B", 1, op=2, R=B, Lin=L0, Cin=256, Cout=128, w_nstride=5, w_kstride=128 * 5, mul=MUL_LRELU_SIGN)
    # ---- generator: src/gan/models.py:46-83 ----
    add("G.pre2.fwd", 6, op=3, R=B, Cin=512, Cout=256 * L0, act=ACT_RELU, bias=True, perm=(256, L0))
    add("G.deconv0.fwd", 6, op=2, R=B, Lin=L0, Cin=256, Cout=128, w_nstride=5, w_kstride=128 * 5, bias=True, out_bf16=0)
    add("G.deconv3.fwd", 6, op=2, R=B, Lin=2 * L0, Cin=128, Cout=64, w_nstride=5, w_kstride=64 * 5, bias=True, out_bf16=0)
    add("G.deconv3.dgrad", 1, op=0, R=B, Lin=4 * L0, Cin=64, Cout=128, ks=5, stride=2, pad=2, w_nstride=64 * 5, w_kstride=5,
        mul=MUL_RELU_SIGN, out_bf16=0)
    add("G.deconv0.dgrad", 1, op=0, R=B, Lin=2 * L0, Cin=128, Cout=256, ks=5, stride=2, pad=2, w_nstride=128 * 5, w_kstride=5,
        mul=MUL_RELU_SIGN)
    add("G.deconv3.wgrad", 1, op=6, R=B, Lin=2 * L0, Cin=128, Cout=64)
    add("G.deconv0.wgrad", 1, op=6, R=B, Lin=L0, Cin=256, Cout=128)
    add("G.pre2.wgrad", 1, op=7, R=B, Cin=512, Cout=256 * L0, perm=(256, L0))
    add("G.pre2.dgrad", 1, op=4, R=B, Cin=512, Cout=256 * L0, mul=MUL_RELU_SIGN, out_bf16=0, perm=(256, L0))
    # ---- frozen emotion discriminator (eval BatchNorm folded into scale/shift): ed_model.py:35-69 ----
    for i, (ci, co) in enumerate(((64, 128), (128, 256), (256, 256)), start=1):
        add(f"ED.conv{i}.fwd", 1, op=0, R=B, Lin=8 * L0, Cin=ci, Cout=co, ks=3, stride=1, pad=1, act=ACT_GELU, bias=True,
            scale=True, aux=True)
        add(f"ED.conv{i}.dgrad", 1, op=1, R=B, Lin=8 * L0, Cin=ci, Cout=co, ks=3, pad=1, scale=True, mul=MUL_VALUE)
    # ---- float32 Linears on kind::tf32 (critic fc.1 over 3B rows, generator MLPs) ----
    add("D.fc.fwd.tf32", 5, op=3, R=R, Cin=256, Cout=256, act=ACT_LRELU, bias=True, in_bf16=0, out_bf16=0, mask_bf16=0, tf32=1)
    add("D.fc.dgrad.tf32", 5, op=4, R=R, Cin=256, Cout=256, in_bf16=0, out_bf16=0, mask_bf16=0, tf32=1)
    add("G.n2l0.fwd.tf32", 6, op=3, R=B, Cin=256, Cout=512, act=ACT_RELU, bias=True, in_bf16=0, out_bf16=0, mask_bf16=0, tf32=1)
    add("G.pre0.fwd.tf32", 6, op=3, R=B, Cin=64, Cout=512, act=ACT_RELU, bias=True, in_bf16=0, out_bf16=1, mask_bf16=1, tf32=1)
    # ---- their weight gradients: float32 operands cast to bf16 copies, reduction over the rows on tc_wgrad_kernel ----
    add("D.fc.wgrad.f32", 5, op=7, R=R, Cin=256, Cout=256, in_bf16=0, out_bf16=0, mask_bf16=0, tf32=1)
    add("G.n2l0.wgrad.f32", 1, op=7, R=B, Cin=256, Cout=512, in_bf16=0, out_bf16=0, mask_bf16=0, tf32=1)
    return S


_FULL_FP32 = False      # Layer(spec with full_fp32=True): operands keep all 24 significand bits (the fp32_tc path)


def _grid(shape, gen, dev, scale=1.0, bf16=True):
    """Random values that are exactly representable in bf16 (so that bf16 / TF32 operand rounding is the identity)."""
    x = torch.randn(shape, generator=gen, device=dev) * scale
    if _FULL_FP32:
        assert not bf16
        return x
    x = x.to(torch.bfloat16)
    return x if bf16 else x.float()


def fp32_layers(B):
    """The same contractions as cycle_layers(B) in float32 storage with full-precision operands: what the fp32 parity mode
    launches, for the six-term bf16 tensor-core form (mg_debug_set("fp32_tc", 1), csrc/gemm_tc.cuh try_split_*)."""
    out = []
    for s in cycle_layers(B):
        if s["tf32"] and s["name"] not in ("D.fc.fwd.tf32", "D.fc.dgrad.tf32", "D.fc.wgrad.f32"):
            continue
        d = dict(s)
        d.update(in_bf16=0, out_bf16=0, mask_bf16=0, tf32=0, full_fp32=True, name=s["name"].replace(".tf32", "").replace(".f32", "") + ".fp32")
        out.append(d)
    return out


class Layer:
    """Device tensors of one spec + run() through the C ABI + reference() in float64."""

    def __init__(self, spec, seed=0, dev="cuda"):
        self.s = s = dict(spec)
        global _FULL_FP32
        _FULL_FP32 = bool(s.get("full_fp32"))
        try:
            self._build(s, seed, dev)
        finally:
            _FULL_FP32 = False

    def _build(self, s, seed, dev):
        g = torch.Generator(device=dev).manual_seed(seed)
        op, R, Cin, Cout = s["op"], s["R"], s["Cin"], s["Cout"]
        Lin = s.get("Lin", 1)
        ib, ob = bool(s["in_bf16"]), bool(s["out_bf16"])
        odt = torch.bfloat16 if ob else torch.float32
        self.dW = None
        wscale = 1.0 / math.sqrt(Cin * max(s["ks"], 1))
        if op == 0:      # conv forward
            Lout = Lin // s["stride"]
            self.x = _grid((R, Lin, Cin), g, dev, bf16=ib)
            # [n][k][t]; the dgrad of a ConvTranspose1d passes the same layout explicitly (its weight is [C_in_T = n][C_out_T = k][t])
            self.W = _grid((Cout, Cin, s["ks"]), g, dev, wscale, bf16=False)
            oshape = (R, Lout, Cout)
        elif op == 1:    # stride-1 conv dgrad: in = dOut [R, L, Cout] -> dIn [R, L, Cin]
            self.x = _grid((R, Lin, Cout), g, dev, bf16=ib)
            self.W = _grid((Cout, Cin, s["ks"]), g, dev, 1.0 / math.sqrt(Cout * s["ks"]), bf16=False)
            oshape = (R, Lin, Cin)
        elif op == 2:    # k5 s2 up-sampling: W[t + n*5 + k*Cout*5] = [k][n][t]
            self.x = _grid((R, Lin, Cin), g, dev, bf16=ib)
            self.W = _grid((Cin, Cout, 5), g, dev, 1.0 / math.sqrt(Cin * 2.5), bf16=False)
            oshape = (R, 2 * Lin, Cout)
        elif op == 3:
            self.x = _grid((R, Cin), g, dev, bf16=ib)
            self.W = _grid((Cout, Cin), g, dev, wscale, bf16=False)
            oshape = (R, Cout)
        elif op == 4:    # linear dgrad: in = dZ [R, Cout] -> dX [R, Cin]
            self.x = _grid((R, Cout), g, dev, bf16=ib)
            self.W = _grid((Cout, Cin), g, dev, 1.0 / math.sqrt(Cout), bf16=False)
            oshape = (R, Cin)
        elif op == 5:    # conv wgrad: in = dOut [R, Lout, Cout], in2 = x [R, Lin, Cin]
            Lout = Lin // s["stride"]
            self.x = _grid((R, Lout, Cout), g, dev, bf16=ib)
            self.x2 = _grid((R, Lin, Cin), g, dev, bf16=ob)
            self.dW = torch.zeros((Cout, Cin, s["ks"]), device=dev)
            oshape = None
        elif op == 6:    # ConvTranspose1d wgrad: in = x [R, Lin, Cin], in2 = dOut [R, 2 Lin, Cout]
            self.x = _grid((R, Lin, Cin), g, dev, bf16=ib)
            self.x2 = _grid((R, 2 * Lin, Cout), g, dev, bf16=ob)
            self.dW = torch.zeros((Cin, Cout, 5), device=dev)
            oshape = None
        else:            # linear wgrad: in = dZ [R, Cout], in2 = A [R, Cin]
            self.x = _grid((R, Cout), g, dev, bf16=ib)
            self.x2 = _grid((R, Cin), g, dev, bf16=ob)
            self.dW = torch.zeros((Cout, Cin), device=dev)
            oshape = None
        self.oshape = oshape
        nout = oshape[-1] if oshape else 0
        self.bias = _grid((nout,), g, dev, 0.5, bf16=False) if s["bias"] else None
        self.scale = (_grid((nout,), g, dev, 0.25, bf16=False).abs() + 0.5) if s["scale"] else None
        self.out = self.aux = self.mask = None
        if oshape:
            self.out = torch.empty(oshape, dtype=odt, device=dev)
            if s["aux"]:
                self.aux = torch.empty(oshape, dtype=odt, device=dev)
            if s["mul"] != MUL_NONE:
                mdt = torch.bfloat16 if s["mask_bf16"] else torch.float32
                self.mask = _grid(oshape, g, dev, bf16=not _FULL_FP32).to(mdt)
                if s["inplace_mask"]:          # the adjoint chain overwrites the saved activation it masks with
                    self.out = self.mask.clone().to(odt)
                    self.mask0 = self.mask
                    self.mask = self.out

    def desc(self):
        s = self.s
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        d = DebugLayer()
        d.op, d.in_bf16, d.mask_bf16, d.tf32 = s["op"], s["in_bf16"], s["mask_bf16"], s["tf32"]
        d.out_bf16 = s["out_bf16"]
        d.R, d.Lin, d.Cin, d.Cout, d.ks, d.stride, d.pad = s["R"], s.get("Lin", 1), s["Cin"], s["Cout"], s["ks"], s["stride"], s["pad"]
        d.act, d.mul_mode, d.accumulate = s["act"], s["mul"], 0
        d.w_nstride, d.w_kstride = s["w_nstride"], s["w_kstride"]
        d.n_perm_q, d.n_perm_p = s["perm"]
        d.inp, d.in2, d.out, d.aux, d.mul_src = p(self.x), p(getattr(self, "x2", None)), p(self.out), p(self.aux), p(self.mask)
        d.W, d.bias, d.col_scale, d.dW = p(getattr(self, "W", None)), p(self.bias), p(self.scale), p(self.dW)
        if getattr(self, "pool", None) is not None:      # fused AdaptiveAvgPool1d(1) of conv forward layers (op 0)
            d.pool_out, d.pool_scale = p(self.pool), 1.0 / float(self.oshape[1])
            d.pool_done = ctypes.c_void_p(self.pool_done.ctypes.data)
        if getattr(self, "colsum", None) is not None:    # fused column sums of up-sampling dgrads (op 2)
            d.colsum_out, d.colsum_samples = p(self.colsum), int(self.colsum_samples)
            d.colsum_done = ctypes.c_void_p(self.colsum_done.ctypes.data)
        if getattr(self, "stats", None) is not None:     # fused BatchNorm statistics of up-sampling layers with float32 out
            d.stats_out = p(self.stats)
            d.stats_done = ctypes.c_void_p(self.stats_done.ctypes.data)
        return d

    def enable_stats(self):
        import numpy as np
        self.stats = torch.zeros(2 * self.oshape[2], device=self.out.device)
        self.stats_done = np.zeros(1, dtype=np.int32)

    def enable_colsum(self, samples):
        """Asks the layer (op 2, bf16 out) to add the column sums of its output over the first `samples` samples to
        self.colsum; self.colsum_done[0] tells whether the kernels that ran did."""
        import numpy as np
        self.colsum = torch.zeros(self.oshape[2], device=self.out.device)
        self.colsum_samples = samples
        self.colsum_done = np.zeros(1, dtype=np.int32)

    def enable_pool(self):
        """Asks the layer (op 0, bf16 out) for the mean over positions as a second output; self.pool_done[0] tells whether
        the kernel that ran produced it."""
        import numpy as np
        self.pool = torch.full((self.oshape[0], self.oshape[2]), float("nan"), device=self.out.device)
        self.pool_done = np.zeros(1, dtype=np.int32)

    def run(self):
        if self.s["inplace_mask"]:
            self.out.copy_(self.mask0)
        if self.dW is not None:
            self.dW.zero_()
        d = self.desc()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(_lib().mg_debug_layer_run(ctypes.byref(d), st))
        return last_launch()

    # ---- float64 reference of the same contraction, for samples [r0, r1) ----
    def reference(self, r0=0, r1=None):
        s = self.s
        op, R = s["op"], s["R"]
        r1 = R if r1 is None else r1
        x = self.x[r0:r1].double()
        q, p_ = s["perm"]
        if op == 0:
            y = F.conv1d(x.permute(0, 2, 1), self.W.double(), stride=s["stride"], padding=s["pad"]).permute(0, 2, 1)
        elif op == 1:
            y = F.conv_transpose1d(x.permute(0, 2, 1), self.W.double(), stride=1, padding=s["pad"]).permute(0, 2, 1)
        elif op == 2:
            y = F.conv_transpose1d(x.permute(0, 2, 1), self.W.double(), stride=2, padding=2, output_padding=1).permute(0, 2, 1)
        elif op == 3:
            y = x @ self.W.double().t()                       # physical columns
            if q:                                             # logical column n <-> physical (n % q) * p + n // q
                n = torch.arange(y.shape[1], device=y.device)
                y = y[:, (n % q) * p_ + n // q]
        elif op == 4:
            W = self.W.double()
            if q:                                             # the reduction index is permuted the same way
                n = torch.arange(W.shape[0], device=W.device)
                W = W[(n % q) * p_ + n // q]
            y = x @ W
        else:
            raise ValueError(op)
        if self.scale is not None:
            y = y * self.scale.double()
        if self.bias is not None:
            b = self.bias.double()
            if op == 3 and q:
                n = torch.arange(b.shape[0], device=b.device)
                b = b[(n % q) * p_ + n // q]
            y = y + b
        aux = None
        if s["act"] == ACT_RELU:
            y = y.clamp_min(0)
        elif s["act"] == ACT_LRELU:
            y = torch.where(y > 0, y, 0.2 * y)
        elif s["act"] == ACT_GELU:
            cdf = 0.5 * (1 + torch.erf(y / math.sqrt(2.0)))
            aux = cdf + y * torch.exp(-0.5 * y * y) / math.sqrt(2 * math.pi)
            y = y * cdf
        if s["mul"] != MUL_NONE:
            m = (self.mask0 if s["inplace_mask"] else self.mask)[r0:r1].double()
            if s["mul"] == MUL_LRELU_SIGN:
                y = y * torch.where(m > 0, 1.0, 0.2)
            elif s["mul"] == MUL_RELU_SIGN:
                y = y * (m > 0)
            else:
                y = y * m
        return y, aux

    def reference_wgrad(self, chunk=1024):
        s = self.s
        op, R = s["op"], s["R"]
        acc = torch.zeros_like(self.dW, dtype=torch.float64)
        for r0 in range(0, R, chunk):
            g, a = self.x[r0:r0 + chunk].double(), self.x2[r0:r0 + chunk].double()
            if op == 5:      # dW[co][ci][t] = sum dOut[r, l, co] x[r, stride*l + t - pad, ci]
                acc += torch.nn.grad.conv1d_weight(a.permute(0, 2, 1), acc.shape, g.permute(0, 2, 1), stride=s["stride"],
                                                   padding=s["pad"])
            elif op == 6:    # dW[ci][co][t] = sum x[r, i, ci] dOut[r, 2i + t - 2, co]: conv wgrad with roles swapped
                acc += torch.nn.grad.conv1d_weight(a.permute(0, 2, 1), (s["Cin"], s["Cout"], 5), g.permute(0, 2, 1), stride=2,
                                                   padding=2)
            else:
                acc += g.t() @ a
        q, p_ = s["perm"]
        if op == 7 and q:        # gradient row of logical column n lives at the physical weight row (n % q) * p + n // q
            n = torch.arange(acc.shape[0], device=acc.device)
            out = torch.empty_like(acc)
            out[(n % q) * p_ + n // q] = acc
            acc = out
        return acc


def compare(got, ref, stored_bf16, abs_allow=4e-6):
    """(max |diff| / max |ref|, worst error in bf16 ulps of the reference element).  The ulp figure first forgives
    abs_allow * max|ref| of absolute error: fp32 accumulation noise is relative to the SUM's terms, not to a
    result that cancelled to nearly zero."""
    got, ref = got.double(), ref.double()
    scale = ref.abs().max().clamp_min(1e-30)
    d = (got - ref).abs()
    rel = (d.max() / scale).item()
    if not stored_bf16:
        return rel, 0.0
    mag = ref.abs().clamp_min(scale * 2.0 ** -40)
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)
    return rel, ((d - abs_allow * scale).clamp_min(0) / ulp).max().item()


def probe_time(fn, iters=5, family=3):
    """Mean device time (ms) of the tensor-core kernel(s) inside fn(), from the library's per-launch CUDA events."""
    L = _native.lib()
    fn()
    torch.cuda.synchronize()
    out = (ctypes.c_double * 4)()
    L.mg_probe_begin(family)
    for _ in range(iters):
        fn()
    L.mg_probe_end(out)
    return out[1] / iters, out[0] / iters
