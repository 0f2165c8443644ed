"""Stand-alone forms of the reference's MLP-type inner blocks, as autograd Functions over the C ABI.

On the hot path NoiseToLatent (src/gan/models.py:20-29), GeneratorDecoder.pre (models.py:46-51) and MLPClassifier
(src/emotion_discriminator/ed_model.py:74-101) run fused inside mg_generator_* / mg_emotion_*.  A caller that invokes the
inner module directly -- or builds the emotion discriminator with input_mode 'latent' (ed_model.py:128-136), which is the
MLPClassifier alone -- gets the same arithmetic as a chain of mg_linear_forward / mg_act_dropout_forward launches, float32,
with a full backward (input, weight and bias gradients).  CUDA tensors only; there is no CPU path.
"""
import torch
import torch.nn as nn

from . import _native

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU = 0, 1, 2, 3


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensors only (the B200 path has no CPU fallback)")


class _LinearFn(torch.autograd.Function):
    """z = x W^T + b"""

    @staticmethod
    def forward(ctx, x, W, b):
        _need_cuda(x, "linear")
        x2 = x.detach().to(torch.float32).reshape(-1, x.shape[-1]).contiguous()
        Wc = W.detach().to(torch.float32).contiguous()
        bc = b.detach().to(torch.float32).contiguous() if b is not None else None
        rows, K, N = x2.shape[0], x2.shape[1], Wc.shape[0]
        z = torch.empty((rows, N), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _native.call("mg_linear_forward", x2.data_ptr(), Wc.data_ptr(), bc.data_ptr() if bc is not None else None,
                         z.data_ptr(), rows, K, N, _stream(x))
        ctx.save_for_backward(x2, Wc)
        ctx.has_bias, ctx.x_shape = b is not None, x.shape
        return z.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dz):
        x2, Wc = ctx.saved_tensors
        rows, K, N = x2.shape[0], x2.shape[1], Wc.shape[0]
        dz2 = dz.detach().to(torch.float32).reshape(rows, N).contiguous()
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dW = torch.zeros_like(Wc) if ctx.needs_input_grad[1] else None
        db = torch.zeros(N, device=x2.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(x2.device):
            _native.call("mg_linear_backward", x2.data_ptr(), Wc.data_ptr(), dz2.data_ptr(), ptr(dx), ptr(dW), ptr(db),
                         rows, K, N, _stream(x2))
        return (dx.reshape(ctx.x_shape) if dx is not None else None), dW, db


class _ActDropoutFn(torch.autograd.Function):
    """h = act(z) * mask * scale   (mask None: no dropout)"""

    @staticmethod
    def forward(ctx, z, mask, scale, act):
        _need_cuda(z, "activation")
        zc = z.detach().to(torch.float32).contiguous()
        mc = mask.detach().to(torch.float32).contiguous() if mask is not None else None
        h = torch.empty_like(zc)
        with torch.cuda.device(z.device):
            _native.call("mg_act_dropout_forward", zc.data_ptr(), mc.data_ptr() if mc is not None else None, float(scale),
                         int(act), h.data_ptr(), zc.numel(), _stream(z))
        ctx.save_for_backward(zc, mc)
        ctx.scale, ctx.act = float(scale), int(act)
        return h

    @staticmethod
    def backward(ctx, dh):
        zc, mc = ctx.saved_tensors
        dhc = dh.detach().to(torch.float32).contiguous()
        dz = torch.empty_like(zc)
        with torch.cuda.device(zc.device):
            _native.call("mg_act_dropout_backward", dhc.data_ptr(), zc.data_ptr(), mc.data_ptr() if mc is not None else None,
                         ctx.scale, ctx.act, dz.data_ptr(), zc.numel(), _stream(zc))
        return dz, None, None, None


def effective_weight(m, x):
    """The weight a layer would use in its own forward.  torch.nn.utils.spectral_norm (what the reference applies when
    use_spectral_norm is set, ed_model.py:29-33,81-86) recomputes `weight` = weight_orig / sigma in a forward PRE-HOOK (one
    power iteration in train mode); the native operators never call the layer, so its pre-hooks are run here.  The returned
    tensor carries the autograd graph back to weight_orig."""
    for hook in m._forward_pre_hooks.values():
        hook(m, (x,))
    return m.weight


def linear(x, W, b=None):
    return _LinearFn.apply(x, W, b)


def act_dropout(z, act, mask=None, scale=1.0):
    return _ActDropoutFn.apply(z, mask, scale, act)


_ACT_OF = {nn.ReLU: ACT_RELU, nn.GELU: ACT_GELU, nn.LeakyReLU: ACT_LRELU}


def run_mlp(seq, x, training, masks=None):
    """Runs an nn.Sequential of Linear / ReLU / GELU / LeakyReLU(0.2) / Dropout children (the reference's MLP blocks) on the
    native operators.  An activation and the nn.Dropout that follows it are ONE launch.  masks: optional list of injected
    0/1 keep-masks, one per nn.Dropout (tests); default: torch.bernoulli draws, as nn.Dropout would make them."""
    mods = list(seq)
    i, drop_i = 0, 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            x = linear(x, effective_weight(m, x), m.bias)
            i += 1
            continue
        act = ACT_NONE
        if type(m) in _ACT_OF:
            if isinstance(m, nn.LeakyReLU) and abs(m.negative_slope - 0.2) > 1e-12:
                raise NotImplementedError("LeakyReLU slope other than 0.2")
            if isinstance(m, nn.GELU) and getattr(m, "approximate", "none") != "none":
                raise NotImplementedError("tanh-approximated GELU")
            act = _ACT_OF[type(m)]
            i += 1
            m = mods[i] if i < len(mods) else None
        mask, scale = None, 1.0
        if isinstance(m, nn.Dropout):
            if training and m.p > 0.0:
                keep = 1.0 - m.p
                mask = masks[drop_i] if masks is not None else torch.bernoulli(torch.full_like(x, keep))
                scale = 1.0 / keep
            drop_i += 1
            i += 1
        elif act == ACT_NONE:
            raise NotImplementedError(f"no native operator for {type(m).__name__} in a stand-alone MLP block")
        if act != ACT_NONE or mask is not None:
            x = act_dropout(x, act, mask, scale)
    return x


# ---------------------------------------------------------------------------------------------------------------------
# conv-type inner blocks (include/melogan_b200.h: mg_convunit_*): Conv1d / ConvTranspose1d [+ BatchNorm1d] + activation
# ---------------------------------------------------------------------------------------------------------------------
UNIT_CONV, UNIT_CONVT = 0, 1
UACT_NONE, UACT_RELU, UACT_GELU, UACT_TANH = 0, 1, 2, 3
_units = {}


def _unit(device):
    import ctypes
    key = torch.device(device).index
    if key not in _units:
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _native.call("mg_convunit_create", ctypes.byref(h))
        _units[key] = h
    return _units[key]


class _ConvUnitFn(torch.autograd.Function):
    """y = act(BN(conv(x))) on channels-last x (R, Lin, Cin); conv: nn.Conv1d or nn.ConvTranspose1d; bn: nn.BatchNorm1d or None"""

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, conv, bn, act):
        _need_cuda(x, "conv unit")
        transposed = isinstance(conv, nn.ConvTranspose1d)
        ks, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if conv.dilation[0] != 1 or conv.groups != 1:
            raise NotImplementedError("conv unit: dilation 1, groups 1")
        if transposed and (ks, stride, pad, conv.output_padding[0]) != (5, 2, 2, 1):
            raise NotImplementedError("conv unit: ConvTranspose1d k5 s2 p2 output_padding 1")
        if not transposed and not ((stride == 1 and 2 * pad == ks - 1) or (ks, stride, pad) == (5, 2, 2)):
            raise NotImplementedError("conv unit: Conv1d stride 1 'same' padding, or k5 s2 p2")
        xc = x.detach().to(torch.float32).contiguous()
        R, Lin, Cin = xc.shape
        Cout = conv.out_channels
        Lout = 2 * Lin if transposed else Lin // stride
        Wc = W.detach().to(torch.float32).contiguous()
        dev = x.device
        z = torch.empty((R, Lout, Cout), device=dev)
        y = torch.empty_like(z)
        gd = torch.empty_like(z) if act == UACT_GELU else None
        mean = torch.empty(Cout, device=dev) if bn is not None else None
        invstd = torch.empty(Cout, device=dev) if bn is not None else None
        train = bool(bn is not None and bn.training)
        if bn is not None and (not bn.track_running_stats or bn.momentum is None or abs(bn.momentum - 0.1) > 1e-12
                               or abs(bn.eps - 1e-5) > 1e-12):
            raise NotImplementedError("conv unit: BatchNorm1d(eps=1e-5, momentum=0.1, track_running_stats=True)")
        p = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            _native.call("mg_convunit_forward", _unit(dev), UNIT_CONVT if transposed else UNIT_CONV, xc.data_ptr(), Wc.data_ptr(),
                         p(b.detach().contiguous() if b is not None else None), R, Lin, Cin, Cout, ks, stride, pad,
                         p(gamma.detach() if gamma is not None else None), p(beta.detach() if beta is not None else None),
                         p(bn.running_mean if bn is not None else None), p(bn.running_var if bn is not None else None),
                         int(train), int(act), z.data_ptr(), y.data_ptr(), p(gd), p(mean), p(invstd), _stream(x))
        if train:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(xc, Wc, z, y, gd, mean, invstd, gamma.detach() if gamma is not None else None)
        ctx.meta = (transposed, ks, stride, pad, int(act), train, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, Wc, z, y, gd, mean, invstd, gamma = ctx.saved_tensors
        transposed, ks, stride, pad, act, train, has_bias = ctx.meta
        R, Lin, Cin = xc.shape
        Cout = y.shape[2]
        dev = xc.device
        dyc = dy.detach().to(torch.float32).contiguous()
        s0, s1 = torch.empty_like(y), torch.empty_like(y)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        dW = torch.zeros_like(Wc)
        db = torch.zeros(Cout, device=dev) if has_bias else None
        dgamma = torch.zeros(Cout, device=dev) if gamma is not None else None
        dbeta = torch.zeros(Cout, device=dev) if gamma is not None else None
        p = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            _native.call("mg_convunit_backward", _unit(dev), UNIT_CONVT if transposed else UNIT_CONV, xc.data_ptr(), Wc.data_ptr(),
                         R, Lin, Cin, Cout, ks, stride, pad, p(gamma), int(train), act, z.data_ptr(), y.data_ptr(), p(gd),
                         p(mean), p(invstd), dyc.data_ptr(), s0.data_ptr(), s1.data_ptr(), p(dx), dW.data_ptr(), p(db),
                         p(dgamma), p(dbeta), _stream(xc))
        return dx, dW, db, dgamma, dbeta, None, None, None


def conv_unit(x_cl, conv, bn=None, act=UACT_NONE):
    """Channels-last (R, L, C) in and out."""
    return _ConvUnitFn.apply(x_cl, effective_weight(conv, x_cl), conv.bias, bn.weight if bn is not None else None,
                             bn.bias if bn is not None else None, conv, bn, act)


_UACT_OF = {nn.ReLU: UACT_RELU, nn.GELU: UACT_GELU, nn.Tanh: UACT_TANH}


def run_conv_stack(seq, x_cl):
    """Runs an nn.Sequential of [Conv1d | ConvTranspose1d] [BatchNorm1d] [ReLU | GELU | Tanh] groups (the reference's conv
    stacks) on channels-last activations (R, L, C); nested nn.Sequential / ConvBlock1D-style children (a module with a
    `.net` Sequential) are flattened."""
    mods = []

    def flat(m):
        for c in m:
            inner = getattr(c, "net", None)
            if isinstance(c, nn.Sequential):
                flat(c)
            elif isinstance(inner, nn.Sequential):
                flat(inner)
            else:
                mods.append(c)
    flat(seq)
    i = 0
    while i < len(mods):
        conv = mods[i]
        if not isinstance(conv, (nn.Conv1d, nn.ConvTranspose1d)):
            raise NotImplementedError(f"no native operator for {type(conv).__name__} in a stand-alone conv stack")
        i += 1
        bn = None
        if i < len(mods) and isinstance(mods[i], nn.BatchNorm1d):
            bn = mods[i]
            i += 1
        act = UACT_NONE
        if i < len(mods) and type(mods[i]) in _UACT_OF:
            if isinstance(mods[i], nn.GELU) and getattr(mods[i], "approximate", "none") != "none":
                raise NotImplementedError("tanh-approximated GELU")
            act = _UACT_OF[type(mods[i])]
            i += 1
        x_cl = conv_unit(x_cl, conv, bn, act)
    return x_cl
