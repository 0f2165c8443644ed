"""Arithmetic of the fp32-on-tensor-cores path (csrc/gemm_tc.cuh: split3, kSplitA / kSplitW, try_split_tapgemm) restated
on the CPU: the three-way bf16 split of a float32 is exact, the six part products kept reproduce the float32 product to
O(2^-24), and the segment orders the kernels use pair the parts as (m,m) (h,l) (l,h) (h,m) (m,h) (h,h).  Host logic only;
the kernels themselves are held to float64 on the GPU by tests/test_fp32_tc_gpu.py."""
import re
from pathlib import Path

import torch

SRC = (Path(__file__).resolve().parents[1] / "melo-gan_b200" / "csrc" / "gemm_tc.cuh").read_text()


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def split3(x):
    h = _bf16(x)
    r = x - h
    m = _bf16(r)
    return h, m, _bf16(r - m)


def _orders():
    a = int(re.search(r"kSplitA = (0x[0-9a-fA-F]+)u", SRC).group(1), 16)
    w = int(re.search(r"kSplitW = (0x[0-9a-fA-F]+)u", SRC).group(1), 16)
    return [(a >> (4 * s)) & 3 for s in range(6)], [(w >> (4 * s)) & 3 for s in range(6)]


def test_three_way_split_is_exact():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 18, generator=g) * torch.exp2(torch.randint(-20, 20, (1 << 18,), generator=g).float())
    # (values within 2^16 of the smallest normal number have subnormal residuals, which some converters flush: not a regime any
    # activation or weight lives in, and the loss is then below 2^-126 absolute)
    x = torch.cat([x, torch.tensor([0.0, -0.0, 1.0, -1.0, 3.0e38, -3.0e38, 1.0e-30, 1.0 + 2.0 ** -23, 1.0 - 2.0 ** -24])])
    h, m, l = split3(x)
    assert torch.equal(h.double() + m.double() + l.double(), x.double())
    # every part is on the bf16 grid (the conversion to the tensor core's operand type is the identity)
    for p in (h, m, l):
        assert torch.equal(_bf16(p), p)


def test_segment_orders_pair_the_six_kept_products_smallest_first():
    a, w = _orders()
    pairs = list(zip(a, w))
    assert sorted(pairs) == sorted([(0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1)])   # hh hm mh hl lh mm; ml lm ll dropped
    assert pairs[-1] == (0, 0)                      # the full-size term comes last (accumulator truncation)
    assert all(sum(p) >= 1 for p in pairs[:-1])


def test_six_products_reproduce_the_float32_product():
    g = torch.Generator().manual_seed(1)
    x, w = torch.randn(1 << 16, generator=g), torch.randn(1 << 16, generator=g)
    a, b = _orders()
    px, pw = split3(x), split3(w)
    got = sum(px[i].double() * pw[j].double() for i, j in zip(a, b))
    exact = x.double() * w.double()
    rel = ((got - exact).abs() / exact.abs().clamp_min(1e-300)).max().item()
    assert rel < 2.0 ** -23, rel                   # dropped: ml + lm + ll <= 2 * 2^-9 * 2^-17 + 2^-34 of |x w|
    # and every kept product is exact in a float32 accumulator (16 significand bits)
    for i, j in zip(a, b):
        p = px[i] * pw[j]
        assert torch.equal(p.double(), px[i].double() * pw[j].double())


def test_k_concatenated_contraction_matches_float64():
    """The layout the tap-GEMM sees: A' = six k-segments per row, W' = the matching six; one long dot product."""
    g = torch.Generator().manual_seed(2)
    A, W = torch.randn(64, 128, generator=g), torch.randn(32, 128, generator=g)
    a, b = _orders()
    pa, pw = split3(A), split3(W)
    A6 = torch.cat([pa[i] for i in a], dim=1).double()
    W6 = torch.cat([pw[j] for j in b], dim=1).double()
    ref = A.double() @ W.double().t()
    err = ((A6 @ W6.t() - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-7, err
