"""Whole-step check of the reductions that moved into the tensor-core epilogues (pooling ws_pool_*, bias-gradient column sums
ws_colsum_*, BatchNorm statistics ws_stats_*): one critic step and one generator step of the bf16 engine at a batch where
the weight-stationary / CTA-pair kernels carry the layers (B = 1024: the same kernel variants as the bench), with the
fusions ON (product) and OFF (`no_fuse`: the separate pool_rows / colreduce passes of round 1), same parameters and inputs.
Both runs execute the same bf16 contractions, so losses, every parameter gradient and the BatchNorm running statistics must
agree to float32 summation order.  This pins the plumbing in csrc/gan.cu (row limits of the bias sums, the flags that skip
the fallback passes, the zeroing of the statistics buffers) that the per-layer tests do not see."""
import pytest
import torch

import tc_layers as TL
from gan_testlib import cuda_batch, make_engine, rel_l2
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu
B = 1024


def _run(no_fuse):
    TL.debug_set("reset", 0)
    TL.debug_set("no_fuse", no_fuse)
    try:
        params = O.make_params(3, fan_in_scale=True)
        batch = O.make_batch(77, B)
        eng, cp, grads = make_engine(B, params, precision="bf16")
        cb = cuda_batch(batch)
        md = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"]).clone()
        gD = {k: v.clone() for k, v in grads["D"].items()}
        # the three conv bias gradients are column sums of the stored dz tensors over the real + fake rows (first 2B samples)
        for key, buf, C in (("conv.4.bias", "d.dz3", 256), ("conv.2.bias", "d.dz2", 128), ("conv.0.bias", "d.dz1", 64)):
            dz = eng.buffer(buf, torch.bfloat16).view(3 * B, -1, C)[:2 * B].double()
            want = dz.sum(dim=(0, 1))
            err = ((gD[key].double() - want).abs().max() / want.abs().max()).item()
            assert err < 1e-4, f"{key}: bias gradient vs column sums of {buf}: {err:.2e} (no_fuse={no_fuse})"
        mg = eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"]).clone()
        gG = {k: v.clone() for k, v in grads["G"].items()}
        gE = {k: v.clone() for k, v in grads["E"].items()}
        bn = {k: cp["G"][k].clone() for k in cp["G"] if "running" in k}
        logits = eng.buffer("ed.logits")[:B * 4].clone()
        notes = eng.buffer("g.notes")[:B * 512 * 4].clone()
        torch.cuda.synchronize()
        eng.close()
        return md, mg, gD, gG, gE, bn, logits, notes
    finally:
        TL.debug_set("reset", 0)


def test_fused_pooling_and_bias_sums_equal_separate_passes():
    """Pooling and bias-gradient sums ON vs OFF, the BatchNorm statistics from the deterministic colreduce pass in both runs
    (the fused statistics are float32 atomics: their last bit, and with it a few bf16 roundings of the generated notes,
    changes from run to run), so the generated notes are bit-identical and what differs is float32 summation order after
    the last bf16 rounding point."""
    a, b = _run(4), _run(7)
    assert torch.equal(a[7], b[7]), "generated notes must not depend on the critic-side fusions"
    assert torch.allclose(a[0], b[0], rtol=2e-5, atol=1e-6), (a[0], b[0])        # loss_d, gp, D(real), D(fake)
    assert torch.allclose(a[1], b[1], rtol=2e-5, atol=1e-6), (a[1], b[1])        # g_adv, g_emo
    assert rel_l2(a[6], b[6]) < 1e-4, "ED logits"      # the Linears behind the pool run as TF32: a last-bit change of a pooled
                                                       # value can flip its 10-bit truncation
    for name, ga, gb in (("D", a[2], b[2]), ("G", a[3], b[3]), ("E", a[4], b[4])):
        for k in ga:
            if name == "G" and k in ("decoder.deconv.0.bias", "decoder.deconv.3.bias"):
                continue                                   # in front of a train-mode BatchNorm: pure cancellation noise
            n = gb[k].double().norm().item()
            e = (ga[k].double() - gb[k].double()).norm().item()
            # weight gradients use float32 atomics in both runs; a 1e-7 change of a pooled value can flip the bf16 rounding
            # of a few dz3 elements
            # (the generator's gradients sit behind the whole ED / critic / decoder backward chain of bf16 tensors)
            tol = 3e-4 if name == "D" else 2e-3
            assert e <= tol * max(n, 1e-12), f"{name}.{k}: fused vs separate passes differ by {e / max(n, 1e-30):.2e}"


def test_fused_batchnorm_statistics_equal_separate_pass():
    """BatchNorm statistics from the deconv epilogues vs the colreduce pass: the running statistics (a direct image of the
    sums) to 1e-6; everything downstream of the bf16-rounded BatchNorm output within bf16 flip noise."""
    a, b = _run(0), _run(4)
    for k in a[5]:
        assert rel_l2(a[5][k], b[5][k]) < 2e-6, k
    assert rel_l2(a[7], b[7]) < 1e-3, "generated notes"
    assert rel_l2(a[6], b[6]) < 2e-3, "ED logits"
    assert torch.allclose(a[0], b[0], rtol=2e-3, atol=1e-4) and torch.allclose(a[1], b[1], rtol=2e-3, atol=1e-4)
    for k in ("decoder.pre.2.weight", "decoder.deconv.0.weight", "decoder.deconv.1.weight", "noise_to_latent.net.0.weight"):
        assert rel_l2(a[3][k], b[3][k]) < 5e-2, k           # mask flips downstream of the bf16-rounded activations
