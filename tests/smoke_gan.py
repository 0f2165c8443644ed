"""Smoke check of the GAN hot path: one small critic step and one generator step on cuda:0 (fp32 mode and bf16
tensor-core mode) compared with the oracle.  Called by __graft_entry__.smoke() only (kept under tests/: the product package never imports oracle/)."""
import torch


def run():
    from melogan import engine as E
    from oracle import gan_oracle as O
    B = 8
    params = O.make_params(4, fan_in_scale=True)
    batch = O.make_batch(40, B)
    ref_d = O.critic_step(O.clone_params(params), batch, {}, update=False)
    ref_g = O.generator_step(O.clone_params(params), batch, {}, update=False)
    for precision, tol in (("fp32", 1e-5), ("bf16", 1e-2)):
        eng = E.GanEngine(B, precision=precision)
        cp = {m: {k: v.clone().cuda() for k, v in P.items()} for m, P in params.items()}
        grads = {m: {k: torch.zeros_like(cp[m][k]) for k in E.GRAD_KEYS[i]} for m, i in (("E", 0), ("G", 1), ("D", 2))}
        for m, i in (("E", 0), ("G", 1), ("D", 2)):
            eng.bind(i, cp[m], grads[m])
        eng.bind(E.MOD_ED, cp["ED"], None)
        cb = {k: v.cuda() for k, v in batch.items()}
        md = eng.critic_step(cb["notes_real"], cb["numeric"], cb["noise_d"], cb["alpha"], cb["mask1_d"], cb["mask2_d"]).cpu()
        mg = eng.generator_step(cb["numeric"], cb["noise_g"], cb["emot_idx"], cb["mask1_g"], cb["mask2_g"]).cpu()
        assert abs(md[0].item() - ref_d["loss_d"].item()) <= tol * abs(ref_d["loss_d"].item()), (precision, md, ref_d["loss_d"])
        assert abs(mg[1].item() - ref_g["loss_g_emo"].item()) <= tol * abs(ref_g["loss_g_emo"].item()), (precision, mg)
        g = grads["D"]["conv.4.weight"].cpu()
        r = ref_d["grads"]["conv.4.weight"]
        err = ((g - r).norm() / r.norm()).item()
        assert err <= (1e-4 if precision == "fp32" else 0.15), (precision, "conv.4.weight grad", err)
        eng.close()
    print("gan smoke ok")
