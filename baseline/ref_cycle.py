"""Reference arm of bench.py: the UNMODIFIED reference modules, driven by the reference's own loop body.

`__graft_entry__.build()` copies the reference's Python sources (src/, config/, tools/) into baseline/_ref when
/root/reference is present (git-ignored, travels to the GPU box with the snapshot).  This module loads
src/gan/{models,feature_encoder,utils}.py and src/emotion_discriminator/ed_model.py from there BY FILE PATH (the
product package uses the same import names) and runs the loop body of src/gan/train_gan.py:183-251 -- which is not
a callable in the reference (it lives under `__main__`) and is therefore restated here statement for statement --
on synthetic batches of the shape in config/gan_config.yaml.  None of this repository's kernels, engine or oracle
is on this path; `pretty_midi` (imported at the top of the reference's utils.py, not installed here) is satisfied
by the recording stand-in of oracle/stubs.

Used for: `bench.py --impl reference` (CPU, all host threads, B = BATCH_SIZE = 32) and the stock-PyTorch-eager
yardstick on the B200 itself (device="cuda", the bench batch, fp32 and bf16 autocast).
"""
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.exists(os.path.join(REF, "src", "gan", "models.py"))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class ReferenceCycle:
    """E_num, G, D, ED and both Adam optimizers built exactly as train_gan.py:85-145 does; cycle() = CRITIC_ITERS critic
    steps on fresh batches + one generator step on the last batch (train_gan.py:168-251)."""

    def __init__(self, device="cpu", seed=42):
        import torch
        import torch.nn as nn
        import torch.optim as optim
        import yaml
        stubs = os.path.join(ROOT, "oracle", "stubs")
        if stubs not in sys.path:
            sys.path.insert(0, stubs)
        self.torch = torch
        models, fe, ed = _load("ref_models", "src/gan/models.py"), _load("ref_fe", "src/gan/feature_encoder.py"), \
            _load("ref_ed", "src/emotion_discriminator/ed_model.py")
        self.utils = _load("ref_utils", "src/gan/utils.py")
        with open(os.path.join(REF, "config", "gan_config.yaml")) as f:
            self.cfg = cfg = yaml.safe_load(f)
        with open(os.path.join(REF, "config", "ed_config.yaml")) as f:
            self.ed_cfg = ed_cfg = yaml.safe_load(f)
        self.device = device = torch.device(device)
        self.utils.seed_everything(seed)
        emb = cfg.get('ENCODER_OUT_DIM', 128)
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            self.E_num = fe.FeatureEncoder(in_dim=cfg.get('NUMERIC_INPUT_DIM', 6), hidden_dims=cfg.get('ENCODER_HIDDEN', [256, 128]),
                                           out_dim=emb).to(device)
            self.G = models.Generator(noise_dim=cfg['NOISE_DIM'], latent_dim=cfg['LATENT_DIM'],
                                      mode=cfg.get('INTEGRATION_MODE', 'conditioning'), max_notes=cfg['MAX_NOTES'],
                                      note_dim=cfg['NOTE_DIM'], numeric_embed_dim=emb).to(device)
            self.D = models.Discriminator(max_notes=cfg['MAX_NOTES'], note_dim=cfg['NOTE_DIM'], numeric_embed_dim=emb).to(device)
            self.ED = ed.EmotionDiscriminator(ed_cfg).to(device)
        self.E_num.apply(self.utils.weights_init); self.G.apply(self.utils.weights_init); self.D.apply(self.utils.weights_init)
        for p in self.ED.parameters():
            p.requires_grad = False
        self.ED.eval()
        betas = (cfg.get('BETA1', 0.5), cfg.get('BETA2', 0.9))
        self.opt_G = optim.Adam(list(self.G.parameters()) + list(self.E_num.parameters()), lr=float(cfg['LR_G']), betas=betas)
        self.opt_D = optim.Adam(self.D.parameters(), lr=float(cfg['LR_D']), betas=betas)
        self.criterion_emo = nn.CrossEntropyLoss()
        self.G.train(); self.E_num.train(); self.D.train()

    def batches(self, B, seed=7):
        torch = self.torch
        g = torch.Generator().manual_seed(seed)
        K, cfg = int(self.cfg.get('CRITIC_ITERS', 5)), self.cfg
        out = []
        for _ in range(K):
            notes = torch.rand((B, cfg['MAX_NOTES'], cfg['NOTE_DIM']), generator=g) * 2 - 1
            numeric = torch.randn((B, cfg.get('NUMERIC_INPUT_DIM', 6)), generator=g)
            numeric[:, 5] = 0.0
            out.append((notes.to(self.device), numeric.to(self.device)))
        labels = (torch.arange(B) % 4).to(self.device)
        return out, labels

    def cycle(self, batches, emot_idx, autocast=None):
        torch, cfg, device = self.torch, self.cfg, self.device
        E_num, G, D_discriminator, D_emotion = self.E_num, self.G, self.D, self.ED
        lambda_gp, lambda_emotion = cfg.get('LAMBDA_GP', 10.0), cfg.get('LAMBDA_EMOTION', 1.0)
        import contextlib
        ctx = (lambda: torch.autocast(device_type=device.type, dtype=autocast)) if autocast is not None else contextlib.nullcontext
        loss_d = loss_g = None
        for batch_idx, (notes_real, numeric_batch) in enumerate(batches):
            bsize = notes_real.size(0)
            encoder_latent = torch.zeros(bsize, cfg['LATENT_DIM'], device=device)
            # ---- train_gan.py:183-205 ----
            self.opt_D.zero_grad()
            with ctx():
                with torch.no_grad():
                    numeric_emb_d = E_num(numeric_batch)
                    noise = torch.randn(bsize, cfg['NOISE_DIM'], device=device)
                    gen_notes_d, _ = G(noise, encoder_latent, numeric_emb_d)
                d_real = D_discriminator(notes_real, numeric_emb_d)
                d_fake = D_discriminator(gen_notes_d.detach(), numeric_emb_d)
                gp = self.utils.compute_gradient_penalty(D_discriminator, notes_real.data, gen_notes_d.data, numeric_emb_d, device)
                loss_d = torch.mean(d_fake) - torch.mean(d_real) + (lambda_gp * gp)
            loss_d.backward()
            self.opt_D.step()
            # ---- train_gan.py:212-251 ----
            if (batch_idx + 1) % int(cfg.get('CRITIC_ITERS', 5)) == 0:
                self.opt_G.zero_grad()
                with ctx():
                    numeric_emb_g = E_num(numeric_batch)
                    noise_g = torch.randn(bsize, cfg['NOISE_DIM'], device=device)
                    gen_notes_g, gen_latent_g = G(noise_g, encoder_latent, numeric_emb_g)
                    d_fake_g = D_discriminator(gen_notes_g, numeric_emb_g)
                    loss_g_adv = -torch.mean(d_fake_g)
                    ed_input = gen_latent_g if self.ed_cfg.get('input_mode', 'notes') == 'latent' else gen_notes_g
                    ed_logits = D_emotion(ed_input)
                    loss_g_emo_cls = self.criterion_emo(ed_logits, emot_idx)
                    loss_g = loss_g_adv + (lambda_emotion * loss_g_emo_cls)
                loss_g.backward()
                self.opt_G.step()
        return loss_d, loss_g


def time_cycles(device="cpu", B=32, steps=10, warmup=2, autocast=None, threads=None):
    """rolls/s of the reference cycle.  CPU: wall clock; CUDA: events around the timed cycles."""
    import torch
    if device == "cpu" and threads:
        torch.set_num_threads(threads)
    rc = ReferenceCycle(device)
    batches, labels = rc.batches(B)
    K = len(batches)
    for _ in range(warmup):
        rc.cycle(batches, labels, autocast)
    if device == "cpu":
        t0 = time.perf_counter()
        for _ in range(steps):
            rc.cycle(batches, labels, autocast)
        el = time.perf_counter() - t0
    else:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            rc.cycle(batches, labels, autocast)
        e1.record()
        torch.cuda.synchronize()
        el = e0.elapsed_time(e1) * 1e-3
    return {"rolls_per_s": K * B * steps / el, "ms_per_cycle": 1e3 * el / steps, "B": B, "cycles": steps}
