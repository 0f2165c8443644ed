#!/usr/bin/env python3
"""Train the VAE (BASELINE config #2) on B200.

    python -m src.ae.train_ae [--config config/ae_config.yaml]

Same flow as the reference's src/ae/train_ae.py (vae_loss, train, main): per batch forward -> vae_loss -> zero_grad ->
backward -> clip_grad_norm_(1.0) -> AdamW, beta warm-up per epoch, ReduceLROnPlateau(factor 0.5, patience 5, min_lr 1e-6)
on the validation loss, ae_best.pth {'epoch','model_state'} / ae_final.pth checkpoints, early stopping.  The model's
whole iteration runs natively (melogan.aux_trainers.VaeTrainer): mg_vae_loss_step (forward + vae_loss + backward) and
mg_adam_step_clipped (gradient-norm clip + AdamW) over flat buffers, replayed as one CUDA graph per step.  `vae_loss` below
is the reference's function, kept for callers that drive the drop-in module themselves.
Data: the reference's MIDIDataset reads per-file .npz archives; this CLI reads the pre-saved <SPLITS_DIR>/<split>/notes.npy
arrays, keeps them on the device and (NORMALIZE_RAW: true) applies the dataset's normalisation there (mg_ae_normalize).  The
probabilistic augmentations are switched off by config/ae_config.yaml (all amplitudes 0) and are not built; TensorBoard and
reconstruction MIDI dumps of the reference loop are not part of the hot path and are left out.
"""
import argparse
import os

import numpy as np
import torch
import torch.nn.functional as F
import yaml

from src.ae.model import VAE


def vae_loss(recon, target, mu, log_var, beta):
    """Loss = MSE(recon, target) + beta * KLD(N(mu, exp(log_var)) || N(0, 1)), KLD averaged over every element."""
    recon_loss = F.mse_loss(recon, target)
    kld_loss = -0.5 * torch.mean(1 + log_var - mu.pow(2) - log_var.exp())
    return recon_loss + beta * kld_loss, recon_loss, kld_loss


def load_rolls(cfg, split, device):
    """<SPLITS_DIR>/<split>/notes.npy as a device-resident float32 tensor.  cfg['NORMALIZE_RAW'] (default false: the GAN
    fast path's arrays are already in model units) applies MIDIDataset.__getitem__'s normalisation of raw rolls
    (reference src/ae/dataset.py:72-89,105) on the device, the whole split in one launch."""
    from melogan import notes as N
    from melogan.aux_trainers import find_split_dir
    x = torch.from_numpy(np.load(os.path.join(find_split_dir(cfg['SPLITS_DIR'], split), "notes.npy")).astype(np.float32))
    x = x.to(device)
    if cfg.get('NORMALIZE_RAW', False):
        x = N.ae_normalize(x, cfg.get('MAX_START_BEAT', 100.0), cfg.get('MAX_DURATION_BEAT', 20.0))
    return x


def train(cfg):
    """train_ae.py:54-199 of the reference on the fast path: melogan.aux_trainers.VaeTrainer (one CUDA graph per step:
    device RNG, fused forward + vae_loss + backward, fused clip_grad_norm_ + AdamW, device loss accumulation)."""
    from melogan.aux_trainers import ReduceOnPlateau, VaeTrainer
    if not torch.cuda.is_available():
        raise SystemExit("train_ae: a CUDA (sm_100a) device is required; this implementation has no CPU fallback")
    device = torch.device("cuda")
    model_dir = cfg.get('CHECKPOINT_DIR', 'models/ae')
    os.makedirs(model_dir, exist_ok=True)
    train_x, val_x = load_rolls(cfg, "train", device), load_rolls(cfg, "val", device)      # resident on the device
    print(f"Train rolls: {len(train_x)}   Val rolls: {len(val_x)}")

    bs = cfg['BATCH_SIZE']
    tr = VaeTrainer(cfg, batch=bs, precision=os.environ.get("MELOGAN_PRECISION", "fp32"), device=device, max_norm=1.0)
    model = tr.model
    scheduler = ReduceOnPlateau(tr.opt, factor=0.5, patience=5, min_lr=1e-6)
    best_val, no_improve = float('inf'), 0
    patience = cfg.get('EARLY_STOP_PATIENCE', 10)
    warm, final_beta = cfg.get('KLD_WARMUP_EPOCHS', 25), float(cfg.get('BETA', 1.0))
    gen = torch.Generator().manual_seed(int(cfg.get('SEED', 0)))
    use_graph = os.environ.get("MELOGAN_NO_GRAPH") is None
    captured_for, s_x = None, None

    for epoch in range(1, cfg['EPOCHS'] + 1):
        beta = final_beta if epoch >= warm else min(final_beta, (epoch / warm) * final_beta)
        tr.beta = beta
        perm = torch.randperm(len(train_x), generator=gen).to(device)                   # shuffle=True, drop_last=True
        for i in range(0, len(train_x) - bs + 1, bs):
            idx = perm[i:i + bs]
            if use_graph and captured_for == (beta, tr.opt.lr):
                torch.index_select(train_x, 0, idx, out=s_x)
                tr.replay()
            else:
                tr.step(train_x[idx].contiguous())                                      # eager (also the warm-up of a capture)
                if use_graph:
                    s_x = tr.capture()                                                   # beta and lr are baked into the graph
                    captured_for = (beta, tr.opt.lr)
        tot = tr.epoch_means()

        val = torch.zeros(3, device=device)
        vb = 0
        for i in range(0, len(val_x), bs):                                               # drop_last=False: the tail batch counts
            val += tr.evaluate(val_x[i:i + bs].contiguous())
            vb += 1
        val = (val / max(1, vb)).tolist()
        scheduler.step(val[0])
        print(f"[Epoch {epoch}] Train: {tot[0]:.6f} (Recon: {tot[1]:.6f}, KLD: {tot[2]:.6f}) | "
              f"Val: {val[0]:.6f} (Recon: {val[1]:.6f}, KLD: {val[2]:.6f})  beta={beta:.2f}")
        if val[0] < best_val:
            best_val, no_improve = val[0], 0
            best_path = os.path.join(model_dir, "ae_best.pth")
            torch.save({'epoch': epoch, 'model_state': model.state_dict()}, best_path)
            print("Saved new best model ->", best_path)
        else:
            no_improve += 1
        if no_improve >= patience:
            print("No improvement for", patience, "epochs. Early stopping.")
            break

    print("Training complete. Best val:", best_val)
    final_model_path = os.path.join(model_dir, "ae_final.pth")
    torch.save(model.state_dict(), final_model_path)
    print("Saved final model:", final_model_path)
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config/ae_config.yaml")
    args = ap.parse_args()
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    train(cfg)


if __name__ == "__main__":
    main()
