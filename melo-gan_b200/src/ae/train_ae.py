#!/usr/bin/env python3
"""Train the VAE (BASELINE config #2) on B200.

    python -m src.ae.train_ae [--config config/ae_config.yaml]

Same flow as the reference's src/ae/train_ae.py (vae_loss, train, main): per batch forward -> vae_loss -> zero_grad ->
backward -> clip_grad_norm_(1.0) -> AdamW, beta warm-up per epoch, ReduceLROnPlateau(factor 0.5, patience 5, min_lr 1e-6)
on the validation loss, ae_best.pth {'epoch','model_state'} / ae_final.pth checkpoints, early stopping.  The model's
forward and backward run in the native kernels (mg_vae_forward / mg_vae_backward); the loss arithmetic, gradient clipping
and the optimizer are torch's, driven exactly as in the reference.
Data: the reference's MIDIDataset reads per-file .npz archives with augmentation (not built); this CLI reads
the pre-saved <SPLITS_DIR>/<split>/notes.npy arrays of the GAN fast path.  TensorBoard and reconstruction MIDI dumps of the
reference loop are not part of the hot path and are left out.
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


def _batches(notes, bs, shuffle, gen, drop_last):
    n = len(notes)
    idx = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
    stop = n - bs + 1 if drop_last else n
    for i in range(0, max(stop, 0), bs):
        yield notes[idx[i:i + bs]]


def train(cfg):
    if not torch.cuda.is_available():
        raise SystemExit("train_ae: a CUDA (sm_100a) device is required; this implementation has no CPU fallback")
    device = torch.device("cuda")
    model_dir = cfg.get('CHECKPOINT_DIR', 'models/ae')
    os.makedirs(model_dir, exist_ok=True)
    load = lambda split: torch.from_numpy(np.load(os.path.join(cfg['SPLITS_DIR'], split, "notes.npy")).astype(np.float32))
    train_x, val_x = load("train"), load("val")
    print(f"Train rolls: {len(train_x)}   Val rolls: {len(val_x)}")

    model = VAE(cfg).to(device)
    with torch.no_grad():
        model.encoder(torch.zeros(1, cfg['MAX_NOTES'], 4, device=device))       # materialise encoder._linear
    optimizer = torch.optim.AdamW(model.parameters(), lr=float(cfg.get('LR', 0.0001)),
                                  weight_decay=float(cfg.get('WEIGHT_DECAY', 0.00001)))
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=0.5, patience=5, min_lr=1e-6)
    best_val, no_improve = float('inf'), 0
    patience = cfg.get('EARLY_STOP_PATIENCE', 10)
    warm, final_beta = cfg.get('KLD_WARMUP_EPOCHS', 25), float(cfg.get('BETA', 1.0))
    gen = torch.Generator().manual_seed(int(cfg.get('SEED', 0)))
    bs = cfg['BATCH_SIZE']

    for epoch in range(1, cfg['EPOCHS'] + 1):
        model.train()
        beta = final_beta if epoch >= warm else min(final_beta, (epoch / warm) * final_beta)
        tot = np.zeros(3)
        nb = 0
        for notes in _batches(train_x, bs, True, gen, drop_last=True):
            notes = notes.to(device)
            recon, z, mu, log_var = model(notes)
            loss, recon_loss, kld_loss = vae_loss(recon, notes, mu, log_var, beta)
            optimizer.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            optimizer.step()
            tot += [loss.item(), recon_loss.item(), kld_loss.item()]
            nb += 1
        tot /= max(1, nb)

        model.eval()
        val = np.zeros(3)
        vb = 0
        with torch.no_grad():
            for notes in _batches(val_x, bs, False, gen, drop_last=False):
                notes = notes.to(device)
                recon, z, mu, log_var = model(notes)
                val += [t.item() for t in vae_loss(recon, notes, mu, log_var, beta=1.0)]
                vb += 1
        val /= max(1, vb)
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
