#!/usr/bin/env python3
"""Encode rolls to VAE latents (mu) on B200 -- the producer of `encoder_feats.npy`.

    python -m src.ae.encode --model data/models/ae/ae_best.pth --notes data/splits/train/notes.npy \\
                            --out_file data/splits/train/encoder_feats.npy [--config config/ae_config.yaml]

Same model handling as the reference's src/ae/encode.py:57-141 (VAE(cfg), lazy encoder linear materialised by a dummy
pass, `model_state` / bare state dict, eval mode, `mu` only, batches of 32, one .npy of shape (N, LATENT_DIM)).  The
forward runs in the native kernels (mg_vae_forward with train = 0: BatchNorm from running statistics).  Input: the
pre-saved notes array of the GAN fast path (the reference walks a manifest of per-file .npz archives).
"""
import argparse
import os

import numpy as np
import torch
import yaml

from src.ae.model import VAE


def encode(model, notes, batch_size=32, device="cuda"):
    """(N, MAX_NOTES, 4) float32 array -> (N, LATENT_DIM) array of posterior means; the tail batch is zero-padded because
    a native context is sized for one batch."""
    model.eval()
    out = []
    with torch.no_grad():
        for i in range(0, len(notes), batch_size):
            chunk = torch.from_numpy(np.ascontiguousarray(notes[i:i + batch_size], dtype=np.float32))
            n = chunk.shape[0]
            if n < batch_size:
                chunk = torch.cat([chunk, torch.zeros((batch_size - n,) + tuple(chunk.shape[1:]))])
            recon, z, mu, log_var = model(chunk.to(device))
            out.append(mu[:n].cpu().numpy())
    return np.concatenate(out, axis=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", type=str, required=True, help="Path to the full ae_best.pth VAE model")
    ap.add_argument("--notes", type=str, required=True, help="(N, MAX_NOTES, 4) notes.npy")
    ap.add_argument("--out_file", type=str, required=True, help="Output .npy file for latents (mu)")
    ap.add_argument("--config", type=str, default="config/ae_config.yaml")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("encode: a CUDA (sm_100a) device is required; this implementation has no CPU fallback")
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    device = torch.device("cuda")
    model = VAE(cfg).to(device)
    with torch.no_grad():
        model.encoder(torch.zeros(1, cfg['MAX_NOTES'], 4, device=device))
    ckpt = torch.load(args.model, map_location=device)
    model.load_state_dict(ckpt['model_state'] if 'model_state' in ckpt else ckpt)
    notes = np.load(args.notes).astype(np.float32)
    print(f"Found {len(notes)} rolls to encode")
    latents = encode(model, notes, device=device)
    os.makedirs(os.path.dirname(os.path.abspath(args.out_file)), exist_ok=True)
    np.save(args.out_file, latents)
    print(f"Saved latents ({latents.shape}) -> {args.out_file}")


if __name__ == "__main__":
    main()
