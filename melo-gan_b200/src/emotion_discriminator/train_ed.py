#!/usr/bin/env python3
"""Train the emotion discriminator (BASELINE config #3) on B200.

    python -m src.emotion_discriminator.train_ed [--config config/ed_config.yaml]

Same functions and flow as the reference's src/emotion_discriminator/train_ed.py (accuracy, save_checkpoint,
run_epoch, load_yaml, build_optimizer, build_scheduler, main): zero_grad -> logits -> CrossEntropy -> backward ->
AdamW, ReduceLROnPlateau on the validation loss, best / periodic checkpoints {'epoch','model','optimizer','cfg'},
early stopping.  The model's forward and backward run in the native kernels (train mode: mg_emotion_train_*,
eval mode: mg_emotion_forward); the optimizer is torch's, driven exactly as in the reference.
Data: the reference's ed_dataset.py reads per-file .npz archives (not built); this CLI reads the
pre-saved <SPLITS_DIR>/<split>/{notes,emotion}.npy arrays of the GAN fast path.
"""
import argparse
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
import yaml

from src.emotion_discriminator.ed_model import EmotionDiscriminator


def accuracy(pred, target):
    return (pred.argmax(dim=1) == target).float().mean().item()


def save_checkpoint(model, optimizer, epoch, cfg, is_best=False):
    os.makedirs(cfg["checkpoint_dir"], exist_ok=True)
    name = cfg["save_name"] if is_best else f"ed_epoch{epoch}.pth"
    path = os.path.join(cfg["checkpoint_dir"], name)
    torch.save({"epoch": epoch, "model": model.state_dict(), "optimizer": optimizer.state_dict(), "cfg": cfg}, path)
    return path


def run_epoch(model, loader, criterion, optimizer, device, is_train=True):
    model.train() if is_train else model.eval()
    total_loss = total_acc = total_count = 0
    for batch in loader:
        x, y = batch["x"].to(device), batch["y"].to(device)
        if is_train:
            optimizer.zero_grad()
        with torch.set_grad_enabled(is_train):
            logits = model(x)
            loss = criterion(logits, y)
        if is_train:
            loss.backward()
            optimizer.step()
        bs = x.size(0)
        total_loss += loss.item() * bs
        total_acc += accuracy(logits.detach(), y) * bs
        total_count += bs
    return total_loss / total_count, total_acc / total_count


def load_yaml(path):
    with open(path, "r") as f:
        return yaml.safe_load(f)


def build_optimizer(model, cfg):
    o = cfg["optimizer"]
    name, lr = o["name"].lower(), float(o["lr"])
    wd, betas = o.get("weight_decay", 0), tuple(o.get("betas", [0.9, 0.999]))
    if name == "adamw":
        return optim.AdamW(model.parameters(), lr=lr, weight_decay=wd, betas=betas)
    if name == "adam":
        return optim.Adam(model.parameters(), lr=lr, weight_decay=wd, betas=betas)
    raise ValueError(f"Unsupported optimizer {name}")


def build_scheduler(optimizer, cfg):
    sch = cfg.get("scheduler")
    if not sch or sch.get("name") is None:
        return None
    if sch["name"].lower() == "reducelronplateau":
        return optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode=sch.get("mode", "min"), factor=sch.get("factor", 0.5),
                                                    patience=sch.get("patience", 5), threshold=sch.get("threshold", 1e-4))
    raise ValueError(f"Unsupported scheduler {sch['name']}")


class _ArrayLoader:
    """Batches of {'x': (B, max_notes, 4), 'y': (B,)} from pre-saved arrays; drop_last so that the native context
    sees one batch size."""

    def __init__(self, notes, labels, batch_size, shuffle, seed):
        self.x, self.y, self.bs, self.shuffle = torch.from_numpy(notes), torch.from_numpy(labels), batch_size, shuffle
        self.gen = torch.Generator().manual_seed(seed)

    def __iter__(self):
        n = len(self.x)
        idx = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        for i in range(0, n - self.bs + 1, self.bs):
            j = idx[i:i + self.bs]
            yield {"x": self.x[j], "y": self.y[j]}

    def __len__(self):
        return len(self.x) // self.bs


def _load_split(cfg, csv_key):
    from src.gan.utils import emotion_to_index
    split = Path(cfg[csv_key]).stem
    base = os.path.join(os.path.dirname(cfg[csv_key]), split)
    notes = np.load(os.path.join(base, "notes.npy")).astype(np.float32)
    emo = np.load(os.path.join(base, "emotion.npy"), allow_pickle=True)
    return notes, np.array([emotion_to_index(e) for e in emo], dtype=np.int64)


def main(cfg_path):
    cfg = load_yaml(cfg_path)
    if not torch.cuda.is_available():
        raise SystemExit("train_ed: a CUDA (sm_100a) device is required; this implementation has no CPU fallback")
    device = torch.device("cuda")
    torch.manual_seed(cfg.get("seed", 42))
    tr_x, tr_y = _load_split(cfg, "train_split_csv")
    va_x, va_y = _load_split(cfg, "val_split_csv")
    train_loader = _ArrayLoader(tr_x, tr_y, cfg["batch_size"], True, cfg.get("seed", 42))
    val_loader = _ArrayLoader(va_x, va_y, cfg["batch_size"], False, 0)
    model = EmotionDiscriminator(cfg).to(device)
    optimizer, criterion = build_optimizer(model, cfg), nn.CrossEntropyLoss()
    scheduler = build_scheduler(optimizer, cfg)
    best, bad = float("inf"), 0
    for epoch in range(1, cfg["num_epochs"] + 1):
        tl, ta = run_epoch(model, train_loader, criterion, optimizer, device, is_train=True)
        vl, va = run_epoch(model, val_loader, criterion, optimizer, device, is_train=False)
        print(f"Epoch {epoch}: train loss {tl:.4f} acc {ta:.3f} | val loss {vl:.4f} acc {va:.3f}")
        if scheduler is not None:
            scheduler.step(vl)
        if vl < best:
            best, bad = vl, 0
            save_checkpoint(model, optimizer, epoch, cfg, is_best=True)
        else:
            bad += 1
        if epoch % cfg.get("save_freq", 5) == 0:
            save_checkpoint(model, optimizer, epoch, cfg)
        if bad >= cfg.get("early_stopping_patience", 10):
            print("Early stopping.")
            break


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="config/ed_config.yaml")
    main(ap.parse_args().config)
