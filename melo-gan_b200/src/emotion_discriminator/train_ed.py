#!/usr/bin/env python3
"""Train the emotion discriminator (BASELINE config #3) on B200.

    python -m src.emotion_discriminator.train_ed [--config config/ed_config.yaml]

Same functions and flow as the reference's src/emotion_discriminator/train_ed.py (accuracy, save_checkpoint,
run_epoch, load_yaml, build_optimizer, build_scheduler, main): zero_grad -> logits -> CrossEntropy -> backward ->
AdamW, ReduceLROnPlateau on the validation loss, best / periodic checkpoints {'epoch','model','optimizer','cfg'},
early stopping.  main() runs the fast path (melogan.aux_trainers.EdTrainer): train-mode forward, fused cross-entropy
(loss, accuracy, dlogits), backward and AdamW over flat buffers as ONE CUDA graph per step, losses accumulated on the
device; run_epoch / build_optimizer / build_scheduler are the reference's functions for callers that drive the drop-in
module with torch's optimizer themselves (tests/test_ed_train_gpu.py does).
Data: the reference's ed_dataset.py reads per-file .npz archives (not built); this CLI reads the
pre-saved <SPLITS_DIR>/<split>/{notes,emotion}.npy arrays of the GAN fast path and keeps them on the device.
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
    """Batches of {'x': (B, max_notes, 4), 'y': (B,)} from pre-saved arrays.  Training drops the last partial batch (the
    native training context has one batch size); validation keeps it, like the reference's DataLoader."""

    def __init__(self, notes, labels, batch_size, shuffle, seed, drop_last=True):
        self.x, self.y, self.bs, self.shuffle = torch.as_tensor(notes), torch.as_tensor(labels), batch_size, shuffle
        self.drop_last = drop_last
        self.gen = torch.Generator().manual_seed(seed)

    def __iter__(self):
        n = len(self.x)
        idx = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        stop = n - self.bs + 1 if self.drop_last else n
        for i in range(0, max(stop, 0), self.bs):
            j = idx[i:i + self.bs].to(self.x.device)
            yield {"x": self.x[j], "y": self.y[j]}

    def __len__(self):
        return len(self.x) // self.bs if self.drop_last else (len(self.x) + self.bs - 1) // self.bs


def _load_split(cfg, csv_key):
    from melogan.aux_trainers import EdTrainer, find_split_dir
    from src.gan.utils import emotion_to_index
    base = find_split_dir(os.path.dirname(cfg[csv_key]), Path(cfg[csv_key]).stem)
    notes = np.load(os.path.join(base, "notes.npy")).astype(np.float32)
    emo = np.load(os.path.join(base, "emotion.npy"), allow_pickle=True)
    labels = torch.from_numpy(np.array([emotion_to_index(e) for e in emo], dtype=np.int64))
    EdTrainer.check_labels(labels, cfg.get("n_classes", 4))     # nn.CrossEntropyLoss raises on these; so do we, once
    return torch.from_numpy(notes), labels


def main(cfg_path):
    """train_ed.py:86-140 of the reference on the fast path: melogan.aux_trainers.EdTrainer (train-mode forward, fused
    cross-entropy, backward and AdamW as one CUDA graph per step; losses accumulated on the device)."""
    from melogan.aux_trainers import EdTrainer, ReduceOnPlateau
    cfg = load_yaml(cfg_path)
    if not torch.cuda.is_available():
        raise SystemExit("train_ed: a CUDA (sm_100a) device is required; this implementation has no CPU fallback")
    device = torch.device("cuda")
    torch.manual_seed(cfg.get("seed", 42))
    (tr_x, tr_y), (va_x, va_y) = _load_split(cfg, "train_split_csv"), _load_split(cfg, "val_split_csv")
    tr_x, tr_y, va_x, va_y = (t.to(device) for t in (tr_x, tr_y, va_x, va_y))       # the data set stays on the device
    bs = cfg["batch_size"]
    train_loader = _ArrayLoader(tr_x, tr_y, bs, True, cfg.get("seed", 42), drop_last=True)
    val_loader = _ArrayLoader(va_x, va_y, bs, False, 0, drop_last=False)
    tr = EdTrainer(cfg, batch=bs, precision=os.environ.get("MELOGAN_PRECISION", "fp32"), device=device)
    model = tr.model
    sch = cfg.get("scheduler") or {}
    scheduler = None
    if sch.get("name"):
        if sch["name"].lower() != "reducelronplateau":
            raise ValueError(f"Unsupported scheduler {sch['name']}")
        scheduler = ReduceOnPlateau(tr.opt, factor=sch.get("factor", 0.5), patience=sch.get("patience", 5),
                                    threshold=sch.get("threshold", 1e-4))
    metric = cfg.get("metric_for_best", "val_loss")
    if metric not in ("val_loss", "val_acc"):
        raise ValueError(f"Unsupported metric_for_best {metric}")
    use_graph = os.environ.get("MELOGAN_NO_GRAPH") is None
    captured_lr, statics = None, None
    best, bad = float("inf"), 0
    for epoch in range(1, cfg["num_epochs"] + 1):
        for batch in train_loader:
            if use_graph and captured_lr == tr.opt.lr:
                statics[0].copy_(batch["x"]); statics[1].copy_(batch["y"])
                tr.replay()
            else:
                tr.step(batch["x"].contiguous(), batch["y"].contiguous())
                if use_graph:
                    statics, captured_lr = tr.capture(), tr.opt.lr
        tl, ta = tr.epoch_means()
        acc = torch.zeros(2, device=device)
        n = 0
        for batch in val_loader:
            acc += tr.evaluate(batch["x"].contiguous(), batch["y"].contiguous()) * len(batch["y"])
            n += len(batch["y"])
        vl, va = (acc / max(n, 1)).tolist()
        print(f"Epoch {epoch}: train loss {tl:.4f} acc {ta:.3f} | val loss {vl:.4f} acc {va:.3f}")
        if scheduler is not None:
            scheduler.step(vl)
        score = vl if metric == "val_loss" else -va
        if score < best:
            best, bad = score, 0
            save_checkpoint(model, tr.opt, epoch, cfg, is_best=True)
        else:
            bad += 1
        if epoch % cfg.get("save_freq", 5) == 0:
            save_checkpoint(model, tr.opt, epoch, cfg)
        if bad >= cfg.get("early_stopping_patience", 10):
            print("Early stopping.")
            break


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="config/ed_config.yaml")
    main(ap.parse_args().config)
