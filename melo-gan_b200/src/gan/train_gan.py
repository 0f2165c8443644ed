#!/usr/bin/env python3
"""Train MELO-GAN (WGAN-GP + numeric conditioning + frozen emotion discriminator) on B200.

    python -m src.gan.train_gan [--config config/gan_config.yaml] [--ed_config config/ed_config.yaml]
                                [--ed_ckpt data/models/ed/ed_best.pth]

Same CLI, config schema, loop schedule (every batch a critic step, every CRITIC_ITERS-th batch a generator
step on that batch), epoch log line, TensorBoard scalar names and checkpoint layouts as the reference's
src/gan/train_gan.py; each step body runs as one fused native call through melogan.trainer.GanTrainer
(losses stay on the device until the epoch log).  Extra knobs come from the environment, never from the
YAML: MELOGAN_PRECISION=fp32|bf16; under torchrun the batch is sharded over the ranks (NCCL all-reduce).
"""
import argparse
import glob
import os
from pathlib import Path

import numpy as np
import torch
import yaml

from melogan.trainer import GanTrainer
from src.gan.utils import emotion_to_index


def load_config(path):
    with open(path) as f:
        return yaml.safe_load(f)


_PATH_COLUMNS = ['npz_path', 'processed_file', 'processed', 'full_path', 'filepath', 'file', 'filename', 'file_key']


def _resolve_npz_path(processed_dir, raw_cell, row):
    """Manifest cell -> processed .npz, in the reference's order (src/gan/dataset.py:128-156): the cell itself as a path,
    then a stem search in PROCESSED_DIR, then the row's npz_path column."""
    raw_cell = str(raw_cell)
    candidate = raw_cell if os.path.isabs(raw_cell) else os.path.join(processed_dir, raw_cell)
    if raw_cell.lower().endswith('.npz') and os.path.exists(candidate):
        return candidate
    stem = os.path.splitext(os.path.basename(raw_cell))[0]
    found = sorted(glob.glob(os.path.join(processed_dir, f"*{stem}*.npz")))
    if found:
        return found[0]
    alt = row.get('npz_path', '')
    if isinstance(alt, str) and alt:
        candidate = alt if os.path.isabs(alt) else os.path.join(processed_dir, alt)
        if os.path.exists(candidate):
            return candidate
    return None


def load_split_npz(cfg, split_csv):
    """Slow path of the reference's GANDataset (src/gan/dataset.py:58-112,176-200): one .npz per manifest row with keys
    notes (MAX_NOTES, 4), numeric_features (padded / truncated to NUMERIC_INPUT_DIM, zeros if absent) and mood (else the
    manifest's emotion / mood / label column).  Read ONCE into arrays: the step path only ever sees device tensors."""
    import pandas as pd
    df = pd.read_csv(split_csv)
    col = next((c for c in _PATH_COLUMNS if c in df.columns), None)
    if col is None:
        raise KeyError(f"Split CSV must contain one of {_PATH_COLUMNS}. Found columns: {list(df.columns)}")
    processed_dir, nd = cfg.get('PROCESSED_DIR', 'data/processed'), int(cfg.get('NUMERIC_INPUT_DIM', 6))
    mood_col = next((c for c in ('emotion', 'mood', 'label') if c in df.columns), None)
    notes, numeric, moods = [], [], []
    for _, r in df.iterrows():
        path = _resolve_npz_path(processed_dir, r[col], r)
        if path is None:
            print(f"[WARN] Could not find processed .npz for manifest row: {dict(r)}")
            continue
        with np.load(path, allow_pickle=True) as data:
            notes.append(np.asarray(data['notes'], dtype=np.float32))
            num = np.zeros(nd, dtype=np.float32)
            if 'numeric_features' in data:
                v = np.asarray(data['numeric_features'], dtype=np.float32).flatten()
                num[:min(v.size, nd)] = v[:nd]
            numeric.append(num)
            mood = data['mood'].item() if 'mood' in data and np.asarray(data['mood']).ndim == 0 else None
            moods.append(mood if mood is not None else (r[mood_col] if mood_col else None))
    if not notes:
        raise FileNotFoundError(f"no processed .npz resolved from {split_csv} under {processed_dir}")
    labels = np.array([emotion_to_index(m) for m in moods], dtype=np.int64)
    check_labels(labels, str(split_csv))
    return np.stack(notes), np.stack(numeric), labels


def load_split_arrays(cfg, split_csv):
    """Fast path of the reference's GANDataset (src/gan/dataset.py:30-56): <SPLITS_DIR>/<split>/{notes,emotion,
    numeric_features}.npy; without them, the per-file .npz path (load_split_npz)."""
    splits_dir = cfg.get('SPLITS_DIR', 'data/splits')
    name = Path(split_csv).stem
    base = os.path.join(splits_dir, name)
    paths = [os.path.join(base, f) for f in ("notes.npy", "emotion.npy", "numeric_features.npy")]
    if not all(os.path.exists(p) for p in paths):
        print(f"[INFO] Loading data by reading individual .npz files for split: {name}")
        return load_split_npz(cfg, split_csv)
    notes, emotions, numeric = (np.load(p, allow_pickle=True) for p in paths)
    if not (len(notes) == len(emotions) == len(numeric)):
        raise ValueError("NPY file length mismatch (notes, emotions, numeric_features)")
    labels = np.array([emotion_to_index(e) for e in emotions], dtype=np.int64)
    check_labels(labels, name)
    return notes.astype(np.float32), numeric.astype(np.float32), labels


def check_labels(labels, what, n_classes=4):
    """The generator step feeds these to the fused cross-entropy; nn.CrossEntropyLoss (reference train_gan.py:231) raises
    on a target outside [0, n_classes) -- emotion_to_index gives -1 for an unknown or missing mood -- so do we, once, here."""
    bad = np.nonzero((labels < 0) | (labels >= n_classes))[0]
    if len(bad):
        raise IndexError(f"{what}: Target {int(labels[bad[0]])} is out of bounds (row {int(bad[0])}; {len(bad)} such rows): "
                         f"emotion labels must map to [0, {n_classes})")


def load_encoder_feats(path, n, latent_dim):
    """`latent_feats` of the reference's prepare_dataset / GANDataset (train_gan.py:50-51, dataset.py:47-54,169-171): the
    array at cfg['ENCODER_FEATS_TRAIN'] when it exists and has one row per sample, else (warning) zero latents."""
    if path and os.path.exists(path):
        feats = np.load(path).astype(np.float32)
        if feats.shape[0] == n and feats.ndim == 2 and feats.shape[1] == latent_dim:
            return feats
        print(f"[WARN] latent_feats shape mismatch {feats.shape} vs ({n}, {latent_dim}). Ignoring latent_feats.")
    else:
        print(f"[WARN] latent_feats not found at {path}; using zero latents.")
    return np.zeros((n, latent_dim), dtype=np.float32)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="config/gan_config.yaml", help="Path to the main GAN config")
    ap.add_argument("--ed_config", type=str, default="config/ed_config.yaml", help="Path to the ED config")
    ap.add_argument("--ed_ckpt", type=str, default="data/models/ed/ed_best.pth")
    ap.add_argument("--resume", type=str, default=None, help="gan_epochNNNN.pth to continue from (SURVEY 8f-4)")
    args = ap.parse_args(argv)
    cfg, ed_cfg = load_config(args.config), load_config(args.ed_config)
    if not torch.cuda.is_available():
        raise SystemExit("train_gan: a CUDA (sm_100a) device is required; this implementation has no CPU fallback")

    world, rank, local, pg = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), \
        int(os.environ.get("LOCAL_RANK", "0")), None
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        pg = torch.distributed.group.WORLD
    device = torch.device(f"cuda:{local}")
    print(f"Using main device: {device}")

    notes, numeric, labels = load_split_arrays(cfg, cfg['TRAIN_SPLIT'])
    print(f"Train set size: {len(notes)}")
    ed_state = None
    if os.path.exists(args.ed_ckpt):
        print(f"[INFO] Loading pre-trained Emotion Discriminator from {args.ed_ckpt}")
        ck = torch.load(args.ed_ckpt, map_location="cpu")
        ed_state = ck['model'] if 'model' in ck else ck
    else:
        print(f"[WARN] ED checkpoint not found at {args.ed_ckpt}. ED will be random!")

    B = int(cfg.get('BATCH_SIZE', 32))
    if B % world:
        raise SystemExit(f"BATCH_SIZE {B} must be divisible by the world size {world}")
    tr = GanTrainer(cfg, ed_cfg, batch=B // world, precision=os.environ.get("MELOGAN_PRECISION", "fp32"), device=device,
                    ed_state_dict=ed_state, process_group=pg, seed_offset=rank,
                    sync_bn=os.environ.get("MELOGAN_SYNC_BN") == "1" or bool(cfg.get("SYNC_BN", False)))
    d_notes, d_numeric, d_labels = (torch.from_numpy(a).to(device) for a in (notes, numeric, labels))   # 7 MB: resident
    d_cond = None
    if tr.cond_dim:                             # INTEGRATION_MODE 'conditioning': AE latents from src/ae/encode.py
        d_cond = torch.from_numpy(load_encoder_feats(cfg.get('ENCODER_FEATS_TRAIN'), len(notes), tr.cond_dim)).to(device)
    start_epoch = 1
    if args.resume:
        ck = torch.load(args.resume, map_location=device)
        tr.load_state_dict(ck)
        start_epoch = int(ck.get('epoch', 0)) + 1
        print(f"[INFO] Resumed from {args.resume} (epoch {start_epoch - 1})")

    writer = None
    if rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(log_dir=cfg['LOG_DIR'])
        except Exception:
            writer = None
        os.makedirs(cfg['CHECKPOINT_DIR'], exist_ok=True)
        os.makedirs(cfg['SAMPLE_DIR'], exist_ok=True)
    critic_iters = int(cfg.get('CRITIC_ITERS', 5))
    gen = torch.Generator(device="cpu").manual_seed(int(cfg.get("SEED", 42)))
    print("Starting WGAN-GP Training with Emotion Guidance...")
    steps = len(notes) // B                       # drop_last=True
    # Every CRITIC_ITERS consecutive batches (train_gan.py:168-251: a critic step each, a generator step on the last) are
    # ONE CUDA-graph replay over static input buffers -- at the reference's batch 32 the cycle is ~700 launches of a few
    # microseconds each, i.e. launch-bound when issued eagerly.  The first cycle runs eagerly (allocates scratch), then the
    # cycle is captured; the batches left over at the end of an epoch are critic steps only, as in the reference loop.
    use_graph, captured = os.environ.get("MELOGAN_NO_GRAPH") is None, False

    def batch_index(perm, b):
        return perm[b * B:(b + 1) * B].view(world, -1)[rank]

    def eager_batch(perm, b, with_g):
        idx = batch_index(perm, b)
        real, num = d_notes[idx].contiguous(), d_numeric[idx].contiguous()
        cond = d_cond[idx].contiguous() if d_cond is not None else None
        tr.critic_step(real, num, cond=cond)
        if with_g:
            tr.generator_step(num, d_labels[idx].contiguous(), cond=cond)

    for epoch in range(1, cfg['EPOCHS'] + 1):
        perm = torch.randperm(len(notes), generator=gen).to(device)          # shuffle=True, same order on every rank
        if epoch < start_epoch:
            continue                                                          # replay the shuffle stream up to the resume point
        b = 0
        while b < steps:
            if b + critic_iters > steps:                                      # tail of the epoch: critic steps only
                eager_batch(perm, b, False)
                b += 1
                continue
            if captured:
                for i in range(critic_iters):
                    idx = batch_index(perm, b + i)
                    torch.index_select(d_notes, 0, idx, out=tr.s_reals[i])
                    torch.index_select(d_numeric, 0, idx, out=tr.s_numerics[i])
                    if d_cond is not None:
                        torch.index_select(d_cond, 0, idx, out=tr.s_conds[i])
                torch.index_select(d_labels, 0, idx, out=tr.s_labels)
                tr.replay_cycle()
            else:
                for i in range(critic_iters):
                    eager_batch(perm, b + i, i == critic_iters - 1)
                if use_graph:
                    tr.capture_cycle()
                    captured = True
            b += critic_iters
        d_loss, g_adv, g_emo = tr.epoch_means()
        saving = epoch % cfg.get('SAVE_FREQ', 5) == 0
        ck_state = tr.state_dict() if saving else None        # collective: averages the BatchNorm running stats over ranks
        if rank == 0:
            print(f"Epoch {epoch}/{cfg['EPOCHS']} | D_loss: {d_loss:.4f} | G_adv: {g_adv:.4f} | G_emo: {g_emo:.4f}")
            if writer is not None:
                writer.add_scalar("Loss/Critic", d_loss, epoch)
                writer.add_scalar("Loss/Generator_Adv", g_adv, epoch)
                writer.add_scalar("Loss/Generator_Emo", g_emo, epoch)
            if saving:
                ck = {'epoch': epoch}
                ck.update(ck_state)
                torch.save(ck, os.path.join(cfg['CHECKPOINT_DIR'], f"gan_epoch{epoch:04d}.pth"))
    tr.sync_bn_running_stats_()
    if rank == 0:
        torch.save({'G': tr.G.state_dict(), 'E_num': tr.E_num.state_dict()},
                   os.path.join(cfg['CHECKPOINT_DIR'], "gan_final.pth"))
        if writer is not None:
            writer.close()
        print("Training Complete.")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
