"""End-to-end smoke of the drop-in command lines on synthetic data (GPU box): ED training -> VAE training -> latent
encoding -> GAN training in 'conditioning' mode with resume -> roll_to_midi.  Writes everything under a temp dir."""
import os, subprocess, sys, tempfile
import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "melo-gan_b200")
tmp = tempfile.mkdtemp(prefix="melogan_cli_")
rng = np.random.default_rng(0)
emos = np.array(["happy", "sad", "angry", "calm"])
for split, n in (("train", 160), ("val", 64)):
    d = os.path.join(tmp, "splits", f"{split}_split")
    os.makedirs(d)
    np.save(os.path.join(d, "notes.npy"), rng.uniform(-1, 1, (n, 512, 4)).astype(np.float32))
    np.save(os.path.join(d, "emotion.npy"), emos[rng.integers(0, 4, n)])
    np.save(os.path.join(d, "numeric_features.npy"), rng.normal(size=(n, 6)).astype(np.float32))
    os.symlink(d, os.path.join(tmp, "splits", split))          # train_ae / encode read <SPLITS_DIR>/<split>/notes.npy
    open(os.path.join(tmp, "splits", f"{split}_split.csv"), "w").write("npz_path\n")

def cfg_of(name, **over):
    c = yaml.safe_load(open(os.path.join(PKG, "config", name)))
    c.update(over)
    p = os.path.join(tmp, name)
    yaml.safe_dump(c, open(p, "w"))
    return p

def run(*args):
    print("+", " ".join(args), flush=True)
    subprocess.check_call([sys.executable, *args], cwd=PKG, env=dict(os.environ, PYTHONPATH=PKG))

S = os.path.join(tmp, "splits")
ed = cfg_of("ed_config.yaml", num_epochs=2, batch_size=32, checkpoint_dir=os.path.join(tmp, "ed"),
            train_split_csv=os.path.join(S, "train_split.csv"), val_split_csv=os.path.join(S, "val_split.csv"))
run("-m", "src.emotion_discriminator.train_ed", "--config", ed)
ae = cfg_of("ae_config.yaml", EPOCHS=2, LATENT_DIM=64, SPLITS_DIR=S, CHECKPOINT_DIR=os.path.join(tmp, "ae"))
run("-m", "src.ae.train_ae", "--config", ae)
run("-m", "src.ae.encode", "--model", os.path.join(tmp, "ae", "ae_best.pth"), "--notes", os.path.join(S, "train", "notes.npy"),
    "--out_file", os.path.join(S, "train_split", "encoder_feats.npy"), "--config", ae)
gan = cfg_of("gan_config.yaml", EPOCHS=2, SAVE_FREQ=1, INTEGRATION_MODE="conditioning", SPLITS_DIR=S,
             TRAIN_SPLIT=os.path.join(S, "train_split.csv"), CHECKPOINT_DIR=os.path.join(tmp, "gan"),
             LOG_DIR=os.path.join(tmp, "gan_logs"), SAMPLE_DIR=os.path.join(tmp, "gan_samples"))
ed_ck = os.path.join(tmp, "ed", yaml.safe_load(open(ed))["save_name"])
run("-m", "src.gan.train_gan", "--config", gan, "--ed_config", ed, "--ed_ckpt", ed_ck)
gan3 = cfg_of("gan_config.yaml", EPOCHS=3, SAVE_FREQ=1, INTEGRATION_MODE="conditioning", SPLITS_DIR=S,
              TRAIN_SPLIT=os.path.join(S, "train_split.csv"), CHECKPOINT_DIR=os.path.join(tmp, "gan"),
              LOG_DIR=os.path.join(tmp, "gan_logs"), SAMPLE_DIR=os.path.join(tmp, "gan_samples"))
run("-m", "src.gan.train_gan", "--config", gan3, "--ed_config", ed, "--ed_ckpt", ed_ck, "--resume",
    os.path.join(tmp, "gan", "gan_epoch0002.pth"))
roll = os.path.join(tmp, "roll.npy")
np.save(roll, rng.uniform(0, 1, (512, 4)).astype(np.float32))
run(os.path.join("tools", "roll_to_midi.py"), roll)
print("cli smoke ok:", sorted(os.listdir(os.path.join(tmp, "gan"))))
