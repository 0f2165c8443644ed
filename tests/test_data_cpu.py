"""the on-disk formats either side of the step -- the .npy fast path and the per-file .npz slow path of the
reference's GANDataset (src/gan/dataset.py:30-56,58-112,176-200) -- read into the arrays the trainer keeps resident."""
import os

import numpy as np
import pytest

from src.gan.train_gan import load_split_arrays, load_split_npz


def _write_split(tmp, n=6):
    rng = np.random.default_rng(1)
    proc = tmp / "processed"
    proc.mkdir()
    rows, want = ["filename,emotion"], []
    emos = ["happy", "sad", "angry", "calm", "sad", "happy"]
    for i in range(n):
        notes = rng.uniform(-1, 1, (512, 4)).astype(np.float32)
        kw = {"notes": notes, "tempo": 120.0, "filename": f"song{i}.mid"}
        if i % 3 != 2:
            kw["numeric_features"] = rng.normal(size=6 if i % 2 == 0 else 4).astype(np.float32)   # 4 -> zero-padded to 6
        if i == 1:
            kw["mood"] = "calm"                       # the archive's own mood wins over the manifest column
        np.savez(proc / f"song{i}_processed.npz", **kw)
        rows.append(f"song{i}.mid,{emos[i]}")
        want.append((notes, kw.get("numeric_features"), "calm" if i == 1 else emos[i]))
    (tmp / "splits").mkdir()
    csv = tmp / "splits" / "train_split.csv"
    csv.write_text("\n".join(rows) + "\n")
    return {"PROCESSED_DIR": str(proc), "SPLITS_DIR": str(tmp / "splits"), "NUMERIC_INPUT_DIM": 6}, str(csv), want


def test_npz_slow_path_matches_the_archives(tmp_path):
    cfg, csv, want = _write_split(tmp_path)
    notes, numeric, labels = load_split_npz(cfg, csv)
    assert notes.shape == (6, 512, 4) and numeric.shape == (6, 6) and labels.dtype == np.int64
    idx = {"happy": 0, "sad": 1, "angry": 2, "calm": 3}
    for i, (n, f, mood) in enumerate(want):
        np.testing.assert_array_equal(notes[i], n)
        exp = np.zeros(6, np.float32)
        if f is not None:
            exp[:f.size] = f
        np.testing.assert_array_equal(numeric[i], exp)
        assert labels[i] == idx[mood]


def test_fast_path_is_preferred_and_falls_back(tmp_path):
    cfg, csv, want = _write_split(tmp_path)
    a = load_split_arrays(cfg, csv)                        # no .npy arrays yet -> slow path
    d = tmp_path / "splits" / "train_split"
    d.mkdir()
    np.save(d / "notes.npy", a[0][::-1].copy())
    np.save(d / "emotion.npy", np.array(["calm"] * 6))
    np.save(d / "numeric_features.npy", a[1])
    b = load_split_arrays(cfg, csv)                        # pre-saved arrays win (dataset.py:30-56)
    np.testing.assert_array_equal(b[0], a[0][::-1])
    assert (b[2] == 3).all()
    (tmp_path / "splits" / "bad.csv").write_text("x,y\n1,2\n")
    with pytest.raises(KeyError):
        load_split_npz(cfg, str(tmp_path / "splits" / "bad.csv"))


def test_labels_outside_the_classes_raise_like_cross_entropy():
    """emotion_to_index returns -1 for unknown moods; nn.CrossEntropyLoss raises on such targets."""
    import pytest
    from src.gan import train_gan as T
    T.check_labels(np.array([0, 1, 2, 3]), "ok")
    with pytest.raises(IndexError):
        T.check_labels(np.array([0, -1, 2]), "bad")
    with pytest.raises(IndexError):
        T.check_labels(np.array([0, 4]), "bad")


def test_encoder_feats_follow_the_config_key_and_fall_back_to_zeros(tmp_path):
    """cfg['ENCODER_FEATS_TRAIN'] like the reference's prepare_dataset; missing file / length mismatch ->
    zero latents with a warning (dataset.py:47-54,169-171)."""
    from src.gan import train_gan as T
    f = tmp_path / "encoder_feats.npy"
    np.save(f, np.arange(12, dtype=np.float32).reshape(3, 4))
    assert np.array_equal(T.load_encoder_feats(str(f), 3, 4), np.arange(12, dtype=np.float32).reshape(3, 4))
    assert not T.load_encoder_feats(str(f), 5, 4).any() and T.load_encoder_feats(str(f), 5, 4).shape == (5, 4)
    assert not T.load_encoder_feats(str(tmp_path / "missing.npy"), 2, 4).any()
    assert not T.load_encoder_feats(None, 2, 4).any()


def test_split_directory_convention_and_plateau_scheduler(tmp_path):
    import torch
    from melogan.aux_trainers import ReduceOnPlateau, find_split_dir
    (tmp_path / "train_split").mkdir()
    (tmp_path / "val").mkdir()
    assert find_split_dir(str(tmp_path), "train").endswith("train_split")
    assert find_split_dir(str(tmp_path), "train_split").endswith("train_split")
    assert find_split_dir(str(tmp_path), "val_split").endswith("val")

    class Opt:
        lr = 1e-3
    o = Opt()
    mine = ReduceOnPlateau(o, factor=0.5, patience=2, threshold=1e-4, min_lr=1e-4)
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.SGD([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(topt, factor=0.5, patience=2, threshold=1e-4, min_lr=1e-4)
    for m in [1.0, 0.9, 0.95, 0.95, 0.95, 0.95, 0.8, 0.81, 0.82, 0.83, 0.84, 0.85, 0.86, 0.87, 0.88, 0.89, 0.9, 0.91]:
        mine.step(m); ref.step(m)
        assert abs(o.lr - topt.param_groups[0]["lr"]) < 1e-12, (m, o.lr, topt.param_groups[0]["lr"])


def test_conv_encoder_dummy_pass_updates_running_stats_like_the_reference():
    """reference src/ae/train_ae.py:75-77 + model.py:27-44: the zero dummy pass runs the conv stack twice in train mode."""
    import torch
    import torch.nn as nn
    from src.ae.model import ConvEncoder
    torch.manual_seed(0)
    enc = ConvEncoder(in_channels=4, latent_dim=8, hidden_dim=32)
    ref = nn.Sequential(*[m for ci, co in ((4, 32), (32, 64), (64, 128))
                          for m in (nn.Conv1d(ci, co, 5, 2, 2), nn.BatchNorm1d(co), nn.ReLU(inplace=True))])
    ref.load_state_dict(enc.conv.state_dict())
    with torch.no_grad():
        enc(torch.zeros(1, 64, 4))
        for _ in range(2):                       # forward's own pass + the one inside build_linear
            ref(torch.zeros(1, 4, 64))
    for k, v in ref.state_dict().items():
        got = enc.conv.state_dict()[k]
        assert torch.allclose(got.float(), v.float(), rtol=1e-5, atol=5e-6), k     # (x - mean) / sqrt(0 + eps) amplifies fp32 rounding noise of the reference pass to ~1e-6
    assert enc._linear is not None and enc._linear[1].in_features == 128 * 8
