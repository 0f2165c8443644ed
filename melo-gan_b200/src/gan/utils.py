"""GAN training / inference utilities, B200-native.

Drop-in for the reference's src/gan/utils.py (same names, signatures and constants).  The two functions
on the hot path run on the GPU:
  compute_gradient_penalty  -> mg_gradient_penalty: critic forward on the interpolates, d(score)/d(x),
                               the penalty and its double backward in fused sm_100a kernels
  save_piano_roll_to_midi   -> mg_extract_notes_gan: velocity gate, quantisation, scale snap and the
                               sequential float32 onset clock, bit-exact with the reference's row loop
"""
import os
import random

import numpy as np
import torch

from melogan import engine as E
from melogan import midi as _midi
from melogan import notes as _notes
from melogan import runtime as R

# --- musical scale tables (reference utils.py:14-28) ---
SCALES = {k: list(v) for k, v in _notes.SCALES.items()}
NOTE_NAMES = ['C', 'C#', 'D', 'D#', 'E', 'F', 'F#', 'G', 'G#', 'A', 'A#', 'B']
MAX_BEAT_TIME = 4.0

# General MIDI programs the demo can ask for (stands in for pretty_midi.instrument_name_to_program)
_GM_PROGRAMS = {"Acoustic Grand Piano": 0, "Bright Acoustic Piano": 1, "Electric Grand Piano": 2, "Electric Piano 1": 4,
                "Harpsichord": 6, "Celesta": 8, "Music Box": 10, "Vibraphone": 11, "Church Organ": 19,
                "Acoustic Guitar (nylon)": 24, "Acoustic Guitar (steel)": 25, "Electric Guitar (clean)": 27,
                "Overdriven Guitar": 29, "Distortion Guitar": 30, "Acoustic Bass": 32, "Violin": 40, "Cello": 42,
                "String Ensemble 1": 48, "Trumpet": 56, "Flute": 73, "Pad 1 (new age)": 88}


def seed_everything(seed=42):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def weights_init(m):
    """N(0, 0.02) weights / zero biases for every module whose class name contains 'Conv' or 'Linear'."""
    cls = m.__class__.__name__
    if 'Conv' in cls or 'Linear' in cls:
        w = getattr(m, 'weight', None)
        if isinstance(w, torch.Tensor):
            torch.nn.init.normal_(w.data, 0.0, 0.02)
        if getattr(m, 'bias', None) is not None:
            torch.nn.init.constant_(m.bias.data, 0.0)


def load_ae_decoder_into_generator(ae_ckpt_path, generator):
    """Copies shape-compatible 'decoder.*' tensors of an AE checkpoint into generator.decoder (reference :47-61)."""
    if not os.path.exists(ae_ckpt_path):
        print(f"[WARN] AE full checkpoint not found at {ae_ckpt_path}")
        return False
    state = torch.load(ae_ckpt_path, map_location='cpu').get('model_state', None)
    if state is None:
        return False
    own = generator.decoder.state_dict()
    cut = len('decoder.')
    picked = {k[cut:]: v for k, v in state.items()
              if k.startswith('decoder.') and k[cut:] in own and own[k[cut:]].shape == v.shape}
    own.update(picked)
    generator.decoder.load_state_dict(own)
    print(f"[INFO] loaded {len(picked)} decoder params from AE ckpt into generator.decoder")
    return True


def emotion_to_index(emotion):
    """happy=0, sad=1, angry=2, calm=3; one-hot / index inputs pass through; anything else -> -1."""
    if emotion is None:
        return -1
    if isinstance(emotion, (list, tuple, np.ndarray)):
        arr = np.array(emotion)
        return int(np.argmax(arr)) if (arr.ndim == 1 and arr.size == 4) else int(arr)
    if isinstance(emotion, str):
        return {'happy': 0, 'sad': 1, 'angry': 2, 'calm': 3}.get(emotion.lower(), -1)
    try:
        return int(emotion)
    except Exception:
        return -1


class _GradientPenaltyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, D, real, fake, emb, alpha, *params):
        eng = D._engine(real)
        P = R.params_of(D, E.D_KEYS)
        G = R.fresh_grads(P, E.D_KEYS)
        eng.bind(E.MOD_D, P, G)
        m = eng.gradient_penalty(R.as_f32c(real), R.as_f32c(fake), R.as_f32c(emb) if emb is not None else None,
                                 R.as_f32c(alpha))
        ctx.D = D
        ctx.grads = [G[k] for k in E.D_KEYS]       # d(GP)/d(theta), from the fused double backward
        return m[1].clone()

    @staticmethod
    def backward(ctx, dgp):
        named = dict(ctx.D.named_parameters())
        grads = tuple((g * dgp) if named[k].requires_grad else None for k, g in zip(E.D_KEYS, ctx.grads))
        return (None, None, None, None, None) + grads


def compute_gradient_penalty(D, real_samples, fake_samples, numeric_embedding, device):
    """WGAN-GP penalty mean((||d D(x_hat)/d x_hat||_2 - 1)^2), x_hat = a*real + (1-a)*fake, a ~ U[0,1) per sample.
    Same RNG draw as the reference (torch.rand(B, 1, 1, device=device)); differentiable w.r.t. D's parameters."""
    alpha = torch.rand(real_samples.size(0), 1, 1, device=device)
    from .models import Discriminator
    if isinstance(D, Discriminator) and real_samples.is_cuda:
        named = dict(D.named_parameters())
        return _GradientPenaltyFn.apply(D, real_samples, fake_samples, numeric_embedding, alpha.reshape(-1),
                                        *[named[k] for k in E.D_KEYS])
    raise NotImplementedError("compute_gradient_penalty runs on the native critic (src.gan.models.Discriminator) with "
                              "CUDA inputs; there is no CPU / generic-autograd fallback")


def extract_notes(notes_array, bpm=120.0, scale='major', root_key=0):
    """Batched GPU form of the row loop: (R, T, 4) or (T, 4) array/tensor -> melogan.notes.NoteBatch."""
    x = notes_array
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if x.dim() == 2:
        x = x.unsqueeze(0)
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("note extraction runs on a CUDA (sm_100a) device; there is no CPU fallback")
        x = x.cuda()
    return _notes.extract_notes_gan(x.to(torch.float32).contiguous(), bpm=max(60, min(bpm, 180)), scale=scale,
                                    root_key=root_key)


def save_piano_roll_to_midi(notes_array, output_path, fs=100, bpm=120.0, scale='major', root_key=0,
                            instrument_name='Acoustic Grand Piano'):
    """GAN output (T, 4) with columns (pitch, velocity, duration, step), normalised to about [-1, 1] -> MIDI file,
    with scale snapping, dynamic tempo and instrument selection (reference utils.py:95-161)."""
    bpm = max(60, min(bpm, 180))
    program = _GM_PROGRAMS.get(instrument_name)
    if program is None:
        print(f"[WARN] Instrument '{instrument_name}' not found. Defaulting to Piano.")
        program = 0
    batch = extract_notes(notes_array, bpm=bpm, scale=scale, root_key=root_key)
    n = int(batch.counts[0])
    vel, pit = batch.velocity[0, :n].cpu().tolist(), batch.pitch[0, :n].cpu().tolist()
    st, en = batch.start[0, :n].cpu().tolist(), batch.end[0, :n].cpu().tolist()
    _midi.write_midi(output_path, zip(vel, pit, st, en), bpm=bpm, program=program)
    print(f"[INFO] Saved MIDI ({instrument_name} | {NOTE_NAMES[root_key]} {scale}) to {output_path}")
