"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np

from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SURFACE_A = ["tiny_shared_bn", "tiny_unshared_nobn", "tiny_spk2_tanh", "cfg1_baseline_m_1s"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["cfg"] = json.loads(str(d["cfg"]))
    d["seed"] = int(d["seed"])
    return d


def unpack(bits, H):
    """inverse of make_golden.pack: packed uint8 [..., ceil(H/8)] -> float32 {0,1} [..., H]."""
    return np.unpackbits(bits, axis=-1)[..., :H].astype(np.float32)


def golden_params(g):
    return synth.make_params(g["cfg"], g["seed"])


def spike_flip_stats(a, b):
    """fraction of differing spikes and index of the first frame with a difference (or -1)."""
    diff = a != b
    frac = float(diff.mean())
    first = -1
    if diff.any():
        first = int(np.argmax(diff.reshape(diff.shape[0], -1).any(axis=1)))
    return frac, first


def load_golden_weights(name):
    z = np.load(os.path.join(GOLDEN, name + "_weights.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def loss_waveforms(n=2, length=12000, seed=20220815):
    """Seeded (estimate, clean) waveform pair of the loss fixture (tests/golden/make_golden_loss.py)."""
    rs = np.random.RandomState(seed)
    t = np.arange(length) / 16000.0
    clean = (0.1 * np.sin(2 * np.pi * (200 + 300 * t) * t))[None, :].repeat(n, 0) * rs.uniform(0.5, 1.5, (n, 1))
    clean = clean.astype(np.float32)
    est = (clean + 0.03 * rs.standard_normal(clean.shape)).astype(np.float32)
    return est, clean
