"""Deterministic synthetic parameters and inputs -- TEST INFRASTRUCTURE ONLY (see gsn_oracle.py header).

Weights are drawn with numpy's legacy RandomState so that a fixture only has to record a seed: the
golden generator loads them into the reference model with `load_state_dict(strict=True)`, the tests
load the very same arrays into the B200 facade and into the oracle.  Distributions follow the
reference initialisers (U(+-1/sqrt(H)) for the cell, ESN:126-130; nn.Linear default for proj), except
BatchNorm, whose affine and running statistics get non-trivial values so eval-mode BN is exercised
(scale = weight/sqrt(var+eps) is kept <= 1 so the leaky membrane recursion stays bounded).
"""
from __future__ import annotations

import math

import numpy as np

# recipes/intel_ndns/spiking_fullsubnet/baseline_m.toml:36-56 (surface A, "M")
CFG_M = dict(
    n_fft=512, hop_length=128, win_length=512, fdrc=0.5, fb_input_size=64, fb_hidden_size=320,
    fb_num_layers=2, fb_proj_size=64, fb_output_activate_function=False, sb_hidden_size=224,
    sb_num_layers=2, freq_cutoffs=[0, 32, 128, 256], df_orders=[5, 3, 1],
    center_freq_sizes=[4, 32, 64], neighbor_freq_sizes=[15, 15, 15], use_pre_layer_norm_fb=True,
    use_pre_layer_norm_sb=True, bn=True, shared_weights=True, sequence_model="GSN", num_spks=1,
)
# recipes/intel_ndns/spiking_fullsubnet_freeze_phase/baseline_s.toml:45-65 expressed on surface A
# (SURVEY.md section 8 size table): "spiking_fullsubnet-S"
CFG_S = dict(CFG_M, fb_hidden_size=240, sb_hidden_size=160, df_orders=[3, 1, 1])
# .../baseline_l.toml (zoo config__2023_07_27 toml:78-99) expressed on surface A: "-L"
CFG_L = dict(CFG_M, fb_hidden_size=320, sb_hidden_size=256, freq_cutoffs=[0, 32, 128, 192, 256],
             df_orders=[5, 3, 1, 1], center_freq_sizes=[2, 4, 32, 64],
             neighbor_freq_sizes=[15, 15, 15, 15])
# .../baseline_xl.toml: M sizes with unshared gate weights
CFG_XL = dict(CFG_M, shared_weights=False)
# recipes/intel_ndns/cirm_gsn/default.toml:36-50
CFG_CIRM = dict(n_fft=512, hop_length=128, win_length=512, fdrc=0.5, input_size=257, hidden_size=268,
                num_layers=4, proj_size=257, output_activate_function=False, df_order=3,
                use_pre_layer_norm_fb=True, bn=True, shared_weights=True, sequence_model="GSN",
                num_spks=1)

CONFIGS = {"M": CFG_M, "S": CFG_S, "L": CFG_L, "XL": CFG_XL}


def tiny_cfg(**over):
    """A structurally complete surface-A config small enough for per-step golden traces."""
    cfg = dict(
        n_fft=64, hop_length=16, win_length=64, fdrc=0.5, fb_input_size=8, fb_hidden_size=48,
        fb_num_layers=2, fb_proj_size=8, fb_output_activate_function=False, sb_hidden_size=40,
        sb_num_layers=2, freq_cutoffs=[0, 8, 24, 32], df_orders=[3, 2, 1],
        center_freq_sizes=[2, 4, 8], neighbor_freq_sizes=[3, 3, 3], use_pre_layer_norm_fb=True,
        use_pre_layer_norm_sb=True, bn=True, shared_weights=True, sequence_model="GSN", num_spks=1,
    )
    cfg.update(over)
    return cfg


def _seq_model_params(rs, prefix, K, H, L, P, shared, bn, ln):
    p = {}
    f32 = np.float32
    if ln:
        p[prefix + "pre_layer_norm.weight"] = rs.uniform(0.7, 1.3, K).astype(f32)
        p[prefix + "pre_layer_norm.bias"] = rs.normal(0, 0.1, K).astype(f32)
    g = 1 if shared else 2
    s = 1.0 / math.sqrt(H)
    for l in range(L):
        q = f"{prefix}sequence_model.layers.{l}.cell."
        kin = K if l == 0 else H
        p[q + "weight_ih"] = rs.uniform(-s, s, (g * H, kin)).astype(f32)
        p[q + "weight_hh"] = rs.uniform(-s, s, (g * H, H)).astype(f32)
        p[q + "bias_ih"] = rs.uniform(-s, s, 2 * H).astype(f32)
        if bn:
            p[q + "batchnorm.weight"] = rs.uniform(0.6, 1.0, H).astype(f32)
            p[q + "batchnorm.bias"] = rs.normal(0, 0.1, H).astype(f32)
            p[q + "batchnorm.running_mean"] = rs.normal(0, 0.1, H).astype(f32)
            p[q + "batchnorm.running_var"] = rs.uniform(1.0, 2.0, H).astype(f32)
            p[q + "batchnorm.num_batches_tracked"] = np.asarray(7, dtype=np.int64)
    if P > 0:
        p[prefix + "proj.weight"] = rs.uniform(-s, s, (P, H)).astype(f32)
        p[prefix + "proj.bias"] = rs.uniform(-s, s, P).astype(f32)
    return p


def make_params(cfg, seed):
    """state_dict (numpy) for surface A `SpikingFullSubNet(**cfg)` (MSF:349-413)."""
    rs = np.random.RandomState(seed)
    shared, bn = cfg.get("shared_weights", False), cfg.get("bn", False)
    S = cfg.get("num_spks", 1)
    p = _seq_model_params(rs, "fb_model.", cfg["fb_input_size"], cfg["fb_hidden_size"],
                          cfg["fb_num_layers"], cfg["fb_proj_size"], shared, bn,
                          cfg.get("use_pre_layer_norm_fb", True))
    for i, (ctr, nbr, df) in enumerate(zip(cfg["center_freq_sizes"], cfg["neighbor_freq_sizes"],
                                           cfg["df_orders"])):
        p.update(_seq_model_params(rs, f"sb_model.sb_models.{i}.", 2 * ctr + 2 * nbr,
                                   cfg["sb_hidden_size"], cfg["sb_num_layers"], 2 * ctr * df * S,
                                   shared, bn, cfg.get("use_pre_layer_norm_sb", True)))
    return p


def make_params_cirm(cfg, seed):
    """state_dict (numpy) for `cirm_gsn.Model(**cfg)` (CGN:162-204)."""
    rs = np.random.RandomState(seed)
    P = cfg["proj_size"] * cfg.get("num_spks", 2) * cfg["df_order"] * 2
    return _seq_model_params(rs, "fb_model.", cfg["input_size"], cfg["hidden_size"], cfg["num_layers"],
                             P, cfg.get("shared_weights", False), cfg.get("bn", False),
                             cfg.get("use_pre_layer_norm_fb", True))


def make_wave(batch, num_samples, seed, sr=16000):
    """Synthetic noisy 'speech' (SURVEY.md 8d): white noise + a slow chirp so magnitudes are not
    white.  float32 [batch, num_samples]."""
    rs = np.random.RandomState(seed)
    t = np.arange(num_samples, dtype=np.float64) / sr
    x = 0.05 * rs.standard_normal((batch, num_samples))
    f0 = rs.uniform(150, 400, (batch, 1))
    f1 = rs.uniform(1000, 3000, (batch, 1))
    dur = max(num_samples / sr, 1e-3)
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    x = x + 0.02 * np.sin(phase) * (1 + np.sin(2 * np.pi * 3.0 * t))
    return x.astype(np.float32)


def make_mag(batch, num_bins, num_frames, seed):
    """Synthetic magnitude spectrogram (no STFT needed): |N(0,1)| * smooth envelope. float32."""
    rs = np.random.RandomState(seed)
    env = 0.2 + rs.uniform(0, 1, (batch, num_bins, 1)) * np.linspace(1.0, 0.3, num_bins)[None, :, None]
    return (np.abs(rs.standard_normal((batch, num_bins, num_frames))) * env).astype(np.float32)


# model_zoo/intel_ndns/spike_fsb/baseline_s/baseline_s.toml [model_g.args] (surface B, zoo "S")
CFG_ZOO_S = dict(sr=16000, fdrc=0.5, n_fft=512, fb_freqs=64, hop_length=128, win_length=512, num_freqs=256,
                 sequence_model="GSU", fb_hidden_size=240, fb_output_activate_function=False,
                 freq_cutoffs=[32, 128], sb_df_orders=[3, 1, 1], sb_num_center_freqs=[4, 32, 64],
                 sb_num_neighbor_freqs=[15, 15, 15], fb_num_center_freqs=[4, 32, 64],
                 fb_num_neighbor_freqs=[0, 0, 0], sb_hidden_size=160, sb_output_activate_function=False,
                 norm_type="offline_laplace_norm", shared_weights=True, bn=True)


# model_zoo/intel_ndns/spike_fsb/baseline_l/config__2023_07_27--22_13_36.toml [model_g.args] (surface B, zoo "L")
CFG_ZOO_L = dict(CFG_ZOO_S, fb_hidden_size=320, sb_hidden_size=256, freq_cutoffs=[32, 128, 192],
                 sb_df_orders=[5, 3, 1, 1], sb_num_center_freqs=[2, 4, 32, 64],
                 sb_num_neighbor_freqs=[15, 15, 15, 15], fb_num_center_freqs=[2, 4, 32, 64],
                 fb_num_neighbor_freqs=[0, 0, 0, 0])


def tiny_cfg_b(**over):
    """Structurally complete surface-B (`Separator`) config, small enough for golden traces."""
    cfg = dict(sr=16000, fdrc=0.5, n_fft=64, fb_freqs=8, hop_length=16, win_length=64, num_freqs=32,
               sequence_model="GSU", fb_hidden_size=48, fb_output_activate_function=False, freq_cutoffs=[8, 24],
               sb_df_orders=[3, 2, 1], sb_num_center_freqs=[2, 4, 8], sb_num_neighbor_freqs=[3, 3, 3],
               fb_num_center_freqs=[2, 4, 8], fb_num_neighbor_freqs=[0, 0, 0], sb_hidden_size=40,
               sb_output_activate_function=False, norm_type="offline_laplace_norm", shared_weights=True, bn=True)
    cfg.update(over)
    return cfg


def make_params_b(cfg, seed):
    """state_dict (numpy) for surface B `Separator(**cfg)`: surface-A layout without LayerNorm and with
    `fc_output_layer` in place of `proj`."""
    rs = np.random.RandomState(seed)
    shared, bn = cfg.get("shared_weights", False), cfg.get("bn", False)
    p = _seq_model_params(rs, "fb_model.", cfg["fb_freqs"], cfg["fb_hidden_size"], 2, cfg["fb_freqs"], shared, bn,
                          False)
    for i, (ctr, nbr, fc, fn, df) in enumerate(zip(cfg["sb_num_center_freqs"], cfg["sb_num_neighbor_freqs"],
                                                   cfg["fb_num_center_freqs"], cfg["fb_num_neighbor_freqs"],
                                                   cfg["sb_df_orders"])):
        p.update(_seq_model_params(rs, f"sb_model.sb_models.{i}.", (ctr + 2 * nbr) + (fc + 2 * fn),
                                   cfg["sb_hidden_size"], 2, 2 * ctr * df, shared, bn, False))
    return {k.replace(".proj.", ".fc_output_layer."): v for k, v in p.items()}
