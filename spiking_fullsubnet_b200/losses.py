"""Loss terms of the intel_ndns recipes: `freq_MAE`, `mag_MAE` (audiozen/loss.py:138-190)
and `SISNRLoss` (audiozen/loss.py:11-40), combined as recipes/intel_ndns/spiking_fullsubnet/trainer.py:33-37 does.

The reference evaluates `freq_MAE` and `mag_MAE` separately, i.e. FOUR 2048-point STFTs per step for two signals;
`ndns_training_loss` computes each signal's STFT once and takes all three spectral means from the same pair of
spectrograms.  Same torch.stft calls and the same expressions, so the values equal the reference's bit for bit;
the gradient flows through torch.stft's own backward.  (Device-agnostic torch code: this is not the GSN hot path,
whose kernels have no CPU form.)
"""
from __future__ import annotations

import torch
import torch.nn as nn

_WINDOWS = {}


def _hann(win, device):
    key = (win, str(device))
    if key not in _WINDOWS:
        _WINDOWS[key] = torch.hann_window(win).to(device).float()
    return _WINDOWS[key]


def _spec(y, win, stride):
    """torch.stft exactly as audiozen/loss.py:139-145 calls it (center=True, reflect pad, periodic hann)."""
    return torch.stft(y.reshape(-1, y.shape[-1]), n_fft=win, hop_length=stride, window=_hann(win, y.device),
                      return_complex=True)


def _band_limited(fn, est_spec, tgt_spec, srs, sudo_sr):
    """The sample-rate-aware branch (audiozen/loss.py:156-164, 183-189): per item, bins below sr/sudo_sr only."""
    loss = 0
    for i, sr in enumerate(srs):
        max_freq = int(est_spec.shape[-2] * sr / sudo_sr)
        loss = loss + fn(est_spec[i][:max_freq], tgt_spec[i][:max_freq])
    return loss / len(srs)


def _freq_term(e, t):
    return (e.real - t.real).abs().mean() + (e.imag - t.imag).abs().mean()


def _mag_term(e, t):
    return (e.abs() - t.abs()).abs().mean()


def freq_MAE(estimation, target, win=2048, stride=512, srs=None, sudo_sr=None):
    """audiozen/loss.py:138-164."""
    e, t = _spec(estimation, win, stride), _spec(target, win, stride)
    return _freq_term(e, t) if srs is None else _band_limited(_freq_term, e, t, srs, sudo_sr)


def mag_MAE(estimation, target, win=2048, stride=512, srs=None, sudo_sr=None):
    """audiozen/loss.py:167-190."""
    e, t = _spec(estimation, win, stride), _spec(target, win, stride)
    return _mag_term(e, t) if srs is None else _band_limited(_mag_term, e, t, srs, sudo_sr)


class SISNRLoss(nn.Module):
    """audiozen/loss.py:11-40 (mean SI-SNR in dB over the batch; `return_neg` flips the sign)."""

    def __init__(self, return_neg=False):
        super().__init__()
        self.return_neg = return_neg

    def forward(self, input, target):
        if not torch.is_tensor(input):
            input = torch.from_numpy(input)
        if not torch.is_tensor(target):
            target = torch.from_numpy(target)
        if input.shape != target.shape:
            raise RuntimeError(f"Dimension mismatch when calculating SI-SNR, {input.shape=} vs {target.shape=}")
        eps = torch.finfo(input.dtype).eps
        s_input = input - torch.mean(input, dim=-1, keepdim=True)
        s_target = target - torch.mean(target, dim=-1, keepdim=True)
        dot = torch.sum(s_target * s_input, dim=-1, keepdim=True)
        proj = dot * s_target / torch.sum(s_target ** 2, dim=-1, keepdim=True)
        e_noise = s_input - proj
        sdr = torch.sum(proj ** 2, dim=-1) / (torch.sum(e_noise ** 2, dim=-1) + eps)
        val = torch.mean(10 * torch.log10(sdr + eps))
        return -val if self.return_neg else val


def ndns_training_loss(enhanced_y, clean_y, win=2048, stride=512):
    """The training loss of recipes/intel_ndns/spiking_fullsubnet/trainer.py:33-37 with each STFT computed once:
    returns the dict the reference's training_step returns (loss, loss_freq_mae, loss_mag_mae, loss_sdr,
    loss_sdr_norm)."""
    e, t = _spec(enhanced_y, win, stride), _spec(clean_y, win, stride)
    loss_freq_mae = _freq_term(e, t)
    loss_mag_mae = _mag_term(e, t)
    loss_sdr = SISNRLoss(return_neg=False)(enhanced_y, clean_y)
    loss_sdr_norm = 0.001 * (100 - loss_sdr)
    return {"loss": loss_freq_mae + loss_mag_mae + loss_sdr_norm, "loss_freq_mae": loss_freq_mae,
            "loss_mag_mae": loss_mag_mae, "loss_sdr": loss_sdr, "loss_sdr_norm": loss_sdr_norm}
