"""Golden fixture for the loss terms (row f3): RUNS THE REFERENCE's audiozen/loss.py (read-only import from
/root/reference, CPU fp32) on a seeded pair of waveforms and stores the values and the gradient of the recipe's
training loss (recipes/intel_ndns/spiking_fullsubnet/trainer.py:33-37) w.r.t. the estimate.

    python tests/golden/make_golden_loss.py        (build container only; the tests read loss_ref.npz)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSN_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
from audiozen.loss import SISNRLoss, freq_MAE, mag_MAE  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import loss_waveforms  # noqa: E402  (the test regenerates the same inputs from the seed)

est, clean = loss_waveforms()
e = torch.from_numpy(est).requires_grad_(True)
c = torch.from_numpy(clean)
loss_freq = freq_MAE(e, c)
loss_mag = mag_MAE(e, c)
loss_sdr = SISNRLoss(return_neg=False)(e, c)
loss = loss_freq + loss_mag + 0.001 * (100 - loss_sdr)
loss.backward()
srs = [16000, 8000]
np.savez_compressed(os.path.join(HERE, "loss_ref.npz"), loss=loss.item(),
                    loss_freq_mae=loss_freq.item(), loss_mag_mae=loss_mag.item(), loss_sdr=loss_sdr.item(),
                    grad=e.grad.numpy(), srs=np.array(srs),
                    freq_srs=freq_MAE(e.detach(), c, srs=srs, sudo_sr=16000).item(),
                    mag_srs=mag_MAE(e.detach(), c, srs=srs, sudo_sr=16000).item(),
                    sisnr_neg=SISNRLoss(return_neg=True)(e.detach(), c).item())
print("wrote loss_ref.npz", loss.item(), loss_freq.item(), loss_mag.item(), loss_sdr.item())
