"""Development tool (GPU): where forward() spends its time outside the network (eager launches, CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, ops  # noqa: E402
from spiking_fullsubnet_b200.modeling import _istft_nosync, _stft  # noqa: E402

DEV = "cuda:0"
cfg = synth.CONFIGS["S"]
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()}, strict=True)
m = m.eval().to(DEV)
B, L = 32, 64000
wave = torch.from_numpy(synth.make_wave(B, L, 21)).to(DEV)


def timed(fn, n=20):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n * 1e3


with torch.no_grad():
    cmp, t_stft = timed(lambda: _stft(wave, m.n_fft, m.hop_length, m.win_length))
    mag, t_abs = timed(lambda: cmp.abs().contiguous())
    m.enable_streaming(True)
    m.enable_cuda_graph(True, frame_chunks=12)
    (projs, fb_all, sb_all), t_net = timed(lambda: m.network(mag))
    S = 1

    def post():
        sre, sim = cmp.real.contiguous(), cmp.imag.contiguous()
        ore = sre.unsqueeze(1).repeat(1, S, 1, 1)
        oim = sim.unsqueeze(1).repeat(1, S, 1, 1)
        cuts, ctrs = m.sb_model.freq_cutoffs, m.sb_model.center_freq_sizes
        lo = 0
        for i, p in enumerate(projs):
            n = (cuts[i + 1] - cuts[i]) // ctrs[i]
            ops.deepfilter_band(p, sre, sim, ore, oim, n, ctrs[i], m.df_orders[i], S, lo)
            lo += n * ctrs[i]
        return torch.complex(ore, oim)

    enh, t_post = timed(post)
    (y, _), t_istft = timed(lambda: (_istft_nosync(enh[:, 0], m.n_fft, m.hop_length, m.win_length, L), enh[:, 0].abs()))
    _, t_fwd = timed(lambda: m(wave))
    print(f"stft {t_stft:.0f} us | abs+contiguous {t_abs:.0f} | network (graph) {t_net:.0f} | split/repeat/deep filter/complex "
          f"{t_post:.0f} | istft+abs {t_istft:.0f} | sum {t_stft + t_abs + t_net + t_post + t_istft:.0f} | forward() graph {t_fwd:.0f}")
