"""Development tool (GPU): kernel-time breakdown of one training step (torch.profiler)."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "L"
cfg = synth.CONFIGS[size]
dev = "cuda"
model = SpikingFullSubNet(**cfg)
model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
model = model.to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
L = 96000
wave = torch.from_numpy(synth.make_wave(32, L, 31)).to(dev)
clean = torch.from_numpy(synth.make_wave(32, L, 41)).to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    enh_y, enh_mag, *_ = model(wave)
    loss = (enh_y - clean).abs().mean() + enh_mag.mean()
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
