"""Development tool (GPU): ms/step of the CUDA-graph wavefront schedule vs number of frame chunks.
Usage: python tools/chunk_sweep.py [SIZE] [chunks ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "S"
chunks = [int(a) for a in sys.argv[2:]] or [1, 8, 12, 16, 24, 32]
cfg = synth.CONFIGS[size]
model = SpikingFullSubNet(**cfg)
model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
model = model.eval().cuda()
B, T = 32, 501
mag = torch.from_numpy(synth.make_mag(B, 257, T, 11)).cuda()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for n in chunks:
    model.enable_cuda_graph(True, frame_chunks=n)
    with torch.no_grad():
        for _ in range(3):
            model.network(mag)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.network(mag)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    print(f"{size} chunks={n}: {np.mean(ts):.3f} ms/step (min {min(ts):.3f}) -> {B * T / np.mean(ts) * 1e3 / 1e6:.2f} M frames/s",
          flush=True)
