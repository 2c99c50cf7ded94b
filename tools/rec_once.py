"""Development tool (GPU): run one recurrence launch (for ncu captures).  Usage: python tools/rec_once.py R H T [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

R, H, T = map(int, sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rs = np.random.RandomState(0)
s = 1 / np.sqrt(H)
xproj = torch.from_numpy(rs.uniform(-1, 1, (T, R, H)).astype(np.float32)).cuda()
w = torch.from_numpy(rs.uniform(-s, s, (H, H)).astype(np.float32)).cuda()
b = torch.from_numpy(rs.uniform(-s, s, 2 * H).astype(np.float32)).cuda()
bits = ops.spike_bits_buffer((T, R), H, "cuda")
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.layer_recurrence(xproj, w, b, backend="tcgen05", out_bits=bits)
    e1.record()
    torch.cuda.synchronize()
    print(f"R={R} H={H} T={T}: {e0.elapsed_time(e1) * 1e3 / T:.2f} us/frame")
