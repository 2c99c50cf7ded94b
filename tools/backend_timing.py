"""Development tool (GPU): us/frame of one recurrence launch per back end."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

T = 501
for (R, H, shared) in [(32, 240, True), (256, 160, True), (64, 160, True), (1024, 256, True), (64, 320, True),
                       (256, 224, False), (480, 512, True), (2016, 512, True)]:
    rs = np.random.RandomState(0)
    s = 1 / np.sqrt(H)
    gH = H if shared else 2 * H
    xproj = torch.from_numpy(rs.uniform(-1, 1, (T, R, gH)).astype(np.float32)).cuda()
    w = torch.from_numpy(rs.uniform(-s, s, (gH, H)).astype(np.float32)).cuda()
    b = torch.from_numpy(rs.uniform(-s, s, 2 * H).astype(np.float32)).cuda()
    res = {}
    outs = {}
    for be in ("tcgen05", "tcgen05_i8"):
        try:
            for _ in range(2):
                h, _, _ = ops.layer_recurrence(xproj, w, b, shared=shared, backend=be)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                h, _, _ = ops.layer_recurrence(xproj, w, b, shared=shared, backend=be)
            e1.record()
            torch.cuda.synchronize()
            res[be] = e0.elapsed_time(e1) / 3 * 1e3 / T
            outs[be] = h
        except NotImplementedError:
            res[be] = float("nan")
    flips = float((outs["tcgen05"] != outs["tcgen05_i8"]).float().mean()) if len(outs) == 2 else float("nan")
    print(f"R={R} H={H} shared={shared}: " + ", ".join(f"{k} {v:.2f} us/frame" for k, v in res.items()) +
          f"; spikes differing between the two: {flips:.2e}", flush=True)
