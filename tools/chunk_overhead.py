"""Development tool (GPU): fixed cost per recurrence launch vs cost per frame (what a frame chunk of the
wavefront schedule pays): times one launch at several T, with carried state like the chunks, and fits
time = a + b*T.  Usage: python tools/chunk_overhead.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

TS = [8, 16, 31, 62, 125, 250, 501]


def run(R, H, backend="tcgen05"):
    rs = np.random.RandomState(0)
    s = 1 / np.sqrt(H)
    Tm = max(TS)
    xproj = torch.from_numpy(rs.uniform(-1, 1, (Tm, R, H)).astype(np.float32)).cuda()
    w = torch.from_numpy(rs.uniform(-s, s, (H, H)).astype(np.float32)).cuda()
    b = torch.from_numpy(rs.uniform(-s, s, 2 * H).astype(np.float32)).cuda()
    h0 = (torch.rand(R, H, device="cuda") < 0.3).float()
    c0 = torch.randn(R, H, device="cuda")
    out_h = torch.empty((Tm, R, H), device="cuda")
    hT, cT = torch.empty_like(h0), torch.empty_like(c0)
    ws = ops.recurrence_workspace(R, H, True, backend, xproj.device)
    res = []
    for T in TS:
        best = 1e9
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.layer_recurrence(xproj[:T], w, b, h0=h0, c0=c0, out_h=out_h[:T], out_hT=hT, out_cT=cT,
                                 backend=backend, workspace=ws)
            e1.record()
            torch.cuda.synchronize()
            if it:
                best = min(best, e0.elapsed_time(e1) * 1e3)
        res.append(best)
    A = np.stack([np.ones(len(TS)), np.array(TS, dtype=np.float64)], 1)
    (a, bslope), *_ = np.linalg.lstsq(A, np.array(res), rcond=None)
    print(f"R={R} H={H} {backend}: " + ", ".join(f"T={t}: {r:.1f}us" for t, r in zip(TS, res)) +
          f" | fit: {a:.1f} us per launch + {bslope:.3f} us per frame", flush=True)


if __name__ == "__main__":
    for (R, H) in [(32, 240), (256, 160), (96, 160), (64, 160), (32, 320), (256, 224)]:
        run(R, H)
    # the spike-input linears of one chunk (tcgen05, persistent): time vs SM budget
    for (M, K, N) in [(1984, 240, 240), (15872, 160, 160), (5952, 160, 160), (15872, 160, 24)]:
        a = (torch.rand(M, K, device="cuda") < 0.3).float()
        w = torch.randn(N, K, device="cuda") * 0.1
        out = torch.empty(M, N, device="cuda")
        line = []
        for bud in (0, 36, 16, 8, 4):
            best = 1e9
            for it in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.linear(a, w, out=out, spikes=True, sm_budget=bud)
                e1.record()
                torch.cuda.synchronize()
                if it:
                    best = min(best, e0.elapsed_time(e1) * 1e3)
            line.append(f"budget {bud}: {best:.1f}us")
        print(f"linear_spikes M={M} K={K} N={N}: " + ", ".join(line), flush=True)
