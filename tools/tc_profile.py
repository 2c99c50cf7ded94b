"""Development tool (GPU): per-phase cycle breakdown of the tcgen05 recurrence kernel (CTA 0, thread 0).
Usage: python tools/tc_profile.py [R H T]"""
import os
import sys
import time

import numpy as np
import torch

os.environ.setdefault("GSN_TC_PROF", "1")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

NAMES = ["sync_top", "mma_issue", "xp_issue+mma_wait", "ld+math+ballot", "dsmem_send", "bits_wait", "expand",
         "total"]


BACKEND = os.environ.get("GSN_PROFILE_BACKEND", "tcgen05")


def run(R, H, T):
    dev = "cuda"
    rs = np.random.RandomState(0)
    s = 1 / np.sqrt(H)
    xproj = torch.from_numpy(rs.uniform(-1, 1, (T, R, H)).astype(np.float32)).to(dev)
    w = torch.from_numpy(rs.uniform(-s, s, (H, H)).astype(np.float32)).to(dev)
    b = torch.from_numpy(rs.uniform(-s, s, 2 * H).astype(np.float32)).to(dev)
    h0 = (torch.rand(R, H, device=dev) < 0.3).float()
    c0 = torch.randn(R, H, device=dev)
    for _ in range(2):
        ops.layer_recurrence(xproj, w, b, backend=BACKEND, h0=h0, c0=c0, want_state=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.layer_recurrence(xproj, w, b, backend=BACKEND, h0=h0, c0=c0, want_state=True)
    e1.record()
    torch.cuda.synchronize()
    ws, off = ops.LAST_WS[0]
    prof = ws[off:off + 32].view(torch.int64).cpu().numpy()
    ms = e0.elapsed_time(e1)
    print(f"R={R} H={H} T={T}: {ms:.3f} ms total, {ms * 1e3 / T:.2f} us/frame; cycles/frame:",
          {n: round(float(v) / T, 1) for n, v in zip(NAMES, prof)},
          "| cycles per launch:", dict(zip(["alloc+h0", "weights", "state+cluster_sync", "loop", "exit"],
                                           [int(v) for v in prof[8:13]])))


if __name__ == "__main__":
    if len(sys.argv) == 4:
        run(*map(int, sys.argv[1:]))
    else:
        for (R, H) in [(32, 240), (256, 160), (96, 160), (64, 160), (16, 128), (16, 64), (1024, 256), (64, 320)]:
            run(R, H, 501)
