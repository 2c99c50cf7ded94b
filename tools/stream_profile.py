"""Development tool (GPU): per-phase cycles of the streaming recurrence kernel (PROF build: GSN_TC_PROF=1).
CTA 0: epilogue warp 0 / lane 0 and the MMA-issue thread.  Usage: python tools/stream_profile.py [R H T [fused]]"""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("GSN_TC_PROF", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

EPI = ["input(wait+ld)", "mma_wait", "ld+math+ballot", "send+trace", "bits_wait", "expand+arrive", "(send: local+fence+arrive)", "(send: +st.async)"]
ISS = ["wait_B", "issue_hh", "issue_ih(+waits)"]


def run(R, H, T, fused):
    rs = np.random.RandomState(0)
    s = 1 / np.sqrt(H)
    t_ = lambda a: torch.from_numpy(a.astype(np.float32)).cuda()  # noqa: E731
    w, b = t_(rs.uniform(-s, s, (H, H))), t_(rs.uniform(-s, s, 2 * H))
    ws = torch.zeros(64, dtype=torch.int64, device="cuda")
    if fused:
        inb = ops.pack_spikes((torch.rand(T, R, H, device="cuda") < 0.4).float())
        wih = t_(rs.uniform(-s, s, (H, H)))
        kw = dict(in_bits=inb, w_ih=wih)
    else:
        kw = dict(xproj=t_(rs.uniform(-1, 1, (T, R, H))))
    for _ in range(2):
        ops.recurrence_stream(w, b, workspace=ws, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.recurrence_stream(w, b, workspace=ws, **kw)
    e1.record()
    torch.cuda.synchronize()
    pr = ws.cpu().numpy()
    ms = e0.elapsed_time(e1)
    print(f"R={R} H={H} T={T} fused={fused}: {ms * 1e3 / T:.2f} us/frame; epilogue cycles/frame:",
          {n: round(float(v) / T, 1) for n, v in zip(EPI, pr[:8])}, "sum", round(float(pr[:6].sum()) / T, 1),
          "| issuer:", {n: round(float(v) / T, 1) for n, v in zip(ISS, pr[8:11])},
          f"| launch: {int(pr[12])} cycles in {int(pr[13])} ns = {float(pr[12]) / max(1.0, float(pr[13])):.3f} GHz")


if __name__ == "__main__":
    if len(sys.argv) >= 4:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), len(sys.argv) > 4 and sys.argv[4] == "1")
    else:
        for (R, H) in [(256, 160), (32, 240), (64, 320)]:
            run(R, H, 501, False)
        run(256, 160, 501, True)
