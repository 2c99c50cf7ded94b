"""Development tool (GPU): per-phase cycles of the training forward kernel (CTA 0)."""
import os
import sys

os.environ["GSN_TRAIN_PROF"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import _lib, ops  # noqa: E402

NAMES = ["matmul", "gate math+partials", "grid barrier", "reduce partials", "normalise+stores", "total"]
T = 501
for (R, H) in [(32, 240), (256, 160), (768, 256)]:
    rs = np.random.RandomState(0)
    s = 1 / np.sqrt(H)
    dev = "cuda"
    xproj = torch.from_numpy(rs.uniform(-1, 1, (T, R, H)).astype(np.float32)).to(dev)
    w = torch.from_numpy(rs.uniform(-s, s, (H, H)).astype(np.float32)).to(dev)
    b = torch.from_numpy(rs.uniform(-s, s, 2 * H).astype(np.float32)).to(dev)
    g, be = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    rm, rv = torch.zeros(H, device=dev), torch.ones(H, device=dev)
    lib = _lib.load()
    new = lambda *sh: torch.empty(sh, device=dev)  # noqa: E731
    h, c, f, gg, xh, inv = new(T, R, H), new(T, R, H), new(T, R, H), new(T, R, H), new(T, R, H), new(T, H)
    nbytes = lib.gsn_layer_train_workspace_bytes(R, H, 1)
    ws = torch.zeros(nbytes // 4 + 128, device=dev)
    off = ((-ws.data_ptr()) % 256) // 4
    wsv = ws[off:]
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.gsn_layer_train_forward(xproj.data_ptr(), w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr(),
                                               rm.data_ptr(), rv.data_ptr(), h.data_ptr(), c.data_ptr(), f.data_ptr(),
                                               gg.data_ptr(), xh.data_ptr(), inv.data_ptr(), T, R, H, 1, 1, 0.1, 1e-5,
                                               wsv.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
    nb = (R + 7) // 8
    cnt_off = H * H + 8 * nb * H
    prof = wsv[cnt_off:cnt_off + 32].view(torch.int64)[2:8].cpu().numpy()
    print(f"R={R} H={H}: {e0.elapsed_time(e1) * 1e3 / T:.2f} us/frame; cycles/frame:",
          {n: round(float(v) / T) for n, v in zip(NAMES, prof)})
