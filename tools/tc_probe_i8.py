"""Development probe (GPU): tcgen05 kind::i8 with A resident in TMEM -- correctness and cycles per MMA."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import _lib  # noqa: E402

lib = _lib.load_probe()
fn = lib.gsn_tc_probe_i8
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]
for (N, K) in [(16, 32), (16, 160), (32, 256), (64, 320), (16, 512)]:
    for a_signed in (1, 0):
        rs = np.random.RandomState(N + K)
        a = rs.randint(-128, 128, (128, K)) if a_signed else rs.randint(0, 256, (128, K))
        b = (rs.uniform(size=(N, K)) < 0.45).astype(np.int64)
        ref = a @ b.T
        ad = torch.from_numpy(a.astype(np.int32)).cuda()
        bd = torch.from_numpy(b.astype(np.int32)).cuda()
        d = torch.zeros(128, N, dtype=torch.int32, device="cuda")
        tm = torch.zeros(2, dtype=torch.int64, device="cuda")
        _lib.check(fn(ad.data_ptr(), bd.data_ptr(), d.data_ptr(), tm.data_ptr(), N, K, a_signed, 1, None))
        torch.cuda.synchronize()
        err = int(np.abs(d.cpu().numpy().astype(np.int64) - ref).max())
        _lib.check(fn(ad.data_ptr(), bd.data_ptr(), d.data_ptr(), tm.data_ptr(), N, K, a_signed, 3, None))
        torch.cuda.synchronize()
        t = tm.cpu().tolist()
        n = 3 * K // 32
        print(f"N={N} K={K} a_signed={a_signed}: max|err|={err}  ({n} MMAs: issue {t[0] / n:.1f}, complete {t[1] / n:.1f} cyc/MMA)",
              flush=True)
