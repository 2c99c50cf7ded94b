"""Development probe (GPU): runs gsn_tc_selftest for every operand-encoding variant and prints the error
against a float64 CPU product.  Usage on the GPU box: python tools/tc_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import _lib  # noqa: E402


def bf16_exact(rs, shape):
    x = torch.from_numpy(rs.uniform(-1, 1, shape).astype(np.float32))
    return x.to(torch.bfloat16).to(torch.float32)


def run(N, K, a_in_tmem, swap, fp16, binary_b=False):
    lib = _lib.load_probe()
    rs = np.random.RandomState(N * 7 + K)
    a = bf16_exact(rs, (128, K))
    b = bf16_exact(rs, (N, K))
    if fp16:
        a, b = a.to(torch.float16).float(), b.to(torch.float16).float()
    if binary_b:
        b = (b > 0).float()
    ref = (a.double() @ b.double().T).numpy()
    ad, bd = a.cuda(), b.cuda()
    d = torch.zeros(128, N, device="cuda")
    st = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    rc = lib.gsn_tc_selftest(ad.data_ptr(), bd.data_ptr(), d.data_ptr(), st.data_ptr(), N, K, a_in_tmem, swap,
                             fp16, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    torch.cuda.synchronize()
    err = float(np.abs(d.cpu().numpy() - ref).max())
    return int(st.item()), err, float(np.abs(ref).max())


if __name__ == "__main__":
    for (N, K) in [(16, 16), (32, 64), (64, 160), (32, 256), (16, 320)]:
        for a_in_tmem in (0, 1):
            for swap in (0, 1):
                for fp16 in (0, 1):
                    try:
                        st, err, mx = run(N, K, a_in_tmem, swap, fp16)
                        print(f"N={N} K={K} a_in_tmem={a_in_tmem} swap={swap} fp16={fp16}: status={st} "
                              f"max_err={err:.3e} (max|ref|={mx:.2f})", flush=True)
                    except Exception as e:  # noqa: BLE001
                        print(f"N={N} K={K} a_in_tmem={a_in_tmem} swap={swap} fp16={fp16}: EXC {e}", flush=True)
