"""Development tool (GPU): cycles per tcgen05.mma (M=128) for operand sources / N.  Undeclared dev symbol."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import _lib  # noqa: E402

lib = _lib.load_probe()
fn = lib.gsn_tc_mma_timing
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for K in (160, 256):
    for N in (16, 32, 64, 128, 256):
        for a_in_tmem in (0, 1):
            for _ in range(2):
                _lib.check(fn(out.data_ptr(), N, K, 3, a_in_tmem, None))
                torch.cuda.synchronize()
            n = 3 * K // 16
            o = out.cpu().tolist()
            print(f"K={K} N={N} a_in_tmem={a_in_tmem}: {n} MMAs, issue {o[0] / n:.1f} cyc/MMA, complete {o[1] / n:.1f} cyc/MMA")

fn2 = lib.gsn_tc_mma_timing2
fn2.restype = C.c_int
fn2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
for variant, name, n in ((0, "K=160 M=128", 30), (1, "K=160 M=64", 30), (2, "K=240 M=128", 45)):
    for N in (16, 32, 64):
        for _ in range(2):
            _lib.check(fn2(out.data_ptr(), N, variant, None))
            torch.cuda.synchronize()
        o = out.cpu().tolist()
        print(f"unrolled issue, {name} N={N}: {n} MMAs, issue {o[0] / n:.1f} cyc/MMA, complete {o[1] / n:.1f} cyc/MMA")
