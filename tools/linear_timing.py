"""Development tool (GPU): time of the tcgen05 spike-input linear vs rows per CTA (fp32 trace vs bit-packed trace).
Usage: python tools/linear_timing.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402


def t_us(fn, n=8):
    best = 1e9
    for it in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


for (K, N) in [(160, 160)]:
    w = torch.randn(N, K, device="cuda") * 0.1
    for bud in (36, 0):
        for tiles_per_cta in (1, 2, 4, 8, 16):
            slices = (N + 127) // 128
            P = max(1, (bud or 148) // slices)
            M = 64 * P * tiles_per_cta
            a = (torch.rand(M, K, device="cuda") < 0.3).float()
            bits = ops.pack_spikes(a)
            out = torch.empty(M, N, device="cuda")
            f = t_us(lambda: ops.linear(a, w, out=out, spikes=True, sm_budget=bud))
            b = t_us(lambda: ops.linear(a, w, out=out, spikes=True, sm_budget=bud, bits=bits))
            print(f"K={K} N={N} budget={bud} tiles/CTA={tiles_per_cta} (M={M}): fp32 trace {f:.1f} us, bits {b:.1f} us",
                  flush=True)

# per-phase cycles of CTA (0,0) / thread 0 (the kernel accumulates them into the trace-buffer header)
import numpy as np  # noqa: E402
from spiking_fullsubnet_b200 import _lib  # noqa: E402
lib = _lib.load()
buf = torch.zeros(64 + 32 * 64, dtype=torch.uint8, device="cuda")
_lib.check(lib.gsn_trace_set(buf.data_ptr(), buf.numel()))
K, N, M = 160, 160, 64 * 18 * 16
w = torch.randn(N, K, device="cuda") * 0.1
a = (torch.rand(M, K, device="cuda") < 0.3).float()
bits = ops.pack_spikes(a)
out = torch.empty(M, N, device="cuda")
for name, kw in (("bits", dict(bits=bits)), ("fp32", dict())):
    for _ in range(2):
        ops.linear(a, w, out=out, spikes=True, sm_budget=36, **kw)
    torch.cuda.synchronize()
    hdr = buf[:64].cpu().numpy().view(np.uint32)
    n = max(1, int(hdr[7]))
    print(name, "cycles per tile:", dict(zip(["convert", "sync", "mma_wait", "mma_issue", "epilogue"],
                                                [round(int(v) / n) for v in hdr[2:7]])), "tiles", n)
_lib.check(lib.gsn_trace_set(None, 0))
