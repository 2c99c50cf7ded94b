"""Development tool (GPU): us per frame of the spike-input helper stage (gsn_linear_spike_bits_stream) alone, for the
cost model of the streaming plan (modeling._stage_us).  Usage: python tools/stage_k_timing.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

T = 501
ops.stream_preload("cuda:0")
for (R, K, N) in [(96, 160, 64), (256, 160, 24), (192, 256, 24), (192, 256, 256), (128, 256, 256), (32, 240, 240), (8, 320, 320)]:
    bits = ops.pack_spikes((torch.rand(T, R, K, device="cuda") < 0.3).float())
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.zeros(N, device="cuda")
    for ctas in (1, 2):
        c = ctas * ((N + 127) // 128)
        for _ in range(2):
            out = ops.linear_bits_stream(bits, w, b, ctas=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.linear_bits_stream(bits, w, b, ctas=c)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / T
        model = (R / 64.0) * 3 * ((K + 15) // 16) * 75.0 / 1965.0
        print(f"R={R} K={K} N={N} ctas/slice={ctas}: {us:.2f} us/frame (x ctas = {us * ctas:.2f}; cost model {model:.2f})")
