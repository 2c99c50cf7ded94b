"""Development tool (GPU): every stage of the streaming pipeline timed ALONE (no counters: all inputs already there),
us per frame, and checked against the eager kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, ops  # noqa: E402

DEV = "cuda:0"
size = sys.argv[1] if len(sys.argv) > 1 else "S"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 501
cfg = synth.CONFIGS[size]
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()}, strict=True)
m = m.eval().to(DEV)
mag = torch.from_numpy(synth.make_mag(B, 257, T, 11)).to(DEV)


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best * 1e3 / T


with torch.no_grad():
    projs, fb_all, sb_all = m.network(mag)
    cm = ops.compress_mag(mag, 256, cfg["fdrc"])
    plan = m._stream_plan(B)
    fb_act = fb_all[-1]
    alls = [fb_all] + sb_all
    for d, al in zip(plan, alls):
        mm = d["m"]
        H, K, R, C = mm.hidden_size, mm.input_size, d["R"], d["C"]
        cells = [l.cell for l in mm.sequence_model.layers]
        lnw, lnb = mm.pre_layer_norm.weight.detach(), mm.pre_layer_norm.bias.detach()
        x = al[0]
        ref = x.double() @ cells[0].weight_ih.detach().double().t()
        for P in (1, 2, 3):
            xo = torch.empty_like(x)
            xp, us = timeit(lambda: ops.pre_stream(cm, fb_act if d["fb"] else None, d["N"], d["lo"], d["ctr"], d["nbr"],
                                                   cells[0].weight_ih.detach(), lnw, lnb, 1e-5, out_x=xo,
                                                   ctas_per_slice=P))
            print(f"R={R} K={K} H={H}: pre P={P}: {us:.2f} us/frame; max|x - x_eager| {float((xo - x).abs().max()):.1e}; "
                  f"xproj err vs f64 {float((xp.double() - ref).abs().max()):.2e} "
                  f"(fp32 FFMA kernel {float((ops.linear(x, cells[0].weight_ih.detach()).double() - ref).abs().max()):.2e})")
        a, b = cells[0].folded_bn()
        xproj = ops.linear(x, cells[0].weight_ih.detach())
        bits0, us = timeit(lambda: ops.recurrence_stream(cells[0].weight_hh.detach(), cells[0].bias_ih.detach(), a, b,
                                                         xproj=xproj))
        print(f"   rec0: {us:.2f} us/frame; identical {bool(torch.equal(ops.unpack_spikes(bits0, H), al[1]))}")
        a1, b1 = cells[1].folded_bn()
        if d["layers"][1]["fused"]:
            bits1, us = timeit(lambda: ops.recurrence_stream(cells[1].weight_hh.detach(), cells[1].bias_ih.detach(), a1,
                                                             b1, in_bits=bits0, w_ih=cells[1].weight_ih.detach()))
            print(f"   rec1 fused: {us:.2f} us/frame; identical {bool(torch.equal(ops.unpack_spikes(bits1, H), al[2]))}")
        else:
            for P in (1, 2):
                xp1, us = timeit(lambda: ops.linear_bits_stream(bits0, cells[1].weight_ih.detach(), ctas=C * P))
                print(f"   lin P={P}: {us:.2f} us/frame")
            bits1, us = timeit(lambda: ops.recurrence_stream(cells[1].weight_hh.detach(), cells[1].bias_ih.detach(), a1,
                                                             b1, xproj=xp1))
            print(f"   rec1: {us:.2f} us/frame; identical {bool(torch.equal(ops.unpack_spikes(bits1, H), al[2]))}")
        ps = (mm.proj_size + 127) // 128
        for P in (1, 2):
            pr, us = timeit(lambda: ops.linear_bits_stream(bits1, mm.proj.weight.detach(), mm.proj.bias.detach(),
                                                           ctas=ps * P))
            print(f"   proj P={P}: {us:.2f} us/frame; identical {bool(torch.equal(pr, al[3]))}")
