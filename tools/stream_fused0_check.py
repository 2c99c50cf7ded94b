"""Development tool (GPU): the fused layer-0 path (gsn_xplanes_stream -> gsn_recurrence_stream in_planes) against the
separate tensor-core front end (gsn_pre_stream -> gsn_recurrence_stream xproj): spikes must be bit-identical (same x,
same MMA sequence per output element); us per frame of every piece, alone and chained through counters."""
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
    fb_act = fb_all[-1]
    sb = m.sb_model
    models = [dict(m=m.fb_model, N=1, lo=0, ctr=m.fb_input_size, nbr=0, fb=False)]
    for i, mm in enumerate(sb.sb_models):
        lo, hi, ctr = sb.freq_cutoffs[i], sb.freq_cutoffs[i + 1], sb.center_freq_sizes[i]
        models.append(dict(m=mm, N=(hi - lo) // ctr, lo=lo, ctr=ctr, nbr=sb.neighbor_freq_sizes[i], fb=True))
    for d, al in zip(models, [fb_all] + sb_all):
        mm = d["m"]
        H, K, R = mm.hidden_size, mm.input_size, B * d["N"]
        cell = mm.sequence_model.layers[0].cell
        lnw, lnb = mm.pre_layer_norm.weight.detach(), mm.pre_layer_norm.bias.detach()
        a, b = cell.folded_bn()
        w_ih, w_hh, bias = cell.weight_ih.detach(), cell.weight_hh.detach(), cell.bias_ih.detach()
        fbt = fb_act if d["fb"] else None
        geo = (d["N"], d["lo"], d["ctr"], d["nbr"])
        nt = ops._lib.load().gsn_recurrence_stream_tile(R, H, K, 1, 0)
        if nt == 0:
            print(f"R={R} K={K} H={H}: fused layer 0 does not fit tensor memory")
            continue
        xproj_tc = ops.pre_stream(cm, fbt, *geo, w_ih, lnw, lnb, 1e-5, ctas_per_slice=4)
        bits_ref = ops.recurrence_stream(w_hh, bias, a, b, xproj=xproj_tc)
        xop = ops.xplanes_buffer(T, R, K, nt, DEV)
        xo = torch.empty_like(al[0])
        for ctas in (1, 2, 4):
            _, us = timeit(lambda: ops.xplanes_stream(cm, fbt, *geo, nt, xop, lnw, lnb, 1e-5, out_x=None, ctas=ctas))
            print(f"R={R} K={K} H={H} nt={nt}: xplanes ctas={ctas}: {us:.2f} us/frame")
        ops.xplanes_stream(cm, fbt, *geo, nt, xop, lnw, lnb, 1e-5, out_x=xo, ctas=2)
        bits_new, us = timeit(lambda: ops.recurrence_stream(w_hh, bias, a, b, in_planes=xop, w_ih=w_ih, frames_rows=(T, R)))
        same = bool(torch.equal(bits_new, bits_ref))
        flips_eager = int((ops.unpack_spikes(bits_new, H) != al[1]).sum())
        print(f"   rec0 fused (planes): {us:.2f} us/frame; identical to pre_stream path {same}; max|x - x_eager| "
              f"{float((xo - al[0]).abs().max()):.1e}; flips vs eager {flips_eager} of {al[1].numel()}")
        # chained through counters
        cnt = ops.frame_counters(T, DEV, 1)
        s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
        ob = torch.zeros_like(bits_ref)
        for ctas in (1, 2, 4):
            def chained():
                cnt.zero_()
                cur = torch.cuda.current_stream()
                s0.wait_stream(cur)
                s1.wait_stream(cur)
                with torch.cuda.stream(s1):
                    ops.recurrence_stream(w_hh, bias, a, b, in_planes=xop, w_ih=w_ih, frames_rows=(T, R), out_bits=ob,
                                          in_cnt=cnt[0], in_target=R)
                with torch.cuda.stream(s0):
                    ops.xplanes_stream(cm, fbt, *geo, nt, xop, lnw, lnb, 1e-5, out_cnt=cnt[0], ctas=ctas)
                cur.wait_stream(s0)
                cur.wait_stream(s1)
            _, us = timeit(chained)
            print(f"   chained xplanes(ctas={ctas}) -> rec0: {us:.2f} us/frame; identical {bool(torch.equal(ob, bits_ref))}; "
                  f"counters ok {bool((cnt[0] == R).all())}")
