"""Profiling driver for the streaming schedule (run under ncu on the GPU box).

  graph       one CUDA-graph replay of the streaming step at BASELINE configs[1] (S, 32 x 501): profile with
              `ncu --graph-profiling graph` so that the whole graph is ONE workload (kernel-by-kernel replay would
              serialise kernels that wait for each other's frame counters); gives whole-step DRAM bytes.
  standalone  every persistent kernel of the step launched ALONE, without counters, on complete inputs (the launch
              list and the `--set full` capture of gsn::k_recurrence_stream come from this mode).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
DEV = "cuda:0"
B, T = 32, 501
cfg = synth.CONFIGS["S"]
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()}, strict=True)
m = m.eval().to(DEV)
mag = torch.from_numpy(synth.make_mag(B, 257, T, 11)).to(DEV)
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)

with torch.no_grad():
    if mode == "graph":
        m.enable_streaming(True)
        m.enable_cuda_graph(True, frame_chunks=12)
        for _ in range(3):
            flush.fill_(1.0)
            m.network(mag)
            torch.cuda.synchronize()
        print("graph replays done")
    else:
        projs, fb_all, sb_all = m.network(mag)  # eager schedule: complete inputs for every stage
        cm = ops.compress_mag(mag, 256, cfg["fdrc"])
        fb_act = fb_all[-1]
        sb = m.sb_model
        models = [dict(m=m.fb_model, N=1, lo=0, ctr=m.fb_input_size, nbr=0, fb=False)]
        for i, mm in enumerate(sb.sb_models):
            lo, hi, ctr = sb.freq_cutoffs[i], sb.freq_cutoffs[i + 1], sb.center_freq_sizes[i]
            models.append(dict(m=mm, N=(hi - lo) // ctr, lo=lo, ctr=ctr, nbr=sb.neighbor_freq_sizes[i], fb=True))
        plan = m.enable_streaming(True)._stream_plan(B)
        for d, pl, al in zip(models, plan, [fb_all] + sb_all):
            mm = d["m"]
            H, K, R = mm.hidden_size, mm.input_size, B * d["N"]
            cells = [l.cell for l in mm.sequence_model.layers]
            lnw, lnb = mm.pre_layer_norm.weight.detach(), mm.pre_layer_norm.bias.detach()
            fbt = fb_act if d["fb"] else None
            geo = (d["N"], d["lo"], d["ctr"], d["nbr"])
            nt = ops.stream_tile(R, H, K, True)
            xop = ops.xplanes_buffer(T, R, K, nt, DEV)
            flush.fill_(1.0)
            ops.xplanes_stream(cm, fbt, *geo, nt, xop, lnw, lnb, 1e-5, ctas=pl["pre_p"])
            a0, b0 = cells[0].folded_bn()
            flush.fill_(1.0)
            fused1 = pl["layers"][1]["fused"]
            # (one image slot per frame here: alone, the layers do not overlap, so there is no ring to reuse)
            img = ops.spike_image_buffer(T, R, H, DEV) if fused1 else None
            bits0 = ops.recurrence_stream(cells[0].weight_hh.detach(), cells[0].bias_ih.detach(), a0, b0, in_planes=xop,
                                          w_ih=cells[0].weight_ih.detach(), frames_rows=(T, R), img_out=img)
            a1, b1 = cells[1].folded_bn()
            flush.fill_(1.0)
            if fused1:
                bits1 = ops.recurrence_stream(cells[1].weight_hh.detach(), cells[1].bias_ih.detach(), a1, b1,
                                              in_image=img, frames_rows=(T, R), w_ih=cells[1].weight_ih.detach())
            else:
                xp1 = ops.linear_bits_stream(bits0, cells[1].weight_ih.detach(), ctas=pl["C"] * pl["lin_p"])
                flush.fill_(1.0)
                bits1 = ops.recurrence_stream(cells[1].weight_hh.detach(), cells[1].bias_ih.detach(), a1, b1, xproj=xp1)
            flush.fill_(1.0)
            ops.linear_bits_stream(bits1, mm.proj.weight.detach(), mm.proj.bias.detach(),
                                   ctas=((mm.proj_size + 127) // 128) * pl["proj_p"])
            torch.cuda.synchronize()
        print("standalone launches done")
