"""Development tool (GPU): localise the first stage of the chained streaming pipeline whose output differs from the
same stage run alone on complete inputs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, ops  # noqa: E402

DEV = "cuda:0"
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 501
cfg = synth.CFG_S
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()}, strict=True)
m = m.eval().to(DEV).enable_streaming(True)
mag = torch.from_numpy(synth.make_mag(B, 257, T, 11)).to(DEV)


def where(a, b, name):
    d = (a != b)
    if d.dim() == 3:
        d = d.any(dim=2)
    n = int(d.sum())
    if n == 0:
        print(f"   {name}: identical")
        return
    idx = torch.nonzero(d)
    ts = idx[:, 0]
    print(f"   {name}: {n} (frame,row) pairs differ; first frames {sorted(set(ts.tolist()))[:10]}; "
          f"rows at first frame {idx[ts == ts.min()][:, 1].tolist()[:16]}")


with torch.no_grad():
    for rep in range(3):
        projs, fb_all, sb_all = m.network(mag)
        torch.cuda.synchronize()
        cm, counters, results = m._keepalive
        plan = m._stream_plan(B)
        fb_act = results[0][0]
        print(f"rep {rep}")
        for d, res in zip(plan, results):
            proj, outs, bits_all, xproj = res
            mm = d["m"]
            H, R, C = mm.hidden_size, d["R"], d["C"]
            cells = [l.cell for l in mm.sequence_model.layers]
            lnw, lnb = mm.pre_layer_norm.weight.detach(), mm.pre_layer_norm.bias.detach()
            print(f" model R={R}")
            xp = ops.pre_stream(cm, fb_act if d["fb"] else None, d["N"], d["lo"], d["ctr"], d["nbr"],
                                cells[0].weight_ih.detach(), lnw, lnb, 1e-5, ctas_per_slice=d["pre_p"])
            where(xproj, xp, "xproj (pre)")
            a, b = cells[0].folded_bn()
            b0 = ops.recurrence_stream(cells[0].weight_hh.detach(), cells[0].bias_ih.detach(), a, b, xproj=xproj)
            where(bits_all[0], b0, "bits0 (rec0 on the chained xproj)")
            a1, b1 = cells[1].folded_bn()
            if d["layers"][1]["fused"]:
                bb1 = ops.recurrence_stream(cells[1].weight_hh.detach(), cells[1].bias_ih.detach(), a1, b1,
                                            in_bits=bits_all[0], w_ih=cells[1].weight_ih.detach())
            else:
                xp1 = ops.linear_bits_stream(bits_all[0], cells[1].weight_ih.detach(), ctas=C)
                bb1 = ops.recurrence_stream(cells[1].weight_hh.detach(), cells[1].bias_ih.detach(), a1, b1, xproj=xp1)
            where(bits_all[1], bb1, "bits1 (rec1 on the chained bits0)")
            pr = ops.linear_bits_stream(bits_all[1], mm.proj.weight.detach(), mm.proj.bias.detach(), ctas=1)
            where(proj, pr, "proj (on the chained bits1)")
        print("  counters complete:", [int(c.min()) for c in counters], [int(c.max()) for c in counters])
