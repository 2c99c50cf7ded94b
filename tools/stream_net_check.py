"""Development tool (GPU): the streaming pipeline (SpikingFullSubNet.enable_streaming) against the eager schedule:
spike flips per layer, coefficient differences, pre-stage accuracy against float64, ms per step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, ops  # noqa: E402

DEV = "cuda:0"
size = sys.argv[1] if len(sys.argv) > 1 else "S"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 501
graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
cfg = synth.CONFIGS[size]
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()}, strict=True)
m = m.eval().to(DEV)
mag = torch.from_numpy(synth.make_mag(B, 257, T, 11)).to(DEV)


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, min(ts), float(np.median(ts))


with torch.no_grad():
    (pe, fbe, sbe), te, _ = timeit(lambda: m.network(mag))
    pe = [p.clone() for p in pe]
    spikes_e = [fbe[1].clone(), fbe[2].clone()] + [al[1 + l].clone() for al in sbe for l in range(2)]
    x_e = [fbe[0].clone()] + [al[0].clone() for al in sbe]
    # pre-stage accuracy vs float64 on the eager x
    for i, (x, mm) in enumerate(zip(x_e, [m.fb_model] + list(m.sb_model.sb_models))):
        w = mm.sequence_model.layers[0].cell.weight_ih.detach()
        ref = x.double() @ w.double().t()
        a = ops.linear(x, w)
        print(f"model {i}: layer-0 projection fp32 FFMA err vs f64 {float((a.double() - ref).abs().max()):.2e} "
              f"(max|ref| {float(ref.abs().max()):.2f})")
    m.enable_streaming(True)
    if graph:
        m.enable_cuda_graph(True, frame_chunks=12)
    t0 = time.time()
    res = m.network(mag)
    torch.cuda.synchronize()
    print(f"first streaming call ok in {time.time() - t0:.2f} s; plan:",
          [(d["R"], d["nt"], d["fused0"], [l["fused"] for l in d["layers"]], d["pre_p"], d["lin_p"], d["proj_p"])
           for d in m._stream_plan(B)])
    (ps, fbs, sbs), ts, tmed = timeit(lambda: m.network(mag), n=10)
    spikes_s = [fbs[1], fbs[2]] + [al[1 + l] for al in sbs for l in range(2)]
    names = ["fb0", "fb1"] + [f"sb{i}_{l}" for i in range(len(sbs)) for l in range(2)]
    tot = 0
    for n, a, b in zip(names, spikes_e, spikes_s):
        d = int((a != b).sum())
        tot += d
        first = -1
        if d:
            first = int(torch.nonzero((a != b).reshape(a.shape[0], -1).any(dim=1))[0])
        print(f"  {n}: {d} of {a.numel()} spikes differ (first frame {first}); rate {float(b.mean()):.3f}")
    for i, (a, b) in enumerate(zip(pe, ps)):
        print(f"  proj{i}: max|diff| {float((a - b).abs().max()):.3e} of max {float(a.abs().max()):.2f}")
    xs = [fbs[0]] + [al[0] for al in sbs]
    print("  max|x - x_eager|:", [float((a - b).abs().max()) for a, b in zip(x_e, xs)])
    print(f"{size} B={B} T={T}: eager {te:.3f} ms, streaming{' (graph)' if graph else ''} best {ts:.3f} ms median {tmed:.3f} ms "
          f"-> {B * T / ts / 1e3:.2f} M frames/s; total flips {tot}")
