"""Development tool (GPU): streaming pipeline on a small model (for compute-sanitizer runs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet  # noqa: E402

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cfg = synth.tiny_cfg() if which == "tiny" else synth.CONFIGS[which]
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()}, strict=True)
m = m.eval().to(DEV)
mag = torch.from_numpy(synth.make_mag(B, cfg["n_fft"] // 2 + 1, T, 11)).to(DEV)
with torch.no_grad():
    pe, fbe, sbe = m.network(mag)
    torch.cuda.synchronize()
    m.enable_streaming(True)
    print("plan:", [(d["R"], d["m"].input_size, d["m"].hidden_size, d["fused0"], [l["fused"] for l in d["layers"]],
                     d["pre_p"], d["lin_p"], d["proj_p"]) for d in (m._stream_plan(B) or [])], flush=True)
    for it in range(2):
        ps, fbs, sbs = m.network(mag)
        torch.cuda.synchronize()
        names = ["fb"] + [f"sb{i}" for i in range(len(sbs))]
        for n, a, b in zip(names, [fbe] + sbe, [fbs] + sbs):
            print(f"call {it} {n}: x diff {float((a[0] - b[0]).abs().max()):.1e}; flips "
                  f"{[int((a[1 + l] != b[1 + l]).sum()) for l in range(len(a) - 2)]} of {a[1].numel()}; proj diff "
                  f"{float((a[-1] - b[-1]).abs().max()):.1e}", flush=True)
    m.enable_cuda_graph(True, frame_chunks=4)
    for it in range(2):
        ps, fbs, sbs = m.network(mag)
        torch.cuda.synchronize()
        names = ["fb"] + [f"sb{i}" for i in range(len(sbs))]
        for n, a, b in zip(names, [fbe] + sbe, [fbs] + sbs):
            print(f"graph call {it} {n}: x diff {float((a[0] - b[0]).abs().max()):.1e}; flips "
                  f"{[int((a[1 + l] != b[1 + l]).sum()) for l in range(len(a) - 2)]} of {a[1].numel()}; proj diff "
                  f"{float((a[-1] - b[-1]).abs().max()):.1e}", flush=True)
