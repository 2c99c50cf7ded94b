"""Development tool (GPU): kernel timeline of one streaming network() step of surface B (zoo-S Separator, batch 32 x 4 s):
what runs between the full-band pipeline and the sub-band pipeline."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth
from spiking_fullsubnet_b200 import Separator
cfg = synth.CFG_ZOO_S
z = np.load(os.path.join(ROOT, "tests", "golden", "zoo_s_1s_weights.npz"))
m = Separator(**cfg)
m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
m = m.eval().cuda()
m.enable_streaming(True)
wave = torch.from_numpy(synth.make_wave(32, 64000, 21)).cuda()
mag = torch.stft(wave, 512, 128, 512, window=torch.hann_window(512, device="cuda"), return_complex=True,
                 pad_mode="constant").abs().contiguous()
graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
m.enable_cuda_graph(graph)
with torch.no_grad():
    for _ in range(3):
        m.network(mag)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.network(mag)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{(e.time_range.start - t0):9.1f} +{e.time_range.end - e.time_range.start:8.1f}  {e.name[:100]}")
