import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import synth
from spiking_fullsubnet_b200 import SpikingFullSubNet
cfg = synth.CONFIGS["S"]
m = SpikingFullSubNet(**cfg)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
m = m.eval().cuda()
m.enable_streaming(True)
wave = torch.from_numpy(synth.make_wave(32, 64000, 3)).cuda()
with torch.no_grad():
    for _ in range(3): m(wave)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(wave); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{(e.time_range.start - t0):9.1f} +{e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
