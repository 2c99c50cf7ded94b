import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import synth
from spiking_fullsubnet_b200 import SpikingFullSubNet, ops
cfg = synth.CFG_S
params = synth.make_params(cfg, 5)
model = SpikingFullSubNet(**cfg)
model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
model = model.eval().cuda()
mag = torch.from_numpy(synth.make_mag(32, 257, 501, 11)).cuda()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def step():
    with torch.no_grad():
        return model.network(mag)
for conc in (True, False):
    model.sb_model.concurrent_bands = conc
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"conc={conc}: back-to-back 20 steps: enqueue {1e3*(t1-t0)/20:.2f} ms/step, total {1e3*(t2-t0)/20:.2f} ms/step")
    evs = []
    t0 = time.perf_counter()
    for _ in range(20):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ms = [a.elapsed_time(b) for a, b in evs]
    print(f"conc={conc}: flush+events: wall {1e3*(t2-t0)/20:.2f} ms/step, event ms: min {min(ms):.2f} med {sorted(ms)[10]:.2f} max {max(ms):.2f}")
    evs = []
    for _ in range(20):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    print(f"conc={conc}: sync-before-each: event ms: min {min(ms):.2f} med {sorted(ms)[10]:.2f} max {max(ms):.2f}")
