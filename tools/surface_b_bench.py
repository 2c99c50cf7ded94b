"""Development tool (GPU): surface B (`Separator`, the class the model-zoo checkpoints were trained with) with the TRAINED
zoo weights at BASELINE config-2 size (batch 32 x 4 s): ms per network() step, eager and CUDA-graph replay, and per
forward().  Usage: python tools/surface_b_bench.py [S|L] [batch] [seconds]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import Separator  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "S"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sec = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
cfg = synth.CFG_ZOO_S if size == "S" else synth.CFG_ZOO_L
wfile = os.path.join(ROOT, "tests", "golden", "zoo_s_1s_weights.npz" if size == "S" else "zoo_l_weights.npz")
z = np.load(wfile)
DEV = "cuda:0"
m = Separator(**cfg)
m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
m = m.eval().to(DEV)
L = int(sec * 16000)
wave = torch.from_numpy(synth.make_wave(B, L, 21)).to(DEV)
T = 1 + L // cfg["hop_length"]
mag = torch.stft(wave, 512, 128, 512, window=torch.hann_window(512, device=DEV), return_complex=True,
                 pad_mode="constant").abs().contiguous()
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))


with torch.no_grad():
    for conc in (False, True):
        m.sb_model.concurrent_bands = conc
        m.enable_cuda_graph(False)
        e = timed(lambda: m.network(mag))
        m.enable_cuda_graph(True)
        g = timed(lambda: m.network(mag))
        print(f"zoo-{size} Separator, batch {B} x {sec:g} s (T={T}), sub-bands {'concurrent' if conc else 'serial'}: network eager "
              f"{e:.3f} ms, graph {g:.3f} ms -> {B * T / g / 1e3:.2f} M frames/s")
    m.enable_cuda_graph(False)
    f = timed(lambda: m(wave))
    print(f"   forward() eager {f:.3f} ms -> {B * T / f / 1e3:.2f} M frames/s")
    m.enable_streaming(True)
    if m._stream_plan(B) is None:
        print("   streaming: not co-resident")
    else:
        e = timed(lambda: m.network(mag))
        m.enable_cuda_graph(True)
        g = timed(lambda: m.network(mag))
        f = timed(lambda: m(wave))
        print(f"   STREAMING (two pipelines): network eager {e:.3f} ms, graph {g:.3f} ms -> {B * T / g / 1e3:.2f} M frames/s; "
              f"forward() {f:.3f} ms -> {B * T / f / 1e3:.2f} M frames/s")
