"""Round-1 measurement sweep (GPU): BASELINE.json configs 2, 3 and 5 plus the other shipped sizes.
Writes a markdown table (stdout) for profiles/.  Usage: python tools/sweep.py [quick]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gsn_oracle as O  # noqa: E402  (FLOP bookkeeping + synthetic weights only)
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, efficient_spiking_neuron, ops  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                   "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
    if os.path.exists("MEASURED_PEAKS.json") else 1383.9
dev = "cuda"


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def model_row(size, B, seconds, chunks, reps=5):
    cfg = synth.CONFIGS[size]
    T = 1 + int(seconds * 16000) // 128
    m = SpikingFullSubNet(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
    m = m.eval().to(dev)
    mag = torch.from_numpy(synth.make_mag(B, 257, T, 3)).to(dev)
    backs = sorted({ops.pick_backend(r, h, cfg["shared_weights"]) for _, r, _, h, _ in O.model_rows_and_shapes(cfg, B)})
    with torch.no_grad():
        m.enable_cuda_graph(True, frame_chunks=1)
        ms1 = timed(lambda: m.network(mag), reps)
        m.enable_cuda_graph(True, frame_chunks=chunks)
        msw = timed(lambda: m.network(mag), reps)
    flops = O.algorithmic_flops_per_frame(cfg) * B * T
    best = min(ms1, msw)
    print(f"| {size} | {B} x {seconds:g} s (T={T}) | {'+'.join(backs)} | {ms1:.2f} | {msw:.2f} ({chunks} chunks) | "
          f"{B * T / best * 1e3:,.0f} | {flops / best / 1e9:.2f} | {flops / best / 1e9 / PEAK * 100:.3f} % |", flush=True)
    del m
    torch.cuda.empty_cache()


def stack_row(N, H, K=38, B=32, T=501, reps=5):
    R = B * N
    rs = np.random.RandomState(1)
    stack = efficient_spiking_neuron(K, H, 2, shared_weights=True, bn=True).eval().to(dev)
    x = torch.from_numpy(rs.standard_normal((T, R, K)).astype(np.float32)).to(dev)
    with torch.no_grad():
        ms = timed(lambda: stack(x, None), reps)
    flops = R * T * (2 * H * K + 2 * H * H + 2 * 2 * H * H)  # in->h (both layers) + h->h (both layers)
    print(f"| N={N} (R={R}) | H={H} | {ops.pick_backend(R, H, True)} | {ms:.2f} | {R * T / ms * 1e3:,.0f} | "
          f"{B * T / ms * 1e3:,.0f} | {flops / ms / 1e9:.2f} | {flops / ms / 1e9 / PEAK * 100:.3f} % |", flush=True)


if __name__ == "__main__":
    quick = len(sys.argv) > 1
    print(f"peak = {PEAK} TFLOP/s (measured sustained bf16)\n")
    print("| size | batch x clip | recurrence backend | ms/step serial graph | ms/step wavefront graph | frames/s (best) | "
          "TFLOP/s (algorithmic) | of tensor roofline |")
    print("|---|---|---|---|---|---|---|---|")
    model_row("S", 32, 4, 12)      # BASELINE configs[1]
    model_row("M", 32, 4, 12)
    model_row("XL", 32, 4, 12)
    model_row("L", 32, 4, 12)
    if not quick:
        model_row("L", 64, 10, 16, reps=2)   # BASELINE configs[2]
    print("\nconfig 5: 2-layer GSN stack (input projection + recurrence), K=38, batch 32, T=501, eager launches\n")
    print("| sub-bands | hidden | backend | ms | row-frames/s | utterance-frames/s | TFLOP/s | of tensor roofline |")
    print("|---|---|---|---|---|---|---|---|")
    for N in (15, 31, 63):
        for H in (128, 256, 512):
            stack_row(N, H)
