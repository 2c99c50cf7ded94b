"""Development tool (GPU): device-side timeline of one CUDA-graph replay of the hot path (gsn_trace_set)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from spiking_fullsubnet_b200 import SpikingFullSubNet, _lib  # noqa: E402

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = sys.argv[2] if len(sys.argv) > 2 else "S"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 501
cfg = synth.CONFIGS[size]
model = SpikingFullSubNet(**cfg)
model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
model = model.eval().cuda()
mag = torch.from_numpy(synth.make_mag(batch, 257, frames, 11)).cuda()
lib = _lib.load()
buf = torch.zeros(64 + 32 * 16384, dtype=torch.uint8, device="cuda")
_lib.check(lib.gsn_trace_set(buf.data_ptr(), buf.numel()))
if os.environ.get("GSN_TRACE_ALL_CTAS"):  # header word 15 bit 0: one record per recurrence CTA (kind 5)
    buf[60:64] = torch.tensor([1, 0, 0, 0], dtype=torch.uint8, device="cuda")
if os.environ.get("GSN_TIMELINE_STREAM"):
    model.enable_streaming(True)
model.enable_cuda_graph(True, frame_chunks=chunks)
with torch.no_grad():
    model.network(mag)  # capture (+ warm-up)
    torch.cuda.synchronize()
    buf[:4].zero_()
    model.network(mag)
    torch.cuda.synchronize()
raw = buf.cpu().numpy()
n = int(raw[:4].view(np.uint32)[0])
rec = raw[64:64 + 32 * n].view(np.dtype([("t0", "<u8"), ("t1", "<u8"), ("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4")]))
t_min = rec["t0"].min()
names = {1: "linear", 2: "recur", 3: "feat", 4: "lin_tc", 5: "rcta", 6: "pre", 7: "xplane"}
print(f"{n} traced launches, span {(max(rec['t1'].max(), rec['t0'].max()) - t_min) / 1e3:.1f} us")
for r in sorted(rec, key=lambda r: r["t0"]):
    dur = (int(r["t1"]) - int(r["t0"])) / 1e3 if r["t1"] else 0
    extra = f"  SM clock {int(r['c']) * 1024 / dur:.0f} MHz" if (int(r['kind']) == 2 and dur > 0 and os.environ.get("GSN_TIMELINE_STREAM")) else ""
    print(f"{(int(r['t0']) - int(t_min)) / 1e3:9.1f} us  +{dur:7.1f}  {names.get(int(r['kind']), '?'):6s} {int(r['a']):7d} {int(r['b']):5d} {int(r['c']):4d}{extra}")
