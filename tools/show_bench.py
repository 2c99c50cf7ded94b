"""Development tool: one-screen summary of a bench.py JSON line.  Usage: python tools/show_bench.py file.json"""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head -1` is the usual way to call this

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.4g}  launches {d.get('gpu_launches')}")
r = d.get("roofline", {})
print({k: r[k] for k in ("achieved", "frac", "traffic", "kernel_ms_per_step", "share_of_step") if k in r})
for l in r.get("latency_model", []):
    print(f"  rows {l['rows']:4d} H {l['hidden']:4d} fused {l['fused_input']:4d}  {l['us_per_frame']:.3f} us/frame  mma {l['mma_per_frame']}")
