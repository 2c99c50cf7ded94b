"""Development tool: registers / spills per kernel from the ptxas logs of the in-tree build.
Usage: python tools/ptxas_summary.py [substring]"""
import glob
import os
import re
import subprocess
import sys

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spiking_fullsubnet_b200", "csrc", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    cur = None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line) or re.search(r"Function properties for (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and ("spill" in line or "Used" in line):
            if pat in cur or pat in os.path.basename(log):
                name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()[:110]
                print(f"{os.path.basename(log)[:-10]:28s} {name:110s} {line.strip()[:120]}")
