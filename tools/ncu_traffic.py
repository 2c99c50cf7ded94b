"""Build tool: sums dram__bytes_read.sum + dram__bytes_write.sum over the recurrence launches of ONE eager step from
an `ncu --set full` report and writes profiles/r01_ncu_traffic.json (bench.py copies it into roofline.traffic).
Usage: python tools/ncu_traffic.py gpurun_out/r01_prof_rec.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * SCALE[units[col[name]]]


launches = []
for r in rows[2:]:
    if "k_recurrence_tc" not in r[col["Kernel Name"]]:
        continue
    launches.append({"kernel": r[col["Kernel Name"]].split("(")[0], "grid": r[col["Grid Size"]],
                     "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                     {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[col["gpu__time_duration.sum"]]],
                     "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum")})
out = {"source": f"ncu --set full, {len(launches)} recurrence launches of one eager step of bench.py (configs[1]); "
                 f"dram__bytes_read.sum + dram__bytes_write.sum summed over them",
       "dram_bytes_per_step": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches),
       "launches": launches}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_ncu_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
