"""Build tool: markdown table of the `ncu --set full` capture of the streaming recurrence launches.
Usage: python tools/ncu_summary.py gpurun_out/r02b_prof_rec.ncu-rep > table.md"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "duration"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (sm__pipe_tensor_cycles_active)"), ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma instructions % of peak"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("launch__registers_per_thread", "registers"), ("launch__shared_mem_per_block_dynamic", "dyn smem")]
cols = [(c, n) for c, n in cols if c in ci]
modes = {"0": "xproj (bulk copies)", "1": "spike bits + fused W_ih", "2": "bf16x3 operand images + fused W_ih (layer 0)",
         "3": "spike operand image + fused W_ih (layer 1)"}
print("| " + " | ".join(n for _, n in cols) + " |")
print("|" + "---|" * len(cols))
for r in rows[2:]:
    out = []
    for c, _ in cols:
        v = r[ci[c]]
        if c == "Kernel Name":
            v = v.split("(")[0].replace("void ", "")
            a = v[v.index("<") + 1:v.index(">")].split(",") if "<" in v else []
            if len(a) >= 2:
                v += ": " + modes.get(a[1].strip(), "")
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g} {units[ci[c]]}"
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
