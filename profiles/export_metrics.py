"""ncu report -> profiles/<name>.metrics.csv (metric,unit,value of the first kernel in the report), so that the numbers
quoted in DESIGN.md / bench.py can be checked without the binary .ncu-rep (gpurun_out/ is scratch).
usage: python profiles/export_metrics.py gpurun_out/prof_X.ncu-rep [out.csv]"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), os.path.basename(rep).replace(".ncu-rep", ".metrics.csv"))
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
skip = {"ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"}
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        if h in skip or "Triage" in h or v == "" or h.startswith(("smsp__pcsamp", "sass__", "pmsampling", "device__", "nvlink", "pcie", "numa", "syslts", "syslrc", "lrc", "idc", "gcc", "gr__", "gpc", "profiler", "sm__ops_path", "sm__sass_inst_executed_op")):
            continue
        w.writerow([h, u, v])
print(out)
