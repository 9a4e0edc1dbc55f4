#!/usr/bin/env python
"""profiles/r02_traffic.json from a full-size `ncu --set full` capture of the two hot kernels.

On the GPU box (after the same command has exited 0 without ncu):
    ncu --set full --clock-control none --import-source on -k regex:dwtsvd -s 2 -c 2 -o gpurun_out/prof_full3000 \
        python bench.py --steps 2 --warmup 1 --frames 3000 --no-e2e --no-cpu-baseline
Here:
    python scripts/ncu_traffic.py gpurun_out/prof_full3000.ncu-rep 3000 > profiles/r02_traffic.json
"""
import csv
import json
import re
import subprocess
import sys

rep, frames = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
         "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
kernels = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = re.sub(r"<.*", "", d["Kernel Name"].split("(")[0]).replace("void ", "").strip()

    def val(k):
        return float(d[k].replace(",", "")) * scale.get(u[k], 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    kernels[name] = {
        "frames": frames, "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
        "ncu_time_s": val("gpu__time_duration.sum"), "inst_executed": val("smsp__inst_executed.sum"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers": int(val("launch__registers_per_thread")), "grid": int(val("launch__grid_size")),
        "block": int(val("launch__block_size")),
    }
print(json.dumps({"source": f"ncu --set full --clock-control none, python bench.py --steps 2 --warmup 1 --frames {frames} "
                            "--no-e2e --no-cpu-baseline, round 2 (scripts/ncu_traffic.py)", "kernels": kernels}, indent=1))
