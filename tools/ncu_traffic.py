"""DRAM traffic per launch of every kernel in an .ncu-rep -> profiles/r2_traffic.json (read by bench.py).

    python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# kernel symbol -> the name bench.py's per-kernel timers use
ALIAS = {"flow_bin_q_kernel": "flow_bin_f16_kernel", "synth_q_kernel": "synth_f16_kernel", "splat_gather_tiled_kernel": "splat_gather_kernel"}

out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in rows[2:]:
        name = r[ik].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        name = ALIAS.get(name, name)
        b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
        out.setdefault(name, []).append(b)
res = {"source": [os.path.basename(r) for r in sys.argv[1:]], "metric": "dram__bytes_read.sum + dram__bytes_write.sum, per launch (mean over the captured launches)",
       "traffic_bytes_per_launch": {k: sum(v) / len(v) for k, v in out.items()}}
with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
