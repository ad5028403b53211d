"""DRAM bytes per launch of every kernel in an .ncu-rep (``ncu --set full`` capture of one cohort pass) ->
the "cohort" block of profiles/traffic.json, which bench.py reads for ``roofline.traffic``.

    python tools/ncu_traffic.py gpurun_out/r02_pass.ncu-rep profiles/traffic.json
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
acc = collections.defaultdict(list)
for row in rows[2:]:
    d, u = dict(zip(hdr, row)), dict(zip(hdr, units))
    name = re.search(r"gk_\w+(<\d+>)?", d["Kernel Name"])
    if not name:
        continue
    total = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[key]]
        total += float(d[key]) * scale
    acc[name.group(0)].append(total)
doc = json.load(open(out_path))
doc["cohort"] = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for k, v in acc.items()}
json.dump(doc, open(out_path, "w"), indent=1)
print({k: round(sum(v) / len(v) / 1e9, 3) for k, v in acc.items()})
