"""Writes profiles/ncu_traffic.json from an ncu capture of the rollout kernel of THIS build:
  {"src_sha16": sha256 of the library's sources (bench._src_sha16), "traffic_bytes": {workload: dram__bytes_read.sum + dram__bytes_write.sum of one launch}}
bench.py reports roofline.traffic only when the sources of the library it times have the same sha (otherwise null).
  python tools/ncu_traffic.py <workload> <raw page csv made with: ncu -i X.ncu-rep --page raw --csv>"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
workload, path = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
best = None
for vals in rows[2:]:                                     # the rollout_kernel launch with the longest duration
    d = dict(zip(hdr, zip(units, vals)))
    if "rollout_kernel" not in d["Kernel Name"][1]:
        continue
    dur = float(d["gpu__time_duration.sum"][1])
    if best is None or dur > best[0]:
        tr = sum(float(d[k][1]) * UNIT[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        best = (dur, tr, d["gpu__time_duration.sum"][0])
sys.path.insert(0, ROOT)
import bench  # noqa: E402
sha = bench._src_sha16()
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
if out.get("src_sha16") != sha:
    out = {"src_sha16": sha, "traffic_bytes": {}, "duration": {}, "kernel": {}, "kernel_sass_sha16": {}}
out.setdefault("kernel", {})[workload] = bench.BOX_ROLLOUT_KERNEL        # the capture scripts profile box-object workloads
out.setdefault("kernel_sass_sha16", {})[workload] = bench._kernel_sass_sha16(bench.BOX_ROLLOUT_KERNEL)
out["traffic_bytes"][workload] = best[1]
out["duration"][workload] = f"{best[0]} {best[2]}"
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out))
