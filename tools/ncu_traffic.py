#!/usr/bin/env python3
"""profiles/traffic.json from an ncu CSV that carries dram__bytes_read.sum, dram__bytes_write.sum and gpu__time_duration.sum per launch
(tools/ncu_round.sh): the average DRAM bytes per launch of the two hot kernels over whole bench frames, which bench.py reports as
roofline.traffic next to the algorithmic bytes per launch.

  python tools/ncu_traffic.py gpurun_out/TAG_traffic_bench.csv profiles/traffic.json "source note"
"""
import collections
import csv
import io
import json
import sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
per = collections.defaultdict(lambda: collections.defaultdict(float))
unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("vk::", "")
    if name.startswith(("k_shade<2>", "k_shade<(int)2>", "k_shade<2, 0>", "k_shade<(int)2, (bool)0>")):
        key = "k_shade<hero>"
    elif name.startswith("k_trace<0") or name.startswith("k_trace<(bool)0"):
        key = "k_trace"
    else:
        continue
    v = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r["Metric Unit"], 1.0)
    per[key][r["Metric Name"]] += v
    if r["Metric Name"] == "gpu__time_duration.sum":
        per[key]["launches"] += 1
out = {"_source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}
for k, m in per.items():
    n = max(m["launches"], 1)
    out[k] = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / n
    out["_" + k] = {"launches": int(n), "dram_read_bytes_per_launch": m["dram__bytes_read.sum"] / n, "dram_write_bytes_per_launch": m["dram__bytes_write.sum"] / n,
                    "ncu_ms_per_launch": m["gpu__time_duration.sum"] / n}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
