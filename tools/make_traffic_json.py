"""profiles/traffic.json from an ncu metrics pass over ONE batch of the bench workload:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_extend|k_shade' \
      --csv --log-file gpurun_out/traffic.csv python tools/profile_run.py 1 Atrium
Per kernel: number of launches, DRAM bytes (read + write) summed over the launches and per launch - the per-launch figure `bench.py`
reports as roofline.traffic beside the algorithmic bytes of the same launches.
usage: python tools/make_traffic_json.py gpurun_out/traffic.csv profiles/traffic.json [rays_of_the_batch]"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mv, mu, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
acc = collections.defaultdict(lambda: collections.defaultdict(float))
launches = collections.defaultdict(set)
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[kn]))
    key = "k_extend" if "k_extend" in name else ("k_shade" if "k_shade" in name else None)
    if not key:
        continue
    launches[key].add(r[idc])
    acc[key][r[mn]] += float(r[mv].replace(",", "")) * scale.get(r[mu], 1)
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over every k_extend / k_shade launch of one batch of the bench workload "
                 "(python tools/profile_run.py 1 Atrium; profiles/r2_traffic_launches.csv)"}
rays = float(sys.argv[3]) if len(sys.argv) > 3 else None
for key in ("k_extend", "k_shade"):
    n = len(launches[key])
    byt = acc[key]["dram__bytes_read.sum"] + acc[key]["dram__bytes_write.sum"]
    out[key + "_launches"] = n
    out[key + "_dram_bytes_total"] = byt
    out[key + "_dram_bytes_per_launch"] = byt / max(n, 1)
    out[key + "_ms_total_under_ncu"] = acc[key]["gpu__time_duration.sum"]
    if rays:
        out[key + "_dram_bytes_per_ray"] = byt / rays
if rays:
    out["rays_of_the_batch"] = rays
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
