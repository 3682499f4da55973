"""Turn an ncu launch list with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum (the committed recipe:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c <n> --csv \
        --log-file gpurun_out/launches.csv python bench.py --no-secondary --steps 4 --warmup 3

) into profiles/r02_traffic.json: DRAM bytes of ONE train step of this build (the last `launches_per_step` launches of the capture), which
bench.py reports as roofline.traffic as long as the kernel sources have not changed since (sources_sha).

    python tools/traffic_from_ncu.py gpurun_out/launches.csv <launches_per_step> [--batch 65536 --precision bf16 --x-dtype bf16]
"""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("launches_per_step", type=int)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--x-dtype", default="bf16")
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_traffic.json"))
a = ap.parse_args()

rows = list(csv.reader(open(a.csv)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
per = {}
order = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    k = int(r[0])
    if k not in per:
        per[k] = {"name": r[ki]}
        order.append(k)
    per[k][r[mi]] = float(r[vi].replace(",", ""))
# the step ends with the optimiser pass: take the last complete step
ends = [i for i, k in enumerate(order) if "adam_kernel" in per[k]["name"]]
if not ends:
    raise SystemExit("no adam_kernel launch in the capture")
last = ends[-1]
step = order[last - a.launches_per_step + 1:last + 1]
if len(step) != a.launches_per_step or (len(ends) > 1 and ends[-1] - ends[-2] != a.launches_per_step):
    raise SystemExit(f"launches_per_step={a.launches_per_step} does not match the capture (adam launches at {ends[-3:]})")
rd = sum(per[k].get("dram__bytes_read.sum", 0.0) for k in step)
wr = sum(per[k].get("dram__bytes_write.sum", 0.0) for k in step)
us = sum(per[k].get("gpu__time_duration.sum", 0.0) for k in step) / 1e3
from bench import sources_sha  # noqa: E402

rec = dict(traffic_bytes_per_step=int(rd + wr), dram_read_bytes=int(rd), dram_write_bytes=int(wr), launches_per_step=a.launches_per_step,
           serialized_us_per_step=us, batch=a.batch, precision=a.precision, x_dtype=a.x_dtype, sources_sha=sources_sha(),
           source=os.path.basename(a.csv))
with open(a.out, "w") as f:
    json.dump(rec, f, indent=1)
print(json.dumps(rec))
for k in step:
    e = per[k]
    nm = e["name"].replace("psvae::", "").replace("__nv_bfloat16", "bf16").replace("void ", "")[:120]
    print(f"{e.get('gpu__time_duration.sum', 0) / 1e3:8.1f} us  rd {e.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB  wr {e.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB  {nm}")
