"""Per-launch table of the last bench step in an ncu launch list
(`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)."""
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    rows = list(csv.DictReader(l for l in f if l.startswith('"')))
per, order = {}, []
SCALE = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
for r in rows:
    k = r["ID"]
    if k not in per:
        per[k] = {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]}
        order.append(k)
    per[k][r["Metric Name"]] = float(r["Metric Value"]) * SCALE.get(r["Metric Unit"], 1.0)
starts = [i for i, k in enumerate(order) if "split_kernel" in per[k]["name"]]
tot_ms = tot_gb = 0.0
agg = {}
for k in order[starts[-1]:]:
    x = per[k]
    n = re.sub(r"\(.*", "", x["name"]).replace("void ", "").replace("ar::", "")[:44]
    ms = x.get("gpu__time_duration.sum", 0.0)
    gb = x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
    tot_ms += ms
    tot_gb += gb
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += gb
    extra = f"  {gb:7.2f} GB {gb / ms:6.2f} TB/s" if gb and ms else ""
    print(f"{ms:8.3f} ms {x['grid']:>15} {x['block']:>13} {n}{extra}")
    if "ola_kernel" in n or "ola4_kernel" in n:
        break
print(f"total {tot_ms:.2f} ms, {tot_gb:.1f} GB DRAM")
print("--- by kernel")
for n, (c, ms, gb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.2f} ms {100 * ms / tot_ms:5.1f} % {c:3d} x {n}" + (f"  {gb:7.1f} GB" if gb else ""))
