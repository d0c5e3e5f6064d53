"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` capture of
bench.py into profiles/conv_traffic_<tag>.json: DRAM bytes per launch of the conv engine (conv_umma2 + conv_chain kernels)
over ONE chunk batch (the last complete pass in the capture).

  python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/conv_traffic_r01_final.json 1184 1184
"""
import csv
import json
import sys

src, dst, chunks, batch = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
with open(src) as f:
    rows = list(csv.DictReader(l for l in f if l.startswith('"')))
per = {}
order = []
for r in rows:
    k = r["ID"]
    if k not in per:
        per[k] = {"name": r["Kernel Name"]}
        order.append(k)
    v = float(r["Metric Value"])
    unit = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    per[k][r["Metric Name"]] = v
ids = [k for k in order if "split_kernel" in per[k]["name"]]
last = order.index(ids[-1])
sel = []
for k in order[last:]:
    n = per[k]["name"]
    if "ola_kernel" in n or "ola4_kernel" in n:
        break
    if "conv_umma2_kernel" in n or "conv_chain_kernel" in n or "conv_umma_kernel" in n:
        sel.append(per[k])
tot = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in sel)
out = {"config": {"chunks_per_step_per_gpu": chunks, "batch_chunks": batch},
       "kernel": "conv_umma2_kernel + conv_chain_kernel (all variants)", "launches": len(sel),
       "dram_bytes_per_launch_avg": tot / max(1, len(sel)), "dram_bytes_total": tot,
       "source": f"{src} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one chain batch)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
