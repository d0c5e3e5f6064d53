"""Per-launch table of one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(r) if 'split_kernel' in x['Kernel Name']]
s = idx[-1]
e = len(r)
tot = 0.0
for x in r[s:e]:
    n = re.sub(r'\(.*', '', x['Kernel Name']).replace('void ', '').replace('ar::', '')[:48]
    t = float(x['Metric Value']) / 1e6
    tot += t
    print(f"{t:8.3f} ms  {x['Grid Size']:>15} {x['Block Size']:>13} {n}")
    if 'ola_kernel' in n:
        break
print(f"total {tot:.2f} ms")
