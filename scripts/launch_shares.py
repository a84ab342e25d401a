#!/usr/bin/env python
"""Per-kernel shares of an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv):

    python scripts/launch_shares.py gpurun_out/launches.csv profiles/r01z_launches.md "bench.py --steps 4 --warmup 3"
Per-launch times under ncu are cold-cache and serialised, so only the shares are meaningful."""
import csv
import sys
from collections import defaultdict

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [l for l in open(src) if l.startswith('"')]
tot, cnt = defaultdict(int), defaultdict(int)
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    tot[r["Kernel Name"]] += int(float(r["Metric Value"].replace(",", "")))
    cnt[r["Kernel Name"]] += 1
total = sum(tot.values())
with open(dst, "w") as f:
    f.write("# Launch list shares (ncu --metrics gpu__time_duration.sum --clock-control none; %s)\n\n" % cmd)
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
    f.write("| kernel | launches | total ns | share | mean ns |\n|---|---|---|---|---|\n")
    for k in sorted(tot, key=tot.get, reverse=True):
        f.write("| `%s` | %d | %d | %.3f | %d |\n" % (k[:70], cnt[k], tot[k], tot[k] / total, tot[k] // cnt[k]))
print(open(dst).read())
