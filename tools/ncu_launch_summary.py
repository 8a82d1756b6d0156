"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
Usage: python tools/ncu_launch_summary.py LAUNCHES.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot, cnt, grid = defaultdict(float), defaultdict(int), {}
for r in rows:
    name = r[4].replace("solo::", "")
    key = (name, r[8], r[7])
    tot[key] += float(r[14]) / 1e3
    cnt[key] += 1
all_us = sum(tot.values())
print(f"# {len(rows)} launches, {all_us:.1f} us of kernel time (cold-cache, serialised by ncu: shares are meaningful, absolute times are not)")
for key in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[key]:10.1f} us {100 * tot[key] / all_us:5.1f} %  x{cnt[key]:<4d} avg {tot[key] / cnt[key]:8.1f} us  grid {key[1]:<14s} block {key[2]:<13s} {key[0][:110]}")
