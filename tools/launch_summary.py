#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.OrderedDict(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    tot[name] = tot.get(name, 0) + v; cnt[name] += 1
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{v:10.1f} us {100*v/T:5.1f}%  n={cnt[k]:4d}  avg {v/cnt[k]:8.1f} us  {k}")
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
