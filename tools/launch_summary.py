"""Developer helper: per-kernel count / median / total of an `ncu --metrics gpu__time_duration.sum --csv` log."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    tot[r[ki].split("(")[0][-48:]].append(v)
total = sum(sum(v) for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:50s} n={len(v):4d} median {sorted(v)[len(v) // 2]:8.1f} us  total {sum(v):10.0f} us  {100 * sum(v) / total:5.1f} %")
