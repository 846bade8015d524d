"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv, re, sys
from collections import OrderedDict

src, title = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = rows[0]
iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iU = hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iU], 1.0)
    name = re.sub(r"\(.*", "", r[iK]).replace("nlmc::", "").replace("void ", "").strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {title}\n")
print("(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n")
print("| kernel | launches | total (ns) | mean (ns) | share |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.0f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
