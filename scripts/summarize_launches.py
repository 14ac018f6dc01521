"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total ms, share."""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
ours = sum(v for k, (_, v) in agg.items() if not k.startswith("at::") and "cub::" not in k and "nccl" not in k.lower())
print(f"# {sys.argv[1]}: {sum(n for n, _ in agg.values())} launches, {tot:.2f} ms total (cold-cache, serialised)")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"| `{k[:90]}` | {n} | {v:.3f} | {100 * v / tot:.1f} % |")
