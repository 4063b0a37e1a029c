"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of tools/profile_solve.py into a markdown table:
per-kernel totals of the LAST decode in the file.   python tools/summarize_launches.py in.csv out.md [launches_per_solve]"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
per_solve = int(sys.argv[3]) if len(sys.argv) > 3 else 1245
rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))


def ms(r):
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    return {"us": v / 1e3, "ns": v / 1e6, "s": v * 1e3}.get(u, v)


last = rows[-per_solve:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in last:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += ms(r)
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as f:
    f.write(f"ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), last decode of `{src}`: "
            f"{len(last)} launches, {tot:.2f} ms serialised (cold-cache, per-launch; compare SHARES, not absolutes)\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0005:
            continue
        f.write(f"| `{k[:80]}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |\n")
print(open(dst).read())
