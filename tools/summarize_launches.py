"""Summarise an ncu launch list CSV of tools/profile_solve.py (gpu__time_duration.sum and optionally
dram__bytes_read.sum / dram__bytes_write.sum per launch) into a markdown table: per-kernel totals of the LAST decode.
    python tools/summarize_launches.py in.csv out.md [launches_per_solve] [traffic.json]"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
per_solve = int(sys.argv[3]) if len(sys.argv) > 3 else 1245
tjson = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
metrics = sorted({r["Metric Name"] for r in rows})
per_launch = collections.OrderedDict()
for r in rows:
    per_launch.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
launches = list(per_launch.values())[-per_solve:]


def ms(v):
    x, u = v
    return {"us": x / 1e3, "ns": x / 1e6, "s": x * 1e3, "usecond": x / 1e3, "nsecond": x / 1e6, "msecond": x}.get(u, x)


def mb(v):
    x, u = v
    return {"byte": x / 1e6, "Kbyte": x / 1e3, "Mbyte": x, "Gbyte": x * 1e3}.get(u, x)


agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for l in launches:
    name = re.sub(r"\(.*", "", l["name"]).replace("void ", "")
    a = agg[name]
    a[0] += 1
    a[1] += ms(l["gpu__time_duration.sum"])
    if "dram__bytes_read.sum" in l:
        a[2] += mb(l["dram__bytes_read.sum"]) + mb(l["dram__bytes_write.sum"])
tot = sum(v[1] for v in agg.values())
tot_mb = sum(v[2] for v in agg.values())
with open(dst, "w") as f:
    f.write(f"ncu launch list (`--clock-control none`; metrics {', '.join(metrics)}), last decode of `{src}`: "
            f"{len(launches)} launches, {tot:.2f} ms serialised (cold-cache, per-launch; compare SHARES, not absolutes)"
            + (f", DRAM traffic {tot_mb / 1e3:.2f} GB per decode" if tot_mb else "") + "\n\n")
    f.write("| kernel | launches | total ms | share | DRAM MB |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0005:
            continue
        f.write(f"| `{k[:80]}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {v[2]:.0f} |\n")
if tjson and tot_mb:
    json.dump({"dram_bytes_per_decode": tot_mb * 1e6, "source": f"sum of ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(launches)} launches of one cfg2 decode ({src})"}, open(tjson, "w"))
print(open(dst).read())
