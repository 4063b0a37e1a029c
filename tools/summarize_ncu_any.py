"""One table row per profiled launch of any `ncu --set full` reports (read with `ncu -i <rep> --page raw --csv`, no GPU needed):
time, tensor-pipe / XU / issue utilisation, DRAM bytes and throughput, L2 / L1 throughput.

    python tools/summarize_ncu_any.py out.md "title" rep1.ncu-rep[:label1,label2,...] rep2.ncu-rep ...
"""
import csv
import io
import subprocess
import sys

KEYS = [("time us", "gpu__time_duration.sum", 1e-3), ("SM GHz", "sm__cycles_elapsed.avg.per_second", 1e-9),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("DRAM rd MB", "dram__bytes_read.sum", 1e-6), ("DRAM wr MB", "dram__bytes_write.sum", 1e-6),
        ("DRAM GB/s", "dram__bytes.sum.per_second", 1e-9),
        ("L2 % peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("L1/TEX % peak", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("regs", "launch__registers_per_thread", 1)]
UNIT_SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
              "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9, "byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    dst, title = sys.argv[1], sys.argv[2]
    out = [f"# {title}", "", "`ncu --set full --clock-control none --import-source on`, every kernel profiled alone (replayed, caches flushed between",
           "passes): times are cold-cache and near the maximum SM clock; inside a decode the same kernels run power-capped (DESIGN.md).",
           "The `.ncu-rep` files stay in `gpurun_out/` (scratch).  Read with `tools/summarize_ncu_any.py`.", ""]
    for spec in sys.argv[3:]:
        rep, _, lab = spec.partition(":")
        labels = lab.split(",") if lab else []
        hdr, units, data = load(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        out += [f"## {rep.split('/')[-1]}", "", "| # | label | kernel | grid | " + " | ".join(k[0] for k in KEYS) + " |", "|---|---|---|---|" + "---:|" * len(KEYS)]
        for n, row in enumerate(data):
            name = row[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("cfm::", "")
            grid = row[idx["Grid Size"]] if "Grid Size" in idx else ""
            cells = []
            for _, key, scale in KEYS:
                if key not in idx or row[idx[key]] in ("", "n/a"):
                    cells.append("-")
                    continue
                v = float(row[idx[key]].replace(",", "")) * UNIT_SCALE.get(units[idx[key]], 1.0) * scale
                cells.append("%.4g" % v)
            out.append(f"| {n} | {labels[n] if n < len(labels) else ''} | `{name}` | {grid} | " + " | ".join(cells) + " |")
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
