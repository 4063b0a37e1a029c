"""Summarise `ncu --set full` captures of the GEMM and attention kernels into a markdown file for profiles/.

    python tools/summarize_ncu_full.py gemm.ncu-rep attn.ncu-rep out.md [title]

The reports are read with `ncu -i <rep> --page raw --csv` (no GPU needed).  The GEMM report is expected to hold the first 12
GEMM launches of one estimator evaluation of cfg2 (tools/profile_solve.py, `-k regex:gemm_tc --launch-count 12`), the attention
report one full-resolution launch."""
import csv
import io
import subprocess
import sys

GEMM_LABELS = ["conv1 k3 K=256 (STATS)", "res_conv 1x1 (f32)", "conv2 k3 (STATS, pair)", "QKV (STORE)", "out-proj (RESID, TMA reduce-add)",
               "FF1 (SNAKE)", "FF2 (RESID, pair, TMA reduce-add)", "QKV (STORE)", "out-proj (RESID)", "FF1 (SNAKE)",
               "FF2 (RESID+copy, pair)", "down conv k3 s2 (MASK, pair)"]
GEMM_SEL = [0, 1, 2, 3, 4, 5, 6, 10, 11]
GEMM_KEYS = [("time [us]", "gpu__time_duration.sum"), ("SM clock [GHz]", "sm__cycles_elapsed.avg.per_second"),
             ("TMA load L2->SM [TB/s]", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second"),
             ("smem bank reads % of peak", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed"),
             ("smem bank writes % of peak", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed"),
             ("LSU smem wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
             ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
             ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
             ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             ("dram read [MB]", "dram__bytes_read.sum"), ("dram write [MB]", "dram__bytes_write.sum"),
             ("warp inst", "smsp__inst_executed.sum"), ("regs/thread", "launch__registers_per_thread")]
ATTN_KEYS = [("time [us]", "gpu__time_duration.sum"), ("SM clock [GHz]", "sm__cycles_elapsed.avg.per_second"),
             ("cycles", "sm__cycles_elapsed.max"), ("warp inst", "smsp__inst_executed.sum"),
             ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             ("XU (MUFU ex2) pipe % of peak", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
             ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs/thread", "launch__registers_per_thread"),
             ("dram read [MB]", "dram__bytes_read.sum"), ("dram write [MB]", "dram__bytes_write.sum"),
             ("stall samples: wait (fixed-latency dependency)", "smsp__pcsamp_warps_issue_stalled_wait"),
             ("stall samples: long scoreboard (TMEM / mbarrier)", "smsp__pcsamp_warps_issue_stalled_long_scoreboard"),
             ("stall samples: selected (issuing)", "smsp__pcsamp_warps_issue_stalled_selected"),
             ("stall samples: branch resolving (mbarrier spin)", "smsp__pcsamp_warps_issue_stalled_branch_resolving"),
             ("stall samples: math pipe throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle")]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def fmt(v):
    try:
        return "%.5g" % float(v.replace(",", ""))
    except ValueError:
        return v


def main():
    gemm_rep, attn_rep, dst = sys.argv[1], sys.argv[2], sys.argv[3]
    title = sys.argv[4] if len(sys.argv) > 4 else "ncu --set full captures"
    out = [f"# {title} (cfg2, first NFE of a cold decode; `--clock-control none --import-source on`)", "",
           "The `.ncu-rep` files stay in `gpurun_out/` (scratch); these are the numbers read from them with `ncu -i ... --page raw --csv`",
           "(`tools/summarize_ncu_full.py`).  Each kernel is profiled alone (replayed ~40 times), so the SM clock is near its maximum here;",
           "inside a decode the same kernels run power-capped (`sw_power_cap`, ~984 W), see DESIGN.md section 7.", ""]
    hdr, _, data = load(gemm_rep)
    idx = {h: i for i, h in enumerate(hdr)}
    sel = [i for i in GEMM_SEL if i < len(data)]
    out += [f"## {gemm_rep.split('/')[-1]} (first 12 GEMM launches of the estimator)", "",
            "| metric | " + " | ".join(GEMM_LABELS[i] for i in sel) + " |", "|---|" + "---:|" * len(sel),
            "| kernel | " + " | ".join(data[i][idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("cfm::", "") for i in sel) + " |"]
    for name, k in GEMM_KEYS:
        if k in idx:
            out.append(f"| {name} | " + " | ".join(fmt(data[i][idx[k]]) for i in sel) + " |")
    hdr, _, data = load(attn_rep)
    idx = {h: i for i, h in enumerate(hdr)}
    out += ["", f"## {attn_rep.split('/')[-1]} (attention, full resolution, 1536 CTAs)", "", "| metric | value |", "|---|---:|"]
    for name, k in ATTN_KEYS:
        if k in idx:
            out.append(f"| {name} | {fmt(data[0][idx[k]])} |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
