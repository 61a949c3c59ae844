"""Summarise an .ncu-rep (read here, no GPU needed) into a markdown table of the metrics the
roofline discussion uses.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/x.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    for r in rows[2:]:
        print(f"## `{r[idx['Kernel Name']][:100]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in idx:
                print(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[idx[h]]), h.split("stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6])
        print(f"\nwarps stalled per issue-active cycle (top 6): {top}\n")


if __name__ == "__main__":
    main()
