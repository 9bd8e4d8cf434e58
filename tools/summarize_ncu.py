#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small tracked text file under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_matrix_r01c.ncu-rep profiles/r01_matrix_kernel.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# summary of {rep} (ncu --set full --clock-control none), one block per profiled launch"]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append("")
        lines.append(f"kernel: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:75s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(v), h) for h, v in d.items()
                         if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)
        lines.append("  top stall reasons (warps per issue-active cycle):")
        for v, h in stalls[:6]:
            lines.append(f"    {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
            lines.append(f"  dram traffic per launch (read+write): {tot:.6e} bytes")
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
