#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md.

    python scripts/ncu_summary.py gpurun_out/prof_r1a.ncu-rep profiles/r01a_ncu_full.md "note"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/tex %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary: %s" % rep.split("/")[-1], "", note, "",
             "Per launch (ncu replays each kernel ~40x with cold caches; durations are for shares, not bench values).", ""]
    for r in data:
        name = r[idx["Kernel Name"]]
        lines.append("## %s" % name[:110])
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for m, label in METRICS:
            if m in idx and r[idx[m]] != "":
                lines.append("| %s (`%s`) | %s %s |" % (label, m, r[idx[m]], units[idx[m]]))
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            du = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            ur, ud = units[idx["dram__bytes_read.sum"]], units[idx["gpu__time_duration.sum"]]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[ur]
            tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[ud]
            lines.append("| **traffic (dram read+write)** | %.1f MB -> %.0f GB/s over the ncu duration |"
                         % ((rd + wr) * scale / 1e6, (rd + wr) * scale / (du * tscale) / 1e9))
        except Exception:                                             # noqa: BLE001
            pass
        lines.append("")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
