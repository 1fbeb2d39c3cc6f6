#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` report (one kernel launch): the
metrics the roofline discussion needs and the source lines (SASS) where the warps wait.

    python scripts/ncu_summary.py gpurun_out/<tag>/<name>.ncu-rep profiles/r02/<name>.txt "<command that was profiled>"
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = ncu(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    lines = [cmd, "Times under ncu are not bench values (serialised, cold cache, ~40 replays).", ""]
    idx = {h: i for i, h in enumerate(hdr)}
    lines.append(f"kernel: {vals[idx['Kernel Name']]}" if "Kernel Name" in idx else "")
    for k in KEYS:
        if k in idx:
            lines.append(f"{k} [{units[idx[k]]}] = {vals[idx[k]]}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            lines.append(f"{h} = {vals[idx[h]]}")
    src = ncu(rep, "source")
    if len(src) > 3:
        h2, data = src[1], src[2:]
        i_s, i_src = h2.index("# Samples"), h2.index("Source")
        stall = {k: h2.index(k) for k in h2 if k.startswith("stall_") and "Not Issued" not in k}
        tot = sum(int(r[i_s] or 0) for r in data)
        top = sorted(range(len(data)), key=lambda i: -int(data[i][i_s] or 0))[:25]
        lines += ["", f"warp-stall samples: {tot}; the 25 instructions with the most samples "
                      "(share, dominant stall reasons, SASS):"]
        for i in sorted(top):
            r = data[i]
            n = int(r[i_s] or 0)
            big = ", ".join(f"{k[6:]}={int(r[c] or 0)}" for k, c in stall.items()
                            if int(r[c] or 0) > 0.2 * max(n, 1))
            lines.append(f"  {100.0*n/max(tot,1):5.1f}%  {r[i_src].strip()[:64]:64s} {big}")
        sass = " ".join(r[i_src] for r in data)
        marks = {m: sass.count(m) for m in ("UBLKCP", "SYNCS.", "LDG.E.128", "LDG.E.64", "LDS",
                                            "DMUL", "DADD", "DFMA", "HMMA", "UTCHMMA", "LDL", "STL")}
        lines += ["", "SASS mnemonics in this kernel: " + ", ".join(f"{k} {v}" for k, v in marks.items())]
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
