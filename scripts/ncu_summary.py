#!/usr/bin/env python
"""Summarise an ncu report (``ncu --set full``) into a small text file for profiles/: per kernel the duration, DRAM
traffic, L2 traffic, occupancy, issue rate, stall breakdown and the top stalled SASS instructions.
usage: ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.txt [points_per_launch]"""
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    npts = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
            "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]
    lines = []
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("=== " + name)
        vals = {}
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = r[i]
                lines.append("  %-62s %s %s" % (w, r[i], units[i]))
        try:
            rd, wr = float(vals["dram__bytes_read.sum"]), float(vals["dram__bytes_write.sum"])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd *= scale[units[hdr.index("dram__bytes_read.sum")]]
            wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
            if npts:
                lines.append("  DRAM traffic per grid point: read %.1f B, write %.1f B, total %.1f B" % (rd / npts, wr / npts, (rd + wr) / npts))
        except Exception:
            pass
        lines.append("  warp stall reasons (warps per issue-active cycle):")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(r[i] or 0) > 0.05:
                lines.append("    %-28s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], r[i]))
    src = page(rep, "source", ["--print-source", "sass"])
    cur, h = None, None
    kernels = []
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            h = r
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for k in kernels:
        if k["name"] in seen:
            continue
        seen.add(k["name"])
        iS, iSrc = h.index("# Samples"), h.index("Source")
        tot = sum(int(r[iS]) for r in k["rows"]) or 1
        lines.append("=== top stalled SASS of " + k["name"] + " (%d instructions, %d samples)" % (len(k["rows"]), tot))
        for r in sorted(k["rows"], key=lambda x: -int(x[iS]))[:16]:
            lines.append("  %5.1f%%  %s" % (100.0 * int(r[iS]) / tot, r[iSrc].strip()))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
