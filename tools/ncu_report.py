#!/usr/bin/env python
"""Text summary of one kernel of a `ncu --set full [--import-source on]` report for profiles/ (read here, no GPU needed):
    python tools/ncu_report.py gpurun_out/x.ncu-rep [kernel-name-substring] [header line ...] > profiles/x.txt
Prints the launch's key metrics (raw page), stall reasons per issue, and -- if the report holds source -- warp-state samples
and the SASS instructions that collected the most samples."""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
           "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    for line in sys.argv[3:]:
        print(line)
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    row = next((r for r in rows[2:] if want in r[ki]), None)
    if row is None:
        print("(no launch of a kernel matching '%s' in %s)" % (want, rep))
        return
    print("kernel: %s" % row[ki])
    print()
    for m in METRICS:
        if m in hdr:
            print("%-84s %s %s" % (m, row[hdr.index(m)], units[hdr.index(m)]))
    stall = [(h, row[i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if stall:
        print("\nstall reasons per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
        for h, v in sorted(stall, key=lambda kv: -float(kv[1] or 0)):
            if float(v or 0) >= 0.05:
                print("  %-28s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(v)))
    src = page(rep, "source")
    # the source page lists kernels one after another: find the block of the wanted kernel
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1] if len(r) > 1 else "", hdr=None, rows=[])
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and r:
            cur["rows"].append(r)
    blk = next((b for b in blocks if want in b["name"] and b["rows"]), None)
    if blk is None:
        return
    h = blk["hdr"]
    ix = {c: i for i, c in enumerate(h)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    total = sum(f(r, "# Samples") for r in blk["rows"])
    execd = sum(f(r, "Instructions Executed") for r in blk["rows"])
    print("\nwarp instructions executed: %d over %d SASS instructions; warp-state samples: %d" % (execd, len(blk["rows"]), total))
    if total:
        reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = {c: sum(f(r, c) for r in blk["rows"]) for c in reasons}
        for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
            if v:
                print("  %-26s %5.1f %%" % (c, 100.0 * v / total))
        print("\ninstructions with the most samples:")
        for r in sorted(blk["rows"], key=lambda r: -f(r, "# Samples"))[:12]:
            top = max(reasons, key=lambda c: f(r, c))
            print("  %5.1f %%  %-22s %s" % (100.0 * f(r, "# Samples") / total, top, r[ix["Source"]].strip()[:90]))
    print("\nmost executed SASS regions (runs of instructions with the same execution count):")
    runs, cur = [], None
    for i, r in enumerate(blk["rows"]):
        e = int(f(r, "Instructions Executed"))
        if cur and cur[2] == e:
            cur[1] = i
        else:
            cur = [i, i, e]
            runs.append(cur)
    for a, b, e in sorted(runs, key=lambda t: -(t[1] - t[0] + 1) * t[2])[:8]:
        print("  %5.1f %% of executed instructions: %4d instructions x %7d executions, first: %s" %
              (100.0 * (b - a + 1) * e / max(execd, 1), b - a + 1, e, blk["rows"][a][ix["Source"]].strip()[:70]))


if __name__ == "__main__":
    main()
