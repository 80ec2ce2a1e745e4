#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the committed, human-readable summaries under profiles/.
usage: python profiles/summarize.py launches <launches.csv> <out.md>
       python profiles/summarize.py kernel <file.ncu-rep> <out.md> [top_n_source_lines]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / total))
    print(open(out).read())


def kernel(rep, out, top=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name.split("(")[0])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, vals[i], units[i]))
            f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        if len(srows) > 2:
            h = srows[0]
            def col(*names):
                for n in names:
                    for i, x in enumerate(h):
                        if x.strip() == n:
                            return i
                return None
            ci = col("# Samples", "Warp Stall Sampling (All Samples)", "Sampling Data (All)")
            si = col("Source")
            if ci is not None and si is not None:
                lines = []
                for r in srows[1:]:
                    try:
                        lines.append((float(r[ci].replace(",", "") or 0), r[si]))
                    except (ValueError, IndexError):
                        pass
                tot = sum(x for x, _ in lines) or 1.0
                f.write("### hottest source lines (share of warp-stall samples)\n\n```\n")
                for smp, text in sorted(lines, key=lambda t: -t[0])[:top]:
                    f.write("%5.1f%%  %s\n" % (100 * smp / tot, text.strip()[:150]))
                f.write("```\n")
    print(open(out).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 25)
