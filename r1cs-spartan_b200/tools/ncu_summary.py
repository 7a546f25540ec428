"""Condense an Nsight Compute report into the text summary kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python r1cs-spartan_b200/tools/ncu_summary.py raw.csv > profiles/rNN_ncu_<kernel>.txt

Prints, per captured launch, the metrics the roofline discussion in DESIGN.md uses (duration, grid, registers,
integer-pipe activity, DRAM bytes, cache hit rates, active lanes per instruction, local-memory traffic) and every
issue-stall reason.  `--launches` turns a `--metrics gpu__time_duration.sum` launch list (csv) into per-kernel shares.
"""
import collections
import csv
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_elapsed.max",
]


def summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-95s %s %s" % (w, r[i][:220], units[i]))
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                print("%-95s %s" % (h, r[i]))
        print()


def launches(path, marker=None, which=-2):
    """marker: a kernel that runs once per proof (k_pair_diff opens every proof): only the launches between two of its
    occurrences are counted (`which` = index of the first one; -2 = the last complete proof of the capture)."""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [r for r in data if len(r) > mv]
    if marker:
        idx = [i for i, r in enumerate(data) if marker in r[kn]]
        data = data[idx[which]:idx[which + 1]] if which + 1 != 0 else data[idx[which]:]
        print("one proof: the launches from one %s to the next" % marker)
    acc = collections.OrderedDict()
    for r in data:
        if len(r) > mv:
            name = r[kn].split("(")[0]
            acc[name] = acc.get(name, 0.0) + float(r[mv])
    tot = sum(acc.values())
    print("%d launches, %.3f ms" % (len(data), tot / 1e6))
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print("%-70s %9.3f ms %5.1f%%" % (k[:70], v / 1e6, 100 * v / tot))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        summary(sys.argv[1])
