#!/usr/bin/env python
"""Condense ncu output brought back in gpurun_out/ into the small text files kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/r5_launches.csv  profiles/r1_cfg2_bf16_launches.md
    python scripts/ncu_summary.py full     gpurun_out/r5_prof.ncu-rep  profiles/r1_cfg2_bf16_full.md

`launches`: the `--metrics gpu__time_duration.sum` CSV -> per-kernel count / mean / share of the listed time.
`full`    : an `--set full` report -> the handful of raw metrics the roofline argument needs, per launch.
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "launch__shared_mem_per_block_dynamic",
]


def short(name):
    name = name.replace("npml::<unnamed>::", "").replace("npml::", "").replace("void ", "")
    return name.split("(")[0]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
    per = OrderedDict()
    for r in rows:
        per.setdefault(short(r[4]), []).append(float(r[14]) / (1e3 if r[13] in ("ns", "nsecond") else 1.0))
    total = sum(sum(v) for v in per.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary ({} launches, source {})\n\n".format(len(rows), src))
        f.write("Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `{}` | {} | {:.1f} | {:.1f} | {:.1%} |\n".format(k, len(v), sum(v) / len(v), sum(v), sum(v) / total))
        f.write("\ntotal listed device time: {:.1f} us\n".format(total))


def full(src, dst):
    if src.endswith(".csv"):                   # the raw page exported on the GPU box (scripts/profile_round.sh)
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (source {})\n\n".format(src))
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write("## `{}`\n\n| metric | value | unit |\n|---|---|---|\n".format(short(name)))
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| {} | {} | {} |\n".format(k, r[i], units[i]))
            f.write("\n")


def count(src, pattern):
    """number of launches in a launch-list CSV whose kernel name matches `pattern` (for ncu's -s)"""
    import re
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
    print(sum(1 for r in rows if re.search(pattern, r[4])))


if __name__ == "__main__":
    {"launches": launches, "full": full, "count": count}[sys.argv[1]](sys.argv[2], sys.argv[3])
