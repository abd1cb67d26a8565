"""Summarise ncu output for profiles/: a launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals
and share of the step, and an `ncu --set full` report (.ncu-rep) into the few numbers DESIGN.md quotes.
usage: python tools/ncu_summary.py launches.csv [report.ncu-rep ...]"""
import collections
import csv
import re
import subprocess
import sys


def short(name):
    m = re.match(r"(?:void )?(?:xrd::)?(\w+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")).replace("xrd::", "").replace("(int)", "") if m else name[:60]


def launches(path, skip_pids_before=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        t = tot.setdefault(k, [0, 0.0, r[7], r[8]])
        t[0] += 1
        t[1] += float(r[-1]) / 1e6
    total = sum(v[1] for v in tot.values())
    print("| kernel | launches | block | grid (first) | total ms | share |")
    print("|---|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %s | %s | %.3f | %.1f %% |" % (k, v[0], v[2], v[3], v[1], 100 * v[1] / total))
    print("| **all** | %d | | | %.3f | |" % (len(rows), total))


WANT = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per instruction"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("\n**`%s`** (%s)\n" % (short(r[hdr.index("Kernel Name")]), path.split("/")[-1]))
        print("| metric | value |")
        print("|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))


if __name__ == "__main__":
    launches(sys.argv[1])
    for p in sys.argv[2:]:
        report(p)
