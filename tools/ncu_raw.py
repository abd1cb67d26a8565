"""print selected raw metrics of every kernel in an .ncu-rep:  python tools/ncu_raw.py report.ncu-rep [regex]"""
import csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|registers_per_thread|occupancy_limit|warps_active.avg.pct_of_peak_sustained_active|"
                 r"sm__inst_executed.avg.per_cycle_elapsed$|smsp__inst_executed.sum$|dram__bytes_(read|write).sum$|"
                 r"issue_stalled.*per_issue_active|pipe_fma(heavy|lite)?_cycles_active.avg.pct_of_peak_sustained_elapsed|"
                 r"pipe_(lsu|alu|fp64|xu|uniform).*pct_of_peak_sustained_active|bank_conflicts_pipe_lsu_mem_shared.sum|"
                 r"thread_inst_executed_per_inst_executed.ratio|sm__throughput.avg.pct")
for r in rows[2:]:
    name = r[H.index("Kernel Name")] if "Kernel Name" in H else "?"
    print("==", name[:100])
    for i, h in enumerate(H):
        if pat.search(h):
            try:
                v = float(r[i].replace(",", ""))
                if "issue_stalled" in h and v < 0.2:
                    continue
            except ValueError:
                pass
            print("  %-90s %s %s" % (h, r[i], rows[1][i]))
