#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: python tools/ncu_summary.py REPORT.ncu-rep "header line" > profiles/X.txt
Keeps the metrics DESIGN.md argues with (duration, DRAM bytes, pipe utilisation, issue rate, stall reasons, occupancy,
L2 hit rate, shared-memory conflicts); one column per captured kernel launch."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
                  r"launch__(block_size|grid_size|registers_per_thread|occupancy_limit_\w+|waves_per_multiprocessor)|"
                  r"sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|uniform|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|fmaheavy)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__inst_executed\.avg\.per_cycle_(active|elapsed)|smsp__inst_executed\.avg\.per_cycle_active|"
                  r"smsp__issue_active\.avg\.(per_cycle_active|pct_of_peak_sustained_active)|"
                  r"sm__warps_active\.avg\.(per_cycle_active|pct_of_peak_sustained_active)|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
                  r"lts__t_sector_hit_rate\.pct|lts__t_sectors_srcunit_tex_op_(read|write)\.sum|lts__t_bytes\.sum(\.per_second)?|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__inst_executed\.sum|sm__cycles_elapsed\.max|smsp__cycles_active\.avg)$")


def main():
    rep = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else rep
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, data = rows[0], rows[1], rows[2:]
    print("# " + header)
    for col, name in enumerate(names):
        if name == "Kernel Name" or KEEP.match(name):
            vals = [r[col][:44] for r in data]
            print("%s [%s] %s" % (name, units[col], vals))


if __name__ == "__main__":
    main()
