#!/usr/bin/env python
"""Select the metrics worth keeping from an `ncu --page raw --csv` dump of ONE launch -> "metric,unit,value" rows.
Usage: ncu_raw_select.py raw.csv > profiles/<tag>_metrics.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|dram__bytes_(read|write)\.sum($|\.per_second|\.pct)|gpu__time_duration\.sum|"
                  r"launch__(registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit|waves_per_multiprocessor|block_size|grid_size)|"
                  r"sm__pipe_fp64_cycles_active\.avg\.pct|sm__inst_executed_pipe_(fp64|xu|lsu|alu|fma|uniform).*\.avg\.pct|"
                  r"smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|sm__warps_active\.avg\.pct|sm__throughput\.avg\.pct|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum($|\.pct)|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
                  r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|lts__t_bytes\.sum$|sm__cycles_elapsed\.max|"
                  r"smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on\.sum$|sm__sass_thread_inst_executed\.sum$)")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
print("metric,unit,value")
for h, u, v in sorted(zip(hdr, units, vals)):
    if KEEP.search(h):
        print(f"{h},{u},{v}")
