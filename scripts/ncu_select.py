"""Keep the columns of an `ncu --page raw --csv` export that the round summary uses (the full page has ~2000 columns):
identification, duration, DRAM bytes / throughput, L2 sectors, tensor-pipe activity (every column whose name contains
'pipe_tensor'), shared-memory operand fetch of the tensor core, issue activity, launch geometry.  usage: ncu_select.py in.csv out.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size)$|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"pipe_tensor|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct|l1tex__data_pipe_lsu_wavefronts\.avg\.pct|smsp__issue_active\.avg\.pct|"
                  r"sm__throughput\.avg\.pct|launch__(registers_per_thread|shared_mem_per_block_dynamic|grid_size|block_size)$|lts__t_sectors_op_(read|write)\.sum$|"
                  r"sm__warps_active\.avg\.pct")
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = [i for i, name in enumerate(rows[h]) if KEEP.search(name)]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows[h:]:
        if len(r) == len(rows[h]):
            w.writerow([r[i] for i in cols])
print("kept", len(cols), "of", len(rows[h]), "columns,", len(rows) - h - 2, "launches")
