"""Summarise an .ncu-rep (ncu --set full) into the text files kept under profiles/:
   python scripts/ncu_summary.py gpurun_out/r1_tc_conv_15.ncu-rep > profiles/r1_tc_conv_15_summary.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct|lts__throughput\.avg\.pct|"
                  r"launch__(registers_per_thread$|shared_mem_per_block_dynamic|occupancy_limit_|grid_size|block_size|waves)|"
                  r"sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum$|"
                  r"smsp__issue_active\.avg\.pct|sm__inst_executed_pipe_(fma|alu|xu|lsu|uniform|tmem|tc)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_tensor_cycles_active\.avg|sm__mem_tensor_cycles_active\.avg|sm__pipe_tc_cycles_active\.avg|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|smsp__average_warps_issue_stalled_.*_per_issue_active|"
                  r"sm__cycles_elapsed\.avg$|smsp__cycles_active\.avg$)")

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    rec = dict(zip(hdr, vals))
    print(f"Kernel Name [] = {rec.get('Kernel Name')}")
    print(f"Block Size [] = {rec.get('Block Size')}")
    print(f"Grid Size [] = {rec.get('Grid Size')}")
    for h, u, v in sorted(zip(hdr, units, vals)):
        if KEEP.match(h):
            print(f"{h} [{u}] = {v}")
