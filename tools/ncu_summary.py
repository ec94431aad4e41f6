"""Prints the metrics we care about from an .ncu-rep (integer-pipe kernels): python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct"]
import re
# every per-pipe utilisation / instruction-count column this ncu version has (the integer work runs on the FMA-heavy pipe:
# IMAD / IMAD.WIDE; IADD3 / LOP3 / SEL on the ALU pipe)
PIPE = re.compile(r"^(sm__pipe_(fma|fmaheavy|fmalite|alu|xu|fp64)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|lsu|xu|uniform)\.(avg\.pct_of_peak_sustained_active|sum)|"
                  r"smsp__inst_executed_pipe_(fma|fmaheavy|fmalite|alu)\.sum|sm__cycles_elapsed\.max|smsp__cycles_active\.avg)$")
only = sys.argv[2] if len(sys.argv) > 2 else None   # optional substring filter on the kernel name
for idx, r in enumerate(rows[2:]):
    name = r[hdr.index("Kernel Name")]
    if only and only not in name:
        continue
    print(f"---- [{idx}]", name[:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:85s} {r[i][:24]:>24s} {units[i]}")
    for i, h in enumerate(hdr):
        if PIPE.match(h) and h not in WANT:
            print(f"  {h:85s} {r[i][:24]:>24s} {units[i]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v >= 0.25:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:8.2f}")
