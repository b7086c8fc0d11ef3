"""One line per kernel launch from an `ncu --page raw --csv` export:
python tools/ncu_table.py gpurun_out/x_raw.csv [> profiles/x_ncu_summary.txt]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wave%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%")]
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
have = [(k, n) for k, n in KEYS if k in h]
print("kernel".ljust(46) + "".join(f"{n:>14s}" for _, n in have))
print("".ljust(46) + "".join(f"{u[h.index(k)][:13]:>14s}" for k, _ in have))
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d["Kernel Name"].split("(")[0].replace("void b200pci::", "").replace("b200pci::", "")[:44]
    print(f"{name:46s}" + "".join(f"{d[k][:13]:>14s}" for k, _ in have))
