"""Summarise an .ncu-rep: headline metrics, stall reasons, hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [ntop]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 14
if rep.endswith(".csv"):  # already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv)
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
for v in rows[2:]:
    d = dict(zip(h, v))
    print("==", d.get("Kernel Name", "")[:90])
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio"]
    for k in keys:
        if k in d:
            print(f"  {k:78s} {d[k]}")
    st = {}
    for k, x in d.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                st[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(x.replace(",", ""))
            except ValueError:
                pass
    print("  stalls:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda t: -t[1])[:8]))
if "--traffic" in sys.argv:  # machine-readable DRAM traffic of the named kernel -> profiles/r*_traffic.json
    import json
    want = sys.argv[sys.argv.index("--traffic") + 1]
    for v in rows[2:]:
        d = dict(zip(h, v))
        if want in d.get("Kernel Name", ""):
            unit = dict(zip(h, rows[1]))
            def num(k):
                x = float(d[k].replace(",", ""))
                return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit.get(k, "byte"), 1.0)
            print("TRAFFIC_JSON " + json.dumps({
                "kernel": want, "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                "duration_ns_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")), "report": rep}))
            break
if rep.endswith(".csv") or "--no-source" in sys.argv:
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) < 3:
    sys.exit(0)
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    print(f"  {r[ix['Address']][-5:]} {r[ix['Source']][:64]:64s} smp={int(f(r,'# Samples')):6d} "
          f"long={int(f(r,'stall_long_sb'))} short={int(f(r,'stall_short_sb'))} wait={int(f(r,'stall_wait'))} "
          f"math={int(f(r,'stall_math'))} mio={int(f(r,'stall_mio'))} exec={int(f(r,'Instructions Executed'))}")
