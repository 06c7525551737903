"""Summarise `ncu --set full` reports (scripts/gpu_ncu_all.sh) into one text table per kernel: the metrics the roofline
argument needs, read with `ncu -i <rep> --page raw --csv` (no GPU needed).  python scripts/summarize_ncu.py gpurun_out/ncu_r02 > profiles/r02_ncu_summary.txt"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.isfile(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6584.8, "bf16_tflops": 1679.3}
WANT = [
    ("gpu__time_duration.sum", "duration (ns, cold caches, serialised)"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block (B)"), ("launch__waves_per_multiprocessor", "waves / SM"),
    ("dram__bytes_read.sum", "DRAM read (B)"), ("dram__bytes_write.sum", "DRAM written (B)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__t_bytes.sum", "L2 traffic (B)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (HMMA sub-pipe) active, % of ACTIVE cycles"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor instructions, % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (%)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of max)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    vals = rows[-1]
    return dict(zip(hdr, vals)), dict(zip(hdr, rows[1])) if len(rows) > 2 else {}


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


for rep in sorted(glob.glob(os.path.join(sys.argv[1], "*.ncu-rep"))):
    vals, units = raw(rep)
    name = os.path.basename(rep)[:-8]
    print(f"== {name}: {vals.get('Kernel Name', '?')[:90]}")
    for key, label in WANT:
        if key in vals and vals[key] != "":
            print(f"  {label:62s} {vals[key]:>16s} {units.get(key, '')}")
    dur, rd, wr = fnum(vals.get("gpu__time_duration.sum", "")), fnum(vals.get("dram__bytes_read.sum", "")), fnum(vals.get("dram__bytes_write.sum", ""))
    if dur and rd is not None and wr is not None:
        # units: duration ns or us / bytes B, KB, MB depending on ncu's auto-scaling
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rb = rd * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
        wb = wr * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
        ds = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(units.get("gpu__time_duration.sum", "ns"), 1e-9)
        gbs = (rb + wb) / ds / 1e9
        print(f"  {'DRAM bytes / duration':62s} {gbs:16.1f} GB/s = {gbs / PEAKS['hbm_gbs'] * 100:.1f} % of the measured HBM peak ({PEAKS['hbm_gbs']} GB/s)")
    print()
