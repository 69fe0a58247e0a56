"""Summarise an `ncu --set full` report (.ncu-rep) into a small text table: per captured launch the duration,
DRAM bytes (read + write = `traffic`), DRAM / tensor-pipe utilisation, occupancy and registers.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = csv.reader(out.splitlines())
hdr = next(rd)
units = next(rd)
idx = [(hdr.index(k) if k in hdr else -1, short) for k, short in WANT]
print("# " + sys.argv[1] + "  (ncu --set full --clock-control none; per-launch values; units: " +
      ", ".join(f"{s}={units[i]}" for i, s in idx if i >= 0 and units[i]) + ")")
for r in rd:
    cells = []
    for i, short in idx:
        if i < 0:
            continue
        v = r[i]
        if short == "kernel":
            v = v.replace("void <unnamed>::", "").split("(")[0]
        else:
            try:
                v = f"{float(v.replace(',', '')):.2f}"
            except ValueError:
                pass
        cells.append(f"{short}={v}")
    print("  ".join(cells))
