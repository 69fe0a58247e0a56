"""Summarise an `ncu --set full` report (.ncu-rep) into a small text table: per captured launch the duration,
DRAM bytes (read + write = `traffic`), DRAM / tensor-pipe utilisation, occupancy and registers.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
       python scripts/ncu_summary.py gpurun_out/prof.ncu-rep --traffic profiles/edge_traffic.json
           (also writes the per-launch DRAM traffic of the edge forward / backward entry points, which
            bench.py reports as roofline.traffic)"""
import csv
import json
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


if "--traffic" in sys.argv:
    out_path = sys.argv[sys.argv.index("--traffic") + 1]
    rd = csv.reader(out.splitlines())
    hdr = next(rd)
    next(rd)
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per = {}
    for r in rd:
        name = r[ik]
        mb = float(r[ir].replace(",", "")) + float(r[iw].replace(",", ""))
        key = None
        for frag in ("gat_edge_fwd", "gat_edge_bwd_dst", "gat_edge_bwd_src", "gat_att_reduce_kernel",
                     "gat_att_reduce1", "gat_att_reduce2", "gm_head_sum"):
            if frag in name:
                key = frag
        if key:
            per.setdefault(key, []).append(mb * 1e6)
    avg = {k: sum(v) / len(v) for k, v in per.items()}
    res = {"source": sys.argv[1], "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
           "kernels": avg}
    if "gat_edge_fwd" in avg:
        res["isg_gat_edge_fwd"] = avg["gat_edge_fwd"]
    if "gat_edge_bwd_dst" in avg and "gat_edge_bwd_src" in avg:
        res["isg_gat_edge_bwd"] = sum(avg.get(k, 0.0) for k in ("gat_edge_bwd_dst", "gat_edge_bwd_src",
                                                                 "gat_att_reduce_kernel", "gat_att_reduce1",
                                                                 "gat_att_reduce2", "gm_head_sum"))
    json.dump(res, open(out_path, "w"), indent=1)
