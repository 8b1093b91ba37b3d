#!/usr/bin/env python
"""Key metrics of every kernel launch in an ncu report, as JSON lines:
python scripts/ncu_summary.py <report.ncu-rep> > profiles/<name>.jsonl"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_active",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__warps_active.avg.per_cycle_active": "warps_active",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "sm__cycles_active.avg": "sm_cycles_active",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
for r in rows[2:]:
    d = dict(zip(h, r))
    u = dict(zip(h, units))
    rec = {"kernel": d.get("Kernel Name", "")[:80], "id": d.get("ID")}
    for k, name in want.items():
        if k not in d or d[k] == "":
            continue
        try:
            v = float(d[k].replace(",", ""))
        except ValueError:
            continue
        unit = u.get(k, "")
        if unit in scale and name in ("time", "dram_read", "dram_write", "dyn_smem"):
            v *= scale[unit]
        rec[name] = v
    if "dram_read" in rec and "time" in rec and rec["time"] > 0:
        rec["dram_gbs"] = (rec.get("dram_read", 0) + rec.get("dram_write", 0)) / rec["time"] / 1e9
    print(json.dumps(rec))
