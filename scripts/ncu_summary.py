#!/usr/bin/env python
"""Summarises one `ncu --set full` capture (.ncu-rep) into the JSON kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep fill_tma_kernel profiles/r02_ncu_full_fill_tma_cfg2.json \
        events_per_gpu=1000000 tile_events=512 tma_stages=6 responses_per_event=50

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed), keeps the launches whose kernel name contains the
given string, and writes per-launch values of the metrics the DESIGN/roofline discussion uses plus
`dram_bytes_per_launch` (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches, in bytes) and the
`config` the capture was taken with (bench.py matches it against the run it has just timed)."""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, kernel, out = sys.argv[1:4]
    cfg = {}
    for kv in sys.argv[4:]:
        k, v = kv.split("=", 1)
        try:
            cfg[k] = int(v)
        except ValueError:
            cfg[k] = v
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    sel = [r for r in rows[2:] if kernel in r[col["Kernel Name"]]]
    if not sel:
        raise SystemExit(f"no launch of {kernel} in {rep}")
    res = {"Kernel Name": sel[0][col["Kernel Name"]], "launches": len(sel), "source": rep, "config": cfg}
    for m in METRICS:
        if m in col:
            res[m] = {"unit": units[col[m]], "per_launch": [r[col[m]] for r in sel]}

    def total(m):
        vals = [float(x.replace(",", "")) for x in res[m]["per_launch"]]
        return sum(vals) / len(vals) * SCALE.get(res[m]["unit"], 1.0)
    if "dram__bytes_read.sum" in res and "dram__bytes_write.sum" in res:
        res["dram_bytes_per_launch"] = total("dram__bytes_read.sum") + total("dram__bytes_write.sum")
    with open(out, "w") as f:
        json.dump(res, f, indent=0)
    print(out, "launches", len(sel), "dram bytes/launch", res.get("dram_bytes_per_launch"))


if __name__ == "__main__":
    main()
