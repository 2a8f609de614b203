#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small JSON for profiles/: duration, DRAM traffic,
pipe utilisation, issue rate, occupancy and the top warp-stall reasons of each captured launch.

    python tools/ncu_summary.py gpurun_out/prof_pfb.ncu-rep > profiles/r01_pfb_ncu.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_cycles_pct",   # FFMA2 holds the pipe two cycles
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block": "smem_per_block",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        stalls = {}
        for h, u, v in zip(hdr, units, r):
            if h in KEYS and v != "":
                x = float(v.replace(",", ""))
                if KEYS[h].startswith("dram_r") or KEYS[h].startswith("dram_w"):
                    d[KEYS[h] + "_bytes"] = to_bytes(x, u)
                elif KEYS[h] == "duration":
                    d["duration_us"] = x * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                else:
                    d[KEYS[h]] = x
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(v), 3)
        d["traffic_bytes"] = d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)
        d["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        out.append(d)
    json.dump({"source": path, "launches": out}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
