#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here on the CPU box with `ncu -i`) into a small JSON + markdown table.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME"""
import csv
import io
import json
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read_GB",
    "dram__bytes_write.sum": "dram_write_GB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1tex_lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_dmma_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio") and "warps_issue" in h]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]][:90]}
        for m, name in METRICS.items():
            if m in idx and r[idx[m]] != "":
                try:
                    d[name] = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    d[name] = r[idx[m]]
                if units[idx[m]] and name.endswith("_ms") is False and "GB" in name and units[idx[m]].lower().startswith("mbyte"):
                    d[name] /= 1e3
        top = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in stall), reverse=True)[:5]
        d["top_stalls"] = [[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(v, 2)] for v, h in top]
        if "dram_read_GB" in d and "dram_write_GB" in d:
            d["dram_total_GB"] = d["dram_read_GB"] + d["dram_write_GB"]
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    cols = ["duration_ms", "dram_total_GB", "dram_pct_of_peak", "l1tex_lsu_wavefronts_pct", "sm_throughput_pct", "ipc",
            "pipe_alu_pct", "pipe_fp64_pct", "pipe_dmma_pct", "registers", "achieved_occupancy_pct"]
    with open(out + ".md", "w") as f:
        f.write(f"source: `{rep}` (ncu --set full --clock-control none)\n\n| # | kernel | " + " | ".join(cols) + " | top stalls |\n")
        f.write("|" + "---|" * (len(cols) + 3) + "\n")
        for i, d in enumerate(res):
            f.write(f"| {i} | `{d['kernel'][:48]}` | " + " | ".join(
                (f"{d[c]:.3g}" if isinstance(d.get(c), float) else str(d.get(c, ""))) for c in cols) +
                " | " + ", ".join(f"{a} {b}" for a, b in d["top_stalls"][:3]) + " |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
