#!/usr/bin/env python
"""Summarise `ncu --set full` reports into one JSON (committed under profiles/).

    python scripts/ncu_summary.py gpurun_out/prof_r2_*.ncu-rep > profiles/ncu_r2_summary.json

Reads every report with `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps, per profiled
launch, the metrics the roofline discussion uses: duration, DRAM bytes, achieved DRAM throughput,
tensor-pipe activity, issue activity, shared-memory bank conflicts, registers, occupancy, and the
top warp-stall reasons (pc-sampling)."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum": "smem_store_bank_conflicts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum": "smem_load_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum": "smem_store_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum": "smem_load_wavefronts",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


def main():
    out = {}
    for rep in sys.argv[1:]:
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        if r.returncode != 0:
            out[rep] = {"error": r.stderr[-300:]}
            continue
        rows = list(csv.reader(io.StringIO(r.stdout)))
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            rec = dict(zip(hdr, row))
            name = rec.get("Kernel Name", "?")
            entry = {"kernel": name[:90]}
            stalls = {}
            for h, u in zip(hdr, units):
                v = rec.get(h, "")
                if h in KEEP and v != "":
                    entry[KEEP[h]] = num(v)
                    if u:
                        entry[KEEP[h] + "_unit"] = u
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    if v not in ("", "n/a"):
                        stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = num(v)
                if h.startswith("smsp__average_warp_latency_issue_stalled_") and v not in ("", "n/a"):
                    stalls[h[len("smsp__average_warp_latency_issue_stalled_"):].replace(".ratio", "")] = num(v)
            top = sorted(((v, k) for k, v in stalls.items() if isinstance(v, float)), reverse=True)[:5]
            entry["top_stalls_warps_per_issue"] = {k: round(v, 2) for v, k in top}
            out.setdefault(rep.split("/")[-1], []).append(entry)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
