"""Reads an `ncu --set full` report and prints/saves the metrics the design cites:
warp execution efficiency, issue utilisation, L1/L2 hit rates, DRAM bytes, FP32 instruction
counts, stall reasons.   python tools/summarize_ncu.py <report.ncu-rep> [out.json]"""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__occupancy_limit_registers": "occ_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__occupancy_limit_warps": "occ_limit_warps_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_inst (of 32)",
    "smsp__thread_inst_executed_per_inst_executed.pct": "warp_execution_efficiency_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "smsp__issue_active.avg.pct": "issue_slot_utilisation_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum": "thread_fadd",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum": "thread_fmul",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": "thread_ffma",
    "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum": "thread_fp32",
    "sm__sass_thread_inst_executed_op_fp32_pred_on.sum": "thread_fp32_sm",
    "smsp__average_warp_latency_per_inst_issued.ratio": "warp_cycles_per_issued_inst",
}
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k, name in WANT.items():
            if k in hdr:
                u = units[hdr.index(k)]
                d[name] = r[hdr.index(k)] + (f" {u}" if u else "")
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len(STALL):-len("_per_issue_active.ratio")]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        d["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        res.append(d)
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
