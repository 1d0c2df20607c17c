"""Summarise `ncu --set full` reports into small JSON files under profiles/.
  python profiles/ncu_summarize.py gpurun_out/r2_matvec.ncu-rep profiles/r2_ncu_matvec_summary.json
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__occupancy_limit_registers": "occ_limit_regs_ctas",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_ctas",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_subpipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    v = r[i]
                d[KEYS[h]] = v
                if KEYS[h] in ("duration", "dram_bytes_read", "dram_bytes_write"):
                    d[KEYS[h] + "_unit"] = units[i]
        res.append(d)
    json.dump({"source": rep, "how": "ncu --set full --clock-control none (cold, serialised launches)", "kernels": res},
              open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:70], {k: v for k, v in d.items() if k in ("duration", "dram_throughput_pct", "warps_active_pct",
                                                                      "registers", "fp64_pipe_active_pct", "dram_bytes_read")})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
