"""Condense an `ncu --set full` report into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_lifts_final.ncu-rep \
        profiles/r01_lifts_ncu_summary.json --evals 8192 --p 100

The report comes from `ncu --set full --clock-control none --import-source on
-k regex:lifts_mma -s 1 -c 1 python tools/prof_lifts.py` on a B200; bench.py
reads `dram_bytes_per_launch` and `permutation_evaluations_per_launch` from the
JSON to fill `roofline.traffic`.
"""
import argparse
import csv
import io
import json
import subprocess

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
              "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

PICK = {
    "gpu_time_ms": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "grid_size": "launch__grid_size",
    "block_size": "launch__block_size",
    "dynamic_smem_kb": "launch__shared_mem_per_block_dynamic",
    "ctas_per_sm_limit_smem": "launch__occupancy_limit_shared_mem",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "pipe_fp64_cycles_active_pct":
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dmma_pipe_active_pct":
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "inst_executed": "smsp__inst_executed.sum",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_wavefronts_pct_of_peak":
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "stall_barrier_per_issue":
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_short_scoreboard_per_issue":
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_wait_per_issue":
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_math_pipe_throttle_per_issue":
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--evals", type=int, required=True,
                    help="permutation evaluations in the captured launch")
    ap.add_argument("--p", type=int, required=True)
    ap.add_argument("--command", default="python tools/prof_lifts.py")
    args = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"],
                         check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    out = {"kernel": vals[col["Kernel Name"]], "command": args.command,
           "permutation_evaluations_per_launch": args.evals}
    for key, metric in PICK.items():
        if metric not in col:
            continue
        i = col[metric]
        v = float(vals[i].replace(",", ""))
        if key in ("gpu_time_ms", "dram_bytes_read", "dram_bytes_write"):
            v *= UNIT_SCALE.get(units[i], 1.0)
        out[key] = v
    out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    flops = 7.0 / 3.0 * args.p ** 3 * args.evals
    out["algorithmic_flops_per_launch"] = flops
    out["tflops_under_ncu"] = flops / (out["gpu_time_ms"] * 1e-3) / 1e12
    out["note"] = ("ncu --set full --clock-control none; timing under the profiler "
                   "is not a bench number")
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
