"""Condense an .ncu-rep (one kernel) into the JSON summary kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_decoder_tc.ncu-rep profiles/decoder_tc_ncu_summary.json
"""
import csv, io, json, subprocess, sys

KEEP = [
    "Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main(rep, out, extra=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {n: {"unit": u, "value": v} for n, u, v in zip(names, units, vals) if n in KEEP}

    def num(k):
        return float(m[k]["value"].replace(",", ""))

    def to_bytes(k):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m[k]["unit"]]
        return num(k) * scale

    summary = {"kernel": m["Kernel Name"]["value"],
               "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
               "metrics": m}
    if extra:
        summary.update(json.loads(extra))
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct"):
        print(k, m[k]["value"], m[k]["unit"])
    print("dram bytes", summary["dram_bytes_per_launch"])


if __name__ == "__main__":
    main(*sys.argv[1:])
