"""Summarise an ncu report into the small CSV kept under profiles/ (one row per metric, one column per captured launch).

  python tools/ncu_extract.py gpurun_out/x.ncu-rep > profiles/r02_ncu_prof_contract.csv
"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum",
        "nvlrx__bytes.sum", "nvltx__bytes.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(vals))])
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [v[i] for v in vals])


if __name__ == "__main__":
    main()
