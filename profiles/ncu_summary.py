#!/usr/bin/env python
"""Text summaries of an `ncu --set full --import-source on` report of the step kernel (profiles/ncu.sh):

    python profiles/ncu_summary.py gpurun_out/prof_65536.ncu-rep profiles/r01/ncu_step_kernel_f32_65536envs

writes <prefix>.txt (selected raw metrics, one column per profiled launch) and <prefix>_source_by_{inst,stall}.txt
(per source line: warp instructions executed and stall samples, via -lineinfo)."""
import csv
import os
import subprocess
import sys

METRICS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread launch__waves_per_multiprocessor
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__cycles_active.avg sm__cycles_elapsed.max
lts__t_sector_hit_rate.pct sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__occupancy_limit_warps""".split()


def ncu(*args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(rep, prefix):
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(prefix + ".txt", "w") as f:
        f.write(f"# {rep}: kernel = {data[0][hdr.index('Kernel Name')]}; one column per profiled launch\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"{m:78s} {units[i]:16s} " + "  ".join(r[i] for r in data) + "\n")
    acc, fname, launches = {}, "", set()
    for r in csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass").splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
        elif r[0] == "Function Name":
            launches.add(len(launches) if fname == "gpd_kernels.cuh" else -1)
        elif r[0] == "Line No":
            col = {n: k for k, n in reversed(list(enumerate(r)))}
            ci, cs = col["Instructions Executed"], col["Warp Stall Sampling (All Samples)"]
        elif r[0].isdigit():
            try:
                k = (fname, int(r[0]))
                old = acc.get(k, (r[1].strip(), 0, 0))
                acc[k] = (old[0], old[1] + int(r[ci] or 0), old[2] + int(r[cs] or 0))
            except ValueError:
                pass
    lines = [(k[0], k[1], v[0], v[1], v[2]) for k, v in acc.items()]      # summed over the profiled launches
    ti, ts = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
    for key, idx in (("inst", 3), ("stall", 4)):
        with open(f"{prefix}_source_by_{key}.txt", "w") as f:
            f.write(f"sums over the profiled launches: warp instructions {ti}, stall samples {ts}\n")
            for l in sorted(lines, key=lambda l: -l[idx])[:60]:
                f.write(f"{l[0]:16s}:{l[1]:4d} inst={l[3]:8d} ({100 * l[3] / ti:4.1f}%) samp={l[4]:5d} ({100 * l[4] / ts:4.1f}%) | {l[2][:110]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
