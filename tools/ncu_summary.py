"""Text summary of an ncu --set full report (one block per captured launch): duration, pipe / issue utilisation,
occupancy limiters, DRAM bytes, stall-reason shares.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<what>.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (per-launch times under the profiler are not bench values)")
    for r in rows[2:]:
        print("----")
        print("kernel:", r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w} = {r[i]} {units[i]}")
        st = [(float(r[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
        tot = sum(s for s, _ in st) or 1.0
        print("  stall samples: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * s / tot:.1f}%"
                                               for s, h in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
