"""Summarise an `ncu --set full` report (one kernel) into the short text kept under profiles/:
   python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep > profiles/ncu_conv_rNN.txt"""
import csv, io, subprocess, sys

METRICS = ["launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "gpu__time_duration.sum",
           "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
           "sm__inst_issued.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
           "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
           "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tma_ld.sum",
           "smsp__sass_inst_executed_op_tmem_ldt.sum", "sass__inst_executed_global_loads",
           "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
           "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path, every=False):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for vals in rows[2:]:
        name = vals[ix["Kernel Name"]]
        if name in seen and not every:            # one launch per distinct kernel (the first captured)
            continue
        seen.add(name)
        print("kernel:", name)
        for m in METRICS:
            if m in ix:
                print(f"  {m} [{units[ix[m]]}] = {vals[ix[m]]}")
        stalls = []
        for h, i in ix.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(vals[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top warp-stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main(sys.argv[1], every="--all" in sys.argv)
