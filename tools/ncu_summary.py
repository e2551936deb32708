"""Curated summary of an .ncu-rep (reads with `ncu -i ... --page raw --csv`).  usage: ncu_summary.py rep [out.txt]"""
import csv, io, subprocess, sys
KEYS = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
 "launch__occupancy_limit_warps", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
 "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
 "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
 "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active",
 "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
 "l1tex__t_requests_pipe_tex_mem_texture.sum", "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
 "l1tex__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_tex_mem_texture_op_tex_hit_rate.pct",
 "l1tex__f_tex2sm_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
 "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
 "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
 "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
]
def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        lines.append(f"== {r[hdr.index('Kernel Name')][:70]}  (id {r[0]})")
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k or h.endswith("." + k):
                    lines.append(f"  {k} [{units[i]}] = {r[i]}")
                    break
    out = "\n".join(lines)
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(f"# ncu --set full --clock-control none summary of {rep}\n" + out + "\n")
main()
