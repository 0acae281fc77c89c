import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0     # which kernel of the report (0-based)
vals = rows[2 + which]
print("# kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
keys = ["gpu__time_duration.sum","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","sm__warps_active.avg.pct_of_peak_sustained_active",
"smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","dram__bytes_read.sum","dram__bytes_write.sum",
"launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__grid_size","launch__block_size","sm__maximum_warps_per_active_cycle_pct",
"l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum","l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_hit.sum",
"lts__t_sectors_srcunit_tex_op_read.sum","lts__t_sectors_srcunit_tex_op_write.sum","smsp__warps_eligible.avg.per_cycle_active","smsp__warps_active.avg.per_cycle_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","sm__inst_executed_pipe_lsu.sum","launch__shared_mem_per_block_dynamic","launch__shared_mem_per_block_static","launch__occupancy_limit_warps"]
for i,h in enumerate(hdr):
    if h in keys or ("smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h) or "smsp__average_warp_latency_issue_stalled" in h:
        try:
            v=float(vals[i].replace(",",""))
            if v==0: continue
        except: pass
        print(f"{h:90s} {units[i]:12s} {vals[i]}")
