"""Prints the metrics profiles/ cites from an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "sm__sass_inst_executed_op_shared_st.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"== {path}")
    for row in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:80s} {row[i]:>20s} {units[i]}")
        print("  --")
