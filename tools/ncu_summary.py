"""Summarise an ncu --page raw --csv export: python tools/ncu_summary.py raw.csv [filter-substrings...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size',
    'launch__block_size', 'launch__shared_mem_per_block', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma',
    'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_alu_cycles_active', 'sm__pipe_fma_cycles_active', 'dram__bytes_read.sum',
    'dram__bytes_write.sum', 'lts__t_bytes.sum', 'smsp__average_warp', 'smsp__average_warps_issue_stalled',
    'sm__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'l1tex__data_pipe_lsu_wavefronts.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_', 'smsp__thread_inst_executed_per_inst',
    'l1tex__lsu_writeback', 'l1tex__throughput', 'lts__throughput', 'sm__pipe_', 'sm__mio', 'sm__issue_active']
for vals in rows[2:]:
    print("==", vals[hdr.index('Kernel Name')][:70] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want):
            print(f"{h:95s} {v:>18s} {u}")
