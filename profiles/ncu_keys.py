"""Print the headline metrics of an `ncu --page raw --csv` dump (one block per captured launch)."""
import csv
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp_latency_per_inst_issued.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('----')
    for k in KEYS:
        if k in idx:
            print(f"{k} = {r[idx[k]]} {units[idx[k]]}")
    top = sorted(((float(r[idx[h]].replace(',', '')) if r[idx[h]] else 0.0, h) for h in stall), reverse=True)[:6]
    for v, h in top:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {v:.2f}")
