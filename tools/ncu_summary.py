#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote.  usage: ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum', 'smsp__sass_inst_executed_op_global_ld.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio']
out = []
for r in rows[2:]:
  d = dict(zip(hdr, r))
  for k in keep:
    if k in d: out.append(f'{k:75s} {d[k]:>20s} {units[hdr.index(k)]}')
  st = sorted(((float(v), h) for h, v in d.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and v), reverse=True)
  out.append('stall cycles per issued instruction (top): ' + ', '.join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st[:7]))
  out.append('')
txt = '\n'.join(out)
print(txt)
if len(sys.argv) > 2: open(sys.argv[2], 'w').write(txt + '\n')
