#!/usr/bin/env python3
"""Print the key metrics of an `ncu --page raw --csv` export (one block per kernel launch)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(r'gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__warps_active.avg.pct|registers_per_thread$|'
                 r'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)|sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct|'
                 r'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct|smsp__issue_active.avg.pct|'
                 r'smsp__average_warps_issue_stalled_.*_per_issue_active|occupancy_limit_(registers|shared_mem)|bank_conflicts.*shared.sum$|'
                 r'lts__t_bytes.sum$|lts__t_sectors_srcunit_tex_op_read.sum$|smsp__inst_executed.sum$|sm__inst_executed_pipe_(fp64|lsu|alu|fma|fmaheavy|xu|uniform|cbu|adu).sum$|'
                 r'shared_mem_per_block$|local_(load|store)|smsp__thread_inst_executed_per_inst_executed.ratio|lts__t_sector_hit_rate.pct')
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')][:70])
    for i, h in enumerate(hdr):
        if pat.search(h):
            v = r[i]
            try:
                if float(v.replace(',', '')) == 0:
                    continue
            except ValueError:
                pass
            print(f'  {h} [{units[i]}] = {v}')
