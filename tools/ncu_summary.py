"""Summarise an ncu raw CSV export (ncu -i X.ncu-rep --page raw --csv) for profiles/."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
    r'gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__warps_active.avg.pct|registers_per_thread|'
    r'sm__throughput.avg.pct|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate.pct|smsp__inst_executed.sum$|'
    r'issue_active.avg.pct|l1tex__t_(sectors|requests)_pipe_lsu_mem_global_op_ld.sum$|thread_inst_executed_per_inst|'
    r'occupancy|stalled_.*_per_warp_active|pipe_fp64|inst_executed_pipe_(fma|alu|lsu|xu|fp64)|'
    r'l1tex__data_pipe_lsu_wavefronts.sum$|l1tex__throughput|lts__throughput|smsp__cycles_active.avg$|sm__cycles_elapsed.max|dram__throughput')
for r in rows[2:]:
    print('--- kernel', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print('  %-95s %14s %s' % (h, r[i], units[i]))
