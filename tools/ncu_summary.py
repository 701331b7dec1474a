"""Print a compact summary of an .ncu-rep (raw page): time, traffic, pipes, stalls."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print('kernel:', d.get('Kernel Name'), 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor',
            'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
            'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
            'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
    for k in keys:
        if k in d:
            print(f'  {k:75s} {d[k]:>16s} {u[k]}')
    stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                     if k.startswith('smsp__average_warp_latency_issue_stalled') or k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio')
                     if v not in ('', 'n/a')), reverse=True)
    for val, k in stalls[:9]:
        print(f'  STALL {k:70s} {val:10.3f}')
