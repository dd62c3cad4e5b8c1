"""Condense an `ncu --set full --csv --page raw` export into one line per launch."""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], newline='')) if len(r) > 20]
hdr, units, data = rows[0], rows[1], rows[2:]


def col(*subs):
    if len(subs) == 1 and subs[0] in hdr:         # exact metric name first
        return hdr.index(subs[0])
    for i, h in enumerate(hdr):
        if all(s in h for s in subs):
            return i
    return None


C = {
    'time_ms': col('gpu__time_duration.sum'),
    'rdGB': col('dram__bytes_read.sum') if col('dram__bytes_read.sum') is not None else None,
    'wrGB': col('dram__bytes_write.sum'),
    'dram%': col('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    'regs': col('launch__registers_per_thread'),
    'warps%': col('sm__warps_active.avg.pct_of_peak_sustained_active'),
    'l1tex%': col('l1tex__throughput.avg.pct_of_peak_sustained_active'),
    'lts%': col('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
    'fp64%': col('sm__inst_executed_pipe_fp64', 'pct') if col('sm__inst_executed_pipe_fp64', 'pct') is not None
    else col('pipe_fp64', 'pct'),
    'dmma%': col('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'),
    'lsu%': col('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
    'ipc': col('sm__inst_executed.avg.per_cycle_active'),
    'stall_long': col('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
    'stall_short': col('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
    'stall_barrier': col('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
    'grid': col('launch__grid_size'),
    'block': col('launch__block_size'),
}
ki = col('Kernel Name')


def val(r, name):
    i = C[name]
    if i is None:
        return float('nan')
    try:
        v = float(r[i].replace(',', ''))
    except ValueError:
        return float('nan')
    u = units[i]
    if name == 'time_ms':
        v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1, 'msecond': 1, 's': 1e3, 'second': 1e3}.get(u, 1)
    if name in ('rdGB', 'wrGB'):
        v *= {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1}.get(u, 1e-9)
    return v


print('kernel | '+' | '.join(C.keys()))
for r in data:
    name = re.sub(r'\(.*', '', r[ki]).replace('shb::', '')[:90]
    print(name, '|', ' | '.join(f'{val(r, k):.3f}' for k in C))
