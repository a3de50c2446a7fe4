"""Summarise an .ncu-rep: key metrics of each captured launch + per-opcode / per-region instruction and stall-sample histogram."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__waves_per_multiprocessor',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_active.avg',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d['Kernel Name'][:60], d.get('Grid Size'), d.get('Block Size'))
    for k in hdr:
        if k in keys or 'stalled' in k and 'per_issue_active' in k or 'pct_of_peak_sustained_elapsed' in k and ('lts__t' in k or 'l1tex__' in k or 'dram' in k):
            print('   ', k, d[k], rows[1][hdr.index(k)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for b in blocks[1:2]:
    rr = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    h = rr[1]; I = {x: i for i, x in enumerate(h)}
    data = [r for r in rr[2:] if len(r) == len(h)]
    tot = sum(int(r[I['Instructions Executed']]) for r in data); samp = sum(int(r[I['# Samples']]) for r in data)
    print('total warp instr', tot, 'samples', samp, 'sass lines', len(data))
    c = Counter(); s = Counter()
    for r in data:
        op = [o for o in r[I['Source']].split() if not o.startswith('@')][0].split('.')[0]
        c[op] += int(r[I['Instructions Executed']]); s[op] += int(r[I['# Samples']])
    for k, v in s.most_common(14):
        print(f'  {k:10s} instr {c[k]:10d} ({100*c[k]/tot:4.1f}%)  samples {v} ({100*v/samp:4.1f}%)')
    print('top sampled SASS lines:')
    for r in sorted(data, key=lambda r: -int(r[I['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        print('  ', data.index(r), r[I['# Samples']], r[I['Instructions Executed']], r[I['Source']])
