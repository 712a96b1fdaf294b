#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key counters + opcode mix + stall reasons.
    python tools/ncu_summary.py gpurun_out/r01_tile.ncu-rep > profiles/r01_tile_summary.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__cycles_active.avg']
for r in rows[2:]:
    print('== kernel:', r[hdr.index('Kernel Name')], ' grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for h, u, v in zip(hdr, units, r):
        if h in WANT:
            print(f'  {h:70s} {v:>18s} {u}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[start]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ci['Instructions Executed']]) for r in data)
ops = collections.Counter()
for r in data:
    toks = r[ci['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    ops[op.split('.')[0]] += int(r[ci['Instructions Executed']])
print(f'== opcode mix (warp-level instructions executed, total {tot})')
for op, c in ops.most_common(18):
    print(f'  {op:10s} {c:>13d} {100 * c / tot:5.1f}%')
st = collections.Counter()
for r in data:
    for k in ci:
        if k.startswith('stall_') and 'Not Issued' not in k:
            try:
                st[k] += int(r[ci[k]])
            except ValueError:
                pass
tots = sum(st.values())
print('== warp stall samples (all)')
for k, c in st.most_common(10):
    print(f'  {k:28s} {c:>9d} {100 * c / tots:5.1f}%')
