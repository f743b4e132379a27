#!/usr/bin/env python
"""Summaries for profiles/: per-kernel table of an ncu launch list, and the key metrics of an `ncu --set full` report.

    python tools/summarize_ncu.py launches gpurun_out/x_launches.csv > profiles/rNN_launches_x.txt
    python tools/summarize_ncu.py full gpurun_out/x.ncu-rep > profiles/rNN_ncu_full_x.txt
    python tools/summarize_ncu.py traffic gpurun_out/a.ncu-rep[:suffix] ... > profiles/rNN_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warp_latency_issue_stalled_barrier.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[h]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot, seq = {}, []
    for r in rows[h + 2:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        seq.append((r[ik], v))
        t = tot.setdefault(r[ik], [0, 0.0])
        t[0] += 1
        t[1] += v
    total = sum(t for _, t in tot.values())
    print(f'# {path}: {len(seq)} launches, {total / 1e3:.1f} us in total (ncu: serialised, cold caches: compare SHARES)')
    print(f'{"kernel":64s} {"launches":>8s} {"avg us":>9s} {"share":>7s}')
    for nm, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f'{nm[:64]:64s} {c:8d} {t / c / 1e3:9.2f} {100 * t / total:6.1f}%')
    print('\n# launch order (first 40): ' + ', '.join(f'{n.split("(")[0][:28]}={v / 1e3:.1f}' for n, v in seq[:40]))


def full(rep):
    hdr, units, rows = raw(rep)
    ik = hdr.index('Kernel Name')
    seen = {}
    for r in rows:
        seen.setdefault(r[ik], r)           # first captured launch of each kernel
    print(f'# {rep}: ncu --set full --clock-control none; first captured launch of each kernel')
    for nm, r in seen.items():
        print(f'\n== {nm}')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:86s} {r[i]:>16s} {units[i]}')


def traffic(specs):
    out = {}
    for spec in specs:
        rep, _, suffix = spec.partition(':')
        hdr, units, rows = raw(rep)
        ik = hdr.index('Kernel Name')
        for r in rows:
            key = r[ik] + (f'_{suffix}' if suffix else '')
            if key in out:
                continue

            def val(name):
                i = hdr.index(name)
                scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
                return float(r[i].replace(',', '')) * scale
            out[key] = {'dram_bytes_read': val('dram__bytes_read.sum'), 'dram_bytes_write': val('dram__bytes_write.sum'),
                        'duration_us': float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')) *
                        {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[hdr.index('gpu__time_duration.sum')]],
                        'source': rep.split('/')[-1]}
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == '__main__':
    {'launches': lambda a: launches(a[0]), 'full': lambda a: full(a[0]), 'traffic': traffic}[sys.argv[1]](sys.argv[2:])
