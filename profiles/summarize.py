"""Turns what a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py r1            # reads gpurun_out/{launches_r1.csv, prof_*_r1.ncu-rep, bench_r1_*.json}
    python profiles/summarize.py r1b k1        # launches_r1b.csv + gpurun_out/ncu_k1_r1.ncu-rep -> ncu_nearest_r1b.txt (+ stall table)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
SRC = os.path.join(ROOT, 'gpurun_out')

STALLS = ['math_pipe_throttle', 'not_selected', 'wait', 'short_scoreboard', 'dispatch_stall', 'long_scoreboard', 'barrier',
          'branch_resolving', 'mio_throttle', 'lg_throttle', 'no_instruction']

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__occupancy_limit_registers',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def launches(tag):
    path = os.path.join(SRC, 'launches_%s.csv' % tag)
    if not os.path.isfile(path):
        return
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    hdr = None
    for r in csv.reader(open(path)):
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        name = re.sub(r'^void ', '', r[4])
        name = re.sub(r'\(.*', '', name)
        name = re.sub(r'<.*', '', name)
        t = float(r[-1])
        a = agg[name]
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, 'launches_%s_summary.txt' % tag), 'w') as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1 --no-cpu\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms, not absolutes\n")
        f.write("%-28s %6s %12s %7s %10s %10s\n" % ('kernel', 'n', 'total_us', 'share', 'mean_us', 'max_us'))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-28s %6d %12.1f %6.1f%% %10.1f %10.1f\n" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0] / 1e3, v[2] / 1e3))


def full(tag, which, rep=None):
    rep = rep or os.path.join(SRC, 'prof_%s_%s.ncu-rep' % (which, tag))
    if not os.path.isfile(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, 'ncu_%s_%s.txt' % (which, tag)), 'w') as f:
        f.write("# ncu --set full --clock-control none --import-source on ; selected raw metrics per captured launch\n")
        for r in rows[2:]:
            f.write('----\n')
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("%-72s %s %s\n" % (k, r[i][:120], units[i]))
            for st in STALLS:          # warp-cycles stalled per issued instruction, by reason
                k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % st
                if k in hdr:
                    f.write("%-72s %s\n" % ('stall/issue ' + st, r[hdr.index(k)]))


def bench(tag):
    for fn in sorted(os.listdir(SRC)):
        if fn.startswith('bench_%s' % tag) and fn.endswith('.json'):
            lines = [l for l in open(os.path.join(SRC, fn)) if l.strip().startswith('{')]
            if lines:
                with open(os.path.join(OUT, fn), 'w') as f:
                    for l in lines:
                        f.write(json.dumps(json.loads(l), indent=1) + '\n')


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    launches(tag)
    if len(sys.argv) > 2 and sys.argv[2] == 'k1':
        full(tag, 'nearest', os.path.join(SRC, 'ncu_k1_r1.ncu-rep'))
    else:
        for which in ('nearest', 'steer'):
            full(tag, which)
        bench(tag)
    print(sorted(os.listdir(OUT)))
