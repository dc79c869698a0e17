"""Turns gpurun_out/launches.csv (+ an .ncu-rep) into the text summaries committed under profiles/."""
import collections
import csv
import re
import subprocess
import sys

launches, rep, out = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg, total = collections.OrderedDict(), 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'<.*', '', re.sub(r'\(.*', '', r[kn]))
    t = float(r[mv].replace(',', ''))
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += t
    total += t
with open(out, 'w') as f:
    f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n')
    f.write('# command: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e   (3 steps profiled)\n')
    f.write('total %.1f us over %d launches\n' % (total / 1e3, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-48s n=%3d  %10.1f us  %5.1f%%\n' % (k[:48], v[0], v[1] / 1e3, 100 * v[1] / total))
    f.write('\n# ncu --set full captures of gemm_tc_kernel (per launch)\n')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    want = ['Kernel Name', 'launch__grid_size', 'launch__registers_per_thread', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
            'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
    idx = [(w, h.index(w)) for w in want if w in h]
    units = rr[1]
    for r in rr[2:]:
        f.write('---\n')
        for w, i in idx:
            f.write('  %-75s %s %s\n' % (w, r[i], units[i]))
print(open(out).read())
