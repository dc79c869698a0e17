"""Turns the ncu CSVs brought back in gpurun_out/ into the text summaries committed under profiles/.
usage: summarize_ncu.py <launches.csv> <raw.csv> <out.txt> <command string> <kernel regex for traffic> [traffic.json key]"""
import collections
import csv
import json
import os
import re
import sys

launches, rawcsv, out, cmd, kre = sys.argv[1:6]
tkey = sys.argv[6] if len(sys.argv) > 6 else None
rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg, total = collections.OrderedDict(), 0.0
harness = collections.OrderedDict()          # bench.py's own kernels: cuBLAS tf32 peak measurement, torch.rand data generation
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[kn])
    name = re.sub(r'^void ', '', name)
    t = float(r[mv].replace(',', ''))
    if not name.startswith('mmae::'):
        harness.setdefault(name, [0, 0.0]); harness[name][0] += 1; harness[name][1] += t
        continue
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += t
    total += t
with open(out, 'w') as f:
    f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n')
    f.write('# command: %s\n' % cmd)
    f.write('total %.1f us over %d launches\n' % (total / 1e3, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-64s n=%3d  %10.1f us  %5.1f%%\n' % (k[:64], v[0], v[1] / 1e3, 100 * v[1] / total))
    f.write('# bench harness kernels in the same capture (not part of a step): ' +
            '; '.join('%s n=%d %.1f us' % (k[:40], v[0], v[1] / 1e3) for k, v in harness.items()) + '\n')
    f.write('\n# ncu --set full captures (per launch)\n')
    rr = list(csv.reader(open(rawcsv)))
    h, units = rr[0], rr[1]
    want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__inst_executed.sum',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
    idx = [(w, h.index(w)) for w in want if w in h]
    ikn, ird, iwr, it = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('gpu__time_duration.sum')
    tb, tn = 0.0, 0

    def to_bytes(v, u):
        return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]
    for r in rr[2:]:
        if len(r) < len(h):
            continue
        f.write('---\n')
        for w, i in idx:
            f.write('  %-75s %s %s\n' % (w, r[i], units[i]))
        if re.search(kre, r[ikn]):
            tb += to_bytes(r[ird], units[ird]) + to_bytes(r[iwr], units[iwr])
            tn += 1
    if tn:
        f.write('\n# dram traffic of %s: %.4g bytes over %d launches = %.4g bytes per launch\n' % (kre, tb, tn, tb / tn))
        if tkey:
            tp = os.path.join(os.path.dirname(out), 'traffic.json')
            d = json.load(open(tp)) if os.path.exists(tp) else {}
            d[tkey] = tb / tn
            d[tkey + '_source'] = '%s: dram__bytes_read.sum + dram__bytes_write.sum of %s, mean over %d launches' % (os.path.basename(out), kre, tn)
            json.dump(d, open(tp, 'w'), indent=1)
sys.stdout.write(open(out).read()[:3000] + '\n')
