"""Prints the headline fields of a bench.py JSON line (file argument)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
r = d['roofline']
print(d['config']['workload'], 'n_gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'launches', d['gpu_launches'],
      'frac %.3f' % r['frac'], 'of tf32 %s' % r.get('frac_of_tf32_peak'), 'gemm share %s' % r.get('gemm_share_of_step'), d['clocks'])
if d.get('e2e'): print('  e2e %.4g samples/s, %.3f ms/step' % (d['e2e']['value'], d['e2e']['ms_per_step']))
if d.get('api'): print('  api', d['api'])
if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline'])
for k, v in (d.get('others') or {}).items():
    if 'error' in v: print('  other', k, 'ERROR', v['error']); continue
    print('  other %-5s ms/step %.4f value %.4g launches %d frac %.3f hbm %s' % (k, v['ms_per_step'], v['value'], v['gpu_launches'], v['roofline']['frac'],
          (v['roofline'].get('hbm_bound') or {}).get('frac')))
