mkdir -p gpurun_out
MMAE_FUSE_NOISE=1 MMAE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_dump.log 2> gpurun_out/bench_dump.err; echo "dump rc=$?"
python - <<'PY'
import json,re,collections
d=json.loads(open('gpurun_out/bench_dump.log').read().strip().splitlines()[-1])
print('ms/step %.3f gemm TF/s %.1f share %.3f'%(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step']))
agg=collections.OrderedDict()
for l in open('gpurun_out/bench_dump.err'):
    m=re.match(r'\[mmae gemm\] (M=\d+ N=\d+ K=\d+ ta=\d tb=\d splits=\d+)\s+([\d.]+) ms\s+([\d.]+) TFLOP',l)
    if m: agg.setdefault(m.group(1),[]).append((float(m.group(2)),float(m.group(3))))
for k,v in agg.items(): print(k, 'n=%d'%len(v), 'ms', ' '.join('%.3f'%x[0] for x in v[:6]))
PY
