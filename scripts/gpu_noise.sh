timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -q -x 2>&1 | tail -2
for d in 0 1; do
echo "DIRECT=$d"
MMAE_TC2_DIRECT=$d MMAE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_dump.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f gemm TF/s %.1f'%(d['ms_per_step'], d['roofline']['achieved']))"
python - <<'PY'
import re,collections
agg=collections.OrderedDict()
for l in open('gpurun_out/bench_dump.err'):
    m=re.match(r'\[mmae gemm\] (M=\d+ N=\d+ K=\d+ ta=\d tb=\d splits=\d+)\s+([\d.]+) ms\s+([\d.]+) TFLOP',l)
    if m: agg.setdefault(m.group(1),[]).append(float(m.group(2)))
print(' | '.join('%s:%s'%(k.split(' ta')[0].replace('M=65536 ',''), '/'.join('%.3f'%x for x in v[:2])) for k,v in agg.items()))
PY
done
