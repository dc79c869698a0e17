"""SASS opcode histogram per kernel of libmmae_b200.so: the mnemonics that prove tcgen05 / TMEM / TMA code generation
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops).  Usage: python scripts/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'multimodalautoencoder_b200', 'libmmae_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
WATCH = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTCBAR', 'SYNCS', 'UTCATOMSWS', 'REDUX', 'SHFL', 'MUFU',
         'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'RED', 'ATOM']
kern = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        op = m.group(1)
        hist[kern]['_total'] += 1
        for w in WATCH:
            if op.startswith(w):
                hist[kern][w] += 1
                break
demangle = subprocess.run(['c++filt'], input='\n'.join(hist), capture_output=True, text=True).stdout.splitlines()
print('# SASS opcode counts per kernel of libmmae_b200.so (cuobjdump -sass), sm_100a')
print('# %-86s %7s  %s' % ('kernel', 'instrs', 'watched opcodes'))
tot = collections.Counter()
for (k, c), name in zip(hist.items(), demangle):
    short = re.sub(r'\(.*', '', name)[:86]
    items = ' '.join('%s=%d' % (w, c[w]) for w in WATCH if c[w])
    print('%-88s %7d  %s' % (short, c['_total'], items))
    tot.update(c)
print('# library totals: ' + ' '.join('%s=%d' % (w, tot[w]) for w in WATCH if tot[w]))
