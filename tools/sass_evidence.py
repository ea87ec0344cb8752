#!/usr/bin/env python
"""SASS mnemonic counts per source file of libssdk.so's objects (cuobjdump -sass on ssd_keras_b200/_lib/*.o) + the MMA issue
loop of conv_tcgen05_kernel.  Runs without a GPU.   python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
LIB = os.path.join(ROOT, 'ssd_keras_b200', '_lib')
PAT = re.compile(r'\b(UTCHMMA|UTMALDG\.\dD|LDTM\.?|UTCBAR|UBLKCP\.[SG]\.[SG]|SYNCS\.[A-Z0-9.]+|ELECT|CREDUX\.[A-Z0-9.]+|REDUX|F2FP\.BF16\.F32\.PACK_AB|'
                 r'UTCATOMSWS\.[A-Z_.]+|DFMA|MUFU\.RCP64H|STG\.E\.ENL2\.256)\b')
print('# SASS evidence: cuobjdump -sass of the objects linked into ssd_keras_b200/_lib/libssdk.so (sm_100a), mnemonic counts')
print('# UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk '
      '(S.G: global->shared, G.S: shared->global),')
print('# SYNCS.* = mbarrier ops, ELECT = elect.sync, CREDUX = redux.sync, F2FP.BF16.F32.PACK_AB = cvt.rn.bf16x2.f32, UTCATOMSWS = tcgen05.alloc/dealloc, '
      'STG.E.ENL2.256 = st.global.v8.b32')
loop = None
for name in ('conv', 'wgrad', 'encode', 'loss', 'decode'):
    obj = os.path.join(LIB, name + '.o')
    if not os.path.exists(obj):
        continue
    sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    cnt = collections.Counter(m.group(1) for m in PAT.finditer(sass))
    print('\n== %s.cu' % name)
    for k, v in cnt.most_common():
        print('%7d %s' % (v, k))
    if name == 'conv':
        lines = sass.split('\n')
        start = next((i for i, l in enumerate(lines) if 'conv_tcgen05_kernel' in l and 'Function' in l), 0)
        first = next((i for i in range(start, len(lines)) if 'UTCHMMA' in lines[i]), None)
        if first is not None:
            loop = [l for l in lines[first - 6:first + 60] if '/*' in l and not re.match(r'\s*/\* 0x', l)][:44]
if loop:
    print('\n== conv_tcgen05_kernel: the MMA issue loop (around the first UTCHMMA; with the fused weight operand a k-step is TWO UTCHMMA per m-tile)')
    print('\n'.join(loop))
