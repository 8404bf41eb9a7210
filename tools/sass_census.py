#!/usr/bin/env python3
"""SASS instruction census of liborb_b200.so per kernel (cuobjdump -sass): total instructions | selected mnemonics."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else 'orb_slam3_modified_b200/liborb_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', out)), capture_output=True, text=True).stdout.split('\n')
want = ['UTMALDG', 'UBLKCP', 'SYNCS', 'UCGABAR', 'LDG.E.ENL2.256', 'DFMA', 'DADD', 'DMUL', 'SHFL', 'REDUX', 'VIADD', 'VIMNMX', 'VIMNMX3', 'LDS', 'STS', 'LDL', 'STL', 'BAR.SYNC', 'ATOMS', 'POPC',
        'STG.E.ENL2.256', 'LDG.E.128', 'LDG.E.64', 'MUFU.RCP64H', 'FFMA', 'HMMA', 'UTCHMMA']
print('SASS instruction census of liborb_b200.so (cuobjdump -sass, sm_100a), per kernel: total instructions | selected mnemonics')
print('(UTMALDG = TMA tensor load, SYNCS = mbarrier, UCGABAR = cluster barrier, REDUX = warp reduce, LDG.E.ENL2.256 = 256-bit global load,')
print(' DFMA = FP64 FMA, VIADD/VIMNMX = packed-integer SIMD, LDL/STL = local-memory (spill) traffic)\n')
res = []
for i, blk in enumerate(out.split('Function : ')[1:]):
    ins = re.findall(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', blk, re.M)
    c = collections.Counter()
    for m in ins:
        for w in want:
            if m == w or m.startswith(w + '.') or m.startswith(w + '_') or (w == 'LDG.E.ENL2.256' and m.startswith('LDG') and 'ENL2.256' in m) or (w == 'STG.E.ENL2.256' and m.startswith('STG') and '256' in m):
                c[w] += 1
    name = names[i].split('(')[0]
    res.append((len(ins), name, c))
for n, name, c in sorted(res, reverse=True):
    print('%-46s %6d instr | %s' % (name[-46:], n, ', '.join('%s %d' % (k, v) for k, v in c.most_common())))
