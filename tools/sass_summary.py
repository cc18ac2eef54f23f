#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass), the evidence that the
bulk-copy (TMA, UBLKCP) + mbarrier (SYNCS) path and the cp.async (LDGSTS) cell gathers are what ships.

    python tools/sass_summary.py [lib.so] [name filter regex] > profiles/rNN_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'ndpolator_b200/libndpb200.so'
flt = re.compile(sys.argv[2] if len(sys.argv) > 2 else r'k_fused|k_slow|k_nsm|k_build')
MNEMONICS = ['UBLKCP', 'UTMALDG', 'SYNCS', 'LDGSTS', 'LDG', 'STG', 'LDS', 'DADD', 'DMUL', 'DFMA', 'MUFU.RCP64H', 'SHFL', 'BRA']

demangle = {}
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        counts[cur]['_total'] += 1
        for name in MNEMONICS:
            if op == name or op.startswith(name + '.'):
                counts[cur][name] += 1
names = list(counts)
dm = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines() if names else []
res = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
regs = {}
fn = None
for line in res.splitlines():
    m = re.search(r'Function (\S+):', line)
    if m:
        fn = m.group(1)
    m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
    if m and fn:
        regs[fn] = (int(m.group(1)), int(m.group(2)))
print(f'# {lib}: SASS mnemonics per kernel (cuobjdump -sass; sm_100a)')
print('# ' + ' '.join(f'{k:>7}' for k in ['instrs'] + MNEMONICS) + '  regs  kernel')
tot = collections.Counter()
for mangled, pretty in zip(names, dm):
    for k, v in counts[mangled].items():
        tot[k] += v
    if not flt.search(pretty):
        continue
    c = counts[mangled]
    r = regs.get(mangled, ('?', '?'))[0]
    print('  ' + ' '.join(f'{c[k]:7d}' for k in ['_total'] + MNEMONICS) + f'  {r!s:>4}  {pretty[:110]}')
print('# whole library: ' + ', '.join(f'{k} {tot[k]}' for k in MNEMONICS))
