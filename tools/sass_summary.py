#!/usr/bin/env python
"""Per-kernel counts of the Blackwell instructions in libstyle_b200.so (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store (cp.async.bulk.tensor),
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  Writes a markdown table to stdout."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else 'style_transfer_b200/libstyle_b200.so'
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
OPS = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA']
counts, cur, total = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m:
        op = m.group(1)
        counts[cur]['_n'] += 1
        for o in OPS:
            if op == o or op.startswith(o + '.'):
                counts[cur][o] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print('| kernel | SASS instr | ' + ' | '.join(OPS) + ' |')
print('|---|---|' + '---|' * len(OPS))
for (name, c), dn in zip(counts.items(), demangle):
    short = re.sub(r'\(.*', '', dn.replace('(anonymous namespace)::', '')).replace('void ', '').replace('st::', '')
    if not any(c[o] for o in OPS[:6]) and '--all' not in sys.argv:
        continue
    print('| `%s` | %d | ' % (short[:90], c['_n']) + ' | '.join(str(c[o]) for o in OPS) + ' |')
