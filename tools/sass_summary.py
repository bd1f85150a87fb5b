"""Opcode evidence per kernel from the built library (no GPU needed):

    python tools/sass_summary.py [path/to/libgeniconet_b200.so] > profiles/r02_sass_summary.txt

Counts the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, UTMALDG/UTMASTG =
tensor-map TMA, LDGSTS = cp.async, SYNCS = mbarrier, HMMA = legacy mma.sync.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'geniconet_b200', 'libgeniconet_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'SYNCS', 'HMMA', 'ATOM', 'RED', 'LDG', 'STG']
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        per[cur]['_total'] += 1
        for k in KEYS:
            if op == k or op.startswith(k + '.') or op.startswith(k):
                per[cur][k] += 1
                break
print('# cuobjdump -sass %s : instruction counts per kernel (static), sm_100a' % os.path.relpath(lib, ROOT))
print('# %-86s %7s %s' % ('kernel', 'instrs', ' '.join('%7s' % k for k in KEYS)))
tot = collections.Counter()
for fn, c in per.items():
    name = re.sub(r'\(.*', '', demangle(fn)).replace('void ', '')
    print('%-88s %7d %s' % (name[:88], c['_total'], ' '.join('%7d' % c[k] for k in KEYS)))
    tot.update(c)
print('%-88s %7d %s' % ('TOTAL (%d kernels)' % len(per), tot['_total'], ' '.join('%7d' % tot[k] for k in KEYS)))
