"""Counts the Blackwell-specific SASS mnemonics per kernel of libdfb200.so (cuobjdump -sass on stdin or run directly):
UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA loads, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,
DMMA = fp64 mma.sync.  No GPU needed.  Usage: python tools/sass_summary.py > profiles/r01_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'dragonfly_b200', 'libdfb200.so')],
                      capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTC[A-Z]*MMA[\.\w]*|UTMALDG[\.\w]*|UTMASTG[\.\w]*|UTCBAR[\.\w]*|UTCCP[\.\w]*|LDTM[\.\w]*|STTM[\.\w]*|'
                 r'SYNCS[\.\w]*|DMMA[\.\w]*|ELECT|UCGABAR_\w+|UTCATOMSWS[\.\w]*|HMMA[\.\w]*|IMMA[\.\w]*)')
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    continue
  if cur is not None:
    for t in pat.findall(line):
      counts[cur][t] += 1
print('# cuobjdump -sass dragonfly_b200/libdfb200.so (sm_100a), Blackwell-path mnemonics per kernel')
print('# UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA),')
print('# UTCBAR = tcgen05.commit, SYNCS = mbarrier, DMMA = fp64 mma.sync (no f64 kind exists in tcgen05)')
for k, c in counts.items():
  if any(x.startswith(('UTC', 'UTMA', 'DMMA', 'LDTM')) for x in c):
    name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(.*', '', name)
    print(name)
    print('    ' + ', '.join('%s x%d' % (a, b) for a, b in sorted(c.items())))
