"""Opcode census of the shipped library (cuobjdump -sass): per kernel, the counts of the SASS mnemonics that prove which
hardware paths it uses -- UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), UTMAREDG (cp.reduce.async.bulk.tensor), UBLKCP (cp.async.bulk),
UBLKRED (cp.reduce.async.bulk), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 (packed fp32), REDG / RED (global reductions).

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt        (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'prob_unet_mds_b200', 'libprobunet_b200.so')
KEYS = ['UTCHMMA', 'UTMALDG', 'UTMAREDG', 'UBLKCP', 'UBLKRED', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'FFMA2', 'REDG', 'MUFU', 'HMMA']

out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*', '', name).replace('pu::', '')
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['_total'] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f'# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels (sm_100a SASS); columns = instruction counts')
print('kernel\ttotal\t' + '\t'.join(KEYS))
tot = collections.Counter()
for name, c in counts.items():
    print(f'{name[:60]}\t{c["_total"]}\t' + '\t'.join(str(c[k]) for k in KEYS))
    tot.update(c)
print('ALL\t' + str(tot['_total']) + '\t' + '\t'.join(str(tot[k]) for k in KEYS))
