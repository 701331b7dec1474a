"""Dynamic opcode mix and hottest stall sites from an .ncu-rep source page (SASS view)."""
import csv
import collections
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iSmp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix = collections.Counter()
samples = collections.Counter()
tot = 0
insts = []
for r in rows[2:]:
    if len(r) <= iE:
        continue
    src = r[iS].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
    base = op.split('.')[0]
    if base in ('F2F', 'MUFU', 'I2F', 'F2I'):
        base = '.'.join(op.split('.')[:3])
    n = int(r[iE] or 0)
    sm = int(r[iSmp] or 0)
    mix[base] += n
    samples[base] += sm
    tot += n
    insts.append((sm, n, src))
print(f'total warp-instructions executed: {tot}')
ssum = sum(samples.values()) or 1
for op, n in mix.most_common(28):
    print(f'  {op:18s} {n:12d} {100.0 * n / tot:6.2f}%   stall-samples {100.0 * samples[op] / ssum:6.2f}%')
print('hottest stall sites:')
for sm, n, src in sorted(insts, reverse=True)[:14]:
    print(f'  {sm:6d} samples  exec {n:9d}  {src[:90]}')
