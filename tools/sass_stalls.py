"""Decode the scheduling control bits of a cuobjdump -sass listing (Volta+ encoding: bits 105-108 of
the 128-bit instruction word = stall cycles, 110-112 / 113-115 = scoreboard written / read-released,
116-121 = scoreboards waited for) and sum the static stall cycles over an address range.
usage: sass_stalls.py listing.sass LO HI [SKIP_LO SKIP_HI ...]   (hex addresses)"""
import re
import sys
import collections

lines = open(sys.argv[1]).read().split('\n')
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
skip = [(int(a, 16), int(b, 16)) for a, b in zip(sys.argv[4::2], sys.argv[5::2])]
tot = n = waits = 0
by = collections.Counter()
for i, l in enumerate(lines):
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/', l)
    if not m:
        continue
    a = int(m.group(1), 16)
    if a < lo or a > hi or any(x <= a < y for x, y in skip):
        continue
    m2 = re.search(r'/\* 0x([0-9a-f]{16}) \*/', lines[i + 1])
    hiw = int(m2.group(1), 16)
    ctrl = hiw >> 41
    stall = ctrl & 0xf
    wmask = (ctrl >> 11) & 0x3f
    op = m.group(2).split()[0]
    if op.startswith('@'):
        op = m.group(2).split()[1]
    tot += max(stall, 1)
    n += 1
    by[op.split('.')[0]] += max(stall, 1)
    if wmask:
        waits += 1
print(f'instructions {n}  static issue cycles {tot}  ({tot / n:.2f} per instruction)  instructions waiting on a scoreboard {waits}')
for k, v in by.most_common(12):
    print(f'  {k:10s}{v:6d}')
