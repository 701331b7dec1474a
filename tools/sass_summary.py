"""Per-kernel counts of the instructions that characterise the design (packed fp32, cp.async, TMA, mbarrier,
programmatic dependent launch, the fp64-promoted transforms) in the built library — static SASS, cold paths
included.  usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_sass_techniques import sass_by_kernel  # noqa: E402

COLS = ['FFMA2', 'FMUL2', 'FADD2', 'LDGSTS', 'UTMALDG', 'UTMASTG', 'SYNCS', 'ACQBULK', 'PREEXIT', 'F2F', 'DMUL', 'MUFU', 'SHFL', 'BAR']
names = sass_by_kernel()
print('ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents; LDGSTS = cp.async; UTMALDG / UTMASTG = cp.async.bulk.tensor; SYNCS = mbarrier')
print(f'{"kernel":64s} {"instr":>6s} ' + ' '.join(f'{c:>7s}' for c in COLS))
for mangled in sorted(names):
    ops = names[mangled]
    dem = subprocess.run(['c++filt', mangled], capture_output=True, text=True).stdout.strip().replace('j2p::', '').split('(')[0].replace('void ', '')
    print(f'{dem[:64]:64s} {sum(ops.values()):6d} ' + ' '.join(f'{ops[c]:7d}' for c in COLS))
