"""BASELINE config 5 through the command line (measurement aid): N synthetic 1920x1080 Q75 JPEGs,
`jpeg2png -i 100` over all visible GPUs; reports wall time per phase set.  Files come from PIL."""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_codecs import CLI_DIR, make_jpeg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
threads = sys.argv[2] if len(sys.argv) > 2 else None
exe = os.path.join(CLI_DIR, 'jpeg2png')
with tempfile.TemporaryDirectory() as d:
    files = []
    for i in range(n):
        p = os.path.join(d, f'f{i:03d}.jpg')
        open(p, 'wb').write(make_jpeg(1920, 1080, 75, '4:2:0', False, seed=100 + i))
        files.append(p)
    for rep in range(2):
        for f in files:
            png = f[:-4] + '.png'
            if os.path.exists(png):
                os.remove(png)
        cmd = [exe, '-q', '-i', '100'] + (['-t', threads] if threads else []) + files
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, J2P_TRACE='1') if rep == 1 else None)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        print(f'cli batch: {n} x 1920x1080 Q75 4:2:0, -i 100{" -t " + threads if threads else ""}: {dt:.2f} s  '
              f'{n / dt:.1f} files/s  {n * 1920 * 1080 * 100 / dt / 1e6:.0f} Mpix-it/s (JPEG read + solve + PNG write)', flush=True)
        if rep == 1:
            print('\n'.join(l for l in r.stderr.splitlines() if "read+parse" in l or "session create" in l)[:9000], flush=True)
