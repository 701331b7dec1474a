"""The exact shared-reciprocal division of numerics.cuh, restated for the host and checked against
IEEE division on the CPU (tests/qdiv_check.c); the exhaustive GPU-side check is tools/divcheck.cu."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shared_reciprocal_division_is_correctly_rounded(tmp_path):
    exe = tmp_path / 'qdiv_check'
    cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
    subprocess.run([cc, '-O2', '-std=c11', '-ffp-contract=off', '-msse2', '-mfpmath=sse', '-o', str(exe),
                    os.path.join(HERE, 'qdiv_check.c'), '-lm'], check=True)
    r = subprocess.run([str(exe), '30000000'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ' 0 mismatches' in r.stdout


def test_row_guard_property(tmp_path):
    """One magnitude test per loaded value implies the per-numerator guard (tests/rowguard_check.c)."""
    exe = tmp_path / 'rowguard_check'
    cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
    subprocess.run([cc, '-O2', '-std=c11', '-ffp-contract=off', '-msse2', '-mfpmath=sse', '-o', str(exe),
                    os.path.join(HERE, 'rowguard_check.c'), '-lm'], check=True)
    r = subprocess.run([str(exe), '5000000'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ' 0 below 2^-60' in r.stdout
