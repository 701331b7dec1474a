"""Host-side staging threads (jpeg2png_b200/csrc/copy_pool.h): built with g++ and run on the CPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_copy_pool(tmp_path):
    exe = tmp_path / 'copy_pool_check'
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([cxx, '-O2', '-std=c++17', '-pthread', '-o', str(exe), os.path.join(HERE, 'copy_pool_check.cpp')], check=True)
    for threads in ('1', '3', '8'):
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, J2P_COPY_THREADS=threads), timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        assert 'copy_pool ok' in r.stdout
