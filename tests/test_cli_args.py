"""Command-line validation of jpeg2png_b200/cli/jpeg2png: the messages and exit codes of the
reference (jpeg2png.c:205-311; `jpeg2png: <message>` on stderr, exit status 1, utils.c:11-28).
All of these are decided before any device is touched, so they run on the CPU box."""
import os
import subprocess

import pytest

from tests.test_codecs import CLI_DIR, make_jpeg


@pytest.fixture(scope='module')
def exe():
    subprocess.run(['make', '-C', CLI_DIR, 'jpeg2png'], check=True, capture_output=True)
    return os.path.join(CLI_DIR, 'jpeg2png')


def run(exe, *args, cwd=None):
    return subprocess.run([exe, *args], capture_output=True, text=True, cwd=cwd, timeout=60)


@pytest.mark.parametrize('args,message', [
    (['-w', '0.1,0.2,0.3', 'a.jpg'], 'different weights are only possible when using separated components'),      # jpeg2png.c:211
    (['-w', '1,2', 'a.jpg'], 'invalid weight'),                                                                     # :216
    (['-w', 'x', 'a.jpg'], 'invalid weight'),
    (['-p', '1,2', 'a.jpg'], 'invalid probability weight'),                                                        # :228
    (['-i', '10,20,30', 'a.jpg'], 'different iteration counts are only possible when using separated components'),  # :236
    (['-i', '5,6', 'a.jpg'], 'invalid number of iterations'),                                                      # :242
    (['-t', '0', 'a.jpg'], 'invalid number of threads'),                                                           # :251
    (['-o', 'x.png', 'a.jpg', 'b.jpg'], 'must give output file names for all input files or none'),                # :277
    (['no_such_file.jpg'], 'could not open input file `no_such_file.jpg`'),                                        # :288
])
def test_validation_messages(exe, tmp_path, args, message):
    r = run(exe, *args, cwd=str(tmp_path))
    assert r.returncode == 1
    assert r.stderr.strip().splitlines()[-1] == 'jpeg2png: ' + message


def test_refuses_to_overwrite_before_touching_a_device(exe, tmp_path):
    (tmp_path / 'pic.jpg').write_bytes(make_jpeg(32, 32, 50, '4:2:0'))
    (tmp_path / 'pic.png').write_bytes(b'already here')
    r = run(exe, '-q', 'pic.jpg', cwd=str(tmp_path))
    assert r.returncode == 1
    assert r.stderr.strip().splitlines()[-1] == 'jpeg2png: not overwriting output file `pic.png`'    # jpeg2png.c:306
    assert (tmp_path / 'pic.png').read_bytes() == b'already here'


def test_usage_without_arguments(exe):
    r = run(exe)
    assert r.returncode == 1 and r.stdout.startswith('usage: jpeg2png')        # usage() exits with failure, jpeg2png.c:27-116
