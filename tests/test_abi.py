"""CPU tests of the product's C-ABI boundary: the library loads, exports every symbol
include/jpeg2png_b200.h declares, and — with no GPU — FAILS LOUDLY instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from jpeg2png_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'jpeg2png_b200.h')


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(abi.PRODUCT_LIB):
        subprocess.run(['make', '-C', os.path.dirname(abi.PRODUCT_LIB)], check=True, capture_output=True)
    return abi.load_product()


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;', text)
    return sorted(set(n for n in names if n == 'compute' or n.startswith('j2p_')))


def test_header_symbol_list_is_complete():
    assert _declared_functions() == sorted(abi.HEADER_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), f'{name} declared in the header but not exported'


def test_struct_coef_layout_matches_reference():
    """reference jpeg2png.h:7-20 on LP64: 4 unsigned, 2 pointers, 64 uint16 = 160 bytes."""
    assert C.sizeof(abi.Coef) == 160
    assert abi.Coef.data.offset == 16 and abi.Coef.fdata.offset == 24 and abi.Coef.quant_table.offset == 32


def test_version_string(lib):
    assert b'sm_100a' in lib.j2p_version()


@pytest.mark.skipif(os.environ.get('J2P_EXPECT_GPU') == '1', reason='GPU box')
def test_no_gpu_means_error_not_fallback(lib):
    if lib.j2p_device_count() > 0:
        pytest.skip('a CUDA device is present')
    d = abi.FrameDesc()
    d.nchannel = 1
    d.plane_w[0] = d.plane_h[0] = 8
    d.w_samp[0] = d.h_samp[0] = 1
    d.iterations = 1
    s = C.c_void_p()
    rc = lib.j2p_session_create(C.byref(s), 0, C.byref(d))
    assert rc == -3 and not s.value                      # J2P_ERR_NODEVICE
    assert b'no CPU fallback' in lib.j2p_last_error()


def test_compute_dies_like_the_reference_without_gpu(lib):
    """Drop-in error convention (utils.c:11-28): 'jpeg2png: ...' on stderr, EXIT_FAILURE."""
    if lib.j2p_device_count() > 0:
        pytest.skip('a CUDA device is present')
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from jpeg2png_b200 import synth\n"
        "from tests import helpers as H\n"
        "import numpy as np\n"
        "img = synth.random_coefs([(8,8)]*3, [(1,1)]*3, 1)\n"
        "f = [np.zeros((8,8), np.float32)]*3\n"
        "H.run_compute('product', img, [0,1,2], 0.3, [0.001]*3, 2, f)\n"
        "print('survived')\n" % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert r.returncode == 1
    assert r.stderr.startswith('jpeg2png: ')
    assert 'survived' not in r.stdout


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    monkeypatch.setattr(abi, 'PRODUCT_LIB', str(tmp_path / 'nope.so'))
    monkeypatch.setattr(abi, '_product', None)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        abi.load_product()
