"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libjpeg2png_b200.so and is compared BIT FOR BIT with the oracle restatement, with the committed
golden vectors (produced by the unmodified reference), and — when the prebuilt oracle/_ref
travelled with the snapshot — with the compiled reference itself."""
import ctypes as C
import os

import numpy as np
import pytest

from jpeg2png_b200 import abi, synth
from tests import helpers as H
from tests.golden_io import golden_cases, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def lib():
    lib = abi.load_product()
    assert lib.j2p_device_count() > 0, 'no CUDA device visible: the product has no CPU fallback'
    return lib


def _checker():
    """Reference build if it travelled, else the oracle restatement (pinned to it by test_oracle)."""
    return 'ref' if H.have_ref() else 'oracle'


@pytest.mark.parametrize('name', golden_cases())
def test_product_matches_golden(lib, name):
    g = load_case(name)
    out = H.run_compute('product', g['img'], g['channels'], g['weight'], g['pweight'], g['iterations'], g['fdata'])
    H.assert_bit_identical(out, g['out'], f'golden {name}')


@pytest.mark.parametrize('w,h,q,ss,channels,weight,pw,iters', [
    (256, 256, 10, '4:2:0', [0, 1, 2], 0.3, [0.001] * 3, 50),      # BASELINE config 1
    (256, 256, 10, '4:2:0', [0], 0.3, [0.001], 50),                # separate mode (-s), luma
    (256, 256, 10, '4:2:0', [1], 0.3, [0.001], 50),                # separate mode (-s), chroma (resample)
    (200, 120, 30, '4:2:0', [0, 1, 2], 0.3, [0.001] * 3, 30),      # frame larger than the luma grid
    (136, 72, 50, '4:4:4', [0, 1, 2], 0.7, [0.001, 0.0, 0.01], 40),
    (64, 64, 90, '4:4:4', [0, 1, 2], 0.0, [0.0] * 3, 20),          # TV only
    (520, 264, 75, '4:2:0', [0, 1, 2], 0.3, [0.001] * 3, 25),      # several CTAs in both directions, ragged edges
])
def test_product_matches_oracle(lib, w, h, q, ss, channels, weight, pw, iters):
    img = synth.synth_coefs(w, h, q, ss, seed=1234 + w + h)
    f = H.decode_planes(img, channels)
    want = H.run_compute(_checker(), img, channels, weight, pw, iters, f)
    got = H.run_compute('product', img, channels, weight, pw, iters, f)
    H.assert_bit_identical(got, want, 'product vs checker')


def test_product_matches_oracle_random_planes(lib):
    for seed in range(4):
        img = synth.random_coefs([(40, 24), (24, 16), (16, 8)], [(1, 1), (2, 2), (3, 4)], seed)
        f = H.decode_planes(img)
        want = H.run_compute(_checker(), img, [0, 1, 2], 0.4, [0.001] * 3, 12, f)
        got = H.run_compute('product', img, [0, 1, 2], 0.4, [0.001] * 3, 12, f)
        H.assert_bit_identical(got, want, f'random seed {seed}')


@pytest.mark.parametrize('case', ['tiny', 'patches', 'zero_coefs'])
def test_guard_fallback_rows(lib, case):
    """Values outside the proven range of the fast division / roots (numerics.cuh): the gradient
    kernel must fall back to IEEE arithmetic for those rows and still match bit for bit.  compute()
    takes the initial iterate from the caller, so the test plants the values there."""
    img = synth.synth_coefs(200, 72, 40, '4:4:4', seed=4321)
    f = H.decode_planes(img)
    rng = np.random.default_rng(7)
    if case == 'tiny':                       # every FISTA value far below 2^-35, some exactly 0
        f = [(p * np.float32(2.0 ** -60) * (rng.random(p.shape) < 0.8)).astype(np.float32) for p in f]
    elif case == 'patches':                  # a normal image with denormal, tiny and huge islands
        for p, v in zip(f, (1e-42, 3e-13, 4e21)):
            p[10:14, 30:90] = np.float32(v) * rng.standard_normal((4, 60)).astype(np.float32)
            p[40:41, :] = np.float32(v)
            p[50:60, 100:104] = 0.0
    else:                                    # all-zero coefficients, unit tables: tiny values survive the projection
        for pl in img.planes:
            pl.data[:] = 0
            pl.quant[:] = 1
        f = [(rng.standard_normal(p.shape) * 1e-14).astype(np.float32) for p in f]
    for iters in (1, 6):
        fin = [p.copy() for p in f]
        want = H.run_compute(_checker(), img, [0, 1, 2], 0.3, [0.001] * 3, iters, [p.copy() for p in fin])
        got = H.run_compute('product', img, [0, 1, 2], 0.3, [0.001] * 3, iters, [p.copy() for p in fin])
        H.assert_bit_identical(got, want, f'guard fallback {case} x{iters}')


def test_config2_1080p_420_full_length(lib):
    """BASELINE config 2: 1920x1080 Q10 4:2:0, -i 100, all three planes.  Frame 1920x1088, luma
    grid 1080 rows (SURVEY headline 4).  Checked against the compiled reference when present
    (about 15 s of host time), else against the oracle."""
    img = synth.synth_coefs(1920, 1080, 10, '4:2:0', seed=1236)
    assert (img.frame_w, img.frame_h) == (1920, 1088) and img.planes[0].h == 1080
    f = H.decode_planes(img)
    want = H.run_compute(_checker(), img, [0, 1, 2], 0.3, [0.001] * 3, 100, f)
    got = H.run_compute('product', img, [0, 1, 2], 0.3, [0.001] * 3, 100, f)
    H.assert_bit_identical(got, want, 'config 2')


def test_device_decode_matches_oracle(lib):
    """j2p_session_upload(fdata=NULL) runs the conventional decode on the device (jpeg.c:83-92)."""
    img = synth.synth_coefs(104, 72, 35, '4:2:0', seed=77)
    want = H.decode_planes(img)
    d = abi.FrameDesc()
    d.nchannel = 3
    for c, p in enumerate(img.planes):
        d.plane_w[c], d.plane_h[c], d.w_samp[c], d.h_samp[c] = p.w, p.h, p.w_samp, p.h_samp
        d.pweight[c] = 0.001
    d.weight = 0.3
    d.iterations = 0
    s = C.c_void_p()
    assert lib.j2p_session_create(C.byref(s), 0, C.byref(d)) == 0, lib.j2p_last_error()
    try:
        for c, p in enumerate(img.planes):
            data = np.ascontiguousarray(p.data)
            quant = np.ascontiguousarray(p.quant)
            assert lib.j2p_session_upload(s, c, data.ctypes.data, quant.ctypes.data, None) == 0, lib.j2p_last_error()
        W, H_ = lib.j2p_session_width(s), lib.j2p_session_height(s)
        for c, p in enumerate(img.planes):
            out = np.empty((H_, W), np.float32)
            assert lib.j2p_session_download(s, c, out.ctypes.data) == 0
            # aux_init upsampling of the decode (compute.c:295-302)
            yy = np.minimum(np.arange(H_) // p.h_samp, p.h - 1)
            xx = np.minimum(np.arange(W) // p.w_samp, p.w - 1)
            exp = want[c][yy][:, xx]
            assert (H.bits(out) == H.bits(exp)).all()
    finally:
        lib.j2p_session_destroy(s)


def test_session_reset_is_reproducible(lib):
    """Re-arming a resident session and re-running gives the same bits (run-to-run determinism of
    the fixed-order fp64 reduction)."""
    img = synth.synth_coefs(300, 200, 20, '4:2:0', seed=5)
    f = H.decode_planes(img)
    d = abi.FrameDesc()
    d.nchannel = 3
    for c, p in enumerate(img.planes):
        d.plane_w[c], d.plane_h[c], d.w_samp[c], d.h_samp[c] = p.w, p.h, p.w_samp, p.h_samp
        d.pweight[c] = 0.001
    d.weight = 0.3
    d.iterations = 30
    s = C.c_void_p()
    assert lib.j2p_session_create(C.byref(s), 0, C.byref(d)) == 0, lib.j2p_last_error()
    try:
        for c, p in enumerate(img.planes):
            data = np.ascontiguousarray(p.data)
            quant = np.ascontiguousarray(p.quant)
            fd = np.ascontiguousarray(f[c])
            assert lib.j2p_session_upload(s, c, data.ctypes.data, quant.ctypes.data, fd.ctypes.data) == 0
        W, H_ = lib.j2p_session_width(s), lib.j2p_session_height(s)
        runs = []
        for _ in range(3):
            assert lib.j2p_session_iterate(s, 0, 30) == 0, lib.j2p_last_error()
            planes = []
            for c in range(3):
                out = np.empty((H_, W), np.float32)
                assert lib.j2p_session_download(s, c, out.ctypes.data) == 0
                planes.append(out)
            runs.append(planes)
        H.assert_bit_identical(runs[0], runs[1], 'rerun 1')
        H.assert_bit_identical(runs[0], runs[2], 'rerun 2')
        want = H.run_compute('oracle', img, [0, 1, 2], 0.3, [0.001] * 3, 30, f)
        H.assert_bit_identical(runs[0], want, 'session vs oracle')
    finally:
        lib.j2p_session_destroy(s)


def test_objective_log_matches_oracle(lib):
    """-c csv path: the objective terms the reference logs every iteration (compute.c:271-272).
    The sums are re-associated on the GPU, so this is a tolerance test (1e-9 relative on fp64 sums
    of ~1e5 terms is far tighter than the reference's own C-vs-SIMD disagreement on prob_dist)."""
    img = synth.synth_coefs(200, 136, 15, '4:2:0', seed=31)
    f = H.decode_planes(img)
    iters = 12
    planes_o, log_o = H.run_compute('oracle', img, [0, 1, 2], 0.3, [0.001] * 3, iters, f, want_log=True)
    d = abi.FrameDesc()
    d.nchannel = 3
    for c, p in enumerate(img.planes):
        d.plane_w[c], d.plane_h[c], d.w_samp[c], d.h_samp[c] = p.w, p.h, p.w_samp, p.h_samp
        d.pweight[c] = 0.001
    d.weight = 0.3
    d.iterations = iters
    s = C.c_void_p()
    assert lib.j2p_session_create(C.byref(s), 0, C.byref(d)) == 0, lib.j2p_last_error()
    try:
        assert lib.j2p_session_set_logging(s, 1) == 0
        for c, p in enumerate(img.planes):
            data = np.ascontiguousarray(p.data)
            quant = np.ascontiguousarray(p.quant)
            fd = np.ascontiguousarray(f[c])
            assert lib.j2p_session_upload(s, c, data.ctypes.data, quant.ctypes.data, fd.ctypes.data) == 0
        got = np.zeros((iters, 4))
        for i in range(iters):
            assert lib.j2p_session_iterate(s, i, 1) == 0, lib.j2p_last_error()
            o = (C.c_double * 4)()
            assert lib.j2p_session_objective(s, o) == 0, lib.j2p_last_error()
            got[i] = list(o)
        np.testing.assert_allclose(got, log_o, rtol=1e-9, atol=1e-12)
        # logging must not change the iterate
        W, H_ = lib.j2p_session_width(s), lib.j2p_session_height(s)
        planes = []
        for c in range(3):
            out = np.empty((H_, W), np.float32)
            assert lib.j2p_session_download(s, c, out.ctypes.data) == 0
            planes.append(out)
        H.assert_bit_identical(planes, planes_o, 'logged run vs oracle')
    finally:
        lib.j2p_session_destroy(s)


def test_compute_callbacks_logger_and_progressbar(tmp_path):
    """Drop-in contract of compute(): log->iteration / logger_log once per iteration when a CSV
    log is open (compute.c:428, :271-272), progressbar_inc once per iteration (compute.c:449-452).
    Runs in a child process because the callback library must be loaded before the solver."""
    import subprocess
    import sys
    so = tmp_path / 'libcb.so'
    subprocess.run(['/usr/bin/gcc', '-shared', '-fPIC', '-O1', '-o', str(so), os.path.join(os.path.dirname(__file__), 'cb_helper.c')], check=True)
    code = f"""
import ctypes as C, sys
sys.path.insert(0, {H.ROOT!r})
cb = C.CDLL({str(so)!r}, mode=C.RTLD_GLOBAL)
import numpy as np
from jpeg2png_b200 import abi, synth
from tests import helpers as H
lib = abi.load_product()
img = synth.synth_coefs(96, 64, 20, '4:2:0', seed=9)
f = H.decode_planes(img)
_, want = H.run_compute('oracle', img, [0, 1, 2], 0.3, [0.001] * 3, 9, f, want_log=True)
ca = abi.CoefArray(img, [0, 1, 2], f)
libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p
fh = libc.fopen(b'/dev/null', b'w')
lg = abi.Logger(fh, b'x.jpg', 3, 0)
pb = abi.ProgressBar(0, 9)
pw = (C.c_float * 3)(0.001, 0.001, 0.001)
lib.compute(3, ca.arr, C.byref(lg), C.byref(pb), C.c_float(0.3), pw, 9)
assert cb.cb_log_count() == 9, cb.cb_log_count()
assert cb.cb_progress_count() == 9 and pb.current == 9
assert lg.iteration == 8
cb.cb_get_row.argtypes = [C.c_uint, C.POINTER(C.c_double)]
for i in range(9):
    row = (C.c_double * 5)()
    cb.cb_get_row(i, row)
    assert int(row[0]) == i
    np.testing.assert_allclose(list(row)[1:], want[i], rtol=1e-9, atol=1e-12)
print('callbacks ok')
"""
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert r.returncode == 0 and 'callbacks ok' in r.stdout, r.stdout + r.stderr


_TMA_CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from jpeg2png_b200 import abi, synth
from tests import helpers as H
checker = 'ref' if H.have_ref() else 'oracle'
cases = [
    (712, 384, 40, '4:4:4', [0, 1, 2], 0.3, [0.001, 0.002, 0.0], 20),   # 89 blocks wide: ragged last warp tile, one plane without the DCT-distance term
    (520, 264, 75, '4:2:0', [0, 1, 2], 0.3, [0.001] * 3, 25),          # luma through the TMA kernel, chroma through the 2x2 tile kernel
    (1920, 1080, 10, '4:2:0', [0], 0.3, [0.001], 12),                  # coefficient grid shorter than the frame (resample build), one plane
    (64, 64, 90, '4:4:4', [0, 1, 2], 0.0, [0.0] * 3, 10),              # TV only, no DCT-distance term at all
]
for w, h, q, ss, channels, weight, pw, iters in cases:
    img = synth.synth_coefs(w, h, q, ss, seed=4321 + w + h)
    f = H.decode_planes(img, channels)
    want = H.run_compute(checker, img, channels, weight, pw, iters, f)
    got = H.run_compute('product', img, channels, weight, pw, iters, f)
    H.assert_bit_identical(got, want, f'TMA projection {w}x{h} {ss}')
print('tma parity ok', len(cases))
'''


def test_tma_projection_matches_oracle(lib):
    """The opt-in persistent TMA-fed projection kernel (kernels_project_tma.cu: UTMALDG / UTMASTG,
    J2P_PROJ_TMA=1) against the checker, in a child process because the switch is read once per
    process.  The default projection kernels are what every other test in this file exercises."""
    import subprocess
    import sys
    env = dict(os.environ, J2P_PROJ_TMA='1')
    r = subprocess.run([sys.executable, '-c', _TMA_CHILD, H.ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'tma parity ok' in r.stdout
