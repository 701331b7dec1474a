"""Parity at the BENCHMARKED sizes (pytest -m gpu): BASELINE config 3 (3840x2160 4:4:4, the bench
workload) and config 4 (7680x4320 4:2:0) against the compiled reference's own compute()
(compute.c:407-465) — a bounded number of iterations, because the reference needs ~0.7 s per 4K
iteration and ~2 s per 8K iteration on the box's host cores; per-iteration work does not depend on
the iteration count.  Also: compute() re-entrancy as jpeg2png.c:147-152 uses it (three concurrent
calls with nchannel = 1)."""
import ctypes as C
import threading

import numpy as np
import pytest

from jpeg2png_b200 import abi, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _checker():
    return 'ref' if H.have_ref() else 'oracle'


@pytest.fixture(scope='module')
def lib():
    lib = abi.load_product()
    assert lib.j2p_device_count() > 0, 'no CUDA device visible: the product has no CPU fallback'
    return lib


def big_frame(width, height, quality, subsampling, seed):
    """The bench/strip workloads' frame builder: a quarter-size cartoon tiled 4x4 at block level
    (same statistics, a fraction of the host time of a full-size synthesis)."""
    base = synth.synth_coefs(-(-width // 64) * 16, -(-height // 64) * 16, quality, subsampling, seed)
    return synth.tile_coefs(base, 4, 4, width, height)


def test_config3_4k_444_matches_reference(lib):
    """BASELINE config 3 — the frame bench.py times — 20 of its 100 iterations, bit for bit."""
    img = synth.synth_coefs(3840, 2160, 50, '4:4:4', seed=1237)
    f = H.decode_planes(img)
    want = H.run_compute(_checker(), img, [0, 1, 2], 0.3, [0.001] * 3, 20, f)
    got = H.run_compute('product', img, [0, 1, 2], 0.3, [0.001] * 3, 20, f)
    H.assert_bit_identical(got, want, 'config 3 (4K 4:4:4) x20')


def test_config4_8k_420_matches_reference(lib):
    """BASELINE config 4's frame on ONE GPU, 10 iterations, bit for bit (the strip version of the
    same comparison is tests/test_gpu_strips.py::test_8k_strips_match_reference)."""
    img = big_frame(7680, 4320, 10, '4:2:0', 1238)
    assert (img.frame_w, img.frame_h) == (7680, 4320)
    f = H.decode_planes(img)
    want = H.run_compute(_checker(), img, [0, 1, 2], 0.3, [0.001] * 3, 10, f)
    got = H.run_compute('product', img, [0, 1, 2], 0.3, [0.001] * 3, 10, f)
    H.assert_bit_identical(got, want, 'config 4 (8K 4:2:0) x10')


@pytest.mark.parametrize('ss', ['4:2:0', '4:4:4'])
def test_compute_is_reentrant(lib, ss):
    """jpeg2png.c:147-152 (-s mode) calls compute(1, &coef[c], ...) for the three planes from three
    OpenMP threads at once, jpeg2png.c:330 does the same across files.  Three host threads, each its
    own plane, twice over (the second round reuses cached device blocks and pinned buffers): every
    result must equal the serial call and the checker."""
    img = synth.synth_coefs(712, 408, 25, ss, seed=99)
    weights = [0.3, 0.1, 0.0]
    iters = 30
    f = H.decode_planes(img)
    serial = [H.run_compute('product', img, [c], weights[c], [0.001], iters, [f[c]])[0] for c in range(3)]
    want = [H.run_compute(_checker(), img, [c], weights[c], [0.001], iters, [f[c]])[0] for c in range(3)]
    H.assert_bit_identical(serial, want, 'serial -s mode vs checker')
    for rnd in range(2):
        got, errs = [None] * 3, []
        gate = threading.Barrier(3)

        def work(c):
            try:
                gate.wait()
                got[c] = H.run_compute('product', img, [c], weights[c], [0.001], iters, [f[c]])[0]
            except Exception as e:        # noqa: BLE001 — reported below
                errs.append((c, repr(e)))
        ts = [threading.Thread(target=work, args=(c,)) for c in range(3)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errs, errs
        H.assert_bit_identical(got, want, f'concurrent -s mode, round {rnd}')


def test_thread_device_binding(lib):
    """j2p_set_thread_device: the calling thread's compute() calls go to that device (what the -s
    mode of the command line uses to spread the three planes over up to three GPUs).  On a one-GPU
    box only device 0 exists: binding to it must work, binding beyond the count must be refused."""
    n = lib.j2p_device_count()
    assert lib.j2p_set_thread_device(0) == 0
    assert lib.j2p_set_thread_device(n) != 0
    assert lib.j2p_set_thread_device(-1) == 0          # back to the default (J2P_DEVICE or 0)
    img = synth.synth_coefs(128, 96, 30, '4:2:0', seed=5)
    f = H.decode_planes(img)
    want = H.run_compute(_checker(), img, [0], 0.3, [0.001], 8, [f[0]])
    devs = list(range(min(n, 3)))
    got = {}

    def work(dev):
        lib.j2p_set_thread_device(dev)
        got[dev] = H.run_compute('product', img, [0], 0.3, [0.001], 8, [f[0]])
    ts = [threading.Thread(target=work, args=(d,)) for d in devs]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for dev in devs:
        H.assert_bit_identical(got[dev], want, f'device {dev}')
