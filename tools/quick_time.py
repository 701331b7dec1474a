"""A/B timing aid: per-kernel device time of the bench workload for a given build of the library.
usage: python tools/quick_time.py [lib.so ...]   (never a bench number; see bench.py)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpeg2png_b200 import abi, synth  # noqa: E402

# optional leading frame spec: --frame W H Q SUBSAMPLING   (default: the bench workload)
argv = sys.argv[1:]
separate = False
if argv[:1] == ['--separate']:                  # -s mode: three one-plane sessions, one after the other (jpeg2png.c:147-152)
    separate = True
    argv = argv[1:]
frame = (3840, 2160, 50, '4:4:4')
if argv[:1] == ['--frame']:
    frame = (int(argv[1]), int(argv[2]), int(argv[3]), argv[4])
    argv = argv[5:]
libs = argv or [abi.PRODUCT_LIB]
if frame[0] * frame[1] > 3840 * 2160:          # large frames: tile a quarter-size cartoon (seconds instead of half a minute)
    base = synth.synth_coefs(-(-frame[0] // 64) * 16, -(-frame[1] // 64) * 16, frame[2], frame[3], 1237)
    img = synth.tile_coefs(base, 4, 4, frame[0], frame[1])
else:
    img = synth.synth_coefs(frame[0], frame[1], frame[2], frame[3], 1237)
for path in libs:
    lib = abi.declare_product(C.CDLL(path, mode=C.RTLD_LOCAL))
    groups = [[0], [1], [2]] if separate else [[0, 1, 2]]
    tg = tp = ts = 0.
    checksum = 0.
    for chans in groups:
        d = abi.FrameDesc()
        d.nchannel = len(chans)
        for k, c in enumerate(chans):
            p = img.planes[c]
            d.plane_w[k], d.plane_h[k], d.w_samp[k], d.h_samp[k] = p.w, p.h, p.w_samp, p.h_samp
            d.pweight[k] = 0.001
        d.weight = 0.3 if chans[0] == 0 else 0.0      # the command line's defaults: TGV weight on the first plane only in -s mode
        d.iterations = 100
        s = C.c_void_p()
        assert lib.j2p_session_create(C.byref(s), 0, C.byref(d)) == 0, lib.j2p_last_error()
        for k, c in enumerate(chans):
            p = img.planes[c]
            data = np.ascontiguousarray(p.data)
            quant = np.ascontiguousarray(p.quant)
            assert lib.j2p_session_upload(s, k, data.ctypes.data, quant.ctypes.data, None) == 0
        mg, mp = C.c_float(), C.c_float()
        lib.j2p_session_profile(s, 10, C.byref(mg), C.byref(mp))
        lib.j2p_session_profile(s, 40, C.byref(mg), C.byref(mp))
        # whole solves queued back to back (what bench.py's `value` times): wall clock around 3 x 100 iterations
        import time
        lib.j2p_session_iterate(s, 0, 100)
        lib.j2p_session_sync(s)
        t0 = time.perf_counter()
        for _ in range(3):
            lib.j2p_session_iterate(s, 0, 100)
        lib.j2p_session_sync(s)
        ts += (time.perf_counter() - t0) / 300 * 1e6
        lib.j2p_session_profile(s, 40, C.byref(mg), C.byref(mp))       # leaves the session after 40 iterations
        tg += mg.value * 1e3
        tp += mp.value * 1e3
        out = np.empty((int(lib.j2p_session_height(s)), int(lib.j2p_session_width(s))), np.float32)
        lib.j2p_session_download(s, 0, out.ctypes.data)
        checksum += float(np.float64(out).sum())
        lib.j2p_session_destroy(s)
    print(f'{os.path.basename(path):40s} gradient {tg:8.1f} us  project {tp:8.1f} us  sum {tg + tp:8.1f} us  in-solve {ts:8.1f} us/iteration'
          f'{"  (-s: three one-plane solves)" if separate else ""}  checksum {checksum:.6f}')
