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
    d = abi.FrameDesc()
    d.nchannel = 3
    for c, p in enumerate(img.planes):
        d.plane_w[c], d.plane_h[c], d.w_samp[c], d.h_samp[c] = p.w, p.h, p.w_samp, p.h_samp
        d.pweight[c] = 0.001
    d.weight = 0.3
    d.iterations = 100
    s = C.c_void_p()
    assert lib.j2p_session_create(C.byref(s), 0, C.byref(d)) == 0, lib.j2p_last_error()
    for c, p in enumerate(img.planes):
        data = np.ascontiguousarray(p.data)
        quant = np.ascontiguousarray(p.quant)
        assert lib.j2p_session_upload(s, c, data.ctypes.data, quant.ctypes.data, None) == 0
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
    solve_us = (time.perf_counter() - t0) / 300 * 1e6
    lib.j2p_session_profile(s, 40, C.byref(mg), C.byref(mp))       # leaves the session after 40 iterations, as before
    out = np.empty((img.frame_h, img.frame_w), np.float32)
    lib.j2p_session_download(s, 0, out.ctypes.data)
    print(f'{os.path.basename(path):40s} gradient {mg.value*1e3:8.1f} us  project {mp.value*1e3:8.1f} us  sum {(mg.value+mp.value)*1e3:8.1f} us  in-solve {solve_us:8.1f} us/iteration  checksum {float(np.float64(out).sum()):.6f}')
    lib.j2p_session_destroy(s)
