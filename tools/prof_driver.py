"""Short resident-session run of the bench workload, for use under ncu (never a bench number)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpeg2png_b200 import abi, synth  # noqa: E402

w, h, q, ss, iters = 3840, 2160, 50, '4:4:4', 12
lib_path = None
if len(sys.argv) > 2 and sys.argv[1] == '--lib':      # an A/B build (tools/build_variant.sh)
    lib_path = sys.argv[2]
    del sys.argv[1:3]
if len(sys.argv) > 1:
    w, h, q, ss, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
img = synth.synth_coefs(w, h, q, ss, 1237)
lib = abi.declare_product(C.CDLL(lib_path, mode=C.RTLD_LOCAL)) if lib_path else abi.load_product()
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
    assert lib.j2p_session_upload(s, c, data.ctypes.data, quant.ctypes.data, None) == 0, lib.j2p_last_error()
assert lib.j2p_session_iterate(s, 0, iters) == 0, lib.j2p_last_error()
lib.j2p_session_sync(s)
lib.j2p_session_destroy(s)
print('prof_driver done', w, h, ss, iters)
