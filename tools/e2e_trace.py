"""Phase timing of the drop-in compute() on the bench workload (J2P_TRACE=1)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['J2P_TRACE'] = '1'
import numpy as np
from jpeg2png_b200 import abi, synth
import bench
lib = abi.load_product()
img = synth.synth_coefs(3840, 2160, 50, '4:4:4', 1237)
fdata = bench.device_decode(lib, img, 0)
arrays = [abi.CoefArray(img, [0, 1, 2], fdata) for _ in range(3)]
pw = (C.c_float * 3)(0.001, 0.001, 0.001)
lg = abi.Logger(None, b'', 3, 0)
for k in range(3):
    t0 = time.perf_counter()
    lib.compute(3, arrays[k].arr, C.byref(lg), None, C.c_float(0.3), pw, 100)
    print(f'call {k}: {1e3*(time.perf_counter()-t0):.1f} ms total', file=sys.stderr)
