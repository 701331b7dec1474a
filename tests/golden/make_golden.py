"""Generate the golden vectors under tests/golden/ with the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled into oracle/_ref by oracle/Makefile):

    python tests/golden/make_golden.py

Each case_*.npz holds the inputs (`struct coef` contents + solver parameters) and the float planes
the reference's compute() (SIMD build = what ships, reference Makefile:13) returned for them.  The
reference itself has no tests or fixtures, so these files are the pin for machines
where /root/reference does not exist (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from jpeg2png_b200 import synth  # noqa: E402
from tests import helpers as H  # noqa: E402

CASES = {
    # name: (image factory, channels, weight, pweights, iterations)
    'joint420_q10': (lambda: synth.synth_coefs(64, 48, 10, '4:2:0', 101), [0, 1, 2], 0.3, [0.001] * 3, 20),
    'joint444_mixedp': (lambda: synth.synth_coefs(40, 24, 50, '4:4:4', 102), [0, 1, 2], 0.3, [0.001, 0.0, 0.002], 15),
    'sep_chroma_resample': (lambda: synth.synth_coefs(64, 48, 30, '4:2:0', 103), [1], 0.3, [0.001], 15),
    'sep_luma': (lambda: synth.synth_coefs(64, 48, 30, '4:2:0', 103), [0], 0.1, [0.001], 15),
    # 1080p-style edge case: luma grid has fewer rows than the frame (SURVEY headline 4)
    'luma_short_420': (lambda: synth.synth_coefs(48, 40, 20, '4:2:0', 104), [0, 1, 2], 0.3, [0.001] * 3, 15),
    'no_tgv': (lambda: synth.synth_coefs(48, 32, 75, '4:2:0', 105), [0, 1, 2], 0.0, [0.001] * 3, 15),
    'no_prob': (lambda: synth.synth_coefs(48, 32, 25, '4:4:4', 106), [0, 1, 2], 0.3, [0.0] * 3, 15),
    'odd_sampling': (lambda: synth.random_coefs([(48, 24), (16, 16), (24, 8)], [(1, 1), (3, 2), (2, 3)], 107),
                     [0, 1, 2], 0.5, [0.001, 0.002, 0.0005], 12),
    'tiny_8x8': (lambda: synth.random_coefs([(8, 8)] * 3, [(1, 1)] * 3, 108), [0, 1, 2], 0.3, [0.001] * 3, 10),
    'zero_iterations': (lambda: synth.synth_coefs(32, 32, 10, '4:2:0', 109), [0, 1, 2], 0.3, [0.001] * 3, 0),
}


def main():
    assert H.have_ref() or os.path.exists(H.REFERENCE_SRC), 'needs the reference'
    H.build_oracle_libs()
    for name, (factory, channels, weight, pw, iters) in CASES.items():
        img = factory()
        fdata = H.decode_planes(img, channels)
        out = H.run_compute('ref', img, channels, weight, pw, iters, fdata)
        blob = {'channels': np.array(channels), 'weight': np.float32(weight), 'pweight': np.array(pw, np.float32),
                'iterations': np.int64(iters), 'width': np.int64(img.width), 'height': np.int64(img.height)}
        for k, p in enumerate(img.planes):
            blob[f'p{k}_dims'] = np.array([p.w, p.h, p.w_samp, p.h_samp])
            blob[f'p{k}_data'] = p.data
            blob[f'p{k}_quant'] = p.quant
        for k, f in enumerate(fdata):
            blob[f'fdata{k}'] = f
        for k, o in enumerate(out):
            blob[f'out{k}'] = o
        path = os.path.join(HERE, f'case_{name}.npz')
        np.savez_compressed(path, **blob)
        print(name, [o.shape for o in out], os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
