"""Small solves through the C ABI for compute-sanitizer (tools/run_sanitizers.sh): every kernel family once
— joint 4:4:4, joint 4:2:0 with a frame larger than the luma grid, one-plane (-s) solves, odd sampling."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpeg2png_b200 import synth  # noqa: E402
from tests import helpers as H  # noqa: E402

cases = [
    (136, 72, 50, '4:4:4', [0, 1, 2], 0.7, [0.001, 0.0, 0.01], 4),
    (200, 120, 30, '4:2:0', [0, 1, 2], 0.3, [0.001] * 3, 4),
    (256, 256, 10, '4:2:0', [1], 0.3, [0.001], 3),
    (64, 64, 90, '4:4:4', [0], 0.0, [0.0], 3),
]
for w, h, q, ss, channels, weight, pw, iters in cases:
    img = synth.synth_coefs(w, h, q, ss, seed=99 + w)
    f = H.decode_planes(img, channels)
    out = H.run_compute('product', img, channels, weight, pw, iters, f)
    print('ok', w, h, ss, channels, float(out[0].sum()))
img = synth.random_coefs([(40, 24), (24, 16), (16, 8)], [(1, 1), (2, 2), (3, 4)], 1)
out = H.run_compute('product', img, [0, 1, 2], 0.4, [0.001] * 3, 3, H.decode_planes(img))
print('ok random planes', float(out[0].sum()))
