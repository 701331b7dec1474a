"""Load tests/golden/case_*.npz back into a CoefImage + parameters."""
import glob
import os

import numpy as np

from jpeg2png_b200.synth import CoefImage, Plane

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'case_*.npz')))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f'case_{name}.npz'))
    planes = []
    for k in range(3):
        w, h, sw, sh = (int(v) for v in z[f'p{k}_dims'])
        planes.append(Plane(w=w, h=h, w_samp=sw, h_samp=sh, data=z[f'p{k}_data'], quant=z[f'p{k}_quant']))
    img = CoefImage(width=int(z['width']), height=int(z['height']), planes=planes)
    channels = [int(c) for c in z['channels']]
    fdata = [z[f'fdata{k}'] for k in range(len(channels))]
    out = [z[f'out{k}'] for k in range(len(channels))]
    return dict(img=img, channels=channels, weight=float(z['weight']), pweight=[float(v) for v in z['pweight']],
                iterations=int(z['iterations']), fdata=fdata, out=out)
