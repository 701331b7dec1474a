"""The synthetic inputs bench.py and the tools use (jpeg2png_b200/synth.py)."""
import numpy as np
import pytest

from jpeg2png_b200 import synth


def test_grid_sizes_follow_libjpeg():
    """Plane grids are ceil(ceil(dim / samp) / 8) blocks, not MCU padded (jpeg.c:52-63)."""
    img = synth.synth_coefs(1920, 1080, 10, '4:2:0', seed=1)
    assert [(p.w, p.h, p.w_samp, p.h_samp) for p in img.planes] == [(1920, 1080, 1, 1), (960, 544, 2, 2), (960, 544, 2, 2)]
    assert (img.frame_w, img.frame_h) == (1920, 1088)
    for p in img.planes:
        assert p.data.dtype == np.int16 and p.data.size == p.w * p.h and (p.quant > 0).all()


def test_tile_coefs_repeats_blocks():
    base = synth.synth_coefs(64, 48, 30, '4:2:0', seed=3)
    big = synth.tile_coefs(base, 3, 2, 150, 90)
    assert (big.width, big.height) == (150, 90)
    for pb, pt in zip(base.planes, big.planes):
        cw, ch = -(-150 // pb.w_samp), -(-90 // pb.h_samp)
        assert (pt.w, pt.h) == (-(-cw // 8) * 8, -(-ch // 8) * 8)
        bb = pb.data.reshape(pb.h // 8, pb.w // 8, 64)
        tb = pt.data.reshape(pt.h // 8, pt.w // 8, 64)
        for by in range(tb.shape[0]):
            for bx in range(tb.shape[1]):
                assert (tb[by, bx] == bb[by % bb.shape[0], bx % bb.shape[1]]).all()
        assert (pt.quant == pb.quant).all()
    with pytest.raises(ValueError):
        synth.tile_coefs(base, 1, 1, 150, 90)


def test_synthesis_is_deterministic():
    a = synth.synth_coefs(96, 64, 50, '4:4:4', seed=9)
    b = synth.synth_coefs(96, 64, 50, '4:4:4', seed=9)
    assert all((p.data == q.data).all() for p, q in zip(a.planes, b.planes))
