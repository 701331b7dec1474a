"""Synthetic JPEG coefficient planes (producer (ii): direct coefficient synthesis).

There is no network and no libjpeg headers on the build box, so benchmark and parity inputs are
made the way an encoder would make them: a deterministic cartoon-like image (the reference's
target content, reference README.md:43-44) is level-shifted, converted to YCbCr, chroma is
box-averaged to the subsampled grid, each plane is edge-replicated to whole 8x8 blocks,
transformed with an orthonormal 8x8 DCT and quantised with the IJG tables scaled to the requested
quality.  The result is exactly the `struct coef` contents libjpeg's jpeg_read_coefficients would
hand to the reference (jpeg.c:49-78): int16 coefficients, block-major, natural order, block grid
NOT padded to whole MCUs (jpeg.c:52-53).

Pure numpy; nothing here is on the hot path.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# IJG / Annex K base tables, natural (row-major) order.
_BASE_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)
_BASE_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99,
    18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64)


def quant_table(quality: int, chroma: bool) -> np.ndarray:
    """IJG quality scaling (jpeg_quality_scaling + jpeg_add_quant_table, baseline clamp)."""
    quality = max(1, min(100, int(quality)))
    s = 5000 // quality if quality < 50 else 200 - 2 * quality
    base = _BASE_CHROMA if chroma else _BASE_LUMA
    q = (base * s + 50) // 100
    return np.clip(q, 1, 255).astype(np.uint16)


def _dct_matrix() -> np.ndarray:
    k = np.arange(8)[:, None]
    n = np.arange(8)[None, :]
    m = np.cos((2 * n + 1) * k * np.pi / 16) * 0.5
    m[0, :] = np.sqrt(1.0 / 8.0)
    return m


_D = _dct_matrix()


def cartoon_image(width: int, height: int, seed: int) -> np.ndarray:
    """Deterministic RGB float image, HxWx3 in [0,255]: smooth shading + hard edges + mild noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    smooth = 127.0 + 100.0 * np.sin(0.013 * x) * np.cos(0.017 * y)
    edges = 45.0 * np.sign(np.sin(0.071 * x + 0.3) * np.sin(0.083 * y + 1.1))
    checker = 30.0 * ((((x // 64) + (y // 80)) % 2) * 2.0 - 1.0)
    r = smooth + edges
    g = smooth * 0.8 + checker + 20.0
    b = 255.0 - smooth + 0.5 * edges - 0.5 * checker
    img = np.stack([r, g, b], axis=-1)
    img += rng.integers(-8, 9, size=img.shape).astype(np.float64)
    return np.clip(img, 0.0, 255.0)


@dataclasses.dataclass
class Plane:
    """One colour component as the reference's `struct coef` describes it."""
    w: int                 # samples, multiple of 8
    h: int
    w_samp: int
    h_samp: int
    data: np.ndarray       # int16 [blocks*64], block-major natural order
    quant: np.ndarray      # uint16 [64]


@dataclasses.dataclass
class CoefImage:
    width: int             # image size in pixels (what the metric counts)
    height: int
    planes: list           # [Y, Cb, Cr]

    @property
    def frame_w(self) -> int:
        return max(p.w * p.w_samp for p in self.planes)

    @property
    def frame_h(self) -> int:
        return max(p.h * p.h_samp for p in self.planes)


def _block_dct_quantise(plane: np.ndarray, q: np.ndarray) -> np.ndarray:
    h, w = plane.shape
    blocks = plane.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)       # by, bx, r, c
    coef = np.einsum('ur,yxrc,vc->yxuv', _D, blocks, _D, optimize=True)
    quant = np.rint(coef / q.reshape(8, 8).astype(np.float64))
    return np.clip(quant, -32768, 32767).astype(np.int16).reshape(-1)


def synth_coefs(width: int, height: int, quality: int, subsampling: str, seed: int) -> CoefImage:
    """Make Y/Cb/Cr coefficient planes for a `width` x `height` image.

    subsampling: '4:4:4' (1x1 chroma) or '4:2:0' (2x2 chroma).  Grid sizes follow libjpeg:
    plane blocks = ceil(ceil(dim / samp) / 8), not MCU padded (jpeg.c:52-63).
    """
    if subsampling == '4:4:4':
        sw = sh = 1
    elif subsampling == '4:2:0':
        sw = sh = 2
    else:
        raise ValueError(f'unsupported subsampling {subsampling!r}')
    rgb = cartoon_image(width, height, seed)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    ycc = [0.299 * r + 0.587 * g + 0.114 * b - 128.0,
           -0.168736 * r - 0.331264 * g + 0.5 * b,
           0.5 * r - 0.418688 * g - 0.081312 * b]
    planes = []
    for idx, comp in enumerate(ycc):
        fw, fh = (1, 1) if idx == 0 else (sw, sh)
        cw, ch = -(-width // fw), -(-height // fh)
        if fw > 1 or fh > 1:
            pad = np.pad(comp, ((0, ch * fh - height), (0, cw * fw - width)), mode='edge')
            comp = pad.reshape(ch, fh, cw, fw).mean(axis=(1, 3))
        pw, ph = -(-cw // 8) * 8, -(-ch // 8) * 8
        comp = np.pad(comp, ((0, ph - ch), (0, pw - cw)), mode='edge')
        q = quant_table(quality, chroma=idx > 0)
        planes.append(Plane(w=pw, h=ph, w_samp=fw, h_samp=fh,
                            data=_block_dct_quantise(comp, q), quant=q))
    return CoefImage(width=width, height=height, planes=planes)


def random_coefs(plane_dims, samp, seed: int, amplitude: int = 40, qmax: int = 60) -> CoefImage:
    """Adversarial small planes for parity tests: random sparse coefficients, random tables.

    plane_dims: [(w,h)]*n in samples (multiples of 8); samp: [(w_samp,h_samp)]*n.
    """
    rng = np.random.default_rng(seed)
    planes = []
    for (w, h), (fw, fh) in zip(plane_dims, samp):
        n = (w // 8) * (h // 8) * 64
        data = rng.integers(-amplitude, amplitude + 1, size=n)
        decay = np.tile(1.0 / (1.0 + 0.35 * (np.arange(64) // 8 + np.arange(64) % 8) ** 1.5), n // 64)
        data = np.rint(data * decay) * (rng.random(n) < 0.6)
        data = data.astype(np.int16)
        data[::64] = rng.integers(-200, 201, size=n // 64).astype(np.int16)      # DC terms
        quant = rng.integers(1, qmax + 1, size=64).astype(np.uint16)
        planes.append(Plane(w=w, h=h, w_samp=fw, h_samp=fh, data=data, quant=quant))
    fw_ = max(p.w * p.w_samp for p in planes)
    fh_ = max(p.h * p.h_samp for p in planes)
    return CoefImage(width=fw_, height=fh_, planes=planes)


def tile_coefs(base: CoefImage, nx: int, ny: int, width: int, height: int) -> CoefImage:
    """A `width` x `height` image whose coefficient planes are `base`'s blocks repeated nx x ny times
    and cropped to the block grid the larger image needs — a cheap way to get a large synthetic
    frame (seconds instead of the half minute synth_coefs spends on an 8K cartoon)."""
    planes = []
    for p in base.planes:
        cw, ch = -(-width // p.w_samp), -(-height // p.h_samp)
        bw, bh = -(-cw // 8), -(-ch // 8)
        blocks = p.data.reshape(p.h // 8, p.w // 8, 64)
        big = np.tile(blocks, (ny, nx, 1))
        if big.shape[0] < bh or big.shape[1] < bw:
            raise ValueError('the tiled base image does not cover the requested size')
        planes.append(Plane(w=bw * 8, h=bh * 8, w_samp=p.w_samp, h_samp=p.h_samp,
                            data=np.ascontiguousarray(big[:bh, :bw]).reshape(-1), quant=p.quant))
    return CoefImage(width=width, height=height, planes=planes)
