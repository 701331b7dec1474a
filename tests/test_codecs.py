"""CPU tests of the self-contained codecs behind the CLI: the JPEG
coefficient reader (replaces reference jpeg.c:22-80 / libjpeg) and the PNG writer (replaces
png.c:20-78 / libpng).  Ground truth: Pillow's libjpeg-turbo decode of the same files, the IJG
quantisation tables, and the oracle's restatement of the reference colour conversion."""
import ctypes as C
import io
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest
from PIL import Image

from jpeg2png_b200 import abi, synth
from tests import helpers as H

CLI_DIR = os.path.join(H.ROOT, 'jpeg2png_b200', 'cli')


class Jpeg(C.Structure):
    _fields_ = [('w', C.c_uint), ('h', C.c_uint), ('coefs', abi.Coef * 3)]


@pytest.fixture(scope='module')
def codecs():
    subprocess.run(['make', '-C', CLI_DIR, 'libj2pcodecs.so'], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(CLI_DIR, 'libj2pcodecs.so'))
    lib.j2p_read_jpeg_mem.restype = C.c_int
    lib.j2p_read_jpeg_mem.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Jpeg), C.c_char_p, C.c_size_t]
    lib.j2p_ycc_to_rgb.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t]
    return lib


def read_jpeg(lib, data: bytes):
    j = Jpeg()
    err = C.create_string_buffer(256)
    rc = lib.j2p_read_jpeg_mem(data, len(data), C.byref(j), err, 256)
    if rc != 0:
        return None, err.value.decode()
    planes = []
    for c in j.coefs:
        n = c.w * c.h
        d = np.ctypeslib.as_array(c.data, shape=(n,)).copy()
        planes.append(synth.Plane(w=c.w, h=c.h, w_samp=c.w_samp, h_samp=c.h_samp, data=d, quant=np.array(list(c.quant_table), np.uint16)))
        abi.free_ptr(c.data)
    return synth.CoefImage(width=j.w, height=j.h, planes=planes), ''


def make_jpeg(w, h, quality, subsampling, progressive=False, optimize=False, seed=1):
    rgb = synth.cartoon_image(w, h, seed).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(rgb, 'RGB').save(buf, 'JPEG', quality=quality, subsampling=subsampling, progressive=progressive, optimize=optimize)
    return buf.getvalue()


def decode_like_reference(img):
    """conventional decode + nearest-neighbour upsampling, cropped to the image: what the reference
    would write as PNG with 0 iterations (jpeg.c:83-92, compute.c:295-302)."""
    planes = H.decode_planes(img)
    out = []
    for p, f in zip(img.planes, planes):
        yy = np.minimum(np.arange(img.height) // p.h_samp, p.h - 1)
        xx = np.minimum(np.arange(img.width) // p.w_samp, p.w - 1)
        out.append(f[yy][:, xx])
    return out


@pytest.mark.parametrize('w,h,q,ss,prog,opt', [
    (64, 48, 75, '4:4:4', False, False),
    (200, 120, 10, '4:2:0', False, False),
    (73, 59, 50, '4:2:0', False, True),        # odd size, optimised Huffman tables
    (96, 80, 90, '4:2:2', False, False),
    (160, 96, 30, '4:4:4', True, False),       # progressive: DC/AC first + refinement scans
    (131, 77, 60, '4:2:0', True, True),
])
def test_reader_matches_pillow(codecs, w, h, q, ss, prog, opt):
    data = make_jpeg(w, h, q, ss, prog, opt, seed=w + h)
    img, err = read_jpeg(codecs, data)
    assert img is not None, err
    assert (img.width, img.height) == (w, h)
    fw, fh = {'4:4:4': (1, 1), '4:2:0': (2, 2), '4:2:2': (2, 1)}[ss]
    assert [(p.w_samp, p.h_samp) for p in img.planes] == [(1, 1), (fw, fh), (fw, fh)]
    for k, p in enumerate(img.planes):      # block grids as libjpeg reports them, not MCU padded (jpeg.c:52-53)
        sw, sh = (1, 1) if k == 0 else (fw, fh)
        assert p.w == -(-(-(-w // sw)) // 8) * 8 and p.h == -(-(-(-h // sh)) // 8) * 8
    # quantisation tables in natural order = the IJG tables Pillow scaled
    assert (img.planes[0].quant == synth.quant_table(q, chroma=False)).all()
    assert (img.planes[1].quant == synth.quant_table(q, chroma=True)).all()
    # pixels: our coefficients through the reference-style decode vs libjpeg-turbo's decode
    ours = decode_like_reference(img)
    im = Image.open(io.BytesIO(data))
    im.draft('YCbCr', im.size)
    theirs = np.asarray(im.convert('YCbCr') if im.mode != 'YCbCr' else im).astype(np.float64)
    y_ours = np.clip(np.rint(ours[0] + 128.0), 0, 255)
    assert np.abs(y_ours - theirs[..., 0]).max() <= 1.0          # luma: same block grid, IDCT implementations differ by <= 1
    if ss == '4:4:4':
        for c in (1, 2):
            assert np.abs(np.clip(np.rint(ours[c] + 128.0), 0, 255) - theirs[..., c]).max() <= 1.0


def test_reader_rejections(codecs):
    rgb = synth.cartoon_image(32, 32, 3).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(rgb, 'RGB').convert('L').save(buf, 'JPEG')
    img, err = read_jpeg(codecs, buf.getvalue())
    assert img is None and err == 'only 3 component jpegs are supported'            # jpeg.c:34
    good = make_jpeg(64, 64, 50, '4:2:0')
    img, err = read_jpeg(codecs, good[:len(good) // 2])                              # truncated: decodes what is there or fails, never crashes
    assert img is not None or 'corrupt' in err
    img, err = read_jpeg(codecs, b'not a jpeg at all')
    assert img is None and 'SOI' in err
    bad = bytearray(good)
    dqt = bad.find(b'\xff\xdb')
    bad[dqt + 5] = 0                                                                 # zero entry in the first table
    img, err = read_jpeg(codecs, bytes(bad))
    assert img is None and err == 'invalid quantization table'                      # jpeg.c:43


def test_rgb_conversion_matches_reference_restatement(codecs):
    """png.c:39-62: product conversion == oracle restatement, 8 and 16 bit, planes with different strides."""
    ora = H.load_oracle()
    rng = np.random.default_rng(7)
    w, h = 37, 21
    y = (rng.normal(128, 80, (h, w + 3))).astype(np.float32)
    cb = (rng.normal(0, 60, (h, w + 8))).astype(np.float32)
    cr = (rng.normal(0, 60, (h, w))).astype(np.float32)
    for bits in (8, 16):
        n = w * h * 3 * (bits // 8)
        a, b = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        codecs.j2p_ycc_to_rgb(w, h, bits, y.ctypes.data, w + 3, cb.ctypes.data, w + 8, cr.ctypes.data, w, a.ctypes.data, w * 3 * (bits // 8))
        ora.oracle_ycc_to_rgb(w, h, bits, y.ctypes.data, w + 3, cb.ctypes.data, w + 8, cr.ctypes.data, w, b.ctypes.data)
        assert (a == b).all()


def test_png_writer_roundtrip(codecs, tmp_path):
    codecs.j2p_write_png.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fclose.argtypes = [C.c_void_p]
    rng = np.random.default_rng(11)
    w, h = 50, 33
    y = rng.normal(128, 70, (h, w)).astype(np.float32)
    cb = rng.normal(0, 50, (h, w)).astype(np.float32)
    cr = rng.normal(0, 50, (h, w)).astype(np.float32)
    for bits in (8, 16):
        path = str(tmp_path / f'out{bits}.png')
        f = libc.fopen(path.encode(), b'wb')
        assert codecs.j2p_write_png(f, w, h, bits, y.ctypes.data, w, cb.ctypes.data, w, cr.ctypes.data, w) == 0
        libc.fclose(f)
        want = np.zeros(w * h * 3 * (bits // 8), np.uint8)
        H.load_oracle().oracle_ycc_to_rgb(w, h, bits, y.ctypes.data, w, cb.ctypes.data, w, cr.ctypes.data, w, want.ctypes.data)
        raw = open(path, 'rb').read()
        assert raw[:8] == b'\x89PNG\r\n\x1a\n'
        # walk the chunks, check CRCs, inflate IDAT
        pos, idat, ihdr = 8, b'', None
        while pos < len(raw):
            ln = int.from_bytes(raw[pos:pos + 4], 'big')
            typ, body = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + ln]
            assert zlib.crc32(typ + body) == int.from_bytes(raw[pos + 8 + ln:pos + 12 + ln], 'big')
            if typ == b'IHDR':
                ihdr = body
            if typ == b'IDAT':
                idat += body
            pos += 12 + ln
        assert int.from_bytes(ihdr[:4], 'big') == w and int.from_bytes(ihdr[4:8], 'big') == h and ihdr[8] == bits and ihdr[9] == 2
        rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * 3 * (bits // 8))
        assert (rows[:, 0] == 0).all() and (rows[:, 1:].reshape(-1) == want).all()
        if bits == 8:
            assert (np.asarray(Image.open(path)).reshape(-1) == want).all()


def test_png_writer_large_image_pieced_deflate(codecs, tmp_path):
    """> 2 MB of scanlines: the IDAT stream is assembled from independently deflated 1 MB pieces
    (png_writer.c); it must still be ONE valid zlib stream with the right Adler-32 and pixels."""
    codecs.j2p_write_png.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fclose.argtypes = [C.c_void_p]
    rng = np.random.default_rng(21)
    w, h = 1203, 911                                    # 3.3 MB at 8 bit, 6.6 MB at 16 bit; odd sizes on purpose
    yy, xx = np.mgrid[0:h, 0:w]
    y = (128 + 90 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + rng.normal(0, 3, (h, w))).astype(np.float32)
    cb = (40 * np.sign(np.sin(xx / 50.0))).astype(np.float32)
    cr = rng.normal(0, 50, (h, w)).astype(np.float32)
    for bits in (8, 16):
        path = str(tmp_path / f'big{bits}.png')
        f = libc.fopen(path.encode(), b'wb')
        assert codecs.j2p_write_png(f, w, h, bits, y.ctypes.data, w, cb.ctypes.data, w, cr.ctypes.data, w) == 0
        libc.fclose(f)
        want = np.zeros(w * h * 3 * (bits // 8), np.uint8)
        H.load_oracle().oracle_ycc_to_rgb(w, h, bits, y.ctypes.data, w, cb.ctypes.data, w, cr.ctypes.data, w, want.ctypes.data)
        raw = open(path, 'rb').read()
        pos, idat = 8, b''
        while pos < len(raw):
            ln = int.from_bytes(raw[pos:pos + 4], 'big')
            typ, body = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + ln]
            assert zlib.crc32(typ + body) == int.from_bytes(raw[pos + 8 + ln:pos + 12 + ln], 'big')
            if typ == b'IDAT':
                idat += body
            pos += 12 + ln
        d = zlib.decompressobj()
        rows = d.decompress(idat)
        assert d.eof and d.unused_data == b'', 'not exactly one complete zlib stream'
        rows = np.frombuffer(rows, np.uint8).reshape(h, 1 + w * 3 * (bits // 8))
        # images of this size carry the Up filter (type 2) on every row but the first: undo it
        assert rows[0, 0] == 0 and (rows[1:, 0] == 2).all()
        pixels = np.cumsum(rows[:, 1:].astype(np.uint32), axis=0).astype(np.uint8)      # modulo 256
        assert (pixels.reshape(-1) == want).all()
        if bits == 8:
            assert (np.asarray(Image.open(path)).reshape(-1) == want).all()
        assert len(idat) < rows.size                       # (one plane of this image is white noise: no ratio to expect)


@pytest.mark.parametrize('w,h,sampling,ri', [
    (64, 48, [(1, 1), (1, 1), (1, 1)], 0),
    (70, 50, [(2, 2), (1, 1), (1, 1)], 0),          # 4:2:0, MCU padding dropped on the right and bottom
    (70, 50, [(2, 2), (1, 1), (1, 1)], 3),          # restart interval
    (95, 33, [(2, 1), (1, 1), (1, 1)], 5),          # 4:2:2 + restarts
    (48, 40, [(4, 1), (2, 1), (1, 1)], 2),          # unusual factors: chroma planes with different sampling (w_samp 2 and 4)
    (40, 72, [(1, 2), (1, 1), (1, 1)], 0),          # vertical-only subsampling
])
def test_reader_recovers_known_coefficients_exactly(codecs, w, h, sampling, ri):
    """Files written by tests/jpeg_synth.py from KNOWN coefficients: exact round trip, natural
    order, libjpeg's unpadded block grids, every sampling layout the reference accepts."""
    from tests import jpeg_synth
    planes, quants = jpeg_synth.random_planes(w, h, sampling, seed=w * 3 + h)
    data = jpeg_synth.encode_baseline(w, h, sampling, planes, quants, restart_interval=ri)
    assert Image.open(io.BytesIO(data)).size == (w, h)          # Pillow accepts the file too
    maxh, maxv = max(s[0] for s in sampling), max(s[1] for s in sampling)
    img, err = read_jpeg(codecs, data)
    if any((h // (maxv // s[1]) + 7) // 8 != -(-(-(-h * s[1] // maxv)) // 8) or (w // (maxh // s[0]) + 7) // 8 != -(-(-(-w * s[0] // maxh)) // 8)
           for s in sampling):
        assert img is None and 'jpeg invalid coef' in err      # the reference's own size check (jpeg.c:59-64)
        return
    assert img is not None, err
    for c, p in enumerate(img.planes):
        wb, hb = -(-(-(-w * sampling[c][0] // maxh)) // 8), -(-(-(-h * sampling[c][1] // maxv)) // 8)
        assert (p.w, p.h) == (wb * 8, hb * 8)
        assert (p.w_samp, p.h_samp) == (maxh // sampling[c][0], maxv // sampling[c][1])
        want = planes[c][:hb, :wb].reshape(-1)
        assert (p.data == want).all()
        assert (p.quant == quants[c]).all()


def test_reader_survives_mutated_files(codecs):
    """Mutation fuzzing in a separate process (tests/fuzz_reader.py): truncated, bit-flipped and
    spliced files are parsed or rejected with a message, never a crash.  (The same corpus was run
    under AddressSanitizer + UBSan while the reader was hardened.)"""
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fuzz_reader.py'), '500', '5'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'no crash' in r.stdout
