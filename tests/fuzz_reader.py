"""Mutation fuzzing of the JPEG coefficient reader (jpeg2png_b200/cli/jpeg_reader.c), run as a
separate process by tests/test_codecs.py so that a crash shows up as a failed test, not a dead
pytest.  Every mutated file must either parse (with sane plane sizes) or be rejected with a message."""
import ctypes as C
import io
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpeg2png_b200 import abi, synth  # noqa: E402

CLI_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'jpeg2png_b200', 'cli')


class Jpeg(C.Structure):
    _fields_ = [('w', C.c_uint), ('h', C.c_uint), ('coefs', abi.Coef * 3)]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = C.CDLL(os.path.join(CLI_DIR, 'libj2pcodecs.so'))
    lib.j2p_read_jpeg_mem.restype = C.c_int
    lib.j2p_read_jpeg_mem.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Jpeg), C.c_char_p, C.c_size_t]
    rng = np.random.default_rng(seed)
    seeds = []
    for (w, h, q, ss, prog, opt) in [(64, 48, 75, '4:4:4', False, False), (72, 40, 20, '4:2:0', False, True),
                                     (56, 64, 50, '4:2:0', True, False), (48, 48, 90, '4:2:2', True, True)]:
        rgb = synth.cartoon_image(w, h, 3).astype(np.uint8)
        buf = io.BytesIO()
        Image.fromarray(rgb, 'RGB').save(buf, 'JPEG', quality=q, subsampling=ss, progressive=prog, optimize=opt)
        seeds.append(buf.getvalue())
    parsed = rejected = 0
    for it in range(n):
        data = bytearray(seeds[it % len(seeds)])
        kind = rng.integers(0, 5)
        if kind == 0:                                   # truncate
            data = data[:int(rng.integers(2, len(data)))]
        elif kind == 1:                                 # flip a few bytes anywhere
            for _ in range(int(rng.integers(1, 8))):
                data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
        elif kind == 2:                                 # corrupt the header region (markers, lengths, tables)
            for _ in range(int(rng.integers(1, 6))):
                data[int(rng.integers(2, min(len(data), 700)))] = int(rng.integers(0, 256))
        elif kind == 3:                                 # duplicate or drop a chunk
            a, b = sorted(int(x) for x in rng.integers(2, len(data), 2))
            data = data[:a] + data[b:] if rng.random() < 0.5 else data[:b] + data[a:b] + data[b:]
        else:                                           # insert marker-like garbage
            pos = int(rng.integers(2, len(data)))
            data[pos:pos] = bytes([0xFF, int(rng.integers(0xC0, 0xFF)), 0, int(rng.integers(0, 40))])
        j = Jpeg()
        err = C.create_string_buffer(256)
        rc = lib.j2p_read_jpeg_mem(bytes(data), len(data), C.byref(j), err, 256)
        if rc == 0:
            parsed += 1
            assert 0 < j.w <= 65535 and 0 < j.h <= 65535
            for c in j.coefs:
                assert c.w % 8 == 0 and c.h % 8 == 0 and c.w > 0 and c.h > 0 and 1 <= c.w_samp <= 4 and 1 <= c.h_samp <= 4
                np.ctypeslib.as_array(c.data, shape=(c.w * c.h,)).sum()          # touch every coefficient
                abi.free_ptr(c.data)
        else:
            rejected += 1
            assert err.value, 'rejected without a message'
    print(f'fuzz_reader: {n} mutated files, {parsed} parsed, {rejected} rejected, no crash')


if __name__ == '__main__':
    main()
