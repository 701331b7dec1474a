"""A minimal baseline JPEG *encoder from given coefficients* (test infrastructure): lets the reader
tests use files whose quantised coefficients are known exactly, with arbitrary sampling factors and
restart intervals — things Pillow cannot be asked to produce.  Huffman tables are lifted from a
Pillow-written file (the Annex K tables)."""
import io

import numpy as np
from PIL import Image

ZZ = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
      35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def standard_huffman_tables():
    """{(class, id): (bits[16], vals)} parsed from a Pillow JPEG."""
    buf = io.BytesIO()
    Image.new('RGB', (16, 16), (120, 30, 200)).save(buf, 'JPEG', quality=75)
    data = buf.getvalue()
    tables, pos = {}, 2
    while pos < len(data):
        assert data[pos] == 0xFF
        m = data[pos + 1]
        if m == 0xDA:
            break
        ln = int.from_bytes(data[pos + 2:pos + 4], 'big')
        if m == 0xC4:
            s = data[pos + 4:pos + 2 + ln]
            while s:
                tc, th = s[0] >> 4, s[0] & 15
                bits = list(s[1:17])
                n = sum(bits)
                tables[(tc, th)] = (bits, list(s[17:17 + n]))
                s = s[17 + n:]
        pos += 2 + ln
    return tables


def _codes(bits, vals):
    out, code, k = {}, 0, 0
    for l in range(1, 17):
        for _ in range(bits[l - 1]):
            out[vals[k]] = (code, l)
            code += 1
            k += 1
        code <<= 1
    return out


class _Bits:
    def __init__(self):
        self.out, self.acc, self.n = bytearray(), 0, 0

    def put(self, v, n):
        self.acc = (self.acc << n) | (v & ((1 << n) - 1))
        self.n += n
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def _size_bits(v):
    if v == 0:
        return 0, 0
    s = int(abs(v)).bit_length()
    return s, (v if v > 0 else v + (1 << s) - 1)


def encode_baseline(width, height, sampling, planes, quants, restart_interval=0):
    """sampling: [(h,v)]*3; planes[c]: int array [padded_blocks_y][padded_blocks_x][64] NATURAL order,
    MCU-padded grid; quants[c]: 64 natural-order values.  Returns the JPEG bytes."""
    tabs = standard_huffman_tables()
    dc = [_codes(*tabs[(0, 0)]), _codes(*tabs[(0, 1)])]
    ac = [_codes(*tabs[(1, 0)]), _codes(*tabs[(1, 1)])]
    maxh, maxv = max(h for h, _ in sampling), max(v for _, v in sampling)
    mcux, mcuy = -(-width // (8 * maxh)), -(-height // (8 * maxv))
    out = bytearray(b'\xff\xd8')
    for c, q in enumerate(quants[:2] if np.array_equal(quants[1], quants[2]) else quants):
        out += b'\xff\xdb' + (67).to_bytes(2, 'big') + bytes([c]) + bytes(int(q[ZZ[k]]) for k in range(64))
    tq = [0, 1, 1] if np.array_equal(quants[1], quants[2]) else [0, 1, 2]
    out += b'\xff\xc0' + (17).to_bytes(2, 'big') + b'\x08' + height.to_bytes(2, 'big') + width.to_bytes(2, 'big') + b'\x03'
    for c in range(3):
        out += bytes([c + 1, (sampling[c][0] << 4) | sampling[c][1], tq[c]])
    for (tc, th), (bits, vals) in tabs.items():
        out += b'\xff\xc4' + (19 + len(vals)).to_bytes(2, 'big') + bytes([(tc << 4) | th]) + bytes(bits) + bytes(vals)
    if restart_interval:
        out += b'\xff\xdd\x00\x04' + restart_interval.to_bytes(2, 'big')
    out += b'\xff\xda' + (12).to_bytes(2, 'big') + b'\x03' + bytes([1, 0x00, 2, 0x11, 3, 0x11]) + b'\x00\x3f\x00'
    bw = _Bits()
    pred = [0, 0, 0]
    n_mcu, rst = 0, 0
    for my in range(mcuy):
        for mx in range(mcux):
            if restart_interval and n_mcu and n_mcu % restart_interval == 0:
                bw.flush()
                out += bw.out + bytes([0xFF, 0xD0 + (rst & 7)])
                bw = _Bits()
                rst += 1
                pred = [0, 0, 0]
            for c in range(3):
                h, v = sampling[c]
                t = 0 if c == 0 else 1
                for y in range(v):
                    for x in range(h):
                        b = planes[c][my * v + y][mx * h + x]
                        s, bits = _size_bits(int(b[0]) - pred[c])
                        pred[c] = int(b[0])
                        bw.put(*dc[t][s])
                        if s:
                            bw.put(bits, s)
                        run = 0
                        last = max([k for k in range(1, 64) if b[ZZ[k]] != 0], default=0)
                        for k in range(1, last + 1):
                            val = int(b[ZZ[k]])
                            if val == 0:
                                run += 1
                                continue
                            while run > 15:
                                bw.put(*ac[t][0xF0])
                                run -= 16
                            s, bits = _size_bits(val)
                            bw.put(*ac[t][(run << 4) | s])
                            bw.put(bits, s)
                            run = 0
                        if last < 63:
                            bw.put(*ac[t][0x00])
            n_mcu += 1
    bw.flush()
    out += bw.out + b'\xff\xd9'
    return bytes(out)


def random_planes(width, height, sampling, seed):
    rng = np.random.default_rng(seed)
    maxh, maxv = max(h for h, _ in sampling), max(v for _, v in sampling)
    mcux, mcuy = -(-width // (8 * maxh)), -(-height // (8 * maxv))
    planes = []
    for h, v in sampling:
        p = rng.integers(-60, 61, size=(mcuy * v, mcux * h, 64))
        p = (p * (rng.random(p.shape) < 0.25)).astype(np.int16)
        p[..., 0] = rng.integers(-400, 401, size=p.shape[:2])
        planes.append(p)
    quants = [rng.integers(1, 100, 64), rng.integers(1, 100, 64)]
    quants.append(quants[1])
    return planes, quants
