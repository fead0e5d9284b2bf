"""Oracle (test infrastructure only — nothing under fastdet_b200/ may import this): the image decode at the top of the
reference's perform(), server/detector.py:128-133::

    img = Image.open(io.BytesIO(data)); ...; a = np.array(img)

`decode_reference` is those lines verbatim.  The pixels come from libjpeg-turbo behind Pillow (a third-party dependency
of the reference, not in /root/reference; this image: Pillow 12.2 / libjpeg-turbo 3.x, jpeglib 6.2 API) with the
decompressor defaults Pillow leaves in place: JDCT_ISLOW, do_fancy_upsampling = TRUE, no merged upsampling.  The rest of
this file restates that published algorithm in numpy (ITU-T T.81 for the bit stream; jidctint.c `jpeg_idct_islow`,
jdsample.c `h2v1_fancy_upsample` / `h2v2_fancy_upsample`, jdcolor.c `build_ycc_rgb_table` / `ycc_rgb_convert` for the
sample reconstruction) so the native decoder can be checked stage by stage:

    native Huffman coefficients  ==  entropy_decode()           (CPU test, through fd_jpeg_coefficients)
    reconstruct(coefficients)    ==  decode_reference()         (CPU test: pins the restatement against Pillow)
    fd_decode_jpeg on the device ==  decode_reference()         (GPU test, bit-exact)

Parity pinned: tests/test_jpeg.py checks the restatement against Pillow on generated streams (4:4:4 / 4:2:2 / 4:2:0,
several qualities, restart intervals, odd sizes) and against the committed fixture tests/golden/jpeg.npz.
"""
import io

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
                   46, 53, 60, 61, 54, 47, 55, 62, 63])


def decode_reference(data: bytes) -> np.ndarray:
    """server/detector.py:128-133 (without the size check): what the reference feeds on."""
    from PIL import Image
    return np.array(Image.open(io.BytesIO(data)))


# ---------------------------------------------------------------------------------------------- bit stream (T.81)
def parse(data: bytes) -> dict:
    """Markers up to SOS of a baseline, 3-component, single-scan JPEG (T.81 Annex B)."""
    assert data[:2] == b'\xff\xd8', 'no SOI'
    pos = 2
    qt, huff = {}, {}
    hdr = {'restart_interval': 0}
    while True:
        assert data[pos] == 0xFF
        while data[pos] == 0xFF:
            pos += 1
        m = data[pos]
        pos += 1
        length = int.from_bytes(data[pos:pos + 2], 'big')
        seg = data[pos + 2:pos + length]
        pos += length
        if m in (0xC0, 0xC1):
            assert seg[0] == 8
            hdr['height'], hdr['width'] = int.from_bytes(seg[1:3], 'big'), int.from_bytes(seg[3:5], 'big')
            assert seg[5] == 3
            hdr['comps'] = [(seg[6 + 3 * c], seg[7 + 3 * c] >> 4, seg[7 + 3 * c] & 15, seg[8 + 3 * c]) for c in range(3)]
        elif m == 0xDB:
            o = 0
            while o < len(seg):
                pq, tq = seg[o] >> 4, seg[o] & 15
                if pq:
                    vals = [int.from_bytes(seg[o + 1 + 2 * i:o + 3 + 2 * i], 'big') for i in range(64)]
                else:
                    vals = list(seg[o + 1:o + 65])
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = vals
                qt[tq] = t.reshape(8, 8)
                o += 1 + 64 * (pq + 1)
        elif m == 0xC4:
            o = 0
            while o < len(seg):
                tc, th = seg[o] >> 4, seg[o] & 15
                bits = list(seg[o + 1:o + 17])
                total = sum(bits)
                vals = list(seg[o + 17:o + 17 + total])
                # T.81 Annex C: canonical codes in order of increasing length
                table, code, k = {}, 0, 0
                for length_ in range(1, 17):
                    for _ in range(bits[length_ - 1]):
                        table[(length_, code)] = vals[k]
                        code += 1
                        k += 1
                    code <<= 1
                huff[(tc, th)] = table
                o += 17 + total
        elif m == 0xDD:
            hdr['restart_interval'] = int.from_bytes(seg[:2], 'big')
        elif m == 0xDA:
            assert seg[0] == 3
            hdr['scan'] = [(seg[1 + 2 * c], seg[2 + 2 * c] >> 4, seg[2 + 2 * c] & 15) for c in range(3)]
            hdr['scan_off'] = pos
            break
        else:
            assert not (0xC2 <= m <= 0xCF and m != 0xC8), 'not a baseline stream'
    hdr['qt'], hdr['huff'] = qt, huff
    return hdr


class _Bits:
    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def bit(self):
        if self.n == 0:
            b = self.d[self.p]
            self.p += 1
            if b == 0xFF:
                nxt = self.d[self.p]
                assert nxt == 0, 'marker inside entropy-coded data'
                self.p += 1
            self.acc, self.n = b, 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self, expect):
        self.n = 0
        assert self.d[self.p] == 0xFF and self.d[self.p + 1] == 0xD0 + expect, 'restart marker missing'
        self.p += 2


def _symbol(br, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        v = table.get((length, code))
        if v is not None:
            return v
    raise AssertionError('bad Huffman code')


def _extend(v, s):  # T.81 Figure F.12
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def entropy_decode(data: bytes):
    """Huffman decode of the interleaved scan (T.81 F.2.2, E.2).  Returns (hdr, [quantised coefficients of component c
    as int16 [block rows, block cols, 8, 8]]) with planes padded to whole MCUs.  Pure-Python loops: small images."""
    hdr = parse(data)
    hs, vs = hdr['comps'][0][1], hdr['comps'][0][2]
    assert all(c[1] == 1 and c[2] == 1 for c in hdr['comps'][1:])
    mx_n = -(-hdr['width'] // (8 * hs))
    my_n = -(-hdr['height'] // (8 * vs))
    planes = [np.zeros((my_n * vs, mx_n * hs, 64), np.int16), np.zeros((my_n, mx_n, 64), np.int16),
              np.zeros((my_n, mx_n, 64), np.int16)]
    br = _Bits(data, hdr['scan_off'])
    pred = [0, 0, 0]
    ri, rst = hdr['restart_interval'], 0
    mcu = 0
    for my in range(my_n):
        for mx in range(mx_n):
            if ri and mcu and mcu % ri == 0:
                br.restart(rst)
                rst = (rst + 1) & 7
                pred = [0, 0, 0]
            for c in range(3):
                _, td, ta = hdr['scan'][c]
                h_, v_ = (hs, vs) if c == 0 else (1, 1)
                for v in range(v_):
                    for h in range(h_):
                        blk = planes[c][my * v_ + v, mx * h_ + h]
                        s = _symbol(br, hdr['huff'][(0, td)])
                        if s:
                            pred[c] += _extend(br.bits(s), s)
                        blk[0] = pred[c]
                        k = 1
                        while k < 64:
                            rs = _symbol(br, hdr['huff'][(1, ta)])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                            k += 1
            mcu += 1
    return hdr, [p.reshape(p.shape[0], p.shape[1], 8, 8) for p in planes]


# ---------------------------------------------------------------------------------------------- samples (libjpeg)
CONST_BITS, PASS1_BITS = 13, 2
FIX = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
           f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _wrap16(v):
    return ((v + 32768) & 0xFFFF) - 32768


def _wrap32(v):
    return ((v + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)


def _idct_1d(x):
    """One 8-point pass of jpeg_idct_islow (jidctint.c) along the last axis, before the descale.  The _wrap16 calls are
    no-ops for the streams a sane encoder writes; for damaged ones they reproduce the 16-bit paddw/psubw of the SIMD
    versions (jidctint-sse2/avx2.asm) that libjpeg-turbo — hence Pillow — actually executes."""
    F = FIX
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * F['f0_541']
    tmp2 = z1 + z3 * (-F['f1_847'])
    tmp3 = z1 + z2 * F['f0_765']
    z2, z3 = x[..., 0], x[..., 4]
    tmp0 = _wrap16(z2 + z3) << CONST_BITS
    tmp1 = _wrap16(z2 - z3) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, _wrap16(tmp0 + tmp2), _wrap16(tmp1 + tmp3)
    z5 = (z3 + z4) * F['f1_175']
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F['f0_298'], tmp1 * F['f2_053'], tmp2 * F['f3_072'], tmp3 * F['f1_501']
    z1, z2, z3, z4 = z1 * -F['f0_899'], z2 * -F['f2_562'], z3 * -F['f1_961'], z4 * -F['f0_390']
    z3, z4 = z3 + z5, z4 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                     tmp10 - tmp3], axis=-1)


def _descale(x, n):
    return _wrap32(x + (1 << (n - 1))) >> n  # 32-bit lanes


def idct_islow(coef, quant):
    """coef int16 [..., 8, 8] (row-major), quant [8, 8] -> samples u8 [..., 8, 8]."""
    x = _wrap16(coef.astype(np.int64) * quant.astype(np.int64))  # pmullw
    # pass 1: columns (transform along axis -2); packssdw saturates the workspace to 16 bits
    ws = np.clip(_descale(_idct_1d(np.swapaxes(x, -1, -2)), CONST_BITS - PASS1_BITS), -32768, 32767)
    ws = np.swapaxes(ws, -1, -2)
    # the SIMD code's shortcut when rows 1..7 of the (quantised) block are all zero: each column is its DC term shifted
    # left by PASS1_BITS in a 16-bit lane (psllw: wraps where the full pass would have saturated)
    dc_only = ~np.any(coef[..., 1:, :] != 0, axis=(-1, -2))
    short = np.broadcast_to(_wrap16(x[..., :1, :] << PASS1_BITS), x.shape)
    ws = np.where(dc_only[..., None, None], short, ws)
    # pass 2: rows; range limit = clamp(x + 128): the table of the C code for every value a valid stream produces, the
    # saturating packs of the SIMD code beyond (the C table would wrap around there)
    out = _descale(_idct_1d(ws), CONST_BITS + PASS1_BITS + 3)
    return (np.clip(out, -128, 127) + 128).astype(np.uint8)


def planes_from_coefficients(hdr, coefs):
    """De-quantise + IDCT every block; returns the three padded sample planes."""
    out = []
    for c in range(3):
        q = hdr['qt'][hdr['comps'][c][3]]
        s = idct_islow(coefs[c], q)  # [bh, bw, 8, 8]
        bh, bw = s.shape[:2]
        out.append(s.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8))
    return out


def h2v1_fancy(p):
    """jdsample.c h2v1_fancy_upsample on an int array [rows, n]: 3/4 nearer + 1/4 further, ends copied."""
    p = p.astype(np.int64)
    n = p.shape[1]
    out = np.empty((p.shape[0], 2 * n), np.int64)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def h2v2_fancy(p):
    """jdsample.c h2v2_fancy_upsample on [rows, n]: vertical 3:1 sums first, then the horizontal 3:1 with the +8 / +7
    rounding pair; the context rows above the first and below the last row are those rows themselves (jdmainct.c)."""
    p = p.astype(np.int64)
    rows, n = p.shape
    up = np.concatenate([p[:1], p[:-1]], axis=0)
    down = np.concatenate([p[1:], p[-1:]], axis=0)
    colsum = np.empty((2 * rows, n), np.int64)
    colsum[0::2] = 3 * p + up
    colsum[1::2] = 3 * p + down
    left = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
    right = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
    out = np.empty((2 * rows, 2 * n), np.int64)
    out[:, 0::2] = (3 * colsum + left + 8) >> 4
    out[:, 1::2] = (3 * colsum + right + 7) >> 4
    out[:, 0] = (4 * colsum[:, 0] + 8) >> 4
    out[:, -1] = (4 * colsum[:, -1] + 7) >> 4
    return out


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c: 16.16 fixed-point tables; G sums the two chroma products (ONE_HALF folded into the Cb table) before
    the shift."""
    y, cb, cr = y.astype(np.int64), cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    r = y + ((91881 * cr + 32768) >> 16)
    g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
    b = y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def reconstruct(hdr, coefs) -> np.ndarray:
    """Quantised coefficients -> RGB u8 [height, width, 3], as libjpeg's default decompressor."""
    w, h = hdr['width'], hdr['height']
    hs, vs = hdr['comps'][0][1], hdr['comps'][0][2]
    Y, CB, CR = planes_from_coefficients(hdr, coefs)
    cw, ch = -(-w // hs), -(-h // vs)  # downsampled_width / height of the chroma components
    if (hs, vs) == (1, 1):
        up = lambda p: p
    elif (hs, vs) == (2, 1):
        up = lambda p: h2v1_fancy(p[:ch, :cw])
    elif (hs, vs) == (2, 2):
        up = lambda p: h2v2_fancy(p[:ch, :cw])
    else:
        raise AssertionError('sampling outside the device path')
    return ycc_to_rgb(Y[:h, :w], up(CB)[:h, :w], up(CR)[:h, :w])


def decode(data: bytes) -> np.ndarray:
    """The whole restatement: bytes -> RGB."""
    hdr, coefs = entropy_decode(data)
    return reconstruct(hdr, coefs)
