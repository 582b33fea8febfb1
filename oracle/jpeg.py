"""TEST INFRASTRUCTURE - CPU restatement (pure Python + numpy) of the baseline JPEG decode behind the reference's
``Image.open(path).convert('RGB')`` (utils/dataloader.py:34 through ImageFolder's loader,
utils/image_to_graph/image_to_graph_optimized.py:65-68, utils/inference.py:47).

The algorithm lives in a third-party dependency that is absent from /root/reference: Pillow (requirements.txt, unpinned;
12.2.0 here) decodes through libjpeg-turbo (bundled with the wheel; ``PIL.features.version('jpg')`` = 3.x here) with the
library defaults.  Restated from the published sources, stage by stage:

  * marker parsing, canonical Huffman tables, sequential entropy decode with DC prediction, restart intervals and byte
    unstuffing (jdmarker.c, jdhuff.c);
  * dequantisation + the accurate integer inverse DCT, ``jpeg_idct_islow`` (jidctint.c: CONST_BITS 13, PASS1_BITS 2);
  * "fancy" chroma upsampling (jdsample.c: h2v1 / h2v2 triangle filters with their +1/+2 and +8/+7 rounding terms, the
    first / last real rows standing in for rows beyond the image, replication when a component row has <= 2 samples);
  * YCbCr -> RGB with the 16-bit fixed-point tables of jdcolor.c.

PINNED: ``tests/test_oracle_golden.py`` requires this restatement to reproduce Pillow's pixels bit for bit on generated
files (sizes that are not multiples of the MCU, 4:4:4 / 4:2:2 / 4:2:0, greyscale, optimised tables, restart intervals)
and on the reference's shipped JPEGs (``tests/golden/jpeg_files.npz``: file bytes + the pixels the unmodified reference
loader sees).  The device decoder (csrc/jpeg.cu) is then tested against Pillow and this file.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
                   46, 53, 60, 61, 54, 47, 55, 62, 63])


class Unsupported(Exception):
    pass


def parse(data: bytes):
    if data[:2] != b"\xff\xd8":
        raise Unsupported("no SOI")
    pos, quant, huff, frame, dri = 2, {}, {}, None, 0
    while pos + 4 <= len(data):
        if data[pos] != 0xFF:
            raise Unsupported("marker expected")
        m = data[pos + 1]
        if m == 0xFF:
            pos += 1
            continue
        pos += 2
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        n = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + n]
        if m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise Unsupported("precision")
            H, W, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [(seg[6 + 3 * c], seg[7 + 3 * c] >> 4, seg[7 + 3 * c] & 15, seg[8 + 3 * c]) for c in range(nc)]
            frame = (H, W, comps)
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8):
            raise Unsupported("not baseline Huffman")
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                bits = list(seg[i + 1:i + 17])
                cnt = sum(bits)
                vals = list(seg[i + 17:i + 17 + cnt])
                table, code, k = {}, 0, 0
                for length in range(1, 17):
                    for _ in range(bits[length - 1]):
                        table[(length, code)] = vals[k]
                        code += 1
                        k += 1
                    code <<= 1
                huff[(tc, th)] = table
                i += 17 + cnt
        elif m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                if pq:
                    q = [(seg[i + 1 + 2 * k] << 8) | seg[i + 2 + 2 * k] for k in range(64)]
                else:
                    q = list(seg[i + 1:i + 65])
                nat = np.zeros(64, np.int64)
                nat[ZIGZAG] = q
                quant[tq] = nat
                i += 1 + 64 * (pq + 1)
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            if frame is None or ns != len(frame[2]):
                raise Unsupported("scan layout")
            tabs = [(seg[1 + 2 * c], seg[2 + 2 * c] >> 4, seg[2 + 2 * c] & 15) for c in range(ns)]
            return frame, quant, huff, dri, tabs, pos + n
        pos += n
    raise Unsupported("no scan")


class _Bits:
    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n, self.marker = data, pos, 0, 0, False

    def _fill(self):
        while self.n <= 24:
            byte = 0
            if not self.marker and self.p < len(self.d):
                byte = self.d[self.p]
                if byte == 0xFF:
                    nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nxt == 0:
                        self.p += 2
                    else:
                        self.marker, byte = True, 0
                else:
                    self.p += 1
            self.acc = (self.acc << 8) | byte
            self.n += 8

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v

    def restart(self):
        self.acc, self.n = 0, 0
        if self.marker:
            self.p += 2
            self.marker = False
        else:
            while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
                self.p += 1
            self.p += 2


def _symbol(br: _Bits, table) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.get(1)
        s = table.get((length, code))
        if s is not None:
            return s
    return 0


def _extend(r: int, s: int) -> int:
    return r - (1 << s) + 1 if r < (1 << (s - 1)) else r


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(i0, i1, i2, i3, i4, i5, i6, i7, shift):
    z2, z3 = i2, i6
    z1 = (z2 + z3) * 4433
    tmp2 = z1 + z3 * (-15137)
    tmp3 = z1 + z2 * 6270
    tmp0 = (i0 + i4) << 13
    tmp1 = (i0 - i4) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = i7, i5, i3, i1
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * 9633
    tmp0, tmp1, tmp2, tmp3 = tmp0 * 2446, tmp1 * 16819, tmp2 * 25172, tmp3 * 12299
    z1, z2, z3, z4 = z1 * -7373, z2 * -20995, z3 * -16069, z4 * -3196
    z3, z4 = z3 + z5, z4 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return (_descale(tmp10 + tmp3, shift), _descale(tmp11 + tmp2, shift), _descale(tmp12 + tmp1, shift),
            _descale(tmp13 + tmp0, shift), _descale(tmp13 - tmp0, shift), _descale(tmp12 - tmp1, shift),
            _descale(tmp11 - tmp2, shift), _descale(tmp10 - tmp3, shift))


def idct_islow(blocks: np.ndarray, q: np.ndarray) -> np.ndarray:
    """``int [n, 64]`` coefficients (natural order) -> ``uint8 [n, 8, 8]`` samples (jidctint.c jpeg_idct_islow)."""
    c = (blocks.astype(np.int64) * q[None, :]).reshape(-1, 8, 8)
    ws = np.stack(_idct_1d(*[c[:, r, :] for r in range(8)], 11), 1)            # pass 1: columns -> [n, 8(row), 8(col)]
    out = np.stack(_idct_1d(*[ws[:, :, k] for k in range(8)], 18), 2)          # pass 2: rows
    v = out & 1023                                                             # IDCT_range_limit[x & RANGE_MASK]
    return np.where(v < 128, v + 128, np.where(v < 512, 255, np.where(v < 896, 0, v - 896))).astype(np.uint8)


def _upsample(plane: np.ndarray, dw: int, dh: int, hf: int, vf: int, W: int, H: int) -> np.ndarray:
    """jdsample.c on the true ``[dh, dw]`` samples of a component -> ``[H, W]`` (int64)."""
    s = plane[:dh, :dw].astype(np.int64)
    if hf == 1 and vf == 1:
        return s[:H, :W]
    if vf == 1:                                                                # h2v1
        if dw <= 2:
            return np.repeat(s, 2, 1)[:H, :W]
        out = np.zeros((dh, 2 * dw), np.int64)
        prev = np.concatenate([s[:, :1], s[:, :-1]], 1)
        nxt = np.concatenate([s[:, 1:], s[:, -1:]], 1)
        out[:, 0::2] = (3 * s + prev + 1) >> 2
        out[:, 1::2] = (3 * s + nxt + 2) >> 2
        out[:, 0] = s[:, 0]
        out[:, -1] = s[:, -1]
        return out[:H, :W]
    if dw <= 2:                                                                # h2v2 on a very narrow component
        return np.repeat(np.repeat(s, 2, 0), 2, 1)[:H, :W]
    up = np.concatenate([s[:1], s[:-1]], 0)
    down = np.concatenate([s[1:], s[-1:]], 0)
    rows = np.zeros((2 * dh, dw), np.int64)
    rows[0::2] = 3 * s + up
    rows[1::2] = 3 * s + down
    prev = np.concatenate([rows[:, :1], rows[:, :-1]], 1)
    nxt = np.concatenate([rows[:, 1:], rows[:, -1:]], 1)
    out = np.zeros((2 * dh, 2 * dw), np.int64)
    out[:, 0::2] = (3 * rows + prev + 8) >> 4
    out[:, 1::2] = (3 * rows + nxt + 7) >> 4
    out[:, 0] = (4 * rows[:, 0] + 8) >> 4
    out[:, -1] = (4 * rows[:, -1] + 7) >> 4
    return out[:H, :W]


def decode_baseline(data: bytes) -> np.ndarray:
    """JPEG file bytes -> ``uint8 [H, W, 3]``, the pixels of ``Image.open(...).convert('RGB')``."""
    (H, W, comps), quant, huff, dri, tabs, pos = parse(data)
    nc = len(comps)
    if nc not in (1, 3):
        raise Unsupported("components")
    hs, vs = [c[1] for c in comps], [c[2] for c in comps]
    if nc == 1:
        hs, vs = [1], [1]
    elif not (hs[1] == vs[1] == hs[2] == vs[2] == 1 and (hs[0], vs[0]) in ((1, 1), (2, 1), (2, 2))):
        raise Unsupported("sampling")
    mcu_x, mcu_y = -(-W // (8 * hs[0])), -(-H // (8 * vs[0]))
    coefs = [np.zeros((mcu_y * vs[c], mcu_x * hs[c], 64), np.int64) for c in range(nc)]
    br, pred, left = _Bits(data, pos), [0] * nc, dri
    for my in range(mcu_y):
        for mx in range(mcu_x):
            if dri and left == 0:
                br.restart()
                pred, left = [0] * nc, dri
            for c in range(nc):
                dct, act = huff[(0, tabs[c][1])], huff[(1, tabs[c][2])]
                for by in range(vs[c]):
                    for bx in range(hs[c]):
                        blk = coefs[c][my * vs[c] + by, mx * hs[c] + bx]
                        s = _symbol(br, dct)
                        if s:
                            pred[c] += _extend(br.get(s), s)
                        blk[0] = pred[c]
                        k = 1
                        while k < 64:
                            rs = _symbol(br, act)
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                blk[ZIGZAG[k & 63]] = _extend(br.get(s), s)
                            elif r != 15:
                                break
                            else:
                                k += 15
                            k += 1
            if dri:
                left -= 1
    planes = []
    for c in range(nc):
        by, bx = coefs[c].shape[:2]
        px = idct_islow(coefs[c].reshape(-1, 64), quant[comps[c][3]]).reshape(by, bx, 8, 8)
        planes.append(px.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8))
    Y = planes[0][:H, :W].astype(np.int64)
    if nc == 1:
        return np.stack([Y, Y, Y], -1).astype(np.uint8)
    hf, vf = hs[0] // hs[1], vs[0] // vs[1]
    dw, dh = -(-W * hs[1] // hs[0]), -(-H * vs[1] // vs[0])
    cb = _upsample(planes[1], dw, dh, hf, vf, W, H) - 128
    cr = _upsample(planes[2], dw, dh, hf, vf, W, H) - 128
    r = Y + ((91881 * cr + 32768) >> 16)
    g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
    b = Y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)
