"""A tiny FLAC *encoder* used only by the tests to produce bit-exact round-trip inputs for the library's decoder
(constant / verbatim / fixed-predictor subframes, Rice partitions of order 0 and 1, independent and mid/side stereo)."""
import hashlib
import struct

import numpy as np


class _BW:
    def __init__(self):
        self.bits = []

    def w(self, v, n):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def ws(self, v, n):
        self.w(v & ((1 << n) - 1), n)

    def unary(self, q):
        self.bits.extend([0] * q + [1])

    def pad(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self):
        self.pad()
        b = bytearray()
        for i in range(0, len(self.bits), 8):
            x = 0
            for bit in self.bits[i:i + 8]:
                x = (x << 1) | bit
            b.append(x)
        return bytes(b)


def _crc(data, poly, width):
    c = 0
    top = 1 << (width - 1)
    mask = (1 << width) - 1
    for byte in data:
        c ^= byte << (width - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


def _rice(bw, v, k):
    u = (v << 1) if v >= 0 else ((-v) << 1) - 1
    bw.unary(u >> k)
    if k:
        bw.w(u & ((1 << k) - 1), k)


def _subframe(bw, s, bps, mode, order, porder):
    n = len(s)
    bw.w(0, 1)
    if mode == "constant":
        bw.w(0, 6); bw.w(0, 1); bw.ws(int(s[0]), bps)
    elif mode == "verbatim":
        bw.w(1, 6); bw.w(0, 1)
        for v in s:
            bw.ws(int(v), bps)
    else:
        bw.w(8 + order, 6); bw.w(0, 1)
        for v in s[:order]:
            bw.ws(int(v), bps)
        coefs = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}[order]
        res = [int(s[i]) - sum(c * int(s[i - 1 - j]) for j, c in enumerate(coefs)) for i in range(order, n)]
        bw.w(0, 2); bw.w(porder, 4)
        parts = 1 << porder
        idx = 0
        for p in range(parts):
            cnt = (n >> porder) - (order if p == 0 else 0)
            chunk = res[idx:idx + cnt]; idx += cnt
            mean = max(1.0, float(np.mean(np.abs(chunk))) if chunk else 1.0)
            k = min(14, max(0, int(np.log2(mean)) + 1))
            bw.w(k, 4)
            for v in chunk:
                _rice(bw, v, k)


def write_flac(path, pcm, sample_rate=16000, bps=16, blocksize=1024, mode="fixed", order=2, porder=1, mid_side=False):
    """pcm: int array [frames, channels]."""
    pcm = np.asarray(pcm, dtype=np.int64)
    n, ch = pcm.shape
    out = bytearray(b"fLaC")
    md5 = hashlib.md5(pcm.astype("<i2").tobytes()).digest()
    si = _BW()
    si.w(blocksize, 16); si.w(blocksize, 16); si.w(0, 24); si.w(0, 24); si.w(sample_rate, 20); si.w(ch - 1, 3)
    si.w(bps - 1, 5); si.w(n, 36)
    out += bytes([0x80]) + struct.pack(">I", 34)[1:] + si.bytes() + md5
    fnum = 0
    for start in range(0, n, blocksize):
        blk = pcm[start:start + blocksize]
        bs = len(blk)
        hdr = _BW()
        hdr.w(0x3FFE, 14); hdr.w(0, 1); hdr.w(0, 1)
        hdr.w(7, 4)                      # 16-bit explicit block size
        hdr.w(0, 4)                      # sample rate from STREAMINFO
        hdr.w(10 if (mid_side and ch == 2) else ch - 1, 4)
        hdr.w(0, 3); hdr.w(0, 1)         # sample size from STREAMINFO
        assert fnum < 128
        hdr.w(fnum, 8)
        hdr.w(bs - 1, 16)
        hb = hdr.bytes()
        body = _BW()
        if mid_side and ch == 2:
            l, r = blk[:, 0], blk[:, 1]
            _subframe(body, (l + r) >> 1, bps, mode, order, porder if bs % (1 << porder) == 0 else 0)
            _subframe(body, l - r, bps + 1, mode, order, porder if bs % (1 << porder) == 0 else 0)
        else:
            for c in range(ch):
                _subframe(body, blk[:, c], bps, mode, order, porder if bs % (1 << porder) == 0 else 0)
        frame = hb + bytes([_crc(hb, 0x07, 8)]) + body.bytes()
        out += frame + struct.pack(">H", _crc(frame, 0x8005, 16))
        fnum += 1
    open(path, "wb").write(bytes(out))
    return md5
