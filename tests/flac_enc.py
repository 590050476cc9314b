"""A small FLAC ENCODER for the tests of the native FLAC reader (nisqa_b200/csrc/flac.cpp).  TEST INFRASTRUCTURE ONLY.

There is no FLAC encoder or decoder in this environment (no libsndfile / libFLAC / ffmpeg), so the reader is exercised by
round trips: FLAC is lossless, decode(encode(x)) must be x bit for bit.  The encoder is written from the format
specification independently of the decoder's code (Python big integers and a bit list against C++ shifts), and it can
be steered through every construct the decoder implements: CONSTANT / VERBATIM / FIXED (orders 0-4) / LPC subframes,
Rice and Rice2 partitions with and without escape codes, wasted bits, the four channel assignments, 8 / 16 / 24-bit
samples, short last blocks, unknown total length.
"""
import numpy as np


class Bits(object):
    def __init__(self):
        self.b = []

    def put(self, v, n):
        for i in range(n - 1, -1, -1):
            self.b.append((int(v) >> i) & 1)

    def puts(self, v, n):                 # two's complement
        self.put(int(v) & ((1 << n) - 1), n)

    def unary(self, q):
        self.b.extend([0] * int(q) + [1])

    def align(self):
        while len(self.b) % 8:
            self.b.append(0)

    def tobytes(self):
        assert len(self.b) % 8 == 0
        a = np.array(self.b, dtype=np.uint8).reshape(-1, 8)
        return bytes(np.packbits(a, axis=1).reshape(-1).tolist())


def crc8(data):
    c = 0
    for byte in data:
        c ^= byte
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(data):
    c = 0
    for byte in data:
        c ^= byte << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def utf8_number(n):
    if n < 0x80:
        return [n]
    out, lead_bits = [], 6
    tail = []
    while True:
        tail.append(0x80 | (n & 0x3F))
        n >>= 6
        lead_bits -= 1
        if n < (1 << lead_bits):
            break
    k = len(tail) + 1
    lead = ((0xFF << (8 - k)) & 0xFF) | n
    return [lead] + tail[::-1]


def zigzag(r):
    return (r << 1) if r >= 0 else ((-r) << 1) - 1


def write_residual(bw, res, order, blocksize, method=0, porder=0, force_escape=False):
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    bw.put(method, 2)
    bw.put(porder, 4)
    idx = 0
    for part in range(1 << porder):
        count = (blocksize >> porder) - (order if part == 0 else 0)
        seg = [int(v) for v in res[idx:idx + count]]
        idx += count
        zz = [zigzag(v) for v in seg]
        best_k, best_bits = 0, None
        for k in range(0, esc):
            bits = sum((z >> k) + 1 + k for z in zz)
            if best_bits is None or bits < best_bits:
                best_k, best_bits = k, bits
        if force_escape and part % 2 == 0:
            raw = max([1] + [int(abs(v)).bit_length() + 1 for v in seg])
            bw.put(esc, pbits)
            bw.put(raw, 5)
            for v in seg:
                bw.puts(v, raw)
        else:
            bw.put(best_k, pbits)
            for z in zz:
                bw.unary(z >> best_k)
                if best_k:
                    bw.put(z & ((1 << best_k) - 1), best_k)
    assert idx == len(res)


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def write_subframe(bw, s, bps, kind, **opt):
    """s: list of Python ints of one channel (already decorrelated)."""
    n = len(s)
    wasted = 0
    if opt.get("allow_wasted") and any(s):
        while all((v >> wasted) & 1 == 0 for v in s) and wasted < bps - 1:
            wasted += 1
    if wasted:
        s = [v >> wasted for v in s]
    eb = bps - wasted
    if kind == "auto":
        kind = "constant" if all(v == s[0] for v in s) else "fixed"
    if kind == "constant" and not all(v == s[0] for v in s):
        kind = "verbatim"

    def header(t):
        bw.put(0, 1); bw.put(t, 6)
        if wasted:
            bw.put(1, 1); bw.unary(wasted - 1)
        else:
            bw.put(0, 1)

    if kind == "constant":
        header(0); bw.puts(s[0], eb)
    elif kind == "verbatim":
        header(1)
        for v in s:
            bw.puts(v, eb)
    elif kind == "fixed":
        order = opt.get("order")
        if order is None:                       # smallest sum of |residual|
            best = None
            for o in range(0, min(4, n - 1) + 1):
                c = FIXED[o]
                r = [s[i] - sum(c[j] * s[i - 1 - j] for j in range(o)) for i in range(o, n)]
                cost = sum(abs(v) for v in r)
                if best is None or cost < best[0]:
                    best = (cost, o, r)
            _, order, res = best
        else:
            c = FIXED[order]
            res = [s[i] - sum(c[j] * s[i - 1 - j] for j in range(order)) for i in range(order, n)]
        header(8 + order)
        for v in s[:order]:
            bw.puts(v, eb)
        write_residual(bw, res, order, n, opt.get("method", 0), opt.get("porder", 0), opt.get("force_escape", False))
    elif kind == "lpc":
        order, prec = opt.get("order", 8), opt.get("precision", 12)
        x = np.array(s, dtype=np.float64)
        A = np.stack([x[order - 1 - j:n - 1 - j] for j in range(order)], axis=1)
        coef = np.linalg.lstsq(A, x[order:], rcond=None)[0]
        cmax = max(float(np.abs(coef).max()), 1e-9)
        shift = int(min(15, max(0, prec - 1 - int(np.ceil(np.log2(cmax))) - 1)))
        q = [int(np.clip(round(float(c) * (1 << shift)), -(1 << (prec - 1)), (1 << (prec - 1)) - 1)) for c in coef]
        res = [s[i] - (sum(q[j] * s[i - 1 - j] for j in range(order)) >> shift) for i in range(order, n)]
        header(31 + order)
        for v in s[:order]:
            bw.puts(v, eb)
        bw.put(prec - 1, 4)
        bw.puts(shift, 5)
        for c in q:
            bw.puts(c, prec)
        write_residual(bw, res, order, n, opt.get("method", 0), opt.get("porder", 0), opt.get("force_escape", False))
    else:
        raise ValueError(kind)


BS_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14, 32768: 15}
SR_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
SS_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6}


def encode(pcm, sample_rate, bits=16, blocksize=4096, plan=None, declare_length=True, extra_metadata=True):
    """pcm: int array [n] or [n, channels].  plan(frame_index, channel) -> dict(kind=..., order=..., method=..., porder=...,
    force_escape=..., allow_wasted=...) and plan(frame_index, None) -> channel assignment ('indep' | 'ls' | 'rs' | 'ms').
    Returns the bytes of a .flac file."""
    x = np.asarray(pcm)
    if x.ndim == 1:
        x = x[:, None]
    n, ch = x.shape
    plan = plan or (lambda f, c: "indep" if c is None else {"kind": "auto"})
    frames = []
    min_f, max_f = 1 << 30, 0
    fi = 0
    for start in range(0, n, blocksize):
        blk = x[start:start + blocksize]
        bsz = blk.shape[0]
        assign = plan(fi, None) if ch == 2 else "indep"
        chans = [[int(v) for v in blk[:, c]] for c in range(ch)]
        bps_c = [bits] * ch
        code = ch - 1
        if assign == "ls":
            chans = [chans[0], [a - b for a, b in zip(chans[0], chans[1])]]; bps_c = [bits, bits + 1]; code = 8
        elif assign == "rs":
            chans = [[a - b for a, b in zip(chans[0], chans[1])], chans[1]]; bps_c = [bits + 1, bits]; code = 9
        elif assign == "ms":
            chans = [[(a + b) >> 1 for a, b in zip(chans[0], chans[1])], [a - b for a, b in zip(chans[0], chans[1])]]
            bps_c = [bits, bits + 1]; code = 10
        bw = Bits()
        bw.put(0b11111111111110, 14); bw.put(0, 1); bw.put(0, 1)            # sync, reserved, fixed block size stream
        if bsz in BS_CODES:
            bs_code = BS_CODES[bsz]
        elif bsz <= 256:
            bs_code = 6
        else:
            bs_code = 7
        sr_code = SR_CODES.get(sample_rate, 0) if fi % 2 == 0 else 0       # (every other frame defers to STREAMINFO)
        if sample_rate not in SR_CODES and fi % 2 == 0 and sample_rate < 65536:
            sr_code = 13
        bw.put(bs_code, 4); bw.put(sr_code, 4); bw.put(code, 4)
        bw.put(SS_CODES[bits] if fi % 3 else 0, 3); bw.put(0, 1)
        for byte in utf8_number(fi):
            bw.put(byte, 8)
        if bs_code == 6:
            bw.put(bsz - 1, 8)
        elif bs_code == 7:
            bw.put(bsz - 1, 16)
        if sr_code == 13:
            bw.put(sample_rate, 16)
        hdr = bw.tobytes()
        bw.put(crc8(hdr), 8)
        for c in range(ch):
            opt = dict(plan(fi, c))
            kind = opt.pop("kind", "auto")
            write_subframe(bw, chans[c], bps_c[c], kind, **opt)
        bw.align()
        body = bw.tobytes()
        frame = body + bytes([crc16(body) >> 8, crc16(body) & 0xFF])
        frames.append(frame)
        min_f, max_f = min(min_f, len(frame)), max(max_f, len(frame))
        fi += 1
    si = Bits()
    si.put(blocksize, 16); si.put(blocksize, 16); si.put(min_f if frames else 0, 24); si.put(max_f, 24)
    si.put(sample_rate, 20); si.put(ch - 1, 3); si.put(bits - 1, 5); si.put(n if declare_length else 0, 36)
    si.put(0, 128)                                                           # MD5 not computed (0 = unknown)
    si_b = si.tobytes()
    out = b"fLaC"
    if extra_metadata:
        out += bytes([0x00, 0, 0, len(si_b)]) + si_b
        pad = bytes(37)
        out += bytes([0x80 | 1, 0, 0, len(pad)]) + pad                      # a PADDING block closes the metadata
    else:
        out += bytes([0x80, 0, 0, len(si_b)]) + si_b
    return out + b"".join(frames)
