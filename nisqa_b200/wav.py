"""Host-side WAV (and FLAC) ingest for the engine.

Mirrors what ``lb.load(path, sr=None[, mono=False])`` + the channel pick does at reference
``nisqa/NISQA_lib.py:2298-2306`` (libsndfile float conversion; mono = float32 mean over
channels; ``ms_channel`` picks one channel of a multi-channel file), but keeps mono PCM16
files as int16 so the host->device copy is half the size and the ``/32768`` happens on the
device (bit-identical: division by a power of two).

Returns ``(samples, sample_rate)`` where samples is a contiguous 1-D ``int16`` or ``float32``
array.  Any parse problem raises ``ValueError('Could not load file ...')`` exactly like the
reference's bare ``except`` (lib:2305-2306).
"""
import os
import struct

import numpy as np

_INV = {8: np.float32(1.0 / 128.0), 16: np.float32(1.0 / 32768.0), 24: np.float32(1.0 / 8388608.0)}


def _g711_table(alaw):
    """8-bit G.711 code -> 16-bit linear sample (ITU-T G.711 expansion, as libsndfile's alaw.c / ulaw.c)."""
    out = np.zeros(256, np.int16)
    for code in range(256):
        if alaw:
            c = code ^ 0x55
            t, seg = (c & 0x0F) << 4, (c & 0x70) >> 4
            t = t + 8 if seg == 0 else (t + 0x108) << (seg - 1)
            out[code] = t if c & 0x80 else -t
        else:
            c = ~code & 0xFF
            t = (((c & 0x0F) << 3) + 0x84) << ((c & 0x70) >> 4)
            out[code] = 0x84 - t if c & 0x80 else t - 0x84
    return out


_G711 = {6: _g711_table(True), 7: _g711_table(False)}       # WAVE_FORMAT_ALAW / WAVE_FORMAT_MULAW


def _parse(buf):
    if len(buf) < 12 or buf[0:4] != b"RIFF" or buf[8:12] != b"WAVE":
        raise ValueError("not RIFF/WAVE")
    pos, fmt, payload = 12, None, None
    n = len(buf)
    while pos + 8 <= n:
        cid = bytes(buf[pos:pos + 4])
        (csz,) = struct.unpack_from("<I", buf, pos + 4)
        start = pos + 8
        end = min(start + csz, n)
        if cid == b"fmt ":
            tag, ch, sr, _, block_align, bits = struct.unpack_from("<HHIIHH", buf, start)
            if tag == 0xFFFE and end - start >= 26:
                (tag,) = struct.unpack_from("<H", buf, start + 24)
            # same plausibility rules as the native reader (csrc/wavio.cpp parse_header)
            if not (1 <= ch <= 256) or sr < 1 or bits < 8 or bits > 64 or bits % 8 or block_align != ch * (bits // 8):
                raise ValueError("implausible fmt chunk")
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            payload = (start, end)
        pos = start + csz + (csz & 1)
    if fmt is None or payload is None or fmt[1] < 1:
        raise ValueError("missing fmt/data")
    return fmt, payload


def read_wav(path, ms_channel=None):
    """-> (1-D int16 | float32 contiguous array, sample_rate:int)."""
    try:
        with open(path, "rb") as f:
            buf = f.read()
        if buf[:4] == b"fLaC":                    # FLAC goes through the native reader (csrc/flac.cpp): no NumPy twin
            return read_wav_native(path, ms_channel)
        (tag, ch, sr, bits), (s, e) = _parse(memoryview(buf))
        width = bits // 8
        n_frames = (e - s) // (width * ch)
        raw = np.frombuffer(buf, dtype=np.uint8, count=n_frames * ch * width, offset=s)
        if tag == 1 and bits == 16:
            x = raw.view("<i2").reshape(n_frames, ch)
            if ch == 1:
                return np.ascontiguousarray(x[:, 0]), int(sr)
            if ms_channel is not None:
                return np.ascontiguousarray(x[:, ms_channel]), int(sr)
            y = x.astype(np.float32) / np.float32(32768.0)
        elif tag == 1 and bits == 8:
            y = (raw.reshape(n_frames, ch).astype(np.float32) - np.float32(128.0)) * _INV[8]
        elif tag == 1 and bits == 24:
            b = raw.reshape(n_frames * ch, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = (v ^ 0x800000) - 0x800000
            y = (v.astype(np.float32) * _INV[24]).reshape(n_frames, ch)
        elif tag == 1 and bits == 32:
            y = (raw.view("<i4").astype(np.float64) * (1.0 / 2147483648.0)).astype(np.float32)
            y = y.reshape(n_frames, ch)
        elif tag in (6, 7) and bits == 8:       # G.711 A-law / mu-law: 16-bit expansion, then libsndfile's / 32768
            y = (_G711[tag][raw].astype(np.float32) * _INV[16]).reshape(n_frames, ch)
        elif tag == 3 and bits == 32:
            y = raw.view("<f4").astype(np.float32).reshape(n_frames, ch)
        elif tag == 3 and bits == 64:
            y = raw.view("<f8").astype(np.float32).reshape(n_frames, ch)
        else:
            raise ValueError("unsupported WAVE encoding tag=%d bits=%d" % (tag, bits))
        if ch == 1:
            out = y[:, 0]
        elif ms_channel is not None:
            out = y[:, ms_channel]
        else:
            out = np.mean(y.T, axis=0)          # librosa.to_mono on the [ch, n] float32 array
        return np.ascontiguousarray(out, dtype=np.float32), int(sr)
    except Exception:
        raise ValueError("Could not load file {}".format(path))


def probe_wav(path, ms_channel=None):
    """Native header probe (csrc/wavio.cpp) -> (sample_rate, n_frames, channels, kind) with kind
    0 = deliverable as int16, 1 = float32.  ValueError('Could not load file ...') on failure."""
    import ctypes as C
    from . import engine as _e
    lib = _e.load_library()
    sr, nf, ch, kind = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
    rc = lib.nisqa_wav_probe(os.fsencode(path), -1 if ms_channel is None else int(ms_channel),
                             C.byref(sr), C.byref(nf), C.byref(ch), C.byref(kind))
    if rc != 0:
        raise ValueError("Could not load file {}".format(path))
    return sr.value, nf.value, ch.value, kind.value


def decode_wav_into(path, dst, ms_channel=None):
    """Native decode straight into ``dst`` (1-D int16 or float32 array / pinned tensor view with room
    for the clip).  Returns the number of samples written."""
    from . import engine as _e
    lib = _e.load_library()
    fmt = _e.FMT_S16 if dst.dtype == np.int16 else _e.FMT_F32
    n = lib.nisqa_wav_decode(os.fsencode(path), -1 if ms_channel is None else int(ms_channel), fmt,
                             dst.ctypes.data, dst.shape[0])
    if n < 0:
        raise ValueError("Could not load file {}".format(path))
    return int(n)


def probe_batch(paths, ms_channel=None, n_threads=1):
    """One native call for a whole batch -> (sample_rate int32[n], n_frames int64[n], kind int32[n]).
    Raises ValueError('Could not load file <first bad path>')."""
    import ctypes as C
    from . import engine as _e
    lib = _e.load_library()
    n = len(paths)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    sr = np.zeros(n, np.int32); nf = np.zeros(n, np.int64); kind = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
    bad = lib.nisqa_wav_probe_batch(n, arr, -1 if ms_channel is None else int(ms_channel), int(n_threads),
                                    sr.ctypes.data_as(C.POINTER(C.c_int32)), nf.ctypes.data_as(C.POINTER(C.c_int64)),
                                    kind.ctypes.data_as(C.POINTER(C.c_int32)), st.ctypes.data_as(C.POINTER(C.c_int32)))
    if bad != 0:
        i = int(np.flatnonzero(st)[0]) if bad > 0 and st.any() else 0
        raise ValueError("Could not load file {}".format(paths[i]))
    return sr, nf, kind, arr


def decode_batch(path_array, n, dst, offsets, n_frames, ms_channel=None, n_threads=1, paths=None):
    """Decode the whole batch into ``dst`` (1-D int16 / float32 pinned array) at element ``offsets``."""
    import ctypes as C
    from . import engine as _e
    lib = _e.load_library()
    st = np.zeros(n, np.int32)
    offsets = np.ascontiguousarray(offsets, np.int64); caps = np.ascontiguousarray(n_frames, np.int64)
    fmt = _e.FMT_S16 if dst.dtype == np.int16 else _e.FMT_F32
    bad = lib.nisqa_wav_decode_batch(n, path_array, -1 if ms_channel is None else int(ms_channel), fmt,
                                     dst.ctypes.data, offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                     caps.ctypes.data_as(C.POINTER(C.c_int64)), int(n_threads),
                                     st.ctypes.data_as(C.POINTER(C.c_int32)))
    if bad != 0:
        i = int(np.flatnonzero(st)[0]) if bad > 0 and st.any() else 0
        raise ValueError("Could not load file {}".format(paths[i] if paths else "<batch>"))


def read_wav_native(path, ms_channel=None):
    """read_wav through the native reader (same return contract)."""
    sr, nf, _, kind = probe_wav(path, ms_channel)
    out = np.empty(nf, dtype=np.int16 if kind == 0 else np.float32)
    decode_wav_into(path, out, ms_channel)
    return out, sr


def write_wav_pcm16(path, pcm, sr):
    """Write int16 samples, shape [n] (mono) or [n, ch]."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    ch = 1 if pcm.ndim == 1 else pcm.shape[1]
    data = pcm.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, sr, sr * ch * 2, ch * 2, 16) + b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)


def write_wav_f32(path, y, sr):
    y = np.ascontiguousarray(y, dtype="<f4")
    ch = 1 if y.ndim == 1 else y.shape[1]
    data = y.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 3, ch, sr, sr * ch * 4, ch * 4, 32) + b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)
