"""Native FLAC reader (nisqa_b200/csrc/flac.cpp) - round trips through the independent test encoder
tests/flac_enc.py (FLAC is lossless: decode(encode(x)) == x bit for bit), every construct of the format the reader
implements, damaged files, and the loader entry points.  PARITY UNPINNED: no FLAC encoder / decoder exists in this
environment to produce third-party vectors."""
import os

import numpy as np
import pytest

from nisqa_b200 import synth, wav
import flac_enc as FE


def _write(tmp_path, name, data):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(data)
    return p


def _decode_i16(path, ch=None):
    sr, nf, nch, kind = wav.probe_wav(path, ch)
    out = np.empty(nf, np.int16 if kind == 0 else np.float32)
    wav.decode_wav_into(path, out, ch)
    return out, sr, nch, kind


def test_mono_16_bit_speech_round_trip(tmp_path, built_lib):
    pcm = synth.synth_speech_pcm16(5, 1.3, 48000)
    p = _write(tmp_path, "a.flac", FE.encode(pcm, 48000, 16, blocksize=4096))
    got, sr, nch, kind = _decode_i16(p)
    assert (sr, nch, kind) == (48000, 1, 0) and got.dtype == np.int16
    np.testing.assert_array_equal(got, pcm)
    f = np.empty(len(pcm), np.float32)
    wav.decode_wav_into(p, f)
    np.testing.assert_array_equal(f, pcm.astype(np.float32) / np.float32(32768.0))     # libsndfile's int -> float rule
    y, sr2 = wav.read_wav(p)                                                            # the loader's twin dispatches FLAC natively
    assert sr2 == 48000 and np.array_equal(y, pcm)
    assert len(open(p, "rb").read()) < 0.8 * 2 * len(pcm)                               # (it does compress)


def test_every_subframe_type_predictor_and_residual_coding(tmp_path, built_lib):
    rng = np.random.default_rng(1)
    base = synth.synth_speech_pcm16(6, 1.0, 16000).astype(np.int64)
    n = 16 * 1024
    x = np.resize(base, n)
    x[1024:2048] = 123                                    # a constant block
    x[2048:3072] = rng.integers(-30000, 30000, 1024)      # noise: verbatim / high Rice parameters
    x[3072:4096] = (x[3072:4096] // 8) * 8                # three wasted bits
    kinds = [dict(kind="fixed", order=0), dict(kind="constant"), dict(kind="verbatim"), dict(kind="fixed", allow_wasted=True),
             dict(kind="fixed", order=1, porder=2), dict(kind="fixed", order=2, porder=3, method=1), dict(kind="fixed", order=3, porder=1, force_escape=True),
             dict(kind="fixed", order=4, porder=4), dict(kind="lpc", order=1), dict(kind="lpc", order=8, porder=2), dict(kind="lpc", order=32, precision=15, porder=3, method=1),
             dict(kind="lpc", order=12, precision=9, force_escape=True, porder=2), dict(kind="auto"), dict(kind="verbatim", allow_wasted=True),
             dict(kind="fixed", order=2, force_escape=True), dict(kind="auto", allow_wasted=True)]
    p = _write(tmp_path, "k.flac", FE.encode(x, 16000, 16, blocksize=1024, plan=lambda f, c: "indep" if c is None else kinds[f % len(kinds)]))
    got, sr, _, _ = _decode_i16(p)
    assert sr == 16000
    np.testing.assert_array_equal(got, x.astype(np.int16))


def test_stereo_decorrelations_channel_pick_and_mono_mix(tmp_path, built_lib):
    a = synth.synth_speech_pcm16(7, 0.7, 44100).astype(np.int64)
    b = np.roll(a, 37) // 2 + synth.synth_speech_pcm16(8, 0.7, 44100).astype(np.int64) // 3
    st = np.stack([a, b], axis=1)
    modes = ["indep", "ls", "rs", "ms"]
    plan = lambda f, c: modes[f % 4] if c is None else dict(kind=("lpc" if f % 2 else "fixed"), order=(6 if f % 2 else None) if f % 2 else None)
    p = _write(tmp_path, "s.flac", FE.encode(st, 44100, 16, blocksize=1152, plan=plan))
    for ch in (0, 1):
        got, sr, nch, kind = _decode_i16(p, ch)
        assert (sr, nch, kind) == (44100, 2, 0)
        np.testing.assert_array_equal(got, st[:, ch].astype(np.int16))
    mix, _, _, kind = _decode_i16(p, None)
    assert kind == 1
    ref = np.mean((st.astype(np.float32) / np.float32(32768.0)).T, axis=0)               # librosa.to_mono on float32 channels
    np.testing.assert_array_equal(mix, ref.astype(np.float32))


@pytest.mark.parametrize("bits,scale", [(8, 128.0), (24, 8388608.0)])
def test_other_sample_sizes_and_odd_block_sizes(tmp_path, built_lib, bits, scale):
    rng = np.random.default_rng(bits)
    n = 5000
    t = np.arange(n)
    x = (np.sin(t * 0.01) * (2 ** (bits - 1) - 2) * 0.7).astype(np.int64) + rng.integers(-3, 4, n)
    for bs, declare in ((1000, True), (200, False), (4608, True)):        # 16-bit and 8-bit explicit block sizes, unknown length
        p = _write(tmp_path, "b%d_%d.flac" % (bits, bs), FE.encode(x, 22050 if bs == 200 else 11025, bits, blocksize=bs,
                                                                     declare_length=declare, extra_metadata=bs != 200))
        got, sr, _, kind = _decode_i16(p)
        assert kind == 1 and sr == (22050 if bs == 200 else 11025) and len(got) == n
        np.testing.assert_array_equal(got, (x.astype(np.float32) * np.float32(1.0 / scale)).astype(np.float32))


def test_damaged_files_are_could_not_load(tmp_path, built_lib):
    pcm = synth.synth_speech_pcm16(9, 0.3, 16000)
    data = bytearray(FE.encode(pcm, 16000, 16, blocksize=1024))
    for name, bad in (("trunc.flac", bytes(data[:len(data) // 2])), ("nomagic.flac", b"fLaD" + bytes(data[4:])),
                      ("hdr.flac", bytes(data[:90]) + bytes([data[90] ^ 0x10]) + bytes(data[91:])),
                      ("body.flac", bytes(data[:400]) + bytes([data[400] ^ 0x01]) + bytes(data[401:]))):
        p = _write(tmp_path, name, bad)
        with pytest.raises(ValueError, match="Could not load file"):
            wav.read_wav(p)
    p = _write(tmp_path, "ok.flac", bytes(data))
    assert np.array_equal(wav.read_wav(p)[0], pcm)


def test_batch_loader_mixes_wav_and_flac(tmp_path, built_lib):
    from nisqa_b200 import NISQA_lib as NL
    import pandas as pd
    a, b = synth.synth_speech_pcm16(10, 0.5, 48000), synth.synth_speech_pcm16(11, 0.4, 48000)
    wav.write_wav_pcm16(str(tmp_path / "a.wav"), a, 48000)
    _write(tmp_path, "b.flac", FE.encode(b, 48000, 16))
    df = pd.DataFrame({"deg": ["a.wav", "b.flac"]})
    ds = NL.SpeechQualityDataset(df, data_dir=str(tmp_path), filename_column="deg", mos_column="predict_only", ms_sr=None)

    class Pool(object):
        def get(self, slot, nbytes):
            return np.zeros(int(nbytes) + 64, np.uint8)
    clips, srs = NL._load_batch(ds, [0, 1], Pool(), 0, 2)
    assert srs == [48000, 48000] and np.array_equal(clips[0], a) and np.array_equal(clips[1], b)
