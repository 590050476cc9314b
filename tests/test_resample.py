"""Sample-rate conversion of the ingest (reference lib:2300-2304 with ms_sr != None).

The native resampler (csrc/resample.cpp) against the oracle's restatement of librosa 0.8.1 ``resample`` /
resampy ``kaiser_best`` (bit-exact: same table, same order of additions), the restatement against signals whose
band-limited interpolation is known in closed form, and the loader path of ``ms_sr`` checkpoints.  resampy is a
third-party dependency that is not available here: PARITY UNPINNED, like the rest of the front end.
"""
import os

import numpy as np
import pytest

from nisqa_b200 import NISQA_lib as NL
from nisqa_b200 import resample as R
from nisqa_b200 import wav
from oracle import librosa_compat as lb


@pytest.mark.parametrize("so,sn,n", [(48000, 16000, 48000), (16000, 48000, 16000), (44100, 48000, 30011), (8000, 16000, 777),
                                     (48000, 44100, 5000), (22050, 16000, 12345), (48000, 8000, 100), (16000, 16000, 321)])
def test_native_resampler_is_bit_identical_to_the_restatement(built_lib, so, sn, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.1).astype(np.float32)
    y = R.resample(x, so, sn)
    ref = lb.resample(x, so, sn)
    assert y.dtype == np.float32 and len(y) == int(np.ceil(n * sn / so)) == R.out_len(n, so, sn)
    np.testing.assert_array_equal(y, ref)
    # int16 input = the same samples / 32768
    xi = np.round(x * 32767).astype(np.int16)
    np.testing.assert_array_equal(R.resample(xi, so, sn), lb.resample(xi.astype(np.float32) / np.float32(32768.0), so, sn))


def test_filter_table_is_resampys_kaiser_best():
    win, num_table = R.kaiser_best_half_window()
    ref, nt = lb.sinc_window(64, 9, 0.9475937167399596, 14.769656459379492)
    assert num_table == nt == 512 and win.shape == (512 * 64 + 1,)
    np.testing.assert_array_equal(win, ref)
    assert abs(win[0] - 0.9475937167399596) < 1e-15           # rolloff * sinc(0) * kaiser centre (= 1)
    k = np.arange(1, 60)                                        # zeros of sinc(rolloff * t): t = k / rolloff
    idx = np.round(k / 0.9475937167399596 * 512).astype(int)
    assert np.abs(win[idx]).max() < 2e-3 and abs(win[-1]) < 1e-7          # Kaiser taper: ~1e-8 at the 64th zero


@pytest.mark.parametrize("tgt,tol", [(96000, 2e-6), (44100, 6e-4), (16000, 4e-3)])
def test_restatement_interpolates_band_limited_signals(tgt, tol):
    """A sum of sinusoids below both Nyquist limits has a closed-form resampled version.  Up-sampling reproduces
    it to float32 precision; when down-sampling, resampy strides its filter table by int(ratio * 512) instead of
    ratio * 512, which shifts gain and cut-off slightly (a property of the algorithm: a few 1e-3 at 48 -> 16 kHz)."""
    sr, n = 48000, 48000
    t = np.arange(n) / sr
    f = [(440.0, 0.5, 0.0), (3000.0, 0.3, 1.0), (6500.0, 0.1, 0.3)]
    x = sum(a * np.sin(2 * np.pi * fr * t + ph) for fr, a, ph in f).astype(np.float32)
    y = lb.resample(x, sr, tgt)
    tt = np.arange(len(y)) / tgt
    want = sum(a * np.sin(2 * np.pi * fr * tt + ph) for fr, a, ph in f)
    edge = 3000 * tgt // sr
    assert np.abs(y[edge:-edge] - want[edge:-edge]).max() <= tol


class _HostPool(object):                   # stands in for the pinned ring (torch.pin_memory needs CUDA)
    def __init__(self):
        self.buf = None

    def get(self, slot, nbytes):
        self.buf = np.zeros(int(nbytes) + 64, np.uint8)
        return self.buf


def test_loader_converts_to_ms_sr(tmp_path, built_lib, monkeypatch):
    """Dataset with ms_sr = 16000: files at 48 kHz (stereo, mono mix first), 8 kHz A-law and 16 kHz (untouched) come
    out of the batch loader as float32 at 16 kHz, equal to the oracle's lb.load(path, sr=16000).  (The host conversion
    of the loader, NISQA_RESAMPLE=host; by default the clips travel at their own rates and are converted on the device,
    tests/test_gpu_parity.py::test_device_resampler_is_bit_identical_to_the_host_resampler.)"""
    monkeypatch.setenv("NISQA_RESAMPLE", "host")
    import pandas as pd
    import struct
    rng = np.random.default_rng(3)
    wav.write_wav_pcm16(str(tmp_path / "a.wav"), rng.integers(-9000, 9000, (4801, 2)).astype(np.int16), 48000)
    wav.write_wav_pcm16(str(tmp_path / "c.wav"), rng.integers(-9000, 9000, 1601).astype(np.int16), 16000)
    codes = rng.integers(0, 256, 800).astype(np.uint8).tobytes()
    fmt = struct.pack("<HHIIHH", 6, 1, 8000, 8000, 1, 8)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(codes)) + codes
    (tmp_path / "b.wav").write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    df = pd.DataFrame({"deg": ["a.wav", "b.wav", "c.wav"]})
    ds = NL.SpeechQualityDataset(df, data_dir=str(tmp_path), filename_column="deg", mos_column="predict_only",
                                 ms_sr=16000, ms_channel=None)
    for threads in (1, 3):
        clips, srs = NL._load_batch(ds, [0, 1, 2], _HostPool(), 0, threads)
        assert srs == [16000, 16000, 16000]
        for i, name in enumerate(["a.wav", "b.wav", "c.wav"]):
            ref, sr_ref = lb.load(str(tmp_path / name), sr=16000)
            assert sr_ref == 16000 and clips[i].dtype == np.float32
            np.testing.assert_array_equal(clips[i], ref)
            y, sr = ds.load_pcm(i)
            y = y.astype(np.float32) / np.float32(32768.0) if y.dtype == np.int16 else y
            assert sr == 16000
            np.testing.assert_array_equal(y, ref)
    # channel pick + resample == resample of the picked channel (lib:2299-2302)
    ds1 = NL.SpeechQualityDataset(df, data_dir=str(tmp_path), filename_column="deg", mos_column="predict_only",
                                  ms_sr=16000, ms_channel=1)
    clips, _ = NL._load_batch(ds1, [0], _HostPool(), 0, 1)
    y2, _ = lb.load(str(tmp_path / "a.wav"), sr=16000, mono=False)
    np.testing.assert_array_equal(clips[0], y2[1])
    # ms_sr equal to every file's rate: the ordinary (int16) path
    ds2 = NL.SpeechQualityDataset(df.iloc[[2]], data_dir=str(tmp_path), filename_column="deg", mos_column="predict_only",
                                  ms_sr=16000, ms_channel=None)
    clips, srs = NL._load_batch(ds2, [0], _HostPool(), 0, 1)
    assert clips[0].dtype == np.int16 and srs == [16000]


@pytest.mark.parametrize("so,sn,tol", [(16000, 48000, 2e-4), (8000, 16000, 2e-4), (44100, 48000, 2e-4), (48000, 44100, 5e-4),
                                       (48000, 16000, 3e-3)])
def test_restatement_vs_torchaudio_kaiser_sinc(so, sn, tol):
    """Independent implementation of the same filter family: torchaudio's sinc_interp_kaiser with the parameters
    torchaudio documents as the equivalent of resampy's kaiser_best (width 64, rolloff 0.9475937167399596, beta
    14.769656459379492 - its built-in default beta is that very constant).  It evaluates the windowed sinc
    exactly per polyphase branch where resampy interpolates a 512-per-zero-crossing table (a few 1e-5) and, when
    down-sampling, strides that table by int(ratio * 512) (a few 1e-3, see above)."""
    import torch
    import torchaudio.functional as TF
    from nisqa_b200 import synth
    y = synth.synth_speech_pcm16(3, 1.0, so).astype(np.float32) / np.float32(32768.0)
    a = lb.resample(y, so, sn)
    b = TF.resample(torch.from_numpy(y), so, sn, lowpass_filter_width=64, rolloff=0.9475937167399596,
                    resampling_method="sinc_interp_kaiser", beta=14.769656459379492).numpy()
    m = min(len(a), len(b))
    assert abs(len(a) - len(b)) <= 1          # librosa's ceil(n * float ratio) can exceed the exact length by one
    assert np.abs(a[200:m - 200] - b[200:m - 200]).max() <= tol
