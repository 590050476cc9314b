"""Independent cross-checks of the librosa-0.8.1 restatement (oracle/librosa_compat.py).  The
front-end half is "parity unpinned" (real librosa is not installable offline and the reference
holds no vectors for it); these tests bound it against other implementations of the same
published definitions that ARE available offline."""
import os

import numpy as np
import pytest
import scipy.signal
import torch

from oracle import librosa_compat as lb


@pytest.mark.parametrize("sr,fmax", [(48000, 20000), (16000, 8000), (44100, 20000), (8000, 20000)])
def test_mel_filterbank_vs_torchaudio_and_transformers(sr, fmax):
    ours = lb.mel(sr, 4096, n_mels=48, fmin=0.0, fmax=fmax, htk=False, norm="slaney")
    assert ours.dtype == np.float32 and ours.shape == (48, 2049)
    try:
        import torchaudio.functional as AF
        ta = AF.melscale_fbanks(2049, 0.0, float(fmax), 48, sr, norm="slaney", mel_scale="slaney").T.numpy()
        np.testing.assert_allclose(ours, ta, rtol=0, atol=2e-7)
    except ImportError:
        pass
    from transformers.audio_utils import mel_filter_bank
    tr = mel_filter_bank(2049, 48, 0.0, float(fmax), sr, norm="slaney", mel_scale="slaney").T
    np.testing.assert_allclose(ours, tr.astype(np.float32), rtol=0, atol=2e-7)


@pytest.mark.parametrize("win", [960, 320, 882, 441, 160])
def test_window_matches_scipy(win):
    np.testing.assert_allclose(lb.get_window("hann", win, fftbins=True),
                               scipy.signal.get_window("hann", win, fftbins=True), rtol=0, atol=1e-15)


@pytest.mark.parametrize("sr,n", [(48000, 30000), (16000, 9000), (8000, 1500)])
def test_stft_vs_torch_float64(sr, n):
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(n) * 0.1).astype(np.float32)
    hop, win = int(sr * 0.01), int(sr * 0.02)
    ours = lb.stft(y, n_fft=4096, hop_length=hop, win_length=win, window="hann", center=True, pad_mode="reflect")
    assert ours.dtype == np.complex64 and ours.shape == (2049, 1 + n // hop)
    w = torch.from_numpy(lb.get_window("hann", win))
    if n > 2048:   # torch's reflect pad needs pad < n; the repeated-reflection case is numpy-only
        ref = torch.stft(torch.from_numpy(y).double(), 4096, hop_length=hop, win_length=win, window=w,
                         center=True, pad_mode="reflect", return_complex=True).numpy()
        np.testing.assert_allclose(ours, ref.astype(np.complex64), rtol=0, atol=2e-5)
    else:          # n < n_fft/2: numpy re-reflects; check the periodic reflection identity instead
        ypad = np.pad(y, 2048, mode="reflect")
        period = 2 * n - 2
        idx = np.abs(((np.arange(-2048, n + 2048) % period) + period) % period)
        idx = np.where(idx < n, idx, period - idx)
        np.testing.assert_array_equal(ypad, y[idx])


def test_amplitude_to_db_floor_and_clamp():
    S = np.array([[0.0, 1e-6, 1e-4, 1.0, 100.0]], dtype=np.float32)
    db = lb.amplitude_to_db(S.copy(), ref=1.0, amin=1e-4, top_db=80.0)
    assert db.dtype == np.float32
    np.testing.assert_allclose(db[0], [-40.0, -40.0, -40.0, 0.0, 40.0], atol=1e-5)   # max-80 clamp
    db2 = lb.amplitude_to_db(np.array([[0.0, 1e-6, 1e-3]], dtype=np.float32), ref=1.0, amin=1e-4, top_db=80.0)
    np.testing.assert_allclose(db2[0], [-80.0, -80.0, -60.0], atol=1e-4)             # amin floor


def test_load_pcm_conversions(tmp_path):
    from nisqa_b200 import wav
    pcm = (np.arange(-5, 5) * 3000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wav.write_wav_pcm16(p, pcm, 16000)
    y, sr = lb.load(p, sr=None)
    assert sr == 16000 and y.dtype == np.float32
    np.testing.assert_array_equal(y, pcm.astype(np.float32) / 32768.0)
    st = np.stack([pcm, pcm[::-1]], axis=1)
    wav.write_wav_pcm16(p, st, 8000)
    y, sr = lb.load(p, sr=None)
    np.testing.assert_array_equal(y, np.mean(st.T.astype(np.float32) / 32768.0, axis=0))
    y2, _ = lb.load(p, sr=None, mono=False)
    assert y2.shape == (2, 10)


@pytest.mark.parametrize("sr,sec", [(48000, 2.0), (16000, 1.5), (44100, 1.0), (8000, 1.0), (32000, 0.7), (96000, 0.5)])
def test_whole_front_end_vs_transformers_spectrogram(sr, sec):
    """End-to-end cross-check of the restated front end (reflect-padded centred STFT, n_fft 4096 with the
    periodic Hann window of win < n_fft samples, magnitude, Slaney mel, amplitude_to_db with the 80 dB clamp)
    against an independent implementation: transformers.audio_utils.spectrogram / mel_filter_bank /
    amplitude_to_db, which the Hugging Face feature extractors keep equivalent to librosa.  transformers places
    the win-sample frame at the start of the FFT buffer where librosa centres the padded window - a linear
    phase, invisible in the magnitudes - and pads the signal by win/2 instead of n_fft/2, which selects the
    same samples under the window for even win."""
    from transformers import audio_utils as AU
    from oracle import nisqa_oracle as O
    from nisqa_b200 import synth
    import warnings
    args, _ = O.load_checkpoint(os.path.join(O.default_weights_dir(), "nisqa.tar"))
    y = synth.synth_speech_pcm16(5, sec, sr).astype(np.float32) / np.float32(32768.0)
    ref = O.mel_db(y, sr, args)
    hop, win = int(sr * args["ms_hop_length"]), int(sr * args["ms_win_length"])
    assert win % 2 == 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")            # all-zero filters above fmax at low sample rates
        fb = AU.mel_filter_bank(2049, 48, 0.0, float(args["ms_fmax"]), sr, norm="slaney", mel_scale="slaney")
    mel = AU.spectrogram(y.astype(np.float64), AU.window_function(win, "hann", periodic=True), frame_length=win,
                         hop_length=hop, fft_length=4096, power=1.0, center=True, pad_mode="reflect",
                         mel_filters=fb, mel_floor=0.0)
    db = AU.amplitude_to_db(mel, reference=1.0, min_value=1e-4, db_range=80.0)
    assert db.shape == ref.shape == (48, 1 + len(y) // hop)
    assert np.abs(db - ref).max() <= 1e-4          # measured 1e-5 dB
