"""Sample-rate conversion for checkpoints with ``ms_sr != None`` (reference
``nisqa/NISQA_lib.py:2300-2304`` -> ``lb.load(path, sr=ms_sr)`` -> librosa 0.8.1 ``resample`` with
``res_type='kaiser_best'`` = resampy).  The arithmetic runs in the native library (``csrc/resample.cpp``, host
side like the wav decode); this module builds the interpolation table once and wraps the C calls.

None of the shipped checkpoints sets ``ms_sr``; the path exists for user-trained models.
"""
import ctypes as C
import threading

import numpy as np

# resampy's 'kaiser_best' filter (data/kaiser_best.npz = filters.sinc_window with these parameters)
NUM_ZEROS, PRECISION = 64, 9
ROLLOFF, BETA = 0.9475937167399596, 14.769656459379492

_ready = False
_ready_lock = threading.Lock()       # _load_batch runs on several feeder threads: the table is set exactly once


def kaiser_best_half_window():
    """Right half of the Kaiser-windowed sinc, 2**PRECISION samples per zero crossing (float64, 32769 taps)."""
    from scipy.signal.windows import kaiser
    num_bits = 2 ** PRECISION
    n = num_bits * NUM_ZEROS
    sinc_win = ROLLOFF * np.sinc(ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True))
    return np.ascontiguousarray(kaiser(2 * n + 1, BETA)[n:] * sinc_win, dtype=np.float64), num_bits


def _lib():
    global _ready
    from . import engine as _e
    lib = _e.load_library()
    if not _ready:
        with _ready_lock:
            if not _ready:
                win, num_table = kaiser_best_half_window()
                rc = lib.nisqa_resample_set_filter(win.ctypes.data_as(C.POINTER(C.c_double)), win.shape[0], num_table)
                if rc != 0:
                    raise RuntimeError("nisqa_resample_set_filter failed (%d)" % rc)
                _ready = True
    return lib


def out_len(n, sr_orig, sr_new):
    return int(_lib().nisqa_resample_out_len(int(n), int(sr_orig), int(sr_new)))


def resample(x, sr_orig, sr_new, out=None):
    """1-D int16 / float32 ``x`` at ``sr_orig`` -> float32 at ``sr_new`` (``ceil(n * sr_new / sr_orig)`` samples).
    ``out``: optional float32 destination (e.g. a slice of a pinned batch buffer)."""
    lib = _lib()
    xf = np.ascontiguousarray(x.astype(np.float32) / np.float32(32768.0) if x.dtype == np.int16 else x, dtype=np.float32)
    n_out = out_len(xf.shape[0], sr_orig, sr_new)
    if out is None:
        out = np.empty(n_out, dtype=np.float32)
    if out.dtype != np.float32 or not out.flags.c_contiguous or out.shape[0] < n_out:
        raise ValueError("resample: destination must be contiguous float32 with room for %d samples" % n_out)
    got = lib.nisqa_resample_f32(xf.ctypes.data_as(C.POINTER(C.c_float)), xf.shape[0], int(sr_orig), int(sr_new),
                                 out.ctypes.data_as(C.POINTER(C.c_float)), out.shape[0])
    if got < 0:
        raise ValueError("resample failed (%d): %d samples at %d Hz -> %d Hz" % (got, xf.shape[0], sr_orig, sr_new))
    return out[:got]
